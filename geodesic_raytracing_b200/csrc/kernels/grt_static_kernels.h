// Host-visible declarations for the statically compiled kernels (grt_static_kernels.cu).
#pragma once
#include <stddef.h>

#define GRT_FORMAT_RGBA32F 0   // linear float4, what the reference's `render` writes (LINEAR_FRAMEBUFFER)
#define GRT_FORMAT_RGBA8 1     // sRGB-encoded bytes, the reference's screenshot conversion (main.cpp:2788-2797)

#define GRT_MAX_MIP_LEVELS 10  // MIP_LEVELS, graphics_settings.cpp:163

// Compact mip pyramid: level l is lw[l] x lh[l] RGBA8 texels at data + offset[l].
struct GrtTexture {
    const uchar4* data;
    int w, h, levels;
    int lw[GRT_MAX_MIP_LEVELS], lh[GRT_MAX_MIP_LEVELS];
    size_t offset[GRT_MAX_MIP_LEVELS];
};

#ifdef __cplusplus
extern "C" {
#endif
struct GrtFrame;
struct GrtRenderStripe;
void grt_launch_mip_downsample(const void* src, int sw, int sh, void* dst, int dw, int dh, cudaStream_t s);
void grt_launch_mip_requantise(const void* src, void* dst, size_t n, cudaStream_t s);
void grt_launch_render(const void* rdata, const struct GrtFrame* fr, const struct GrtRenderStripe* st, const struct GrtTexture* t1, const struct GrtTexture* t2, void* out,
                       int out_format, cudaStream_t s);
#ifdef __cplusplus
}
#endif
