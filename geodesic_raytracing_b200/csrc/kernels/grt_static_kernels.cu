// Metric-independent sm_100a kernels, compiled by nvcc into libgrt.so:
//   k_render            `render` cl.cl:5453-5846 (EWA anisotropic background filter, redshift colouring)
//   k_mip_*             background mip pyramid, graphics_settings.cpp:152-212
// /root/reference paths.  The background pyramid is stored compactly (44.7 MB for 4096x2048, which
// stays resident in the 126 MB L2) and sampled with the addressing of the reference's padded
// full-size `image2d_array` layers (level i in the top-left corner, rest clamped to the level's last
// row/column, REPEAT + LINEAR sampler), so wrap-around taps hit the same texels as the reference.
#include <cuda_runtime.h>
#include <stdint.h>

#include "grt_device_common.cuh"
#include "grt_static_kernels.h"

// ---- mip pyramid ---------------------------------------------------------------------------------
// (uchar)(f * 255) after the GL float readback f = b / 255 (graphics_settings.cpp:188-203): a
// truncation, so some byte values come out one lower than they went in. Reproduced exactly
// (round-to-nearest IEEE ops, immune to fast-math).
static __device__ __forceinline__ unsigned char requantise(unsigned char b) {
    float f = __fdiv_rn((float)b, 255.0f);
    f = fminf(fmaxf(f, 0.0f), 1.0f);
    return (unsigned char)__fmul_rn(f, 255.0f);
}

__global__ void k_mip_downsample(const uchar4* __restrict__ src, int sw, int sh, uchar4* __restrict__ dst, int dw, int dh) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
    uchar4 a = src[(size_t)y0 * sw + x0], b = src[(size_t)y0 * sw + x1], c = src[(size_t)y1 * sw + x0], d = src[(size_t)y1 * sw + x1];
    uchar4 o;
    o.x = (unsigned char)((a.x + b.x + c.x + d.x + 2) >> 2);
    o.y = (unsigned char)((a.y + b.y + c.y + d.y + 2) >> 2);
    o.z = (unsigned char)((a.z + b.z + c.z + d.z + 2) >> 2);
    o.w = (unsigned char)((a.w + b.w + c.w + d.w + 2) >> 2);
    dst[(size_t)y * dw + x] = o;
}

__global__ void k_mip_requantise(const uchar4* __restrict__ src, uchar4* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uchar4 v = src[i];
    dst[i] = make_uchar4(requantise(v.x), requantise(v.y), requantise(v.z), requantise(v.w));
}

// ---- render ----------------------------------------------------------------------------------------
static __device__ __forceinline__ float4 texel_f(const GrtTexture& t, int level, int x, int y) {
    int lx = min(x, t.lw[level] - 1), ly = min(y, t.lh[level] - 1);
    uchar4 c = __ldg(t.data + t.offset[level] + (size_t)ly * t.lw[level] + lx);
    const float k = 1.0f / 255.0f;
    return make_float4(c.x * k, c.y * k, c.z * k, c.w * k);
}
// read_imagef(NORMALIZED | REPEAT | LINEAR) on full-size layer `level`
static __device__ float4 sample_layer(const GrtTexture& t, float u, float v, int level) {
    level = min(max(level, 0), t.levels - 1);
    float fu = (u - floorf(u)) * t.w - 0.5f, fv = (v - floorf(v)) * t.h - 0.5f;
    float flu = floorf(fu), flv = floorf(fv);
    int i0 = (int)flu, j0 = (int)flv, i1 = i0 + 1, j1 = j0 + 1;
    float a = fu - flu, b = fv - flv;
    if (i0 < 0) i0 += t.w;
    if (i1 > t.w - 1) i1 -= t.w;
    if (j0 < 0) j0 += t.h;
    if (j1 > t.h - 1) j1 -= t.h;
    float4 t00 = texel_f(t, level, i0, j0), t10 = texel_f(t, level, i1, j0), t01 = texel_f(t, level, i0, j1), t11 = texel_f(t, level, i1, j1);
    float w00 = (1 - a) * (1 - b), w10 = a * (1 - b), w01 = (1 - a) * b, w11 = a * b;
    return make_float4(w00 * t00.x + w10 * t10.x + w01 * t01.x + w11 * t11.x, w00 * t00.y + w10 * t10.y + w01 * t01.y + w11 * t11.y,
                       w00 * t00.z + w10 * t10.z + w01 * t01.z + w11 * t11.z, w00 * t00.w + w10 * t10.w + w01 * t01.w + w11 * t11.w);
}
static __device__ float4 read_mipmap(const GrtTexture& t, float u, float v, float lod) {  // cl.cl:5421-5449
    lod = fmaxf(lod, 0.0f);
    u = fmodf(u, 1.0f);
    v = fmodf(v, 1.0f);
    float lo = floorf(lod), hi = ceilf(lod);
    float dlo = exp2f(lo), dhi = exp2f(hi);
    float4 a = sample_layer(t, u / dlo, v / dlo, (int)lo);
    if (hi == lo) return a;   // mix weight is exactly 0
    float4 b = sample_layer(t, u / dhi, v / dhi, (int)hi);
    return lerp4(a, b, lod - lo);
}
static __device__ __forceinline__ float circular_diff1(float f1, float f2) {  // cl.cl:3598-3605, period 1
    const float two_pi = 6.283185307179586f;
    float d = f2 * two_pi - f1 * two_pi;
    float s, c;
    sincosf(d, &s, &c);
    return atan2f(s, c) / two_pi;
}
static __device__ __forceinline__ float srgb_to_lin1(float in) { return in < 0.04045f ? in / 12.92f : powf((in + 0.055f) / 1.055f, 2.4f); }
static __device__ __forceinline__ float lin_to_srgb1(float in) { return in <= 0.0031308f ? in * 12.92f : 1.055f * powf(in, 1.0f / 2.4f) - 0.055f; }
static __device__ __forceinline__ float energy_of(float3 v) { return v.x * 0.2125f + v.y * 0.7154f + v.z * 0.0721f; }
static __device__ __forceinline__ float3 clamp01(float3 v) { return f3(clampf(v.x, 0, 1), clampf(v.y, 0, 1), clampf(v.z, 0, 1)); }
static __device__ float3 redshift_colour(float3 v, float z, int use_old) {  // cl.cl:5371-5413
    float e = energy_of(v);
    float3 result;
    if (z > 0) result = lerp3(v, f3(e / 0.2125f, 0, 0), tanhf(z));
    else {
        float frac = tanhf(1.0f / (1.0f + z) - 1.0f);
        float3 col = lerp3(v, f3(0, 0, e * (float)(1 / 0.0721)), frac);
        if (!use_old) {
            float remaining = energy_of(col) - energy_of(clamp01(col));
            col.x += remaining * (1 / 0.2125f);
            col.y += remaining * (float)(1 / 0.7154);
        }
        result = col;
    }
    return clamp01(result);
}

#define RT_W 32
#define RT_H 8

// One thread per output pixel. The (u,v) records of the 32x8 tile plus a one-pixel apron on the
// right/bottom are staged through shared memory (each record is read by up to three pixels).
__global__ void __launch_bounds__(RT_W * RT_H) k_render(const GrtRenderData* __restrict__ rdata, GrtFrame fr, GrtRenderStripe st, GrtTexture tex1, GrtTexture tex2, void* out, int out_format) {
    __shared__ float4 tile[RT_H + 1][RT_W + 1];
    const int w = fr.full_w, h = fr.full_h;
    const int bx = blockIdx.x * RT_W, by = blockIdx.y * RT_H;   // by: row inside the stripe
    const int tid = threadIdx.y * RT_W + threadIdx.x;
    for (int i = tid; i < (RT_W + 1) * (RT_H + 1); i += RT_W * RT_H) {
        int tx = i % (RT_W + 1), ty = i / (RT_W + 1);
        int gx = bx + tx, lrow = st.out_local0 + by + ty;
        float4 v = make_float4(0, 0, 0, 0);
        if (gx < w && lrow < st.win_local_end) v = rdata[(size_t)lrow * w + gx];
        tile[ty][tx] = v;
    }
    __syncthreads();
    const int sx = bx + threadIdx.x, srow = by + threadIdx.y;
    if (sx >= w || srow >= st.out_rows) return;
    const int sy = st.out_global0 + srow;          // full-frame row
    const int lrow = st.out_local0 + srow;         // local row in rdata
    const float4 rd = tile[threadIdx.y][threadIdx.x];
    const int flags = __float_as_int(rd.w);
    const int terminated = flags & 0xff, side = flags >> 8;
    float4 result;
    if (terminated != 1) result = make_float4(0, 0, 0, 1);
    else {
        const float sxf = rd.x, syf = rd.y;
        // texture-space Jacobian from the right / below neighbours (left / above on the last column / row)
        float2 tr, bl;
        const bool flipx = sx == w - 1, flipy = sy == h - 1;
        if (!flipx) { float4 n = tile[threadIdx.y][threadIdx.x + 1]; tr = make_float2(n.x, n.y); }
        else { float4 n = rdata[(size_t)lrow * w + sx - 1]; tr = make_float2(n.x, n.y); }
        if (!flipy) { float4 n = tile[threadIdx.y + 1][threadIdx.x]; bl = make_float2(n.x, n.y); }
        else { float4 n = rdata[(size_t)(lrow - 1) * w + sx]; bl = make_float2(n.x, n.y); }
        const float bias_frac = 1.3f;
        float du_dx = circular_diff1(sxf, tr.x) / bias_frac, dv_dx = circular_diff1(syf, tr.y) / bias_frac;
        float du_dy = circular_diff1(sxf, bl.x) / bias_frac, dv_dy = circular_diff1(syf, bl.y) / bias_frac;
        if (flipx) { du_dx = -du_dx; dv_dx = -dv_dx; }
        if (flipy) { du_dy = -du_dy; dv_dy = -dv_dy; }
        const GrtTexture& tex = side > 0 ? tex1 : tex2;
        du_dx *= tex.w; du_dy *= tex.w;
        dv_dx *= tex.h; dv_dy *= tex.h;
        // Heckbert EWA ellipse (cl.cl:5570-5623)
        float Ann = dv_dx * dv_dx + dv_dy * dv_dy + 1;
        float Bnn = -2 * (du_dx * dv_dx + du_dy * dv_dy);
        float Cnn = du_dx * du_dx + du_dy * du_dy + 1;
        float F = Ann * Cnn - Bnn * Bnn / 4;
        float A = Ann / F, B = Bnn / F, C = Cnn / F;
        float root = sqrtf((A - C) * (A - C) + B * B);
        float a_prime = (A + C - root) / 2, c_prime = (A + C + root) / 2;
        float majorRadius = rsqrtf(a_prime), minorRadius = rsqrtf(c_prime);
        float theta = atan2f(B, (A - C) / 2);
        majorRadius = fmaxf(majorRadius, 1.0f);
        minorRadius = fmaxf(minorRadius, 1.0f);
        majorRadius = fmaxf(majorRadius, minorRadius);
        float fProbes = 2 * (majorRadius / minorRadius) - 1;
        int iProbes = (int)floorf(fProbes + 0.5f);
        iProbes = min(iProbes, fr.max_probes);
        if (iProbes < fProbes) minorRadius = 2 * majorRadius / (iProbes + 1);
        float lod = log2f(minorRadius);
        const int maxLod = tex.levels - 1;
        if (lod > maxLod) { lod = (float)maxLod; iProbes = 1; }
        if (iProbes <= 1) {
            if (iProbes < 1) lod = (float)maxLod;
            result = read_mipmap(tex, sxf, syf, lod);
        } else {
            float lineLength = 2 * (majorRadius - minorRadius);
            float st, ct;
            sincosf(theta, &st, &ct);
            float du = ct * lineLength / (iProbes - 1), dv = st * lineLength / (iProbes - 1);
            int currentN = (iProbes & 1) ? -2 * ((iProbes - 1) / 2) : -2 * (iProbes / 2) - 1;
            float sU = du / tex.w, sV = dv / tex.h;
            float4 total = make_float4(0, 0, 0, 0);
            float accumulated = 0;
            const float k = (du * du + dv * dv) / (majorRadius * majorRadius);
            for (int cnt = 0; cnt < iProbes; cnt++) {
                float d_2 = (currentN * currentN / 4.0f) * k;
                float weight = expf(-2.0f * d_2);
                float4 fval = read_mipmap(tex, sxf + (currentN / 2.0f) * sU, syf + (currentN / 2.0f) * sV, lod);
                total = total + weight * fval;
                accumulated += weight;
                currentN += 2;
            }
            result = total * (1.0f / accumulated);
        }
        float3 rgb = f3(result.x, result.y, result.z);
        if (fr.redshift) {
            // cl.cl:5697-5838 with LINEAR_FRAMEBUFFER
            float z = rd.z;
            float3 lin = f3(srgb_to_lin1(rgb.x), srgb_to_lin1(rgb.y), srgb_to_lin1(rgb.z));
            float lum = 0.2126f * lin.x + 0.7152f * lin.y + 0.0722f * lin.z;
            // I' = I (lambda_obs / lambda_emit)^3 with lambda_obs = lambda_emit / (1 + z)
            float ratio = 1.0f / (z + 1.0f);
            float new_lum = clampf(ratio * ratio * ratio * lum, 0.0f, 1.0f);
            if (lum > 0.00001f) lin = clamp01(lin * (new_lum / lum));
            rgb = clamp01(redshift_colour(lin, z, fr.use_old_redshift));
        } else rgb = f3(srgb_to_lin1(rgb.x), srgb_to_lin1(rgb.y), srgb_to_lin1(rgb.z));
        result = make_float4(rgb.x, rgb.y, rgb.z, result.w);
    }
    const size_t o = (size_t)st.out_offset_px + (size_t)srow * w + sx;
    if (out_format == GRT_FORMAT_RGBA32F) ((float4*)out)[o] = result;
    else {
        // screenshot path: clamp, lin_to_srgb, clamp, * 255 truncated (main.cpp:2788-2797)
        float c[4] = {result.x, result.y, result.z, result.w};
        unsigned char b[4];
        for (int i = 0; i < 4; i++) {
            float v = clampf(c[i], 0.0f, 1.0f);
            v = clampf(lin_to_srgb1(v), 0.0f, 1.0f);
            b[i] = (unsigned char)(v * 255.0f);
        }
        ((uchar4*)out)[o] = make_uchar4(b[0], b[1], b[2], b[3]);
    }
}

// ---- host launchers (called from device_rt.cpp) ------------------------------------------------------
extern "C" void grt_launch_mip_downsample(const void* src, int sw, int sh, void* dst, int dw, int dh, cudaStream_t s) {
    dim3 block(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    k_mip_downsample<<<grid, block, 0, s>>>((const uchar4*)src, sw, sh, (uchar4*)dst, dw, dh);
}
extern "C" void grt_launch_mip_requantise(const void* src, void* dst, size_t n, cudaStream_t s) {
    k_mip_requantise<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const uchar4*)src, (uchar4*)dst, n);
}
extern "C" void grt_launch_render(const void* rdata, const GrtFrame* fr, const GrtRenderStripe* st, const GrtTexture* t1, const GrtTexture* t2, void* out, int out_format,
                                  cudaStream_t s) {
    dim3 block(RT_W, RT_H), grid((fr->full_w + RT_W - 1) / RT_W, (st->out_rows + RT_H - 1) / RT_H);
    k_render<<<grid, block, 0, s>>>((const GrtRenderData*)rdata, *fr, *st, *t1, *t2, out, out_format);
}
