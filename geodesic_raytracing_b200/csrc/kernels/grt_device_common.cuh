// Shared device-side definitions for the sm_100a kernels: frame description, SoA ray state,
// small vector helpers.  Included both by the NVRTC-compiled metric module
// (grt_metric_kernels.cuh, one cubin per metric) and by the statically compiled,
// metric-independent kernels (grt_static_kernels.cu).  No host headers: NVRTC has none.
#pragma once

// ---- data layout in HBM -----------------------------------------------------------------------
// Ray state is SoA in float4 planes so every access is one coalesced 128-bit load/store per lane
// (the reference keeps a 96-byte AoS `struct lightray`, cl.cl:813-824 / render_state.hpp:8-19):
//   pos[i], vel[i], acc[i]           position, velocity, acceleration
//   quat[i]                          initial_quat, only allocated/touched for constant-theta metrics
//   aux[i] = {ku_uobsu, running_dlambda_dnew, terminated, sy<<16|sx}
// terminated: 0 live or dropped, 1 finished, 2 prepass-culled (as in the reference), 3 padding slot.
struct GrtRays {
    float4* pos;
    float4* vel;
    float4* acc;
    float4* quat;
    int4* aux;
};

// per-pixel shading record, 16 B (the reference's `struct render_data` is 32 B, cl.cl:5066-5074;
// sx, sy are implied by the index): {u, v, z_shift, terminated | side << 8}
typedef float4 GrtRenderData;

// runtime features (dynamic_feature_config, main.cpp:1123-1158) + frame/window geometry
struct GrtFrame {
    int full_w, full_h;     // full frame size: pixel directions and border rules use these
    int y0, y1;             // rows [y0, y1) of the full frame held by the local buffers (stripe + halo)
    int cell;               // 1: one first-pass ray per pixel, 2: one per 2x2 cell (adaptive sampling)
    int grid_w, grid_h;     // first-pass grid (cells) covering the window
    int grid_y0;            // first cell row of the window in full-frame cell coordinates
    int tiles_x;            // 8x4-cell tiles per grid row
    int n_slots;            // tiles_x * tiles_y * 32 first-pass ray slots (padding slots are terminated=3)
    int prepass_w, prepass_h;
    int use_prepass_mask;   // main pass consults the prepass termination mask
    int flip_geodesic_direction;
    int max_probes;
    int out_y0, out_y1;     // rows actually written to the output image (stripe without halo)
    float universe_size, max_acceleration_change, max_precision_radius, min_step;
    float adaptive_sampling_threshold, field_of_view;
    int redshift, reparameterisation, use_old_redshift, adaptive_sampling;
};

struct GrtCamera {
    float4 pos_generic;     // camera position in the metric's native coordinates
    float4 e[4];            // tetrad, contravariant
    float4 quat;            // camera orientation (x, y, z, w)
    float4 basis_speed;     // observer 3-velocity in the tetrad frame
    float4 pos_cart;        // (t, x, y, z)
    float flip;
    int timelike_coordinate;
    int pad0, pad1;
};

// statistics accumulated by the integration kernel
struct GrtStats {
    unsigned long long trial_steps;     // Verlet bodies executed, rejected retries included
    unsigned long long accepted_steps;
    unsigned long long rays_integrated; // rays that entered the loop
    unsigned long long rays_finished;   // terminated == 1
};

#define GRT_PI_F 3.14159265358979323846f

// 8x4 tiled slot order: a warp's 32 rays cover a compact 8x4 block of cells so their step counts
// correlate (row-major order would put 32 horizontally adjacent pixels in a warp).
static __device__ __forceinline__ int grt_slot_of(int gx, int gy, int tiles_x) {
    return (((gy >> 2) * tiles_x + (gx >> 3)) << 5) | ((gy & 3) << 3) | (gx & 7);
}
static __device__ __forceinline__ void grt_cell_of(int slot, int tiles_x, int& gx, int& gy) {
    int tile = slot >> 5, l = slot & 31;
    gx = ((tile % tiles_x) << 3) | (l & 7);
    gy = ((tile / tiles_x) << 2) | (l >> 3);
}

// ---- float3 / float4 helpers ------------------------------------------------------------------
static __device__ __forceinline__ float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
static __device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
static __device__ __forceinline__ float4 f4(float x, float3 v) { return make_float4(x, v.x, v.y, v.z); }
static __device__ __forceinline__ float3 yzw(float4 v) { return make_float3(v.y, v.z, v.w); }
static __device__ __forceinline__ float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static __device__ __forceinline__ float4 operator-(float4 a, float4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static __device__ __forceinline__ float4 operator-(float4 a) { return f4(-a.x, -a.y, -a.z, -a.w); }
static __device__ __forceinline__ float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
static __device__ __forceinline__ float4 operator*(float s, float4 a) { return a * s; }
static __device__ __forceinline__ float4 operator*(float4 a, float4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static __device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
static __device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
static __device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
static __device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
static __device__ __forceinline__ float3 operator*(float s, float3 a) { return a * s; }
static __device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static __device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static __device__ __forceinline__ float3 cross3(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static __device__ __forceinline__ float len3(float3 a) { return sqrtf(dot3(a, a)); }
static __device__ __forceinline__ float3 norm3(float3 a) { return a * (1.0f / len3(a)); }
static __device__ __forceinline__ float4 norm4(float4 a) { return a * (1.0f / sqrtf(dot4(a, a))); }
static __device__ __forceinline__ float lerpf(float a, float b, float t) { return a + (b - a) * t; }
static __device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) { return a + (b - a) * t; }
static __device__ __forceinline__ float3 lerp3(float3 a, float3 b, float t) { return a + (b - a) * t; }
static __device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
static __device__ __forceinline__ float sgnf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static __device__ __forceinline__ bool bad_float(float x) { return !(fabsf(x) <= 3.402823466e+38f); }  // NaN or Inf
static __device__ __forceinline__ bool bad4(float4 v) { return bad_float(v.x) || bad_float(v.y) || bad_float(v.z) || bad_float(v.w); }
static __device__ __forceinline__ float comp(float4 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
