// Shared device-side definitions for the sm_100a kernels: frame description, SoA ray state,
// small vector helpers.  Included both by the NVRTC-compiled metric module
// (grt_metric_kernels.cuh, one cubin per metric) and by the statically compiled,
// metric-independent kernels (grt_static_kernels.cu).  No host headers: NVRTC has none.
#pragma once

#include "grt_device_structs.h"

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
