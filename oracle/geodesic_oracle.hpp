// =============================================================================================
// CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Plain C++ restatement of the reference's OpenCL kernels (/root/reference/cl.cl) for the
// geodesic raytracing hot path, one function per reference function, each citing the
// cl.cl lines it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// / --impl reference legs may build or call this.
//
// PARITY: PINNED AGAINST THE REFERENCE'S OWN DEVICE CODE.  The reference ships no golden vectors or
// known-answer tests for this path, but its kernels can be run here: oracle/ref_build.py
// compiles /root/reference/cl.cl itself as C++ for the host (oracle/ref_shim/: an OpenCL-C-as-C++
// environment plus the -D macro ABI over the generated metric functions) into oracle/_ref/, and
// ref_driver.cpp runs the reference's kernel sequence.  This restatement reproduces those kernels BIT
// FOR BIT - frames (render_data, image, prepass mask, ray / step counts), tetrads, the geodesic camera
// path - on every bundled metric (tests/test_reference_kernels.py, live and against the committed
// fixtures tests/golden/clref_frames.npz).  What stays unpinned is upstream of cl.cl: the symbolic
// layer deps/vec (empty submodule; our emitter supplies the expressions, cross-checked numerically) and
// the GL driver's mip generation (graphics_settings.cpp:161-186).  Analytic anchors as well
// (tests/test_oracle_physics.py): straight rays in Minkowski, the Schwarzschild shadow radius,
// null-norm and conserved quantities in Kerr-Schild, AD-vs-finite-difference partials.
//
// The metric itself comes from the generated header (GRT_METRIC_HEADER): the same emitted
// functions the CUDA kernels compile, with this file using the *reference formulation*
// `grt_accel_ref` (symbolic Christoffel with symbolic inverse = GEO_ACCEL0..3,
// /root/reference/metric.hpp:184-244) in the integration loop, as the reference does.
// Arithmetic is FP32 throughout, like the OpenCL kernels; native_* functions map to the
// correctly rounded libm ones.
// =============================================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifndef GRT_METRIC_HEADER
#error "compile with -DGRT_METRIC_HEADER=\"<generated metric header>\""
#endif

#define GRT_FN static inline
#define GRT_DIV(a, b) ((a) / (b))
#define GRT_SIN(a) sinf(a)
#define GRT_COS(a) cosf(a)
#define GRT_TAN(a) tanf(a)
#define GRT_LOG(a) logf(a)
#define GRT_SQRT(a) sqrtf(a)
#define GRT_EXP(a) expf(a)
#define GRT_TANH(a) tanhf(a)
#define GRT_POW(a, b) powf(a, b)
#define GRT_SIGN(a) ((a) > 0.0f ? 1.0f : ((a) < 0.0f ? -1.0f : 0.0f))
#define GRT_INF INFINITY
#define GRT_NAN NAN
#include GRT_METRIC_HEADER

namespace oracle {

static const float PI_F = 3.14159265358979323846f;

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };

static inline float3 f3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 f4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float4 f4(float x, float3 v) { return float4{x, v.x, v.y, v.z}; }
static inline float3 yzw(float4 v) { return float3{v.y, v.z, v.w}; }
static inline float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
static inline float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
static inline float3 operator*(float s, float3 a) { return a * s; }
static inline float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
static inline float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline float4 operator-(float4 a, float4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline float4 operator-(float4 a) { return f4(-a.x, -a.y, -a.z, -a.w); }
static inline float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
static inline float4 operator*(float s, float4 a) { return a * s; }
static inline float4 operator*(float4 a, float4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline float4 operator/(float4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
static inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static inline float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length(float3 a) { return sqrtf(dot(a, a)); }
static inline float length(float4 a) { return sqrtf(dot(a, a)); }
static inline float3 normalize(float3 a) { return a / length(a); }
static inline float4 normalize(float4 a) { return a / length(a); }
static inline float mix(float a, float b, float t) { return a + (b - a) * t; }
static inline float3 mix(float3 a, float3 b, float t) { return a + (b - a) * t; }
static inline float4 mix(float4 a, float4 b, float t) { return a + (b - a) * t; }
static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
static inline float3 clamp3(float3 v, float lo, float hi) { return f3(clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)); }
static inline float signf(float x) { return x > 0 ? 1.f : (x < 0 ? -1.f : 0.f); }  // OpenCL sign()
static inline bool is_degenerate(float x) { return std::isnan(x) || !std::isfinite(x); }  // cl.cl:64
static inline bool any_degenerate(float4 v) { return is_degenerate(v.x) || is_degenerate(v.y) || is_degenerate(v.z) || is_degenerate(v.w); }
static inline void to_arr(float4 v, float* a) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
static inline float4 from_arr(const float* a) { return f4(a[0], a[1], a[2], a[3]); }

// ---- runtime configuration -------------------------------------------------------------------
// dynamic_feature_config (main.cpp:1123-1158) without the triangle-renderer features
struct Features {
    float universe_size = 20.f;
    float max_acceleration_change = 0.01f;
    float max_precision_radius = 10.f;
    float min_step = 0.000001f;
    float adaptive_sampling_threshold = 32.f;
    float field_of_view = 90.f;
    int redshift = 0;
    int reparameterisation = 0;
    int use_old_redshift = 0;
    int adaptive_sampling = 1;
};

struct lightray {  // cl.cl:813-824
    float4 position, velocity, initial_quat, acceleration;
    float ku_uobsu, running_dlambda_dnew;
    int terminated, sx, sy;
};

struct render_data {  // cl.cl:5066-5074
    float2 tex_coord;
    float z_shift;
    int sx, sy, terminated, side;
};

struct Counters {
    long long trial_steps = 0, accepted_steps = 0;
};

// ---- small helpers ---------------------------------------------------------------------------
// cl.cl:105-123
static inline float3 cartesian_to_polar(float3 in) {
    float r = length(in);
    float theta = acosf(in.z / r);
    float phi = atan2f(in.y, in.x);
    return f3(r, theta, phi);
}
static inline float3 polar_to_cartesian(float3 in) {
    float x = in.x * sinf(in.y) * cosf(in.z);
    float y = in.x * sinf(in.y) * sinf(in.z);
    float z = in.x * cosf(in.y);
    return f3(x, y, z);
}
// cl.cl:121-139
static inline float3 cartesian_velocity_to_polar_velocity(float3 p, float3 v) {
    float r = length(p);
    float repeated_eq = r * sqrtf(1 - (p.z * p.z / (r * r)));
    float rdot = (p.x * v.x + p.y * v.y + p.z * v.z) / r;
    float tdot = ((p.z * rdot) / (r * repeated_eq)) - v.z / repeated_eq;
    float pdot = (p.x * v.y - p.y * v.x) / (p.x * p.x + p.y * p.y);
    return f3(rdot, tdot, pdot);
}
// cl.cl:176-190
static inline float3 rot_quat_norm(float3 point, float4 q) {
    float3 qv = f3(q.x, q.y, q.z);
    float3 t = 2.f * cross(qv, point);
    return point + q.w * t + cross(qv, t);
}
static inline float3 rot_quat(float3 point, float4 quat) { return rot_quat_norm(point, normalize(quat)); }
// cl.cl:192-209
static inline float3 spherical_velocity_to_cartesian_velocity(float3 p, float3 dp) {
    float r = p.x, dr = dp.x, x = p.y, dx = dp.y, y = p.z, dy = dp.z;
    float v1 = -r * sinf(x) * sinf(y) * dy + r * cosf(x) * cosf(y) * dx + sinf(x) * cosf(y) * dr;
    float v2 = sinf(x) * sinf(y) * dr + r * sinf(x) * cosf(y) * dy + r * cosf(x) * sinf(y) * dx;
    float v3 = cosf(x) * dr - r * sinf(x) * dx;
    return f3(v1, v2, v3);
}
// cl.cl:211-237
static inline float3 fix_ray_position_cart(float3 cartesian_pos, float3 cartesian_velocity, float sphere_radius) {
    cartesian_velocity = normalize(cartesian_velocity);
    float3 C = f3(0, 0, 0);
    float a = 1;
    float b = 2 * dot(cartesian_velocity, (cartesian_pos - C));
    float c = dot(C, C) + dot(cartesian_pos, cartesian_pos) - 2 * (dot(cartesian_pos, C)) - sphere_radius * sphere_radius;
    float discrim = b * b - 4 * a * c;
    if (discrim < 0) return cartesian_pos;
    float t0 = (-b - sqrtf(discrim)) / (2 * a);
    float t1 = (-b + sqrtf(discrim)) / (2 * a);
    float my_t = fabsf(t0) < fabsf(t1) ? t0 : t1;
    return cartesian_pos + my_t * cartesian_velocity;
}
// cl.cl:239-263
static inline float3 fix_ray_position(float3 polar_pos, float3 polar_velocity, float sphere_radius) {
    float position_sign = signf(polar_pos.x);
    float3 cpolar_pos = polar_pos;
    cpolar_pos.x = fabsf(cpolar_pos.x);
    polar_velocity.x *= position_sign;
    float3 cartesian_velocity = spherical_velocity_to_cartesian_velocity(cpolar_pos, polar_velocity);
    float3 cartesian_pos = polar_to_cartesian(cpolar_pos);
    float3 new_cart = fix_ray_position_cart(cartesian_pos, cartesian_velocity, sphere_radius);
    float3 new_polar = cartesian_to_polar(new_cart);
#if GRT_CONSTANT_THETA
    new_polar.y = PI_F / 2;
#endif
    new_polar.x *= position_sign;
    return new_polar;
}
// cl.cl:326-350
static inline float lin_to_srgb_single(float in) { return in <= 0.0031308f ? in * 12.92f : 1.055f * powf(in, 1.0f / 2.4f) - 0.055f; }
static inline float srgb_to_lin_single(float in) { return in < 0.04045f ? in / 12.92f : powf((in + 0.055f) / 1.055f, 2.4f); }
static inline float3 srgb_to_lin(float3 in) { return f3(srgb_to_lin_single(in.x), srgb_to_lin_single(in.y), srgb_to_lin_single(in.z)); }
static inline float3 lin_to_srgb(float3 in) { return f3(lin_to_srgb_single(in.x), lin_to_srgb_single(in.y), lin_to_srgb_single(in.z)); }

static inline float4 tensor_contract(const float t16[16], float4 vec) {  // cl.cl:452-462
    float4 res;
    res.x = t16[0] * vec.x + t16[1] * vec.y + t16[2] * vec.z + t16[3] * vec.w;
    res.y = t16[4] * vec.x + t16[5] * vec.y + t16[6] * vec.z + t16[7] * vec.w;
    res.z = t16[8] * vec.x + t16[9] * vec.y + t16[10] * vec.z + t16[11] * vec.w;
    res.w = t16[12] * vec.x + t16[13] * vec.y + t16[14] * vec.z + t16[15] * vec.w;
    return res;
}

// cl.cl:469-558 (symmetric input: only the upper triangle is read)
static inline void metric_inverse(const float m[16], float invOut[16]) {
    float inv[16], det;
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[11] - m[6] * m[6] * m[15] + m[6] * m[7] * m[11] + m[7] * m[6] * m[11] - m[7] * m[7] * m[10];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[11] + m[6] * m[2] * m[15] - m[6] * m[3] * m[11] - m[7] * m[2] * m[11] + m[7] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[11] - m[2] * m[2] * m[15] + m[2] * m[3] * m[11] + m[3] * m[2] * m[11] - m[3] * m[3] * m[10];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[11] - m[5] * m[2] * m[15] + m[5] * m[3] * m[11] + m[7] * m[2] * m[7] - m[7] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[11] + m[1] * m[2] * m[15] - m[1] * m[3] * m[11] - m[3] * m[2] * m[7] + m[3] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[7] - m[1] * m[1] * m[15] + m[1] * m[3] * m[7] + m[3] * m[1] * m[7] - m[3] * m[3] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[6] * m[2] * m[7] + m[6] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[1] * m[2] * m[11] + m[1] * m[3] * m[10] + m[2] * m[2] * m[7] - m[2] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[6] + m[1] * m[1] * m[11] - m[1] * m[3] * m[6] - m[2] * m[1] * m[7] + m[2] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[6] - m[1] * m[1] * m[10] + m[1] * m[2] * m[6] + m[2] * m[1] * m[6] - m[2] * m[2] * m[5];
    inv[4] = inv[1]; inv[8] = inv[2]; inv[12] = inv[3]; inv[9] = inv[6]; inv[13] = inv[7]; inv[14] = inv[11];
    det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    det = 1.0f / det;
    for (int i = 0; i < 16; i++) invOut[i] = inv[i] * det;
}

// cl.cl:560-683: general 4x4 inverse (cofactor expansion; the classic gluInvertMatrix layout)
static inline void matrix_inverse(const float m[16], float invOut[16]) {
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    det = 1.0f / det;
    for (int i = 0; i < 16; i++) invOut[i] = inv[i] * det;
}

// ---- metric access (cl.cl:969-1199): generated functions; the diagonal / big split of the
// reference only changes which entries are non-zero, the full 16/64 arrays are equivalent ----
static inline void calculate_metric_big(float4 pos, float g[16], const grt_params& cfg) {
    float p[4];
    to_arr(pos, p);
    grt_metric(p, cfg, g);
}
static inline void calculate_partial_derivatives_big(float4 pos, float dg[64], const grt_params& cfg) {
    float p[4];
    to_arr(pos, p);
    grt_metric_partials(p, cfg, dg);
}
static inline float4 generic_to_spherical(float4 in, const grt_params& cfg) {  // cl.cl:1202-1216
    float p[4], o[4];
    to_arr(in, p);
    grt_to_polar(p, cfg, o);
    return from_arr(o);
}
static inline float4 generic_velocity_to_spherical_velocity(float4 in, float4 inv, const grt_params& cfg) {  // cl.cl:1218-1237
    float p[4], v[4], o[4];
    to_arr(in, p); to_arr(inv, v);
    grt_to_polar_vel(p, v, cfg, o);
    return from_arr(o);
}
static inline float4 spherical_to_generic(float4 in, const grt_params& cfg) {  // cl.cl:1239-1252
    float p[4], o[4];
    to_arr(in, p);
    grt_from_polar(p, cfg, o);
    return from_arr(o);
}
static inline float4 spherical_velocity_to_generic_velocity(float4 in, float4 inv, const grt_params& cfg) {  // cl.cl:1254-1273
    float p[4], v[4], o[4];
    to_arr(in, p); to_arr(inv, v);
    grt_from_polar_vel(p, v, cfg, o);
    return from_arr(o);
}
static inline float4 get_coordinate_period(const grt_params& cfg) {  // cl.cl:1338-1355
    float o[4];
    grt_period(cfg, o);
    return from_arr(o);
}
static inline float get_distance_to_object(float4 polar, const grt_params& cfg) {  // cl.cl:3377-3387
    float p[4], o[1];
    to_arr(polar, p);
    grt_distance(p, cfg, o);
    return o[0];
}

static inline float4 lower_index_big(float4 vec, const float g[16]) {  // cl.cl:828-848
    float v[4], r[4];
    to_arr(vec, v);
    for (int i = 0; i < 4; i++) {
        float sum = 0;
        for (int j = 0; j < 4; j++) sum += g[i * 4 + j] * v[j];
        r[i] = sum;
    }
    return from_arr(r);
}
static inline float dot_product_big(float4 u, float4 v, const float g[16]) { return dot(lower_index_big(u, g), v); }  // cl.cl:897-902

// cl.cl:1443-1537 (GENERIC_BIG_METRIC) and cl.cl:738-797 (diagonal): numeric Christoffel symbols
// and a^mu = -Gamma^mu_ab v^a v^b, used at ray initialisation.
static inline float4 calculate_acceleration_numeric(float4 lightray_velocity, const float g[16], const float dg[64]) {
#if GRT_CONSTANT_THETA
    lightray_velocity.z = 0;
#endif
    float christoff[64];
#if GRT_BIG_METRIC
    float g_inv[16];
    metric_inverse(g, g_inv);
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 4; k++)
            for (int l = 0; l < 4; l++) {
                float sum = 0;
                for (int m = 0; m < 4; m++) {
                    sum += g_inv[i * 4 + m] * dg[l * 16 + m * 4 + k];
                    sum += g_inv[i * 4 + m] * dg[k * 16 + m * 4 + l];
                    sum -= g_inv[i * 4 + m] * dg[m * 16 + k * 4 + l];
                }
                christoff[i * 16 + k * 4 + l] = 0.5f * sum;
            }
#else
    for (int i = 0; i < 64; i++) christoff[i] = 0;
    for (int i = 0; i < 4; i++) {
        float ginvii = 0.5f * (1 / g[i * 4 + i]);
        for (int m = 0; m < 4; m++) {
            // diagonal partial layout of the reference g_partials[i*4+m] = d_m g_ii
            float adding = ginvii * dg[m * 16 + i * 4 + i];
            christoff[i * 16 + i * 4 + m] += adding;
            christoff[i * 16 + m * 4 + i] += adding;
            christoff[i * 16 + m * 4 + m] -= ginvii * dg[i * 16 + m * 4 + m];
        }
    }
#endif
    float v[4];
    to_arr(lightray_velocity, v);
    float r[4];
    for (int uu = 0; uu < 4; uu++) {
        float sum = 0;
        for (int aa = 0; aa < 4; aa++)
            for (int bb = 0; bb < 4; bb++) sum += v[aa] * v[bb] * christoff[uu * 16 + aa * 4 + bb];
        r[uu] = -sum;
    }
    float4 acceleration = from_arr(r);
#if GRT_CONSTANT_THETA
    acceleration.z = 0;
#endif
    return acceleration;
}

// cl.cl:685-736: Christoffel symbols only (parallel transport)
static inline void get_christoffel_generic(const float g[16], const float dg[64], float christoff[64]) {
#if GRT_BIG_METRIC
    float g_inv[16];
    metric_inverse(g, g_inv);
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 4; k++)
            for (int l = 0; l < 4; l++) {
                float sum = 0;
                for (int m = 0; m < 4; m++) {
                    sum += g_inv[i * 4 + m] * dg[l * 16 + m * 4 + k];
                    sum += g_inv[i * 4 + m] * dg[k * 16 + m * 4 + l];
                    sum -= g_inv[i * 4 + m] * dg[m * 16 + k * 4 + l];
                }
                christoff[i * 16 + k * 4 + l] = 0.5f * sum;
            }
#else
    for (int i = 0; i < 64; i++) christoff[i] = 0;
    for (int i = 0; i < 4; i++) {
        float ginvii = 0.5f * (1 / g[i * 4 + i]);
        for (int m = 0; m < 4; m++) {
            float adding = ginvii * dg[m * 16 + i * 4 + i];
            christoff[i * 16 + i * 4 + m] += adding;
            christoff[i * 16 + m * 4 + i] += adding;
            christoff[i * 16 + m * 4 + m] -= ginvii * dg[i * 16 + m * 4 + m];
        }
    }
#endif
}

// ---- tetrads (cl.cl:1615-1883, 1919-1972, 2072-2114, 2210-2224, 2288-2439) ----------------------
static inline float4 gram_proj_big(float4 u, float4 v, const float g[16]) {
    float top = dot_product_big(u, v, g);
    float bottom = dot_product_big(u, u, g);
    return (top / bottom) * u;
}
static inline float4 normalise_big(float4 in, const float g[16]) {
    float d = dot_product_big(in, in, g);
    return in / sqrtf(fabsf(d));
}
struct orthonormal_basis { float4 v1, v2, v3, v4; };
static inline orthonormal_basis orthonormalise4_metric(float4 i1, float4 i2, float4 i3, float4 i4, const float g[16]) {  // cl.cl:1647-1675
    float4 u1 = i1;
    float4 u2 = i2;
    u2 = u2 - gram_proj_big(u1, u2, g);
    float4 u3 = i3;
    u3 = u3 - gram_proj_big(u1, u3, g);
    u3 = u3 - gram_proj_big(u2, u3, g);
    float4 u4 = i4;
    u4 = u4 - gram_proj_big(u1, u4, g);
    u4 = u4 - gram_proj_big(u2, u4, g);
    u4 = u4 - gram_proj_big(u3, u4, g);
    orthonormal_basis ret;
    ret.v1 = normalise_big(u1, g);
    ret.v2 = normalise_big(u2, g);
    ret.v3 = normalise_big(u3, g);
    ret.v4 = normalise_big(u4, g);
    return ret;
}
static inline int calculate_which_coordinate_is_timelike(const float4 e[4], const float g[16]) {  // cl.cl:1686-1759
    float m[16];
    for (int a = 0; a < 4; a++) to_arr(e[a], m + a * 4);
    float minkowski[16];
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            float sum = 0;
            for (int mu = 0; mu < 4; mu++)
                for (int v = 0; v < 4; v++) sum += g[mu * 4 + v] * m[a * 4 + mu] * m[b * 4 + v];
            minkowski[a * 4 + b] = sum;
        }
    int lowest_index = -1;
    float lowest_index_value = 0;
    for (int i = 0; i < 4; i++)
        if (minkowski[i * 4 + i] < lowest_index_value) { lowest_index = i; lowest_index_value = minkowski[i * 4 + i]; }
    if (lowest_index != -1) return lowest_index;
    return 0;
}
struct frame_basis { float4 v1, v2, v3, v4; int timelike_coordinate; };
static inline frame_basis calculate_frame_basis_with_swap_index(const float g[16], int index_swap) {  // cl.cl:1761-1847
    float4 as_array[4] = {f4(1, 0, 0, 0), f4(0, 1, 0, 0), f4(0, 0, 1, 0), f4(0, 0, 0, 1)};
    float lengths[4];
    for (int i = 0; i < 4; i++) lengths[i] = dot(as_array[i], lower_index_big(as_array[i], g));
    std::swap(as_array[0], as_array[index_swap]);
    std::swap(lengths[0], lengths[index_swap]);
    int indices[4] = {0, 1, 2, 3};
    int first_nonzero = -1;
    float eps = 0.00001f;
    for (int i = 0; i < 4; i++)
        if (!(fabsf(lengths[i] - 0.f) <= eps)) { first_nonzero = i; break; }
    if (first_nonzero == -1) first_nonzero = 0;
    if (first_nonzero != 0) {
        std::swap(as_array[0], as_array[first_nonzero]);
        std::swap(indices[0], indices[first_nonzero]);
    }
    orthonormal_basis result = orthonormalise4_metric(as_array[0], as_array[1], as_array[2], as_array[3], g);
    float4 result_as_array[4] = {result.v1, result.v2, result.v3, result.v4};
    float4 sorted_result[4];
    for (int i = 0; i < 4; i++) sorted_result[indices[i]] = result_as_array[i];
    int which = calculate_which_coordinate_is_timelike(sorted_result, g);
    if (which > 0) std::swap(sorted_result[0], sorted_result[which]);
    frame_basis r;
    r.v1 = sorted_result[0]; r.v2 = sorted_result[1]; r.v3 = sorted_result[2]; r.v4 = sorted_result[3];
    r.timelike_coordinate = which == -1 ? 0 : which;
    return r;
}
static inline frame_basis calculate_frame_basis(const float g[16]) {  // cl.cl:1849-1861
    frame_basis f1 = calculate_frame_basis_with_swap_index(g, 0);
    if (f1.timelike_coordinate == 0) return f1;
    return calculate_frame_basis_with_swap_index(g, f1.timelike_coordinate);
}
static inline void calculate_lorentz_boost(float4 time_basis, float4 observer_velocity, const float g[16], float coeff_out[16]) {  // cl.cl:1919-1972
    float4 lowered_time_basis = lower_index_big(time_basis, g);
    float4 lowered_observer_velocity = lower_index_big(observer_velocity, g);
    float T[4], lT[4], uobs[4], luobs[4];
    to_arr(time_basis, T); to_arr(lowered_time_basis, lT); to_arr(observer_velocity, uobs); to_arr(lowered_observer_velocity, luobs);
    float lorentz_factor = -dot(lowered_time_basis, observer_velocity);
    for (int u = 0; u < 4; u++)
        for (int v = 0; v < 4; v++)
            coeff_out[u * 4 + v] = (u == v ? 1.f : 0.f) + (1 / (1 + lorentz_factor)) * (T[u] + uobs[u]) * (lT[v] + luobs[v]) - 2 * uobs[u] * lT[v];
}
static inline void get_tetrad_inverse(const float4 e_hi[4], float4 e_lo[4]) {  // cl.cl:2072-2093
    float m[16] = {e_hi[0].x, e_hi[1].x, e_hi[2].x, e_hi[3].x,
                   e_hi[0].y, e_hi[1].y, e_hi[2].y, e_hi[3].y,
                   e_hi[0].z, e_hi[1].z, e_hi[2].z, e_hi[3].z,
                   e_hi[0].w, e_hi[1].w, e_hi[2].w, e_hi[3].w};
    float inv[16];
    matrix_inverse(m, inv);
    for (int i = 0; i < 4; i++) e_lo[i] = f4(inv[i * 4 + 0], inv[i * 4 + 1], inv[i * 4 + 2], inv[i * 4 + 3]);
}
static inline float4 coordinate_to_tetrad_basis(float4 vec_up, const float4 e_lo[4]) {  // cl.cl:2096-2108
    return f4(dot(e_lo[0], vec_up), dot(e_lo[1], vec_up), dot(e_lo[2], vec_up), dot(e_lo[3], vec_up));
}
static inline float4 tetrad_to_coordinate_basis(float4 vec_up, const float4 e_hi[4]) {  // cl.cl:2112-2115
    return vec_up.x * e_hi[0] + vec_up.y * e_hi[1] + vec_up.z * e_hi[2] + vec_up.w * e_hi[3];
}
static inline float4 get_timelike_vector(float3 cartesian_basis_speed, float time_direction, const float4 e[4]) {  // cl.cl:2210-2224
    float v2 = dot(cartesian_basis_speed, cartesian_basis_speed);
    float Y = 1 / sqrtf(1 - v2);
    float4 bT = time_direction * Y * e[0];
    float4 bX = Y * cartesian_basis_speed.x * e[1];
    float4 bY = Y * cartesian_basis_speed.y * e[2];
    float4 bZ = Y * cartesian_basis_speed.z * e[3];
    return bT + bX + bY + bZ;
}
struct ortho_result { float3 v1, v2, v3; };
static inline float3 project(float3 u, float3 v) { return (dot(u, v) / dot(u, u)) * u; }
static inline ortho_result orthonormalise(float3 i1, float3 i2, float3 i3) {  // cl.cl:1549-1567
    float3 u1 = i1, u2 = i2, u3 = i3;
    u2 = u2 - project(u1, u2);
    u3 = u3 - project(u1, u3);
    u3 = u3 - project(u2, u3);
    return ortho_result{normalize(u1), normalize(u2), normalize(u3)};
}

// cl.cl:2288-2439
static inline void calculate_tetrads(float4 at_metric, float3 cartesian_basis_speed, float4 e_out[4], const grt_params& cfg, int should_orient) {
    float4 polar_camera = generic_to_spherical(at_metric, cfg);
    if (any_degenerate(at_metric)) {
        e_out[0] = f4(1, 0, 0, 0); e_out[1] = f4(0, 1, 0, 0); e_out[2] = f4(0, 0, 1, 0); e_out[3] = f4(0, 0, 0, 1);
        return;
    }
    float g[16];
    calculate_metric_big(at_metric, g, cfg);
    frame_basis basis = calculate_frame_basis(g);
    float4 e[4] = {basis.v1, basis.v2, basis.v3, basis.v4};
    if (should_orient) {
        float3 apolar = yzw(polar_camera);
        apolar.x = fabsf(apolar.x);
        float3 cart_camera = polar_to_cartesian(apolar);
        float4 e_lo[4];
        get_tetrad_inverse(e, e_lo);
        float3 sx = cartesian_velocity_to_polar_velocity(cart_camera, f3(1, 0, 0));
        float3 sy = cartesian_velocity_to_polar_velocity(cart_camera, f3(0, 1, 0));
        float3 sz = cartesian_velocity_to_polar_velocity(cart_camera, f3(0, 0, 1));
        if (polar_camera.y < 0) { sx.x = -sx.x; sy.x = -sy.x; sz.x = -sz.x; }
        float4 gx = spherical_velocity_to_generic_velocity(polar_camera, f4(0, sx), cfg);
        float4 gy = spherical_velocity_to_generic_velocity(polar_camera, f4(0, sy), cfg);
        float4 gz = spherical_velocity_to_generic_velocity(polar_camera, f4(0, sz), cfg);
        float4 tE1 = coordinate_to_tetrad_basis(gy, e_lo);
        float4 tE2 = coordinate_to_tetrad_basis(gx, e_lo);
        float4 tE3 = coordinate_to_tetrad_basis(gz, e_lo);
        ortho_result result = orthonormalise(yzw(tE1), yzw(tE2), yzw(tE3));
        float4 basis1 = f4(0, result.v1), basis2 = f4(0, result.v2), basis3 = f4(0, result.v3);
        float4 x_out = tetrad_to_coordinate_basis(basis2, e);
        float4 y_out = tetrad_to_coordinate_basis(basis1, e);
        float4 z_out = tetrad_to_coordinate_basis(basis3, e);
        e[1] = x_out; e[2] = y_out; e[3] = z_out;
    }
    {
        float4 observer_velocity = get_timelike_vector(cartesian_basis_speed, 1, e);
        float lorentz[16];
        calculate_lorentz_boost(e[0], observer_velocity, g, lorentz);
        e[0] = observer_velocity;
        e[1] = tensor_contract(lorentz, e[1]);
        e[2] = tensor_contract(lorentz, e[2]);
        e[3] = tensor_contract(lorentz, e[3]);
    }
    for (int i = 0; i < 4; i++) e_out[i] = e[i];
}

// cl.cl:2441-2481
static inline void boost_tetrad(float4 at_metric, float3 basis_speed, float4 e[4], const grt_params& cfg) {
    float4 observer_velocity = get_timelike_vector(basis_speed, 1, e);
    float lorentz[16], g[16];
    calculate_metric_big(at_metric, g, cfg);
    calculate_lorentz_boost(e[0], observer_velocity, g, lorentz);
    e[0] = observer_velocity;
    e[1] = tensor_contract(lorentz, e[1]);
    e[2] = tensor_contract(lorentz, e[2]);
    e[3] = tensor_contract(lorentz, e[3]);
}

// ---- ray setup (cl.cl:1974-2059, 2949-3065, 3143-3251, 6018-6034) --------------------------------
static inline float4 aa_to_quat(float3 axis, float angle) {  // cl.cl:1976-1984
    float4 q = f4(axis.x * sinf(angle / 2), axis.y * sinf(angle / 2), axis.z * sinf(angle / 2), cosf(angle / 2));
    return normalize(q);
}
static inline float3 rejection(float3 my_vector, float3 basis) { return normalize(my_vector - dot(my_vector, basis) * basis); }  // cl.cl:306-309
static inline float4 get_theta_adjustment_quat(float3 pixel_direction, float4 polar_camera_in, float angle_sign) {  // cl.cl:2015-2042
    if (length(pixel_direction) < 0.00001f) pixel_direction = f3(0, 1, 0);
    float3 apolar = yzw(polar_camera_in);
    apolar.x = fabsf(apolar.x);
    float3 cartesian_camera_pos = polar_to_cartesian(apolar);
    float3 bx = normalize(pixel_direction);
    float3 by = normalize(-cartesian_camera_pos);
    bx = normalize(rejection(bx, by));
    float3 plane_n = -normalize(cross(bx, by));
    float angle_to_flat = acosf(dot(plane_n, f3(0, 0, 1)));
    return aa_to_quat(normalize(cross(plane_n, f3(0, 0, 1))), angle_to_flat * angle_sign);
}
static inline float3 calculate_pixel_direction(int cx, int cy, float width, float height, float4 camera_quat, const Features& dfg) {  // cl.cl:2044-2059
    float fov = dfg.field_of_view;
    float fov_rad = (fov / 360.f) * 2 * PI_F;
    float nonphysical_plane_half_width = width / 2;
    float nonphysical_f_stop = nonphysical_plane_half_width / tanf(fov_rad / 2);
    float3 pixel_direction = f3(cx - width / 2, cy - height / 2, nonphysical_f_stop);
    pixel_direction = normalize(pixel_direction);
    pixel_direction = rot_quat(pixel_direction, camera_quat);
    return pixel_direction;
}
static inline int should_early_terminate(int x, int y, int width, int height, const int* termination_buffer) {  // cl.cl:2061-2070
    if (x < 0 || y < 0 || x > width - 1 || y > height - 1) return 0;
    return termination_buffer[y * width + x] == 1;
}

struct corrected_lightray { float4 position, velocity, inverse_quat; };
static inline corrected_lightray correct_lightray(float4 position, float4 velocity, const grt_params& cfg) {  // cl.cl:2949-2997
    float4 extra_quat = f4(0, 0, 0, 1);
#if GRT_CONSTANT_THETA
    {
        float4 polar_pos = generic_to_spherical(position, cfg);
        float4 pos_spherical = generic_to_spherical(position, cfg);
        float4 vel_spherical = generic_velocity_to_spherical_velocity(position, velocity, cfg);
        float fsign = signf(pos_spherical.y);
        pos_spherical.y = fabsf(pos_spherical.y);
        float3 pos_cart = polar_to_cartesian(yzw(pos_spherical));
        float3 vel_cart = spherical_velocity_to_cartesian_velocity(yzw(pos_spherical), yzw(vel_spherical));
        float4 quat = get_theta_adjustment_quat(vel_cart, polar_pos, 1);
        extra_quat = get_theta_adjustment_quat(vel_cart, polar_pos, -1);
        pos_cart = rot_quat(pos_cart, quat);
        vel_cart = rot_quat(vel_cart, quat);
        float3 next_pos_spherical = cartesian_to_polar(pos_cart);
        float3 next_vel_spherical = cartesian_velocity_to_polar_velocity(pos_cart, vel_cart);
        if (fsign < 0) next_pos_spherical.x = -next_pos_spherical.x;
        float4 next_pos_generic = spherical_to_generic(f4(pos_spherical.x, next_pos_spherical), cfg);
        float4 next_vel_generic = spherical_velocity_to_generic_velocity(f4(pos_spherical.x, next_pos_spherical), f4(vel_spherical.x, next_vel_spherical), cfg);
        position = next_pos_generic;
        velocity = next_vel_generic;
    }
#endif
    (void)cfg;
    return corrected_lightray{position, velocity, extra_quat};
}

// cl.cl:3000-3065 (render ray: ku_uobsu = g(k, u_obs)) and cl.cl:3067-3115 (trace ray: ku_uobsu = 1)
static inline lightray geodesic_to_ray(int cx, int cy, float4 position, float4 velocity, const float4* observer_velocity, const grt_params& cfg) {
    corrected_lightray corrected = correct_lightray(position, velocity, cfg);
    position = corrected.position;
    velocity = corrected.velocity;
#if GRT_CONSTANT_THETA
    position.z = PI_F / 2;
    velocity.z = 0;
#endif
    float g[16], dg[64];
    calculate_metric_big(position, g, cfg);
    calculate_partial_derivatives_big(position, dg, cfg);
    // fix_light_velocity_big returns its input unchanged (cl.cl:1401-1403)
    float4 lightray_acceleration = calculate_acceleration_numeric(velocity, g, dg);
    lightray ray;
    ray.position = position;
    ray.velocity = velocity;
    ray.acceleration = lightray_acceleration;
    ray.initial_quat = corrected.inverse_quat;
    ray.running_dlambda_dnew = 1;
    ray.terminated = 0;
    if (observer_velocity) {
        float4 uobsu_lower = lower_index_big(*observer_velocity, g);
        ray.ku_uobsu = dot(velocity, uobsu_lower);
    } else ray.ku_uobsu = 1;
    ray.sx = cx;
    ray.sy = cy;
    return ray;
}

// cl.cl:6018-6034
static inline float4 cart_to_generic(float4 position_cart, float flip, const grt_params& cfg) {
    float3 polar = cartesian_to_polar(yzw(position_cart));
    if (flip > 0) polar.x = -polar.x;
    return spherical_to_generic(f4(position_cart.x, polar), cfg);
}

// one pixel of init_rays_generic / the ray construction of handle_adaptive_sampling (cl.cl:3159-3211, 5301-5327)
static inline lightray make_pixel_ray(int cx, int cy, int width, int height, float4 camera_generic, float4 camera_quat, const float4 e[4],
                                      int flip_geodesic_direction, const grt_params& cfg, const Features& dfg, bool apply_forward_flag) {
    float3 pixel_direction = calculate_pixel_direction(cx, cy, (float)width, (float)height, camera_quat, dfg);
    pixel_direction = normalize(pixel_direction);
    float4 pixel_x = pixel_direction.x * e[1];
    float4 pixel_y = pixel_direction.y * e[2];
    float4 pixel_z = pixel_direction.z * e[3];
    float4 pixel_t = -e[0];
#if GRT_FORWARD_GEODESIC_PATH
    if (apply_forward_flag) pixel_t = e[0];
#endif
    (void)apply_forward_flag;
    if (flip_geodesic_direction) pixel_t = -pixel_t;
    float4 lightray_velocity = pixel_x + pixel_y + pixel_z + pixel_t;
    return geodesic_to_ray(cx, cy, camera_generic, lightray_velocity, &e[0], cfg);
}

// cl.cl:3143-3251. Returns the number of rays written (metric_ray_count).
static inline int init_rays_generic(float4 camera_generic, float4 camera_quat, lightray* metric_rays, int width, int height,
                                    const int* termination_buffer, int prepass_width, int prepass_height, int flip_geodesic_direction,
                                    const float4 e[4], const grt_params& cfg, const Features& dfg, int i_am_prepass) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int id = 0; id < width * height; id++) {
        const int cx = id % width;
        const int cy = id / width;
        lightray ray = make_pixel_ray(cx, cy, width, height, camera_generic, camera_quat, e, flip_geodesic_direction, cfg, dfg, true);
        if (prepass_width != width && prepass_height != height) {
            float fx = (float)cx / width;
            float fy = (float)cy / height;
            int lx = (int)roundf(fx * prepass_width);
            int ly = (int)roundf(fy * prepass_height);
            if (should_early_terminate(lx - 1, ly, prepass_width, prepass_height, termination_buffer) &&
                should_early_terminate(lx, ly, prepass_width, prepass_height, termination_buffer) &&
                should_early_terminate(lx + 1, ly, prepass_width, prepass_height, termination_buffer) &&
                should_early_terminate(lx, ly - 1, prepass_width, prepass_height, termination_buffer) &&
                should_early_terminate(lx, ly + 1, prepass_width, prepass_height, termination_buffer))
                ray.terminated = 2;
        }
        if (i_am_prepass || !dfg.adaptive_sampling) metric_rays[id] = ray;
        else {
            if ((cx % 2) != 0 || (cy % 2) != 0) continue;
            metric_rays[(cy / 2) * (width / 2) + cx / 2] = ray;
        }
    }
    return (i_am_prepass || !dfg.adaptive_sampling) ? height * width : (height * width) / 4;
}

// ---- integrator (cl.cl:3273-3346, 3400-3456, 3954-4247) --------------------------------------------
static inline void step_verlet(float4 position, float4 velocity, float4 acceleration, float ds, float4* position_out, float4* velocity_out,
                               float4* acceleration_out, float* dLambda_dNew, const grt_params& cfg, const Features& dfg) {
    float4 next_position = position + velocity * ds + 0.5f * acceleration * ds * ds;
    float4 intermediate_next_velocity = velocity + acceleration * ds;
    float4 next_acceleration;
    {
        // GEO_ACCEL0..3 (with v3 = pi/2, iv3 = 0, a.z = 0 under GENERIC_CONSTANT_THETA - folded in by the emitter)
        float p[4], v[4], a[4];
        to_arr(next_position, p);
        to_arr(intermediate_next_velocity, v);
        grt_accel_ref(p, v, cfg, a);
        next_acceleration = from_arr(a);
    }
    float4 next_velocity = velocity + 0.5f * (acceleration + next_acceleration) * ds;
    float max_divisor = fmaxf(fmaxf(fabsf(next_velocity.x), fabsf(next_velocity.y)), fmaxf(fabsf(next_velocity.z), fabsf(next_velocity.w)));
    float K = 1 / max_divisor;
    if (!dfg.reparameterisation) K = 1;
    if (dLambda_dNew) *dLambda_dNew = K;
    *position_out = next_position;
    *velocity_out = next_velocity * K;
    *acceleration_out = next_acceleration * K * K;
}

enum ds_result { DS_NONE, DS_SKIP, DS_RETURN };
#define I_HATE_COMPUTERS (256 * 256)

static inline float acceleration_to_precision(float4 acceleration, float max_acceleration, float* next_ds_out) {  // cl.cl:3400-3429
    float uniform_coordinate_precision_divisor = fmaxf(fmaxf(GRT_W_V1, GRT_W_V2), fmaxf(GRT_W_V3, GRT_W_V4));
    float current_acceleration_err = length(acceleration * f4(GRT_W_V1, GRT_W_V2, GRT_W_V3, GRT_W_V4)) * 0.01f;
    current_acceleration_err /= uniform_coordinate_precision_divisor;
    float experienced_acceleration_change = current_acceleration_err;
    float err = max_acceleration;
    float max_timestep = 100000;
    float diff = experienced_acceleration_change * I_HATE_COMPUTERS;
    if (diff < err * I_HATE_COMPUTERS / powf(max_timestep, 2)) diff = err * I_HATE_COMPUTERS / powf(max_timestep, 2);
    float next_ds = sqrtf(((err * I_HATE_COMPUTERS) / diff));
    *next_ds_out = next_ds;
    return diff;
}

static inline int calculate_ds_error(float current_ds, float4 next_acceleration, float max_acceleration, float* next_ds_out, const Features& dfg) {  // cl.cl:3431-3456
    float next_ds = 0;
    float diff = acceleration_to_precision(next_acceleration, max_acceleration, &next_ds);
    next_ds = 0.99f * current_ds * clampf(next_ds / current_ds, 0.3f, 2.f);
    float min_step = dfg.min_step;
    next_ds = fmaxf(next_ds, min_step);
    *next_ds_out = next_ds;
    float err = max_acceleration;
#if GRT_SINGULARITY_DETECTION
    if (next_ds == min_step && (diff / I_HATE_COMPUTERS) > err * 10000) return DS_RETURN;
#endif
    (void)err; (void)diff;
    if (next_ds < current_ds / 1.95f) return DS_SKIP;
    return DS_NONE;
}

static inline float linear_val(float value, float min_val, float max_val, float val_at_min, float val_at_max) {  // cl.cl:799-811
    value = clampf(value, min_val, max_val);
    float mixd = (value - min_val) / (max_val - min_val);
    return mix(val_at_min, val_at_max, mixd);
}

// one work-item of do_generic_rays (cl.cl:3954-4247), triangle-rendering branches omitted
static inline void do_generic_ray(lightray* ray, const grt_params& cfg, const Features& dfg, Counters* cnt) {
    if (ray->terminated == 2) return;
    float4 position = ray->position;
    float4 velocity = ray->velocity;
    float4 acceleration = ray->acceleration;
    float f_in_x = fabsf(velocity.x);
#if GRT_CONSTANT_THETA
    position.z = PI_F / 2;
    velocity.z = 0;
    acceleration.z = 0;
#endif
    float next_ds = 0.00001f;
#if GRT_ADAPTIVE_PRECISION
    (void)acceleration_to_precision(acceleration, dfg.max_acceleration_change, &next_ds);
#endif
    float subambient_precision = 0.5f;
    float ambient_precision = 0.2f;
    int loop_limit = 4096 * 4;
    float running_dlambda_dnew = 1;
    (void)f_in_x;
    for (int i = 0; i < loop_limit; i++) {
#if GRT_CONSTANT_THETA
        position.z = PI_F / 2;
        velocity.z = 0;
        acceleration.z = 0;
#endif
        float new_max = dfg.max_precision_radius;
        float new_min = 3;
        float4 polar_position = generic_to_spherical(position, cfg);
#if GRT_CONSTANT_THETA
        polar_position.z = PI_F / 2;
#endif
        float r_value = get_distance_to_object(polar_position, cfg);
        float ds = linear_val(fabsf(r_value), new_min, new_max, ambient_precision, subambient_precision);
#if GRT_ADAPTIVE_PRECISION
        ds = next_ds;
#endif
        if (fabsf(r_value) < new_max) ds = fminf(ds, ambient_precision);
        else ds = 0.1f * (fabsf(r_value) - new_max) + ambient_precision;
        bool should_terminate = fabsf(polar_position.y) >= dfg.universe_size;
#if GRT_SINGULAR
        should_terminate |= fabsf(polar_position.y) < GRT_SINGULAR_TERMINATOR;
#endif
#if GRT_HAS_CYLINDRICAL_SINGULARITY
        if (position.y < GRT_CYLINDRICAL_TERMINATOR) return;
#endif
#if !GRT_UNCONDITIONALLY_NONSINGULAR
        if (fabsf(velocity.x / running_dlambda_dnew) > 1000 + f_in_x && fabsf(acceleration.x / running_dlambda_dnew) > 100) return;
#endif
        if (should_terminate) {
            ray->position = position;
            ray->velocity = velocity;
            ray->running_dlambda_dnew = running_dlambda_dnew;
            ray->terminated = 1;
            return;
        }
        float dLambda_dNew = 1;
        float4 next_position, next_velocity, next_acceleration;
        step_verlet(position, velocity, acceleration, ds, &next_position, &next_velocity, &next_acceleration, &dLambda_dNew, cfg, dfg);
        if (cnt) cnt->trial_steps++;
        running_dlambda_dnew *= dLambda_dNew;
#if GRT_ADAPTIVE_PRECISION
        if (fabsf(r_value) < new_max) {
            int res = calculate_ds_error(ds, next_acceleration, dfg.max_acceleration_change, &next_ds, dfg);
            if (res == DS_RETURN) return;
            if (res == DS_SKIP) { i--; continue; }
        }
#endif
        if (cnt) cnt->accepted_steps++;
        position = next_position;
        velocity = next_velocity;
        acceleration = next_acceleration;
        if (any_degenerate(position) || any_degenerate(velocity) || any_degenerate(acceleration)) return;
    }
}

static inline void do_generic_rays(lightray* rays, int count, const grt_params& cfg, const Features& dfg, Counters* total, int* per_ray_steps = nullptr) {
    long long trial = 0, accepted = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : trial, accepted)
    for (int id = 0; id < count; id++) {
        Counters c;
        do_generic_ray(&rays[id], cfg, dfg, &c);
        trial += c.trial_steps;
        accepted += c.accepted_steps;
        if (per_ray_steps) per_ray_steps[id] = (int)c.trial_steps;
    }
    if (total) { total->trial_steps += trial; total->accepted_steps += accepted; }
}

// ---- prepass (cl.cl:4997-5020) ---------------------------------------------------------------
static inline void clear_termination_buffer(int* termination_buffer, int width, int height) {
    for (int id = 0; id < width * height; id++) termination_buffer[id] = 1;
}
static inline void calculate_singularities(const lightray* finished_rays, int finished_count, int* termination_buffer, int width, int height) {
    (void)height;
    for (int id = 0; id < finished_count; id++) {
        int sx = id % width, sy = id / width;
        termination_buffer[sy * width + sx] = !finished_rays[id].terminated;
    }
}

// ---- shading data (cl.cl:5024-5213) ------------------------------------------------------------
static inline float4 get_intersection_position(const lightray& ray, const grt_params& cfg, const Features& dfg) {  // cl.cl:5024-5064
    float4 position = generic_to_spherical(ray.position, cfg);
    float4 velocity = generic_velocity_to_spherical_velocity(ray.position, ray.velocity, cfg);
#if GRT_CONSTANT_THETA
    position.z = PI_F / 2;
    velocity.z = 0;
#endif
    if (fabsf(position.y) >= dfg.universe_size) {
        float3 np = fix_ray_position(yzw(position), yzw(velocity), dfg.universe_size);
        position = f4(position.x, np);
    }
#if GRT_SINGULAR && GRT_TRAVERSABLE_EVENT_HORIZON
    if (fabsf(position.y) < GRT_SINGULAR_TERMINATOR) {
        float3 np = fix_ray_position(yzw(position), yzw(velocity), GRT_SINGULAR_TERMINATOR);
        position = f4(position.x, np);
    }
#endif
    float3 npolar = yzw(position);
#if GRT_CONSTANT_THETA
    {
        float3 cart_pos = polar_to_cartesian(yzw(position));
        cart_pos = rot_quat(cart_pos, ray.initial_quat);
        npolar = cartesian_to_polar(cart_pos);
    }
#endif
    return f4(position.x, npolar);
}
static inline float3 angle_to_vec(float2 a) { return polar_to_cartesian(f3(1.f, a.x, a.y)); }  // cl.cl:5076-5079
static inline float2 angle_to_tex(float2 angle) {  // cl.cl:5081-5100
    float thetaf = fmodf(angle.x, 2 * PI_F);
    float phif = angle.y;
    if (thetaf >= PI_F) { phif += PI_F; thetaf -= PI_F; }
    phif = fmodf(phif, 2 * PI_F);
    float sxf = (phif) / (2 * PI_F);
    float syf = thetaf / PI_F;
    sxf += 0.5f;
    return float2{sxf, syf};
}
static inline float2 tex_to_angle(float2 tex) {  // cl.cl:5102-5109
    tex.x -= 0.5f;
    tex.x *= 2 * PI_F;
    tex.y *= PI_F;
    return tex;
}
static inline render_data interpolate_render_data(const render_data& ray1, const render_data& ray2) {  // cl.cl:5111-5133
    float2 a1 = tex_to_angle(ray1.tex_coord);
    float2 a2 = tex_to_angle(ray2.tex_coord);
    float3 v1 = angle_to_vec(float2{a1.y, a1.x});
    float3 v2 = angle_to_vec(float2{a2.y, a2.x});
    float3 vc = (v1 + v2) / 2.f;
    float3 fangle = cartesian_to_polar(vc);
    float2 tc = angle_to_tex(float2{fangle.y, fangle.z});
    render_data out;
    out.tex_coord = tc;
    out.z_shift = (ray1.z_shift + ray2.z_shift) / 2.f;
    out.terminated = ray1.terminated;
    out.sx = (ray1.sx + ray2.sx) / 2;
    out.sy = (ray1.sy + ray2.sy) / 2;
    out.side = (ray1.side + ray2.side) / 2;
    return out;
}

// cl.cl:5135-5213
static inline void calculate_render_data(const lightray* rays_in, int rays_in_count, render_data* rdata, int width, int height,
                                         const grt_params& cfg, const Features& dfg) {
    (void)height;
#pragma omp parallel for schedule(dynamic, 256)
    for (int gid = 0; gid < rays_in_count; gid++) {
        const lightray* ray = &rays_in[gid];
        render_data dat;
        dat.terminated = ray->terminated;
        dat.sx = ray->sx;
        dat.sy = ray->sy;
        dat.z_shift = 0;
        dat.tex_coord = float2{0, 0};
        dat.side = 1;
        int id = dat.sy * width + dat.sx;
        if (ray->terminated != 1) { rdata[id] = dat; continue; }
        float4 position = get_intersection_position(*ray, cfg, dfg);
        float4 generic_velocity = ray->velocity / ray->running_dlambda_dnew;
        dat.side = generic_to_spherical(ray->position, cfg).y < 0 ? 0 : 1;
        float r_value = position.y;
#if !GRT_TRAVERSABLE_EVENT_HORIZON
        if (fabsf(r_value) <= 1) { rdata[id] = dat; continue; }
#endif
        (void)r_value;
        float4 generic_position = ray->position;
        float4 fe[4];
        calculate_tetrads(generic_position, f3(0, 0, 0), fe, cfg, 0);
        float g[16];
        calculate_metric_big(generic_position, g, cfg);
        float4 obvs_low = lower_index_big(fe[0], g);
        float z_shift = (dot(generic_velocity, obvs_low) / ray->ku_uobsu) - 1;
        z_shift = fmaxf(z_shift, -0.999f);
        dat.z_shift = z_shift;
        dat.tex_coord = angle_to_tex(float2{position.z, position.w});
        rdata[id] = dat;
    }
}

// ---- adaptive sampling (cl.cl:5215-5345) --------------------------------------------------------
static inline float angle_between_angles2(float2 a1, float2 a2) {
    float3 v1 = angle_to_vec(a1), v2 = angle_to_vec(a2);
    return acosf(clampf(dot(v1, v2), -1.f, 1.f));
}
// Returns the number of extra rays appended to unprocessed_rays_out. The reference appends with
// atomic_add (order undefined); here cells are visited in row-major order.
static inline int handle_adaptive_sampling(const lightray* rays_in, render_data* rdat, lightray* unprocessed_rays_out,
                                           float4 camera_generic, float4 camera_quat, const float4 e[4], int width, int height,
                                           const grt_params& cfg, const Features& dfg) {
    int out_count = 0;
    const int hw = width / 2, hh = height / 2;
    std::vector<unsigned char> sample_flags((size_t)hw * hh);
#pragma omp parallel for schedule(dynamic, 4)
    for (int sy = 0; sy < hh; sy++)
        for (int sx = 0; sx < hw; sx++) {
            bool should_sample = true;
            if (sx != 0 && sx != hw - 1 && sy != 0 && sy != hh - 1) {
                int lx = sx - 1, rx = sx + 1, ly = sy - 1, ry = sy + 1;
                const lightray& centre = rays_in[sy * hw + sx];
                const lightray& left = rays_in[sy * hw + lx];
                const lightray& right = rays_in[sy * hw + rx];
                const lightray& up = rays_in[ly * hw + sx];
                const lightray& down = rays_in[ry * hw + sx];
                const lightray& down_right = rays_in[ry * hw + rx];
                float4 lpos = get_intersection_position(left, cfg, dfg);
                float4 rpos = get_intersection_position(right, cfg, dfg);
                float4 upos = get_intersection_position(up, cfg, dfg);
                float4 dpos = get_intersection_position(down, cfg, dfg);
                float x_error = fabsf(angle_between_angles2(float2{lpos.z, lpos.w}, float2{rpos.z, rpos.w}));
                float y_error = fabsf(angle_between_angles2(float2{dpos.z, dpos.w}, float2{upos.z, upos.w}));
                // `... / 2 * M_PI` multiplies by pi (precedence), M_PI is a double in OpenCL C
                float relative_angular_error = (float)(((x_error + x_error + y_error + y_error) / 4.f) / 2 * 3.14159265358979323846);
                float fov = dfg.field_of_view;
                float fov_angle_pi = (float)(fov * 2 * 3.14159265358979323846 / 360.f);
                float rough_angular_change_per_pixel = fov_angle_pi / width;
                should_sample = relative_angular_error >= rough_angular_change_per_pixel * dfg.adaptive_sampling_threshold;
                if (centre.terminated != left.terminated || centre.terminated != right.terminated || centre.terminated != up.terminated ||
                    centre.terminated != down.terminated || centre.terminated != down_right.terminated)
                    should_sample = true;
            }
            sample_flags[(size_t)sy * hw + sx] = should_sample;
        }
    for (int sy = 0; sy < hh; sy++)
        for (int sx = 0; sx < hw; sx++) {
            if (sample_flags[(size_t)sy * hw + sx]) {
                int base_sx = sx * 2, base_sy = sy * 2;
                int ray_pos[3][2] = {{base_sx + 1, base_sy}, {base_sx, base_sy + 1}, {base_sx + 1, base_sy + 1}};
                for (int i = 0; i < 3; i++)
                    unprocessed_rays_out[out_count++] = make_pixel_ray(ray_pos[i][0], ray_pos[i][1], width, height, camera_generic, camera_quat, e, 0, cfg, dfg, false);
            } else {
                const lightray& my_ray = rays_in[sy * hw + sx];
                int lsx = my_ray.sx, lsy = my_ray.sy;
                render_data cdata = rdat[lsy * width + lsx];
                render_data rdata = rdat[lsy * width + lsx + 2];
                render_data ddata = rdat[(lsy + 2) * width + lsx];
                render_data drdata = rdat[(lsy + 2) * width + lsx + 2];
                rdat[lsy * width + lsx + 1] = interpolate_render_data(cdata, rdata);
                rdat[(lsy + 1) * width + lsx] = interpolate_render_data(cdata, ddata);
                rdat[(lsy + 1) * width + lsx + 1] = interpolate_render_data(cdata, drdata);
            }
        }
    return out_count;
}

// ---- background texture (graphics_settings.cpp:152-212) and render (cl.cl:5347-5846) --------------
// Mip pyramid stored compactly; sampling emulates the reference's padded full-size array layers.
struct MipImage {
    int width = 0, height = 0, levels = 0;
    std::vector<std::vector<uint8_t>> level;  // RGBA8
    std::vector<int> lw, lh;
};
static inline uint8_t quantise_like_reference(float f01) {  // (uchar)(in * 255) after GL float readback
    f01 = clampf(f01, 0.f, 1.f);
    return (uint8_t)(f01 * 255);
}
static inline MipImage build_mip_image(const uint8_t* rgba, int w, int h) {
    MipImage img;
    img.width = w;
    img.height = h;
    int max_mips = (int)floorf(log2f((float)std::min(w, h))) + 1;
    img.levels = std::min(max_mips, 10);
    // GL RGBA8 mip chain by 2x2 box filter, round to nearest (glGenerateMipmap behaviour assumed)
    std::vector<std::vector<uint8_t>> gl;
    std::vector<int> ws, hs;
    gl.emplace_back(rgba, rgba + (size_t)w * h * 4);
    ws.push_back(w); hs.push_back(h);
    for (int l = 1; l < img.levels; l++) {
        int pw = ws[l - 1], ph = hs[l - 1];
        int cw = std::max(1, pw / 2), ch = std::max(1, ph / 2);
        std::vector<uint8_t> cur((size_t)cw * ch * 4);
        const std::vector<uint8_t>& prev = gl[l - 1];
        for (int y = 0; y < ch; y++)
            for (int x = 0; x < cw; x++)
                for (int c = 0; c < 4; c++) {
                    int x0 = std::min(2 * x, pw - 1), x1 = std::min(2 * x + 1, pw - 1), y0 = std::min(2 * y, ph - 1), y1 = std::min(2 * y + 1, ph - 1);
                    int s = prev[((size_t)y0 * pw + x0) * 4 + c] + prev[((size_t)y0 * pw + x1) * 4 + c] + prev[((size_t)y1 * pw + x0) * 4 + c] + prev[((size_t)y1 * pw + x1) * 4 + c];
                    cur[((size_t)y * cw + x) * 4 + c] = (uint8_t)((s + 2) >> 2);
                }
        gl.push_back(cur);
        ws.push_back(cw); hs.push_back(ch);
    }
    for (int l = 0; l < img.levels; l++) {
        // the reference divides the base size (int)(swidth / pow(2, i)), equal to the GL level size
        std::vector<uint8_t> q(gl[l].size());
        for (size_t i = 0; i < q.size(); i++) q[i] = quantise_like_reference(gl[l][i] / 255.f);
        img.level.push_back(q);
        img.lw.push_back(ws[l]);
        img.lh.push_back(hs[l]);
    }
    return img;
}
// read_imagef(image2d_array, NORMALIZED | REPEAT | LINEAR, (u, v, layer)) on a full-size layer
// whose top-left (lw x lh) holds the mip level and whose remainder repeats the level's last
// row/column (graphics_settings.cpp:186-203)
static inline float4 sample_layer(const MipImage& img, float u, float v, int layer) {
    layer = std::min(std::max(layer, 0), img.levels - 1);
    const int W = img.width, H = img.height;
    const int lw = img.lw[layer], lh = img.lh[layer];
    const uint8_t* data = img.level[layer].data();
    float fu = (u - floorf(u)) * W, fv = (v - floorf(v)) * H;
    int i0 = (int)floorf(fu - 0.5f), j0 = (int)floorf(fv - 0.5f);
    int i1 = i0 + 1, j1 = j0 + 1;
    float a = (fu - 0.5f) - floorf(fu - 0.5f), b = (fv - 0.5f) - floorf(fv - 0.5f);
    if (i0 < 0) i0 += W;
    if (i1 > W - 1) i1 -= W;
    if (j0 < 0) j0 += H;
    if (j1 > H - 1) j1 -= H;
    auto texel = [&](int x, int y) {
        int lx = std::min(x, lw - 1), ly = std::min(y, lh - 1);
        const uint8_t* p = data + ((size_t)ly * lw + lx) * 4;
        return f4(p[0] / 255.f, p[1] / 255.f, p[2] / 255.f, p[3] / 255.f);
    };
    float4 t00 = texel(i0, j0), t10 = texel(i1, j0), t01 = texel(i0, j1), t11 = texel(i1, j1);
    return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}
static inline float4 read_mipmap(const MipImage& mipmap1, const MipImage& mipmap2, int side, float2 pos, float lod) {  // cl.cl:5421-5449
    lod = fmaxf(lod, 0.f);
    pos.x = fmodf(pos.x, 1.f);
    pos.y = fmodf(pos.y, 1.f);
    float mip_lower = floorf(lod), mip_upper = ceilf(lod);
    float lower_divisor = powf(2.f, mip_lower), upper_divisor = powf(2.f, mip_upper);
    float lower_weight = lod - mip_lower;
    const MipImage& img = side >= 1 ? mipmap1 : mipmap2;
    float4 lo = sample_layer(img, pos.x / lower_divisor, pos.y / lower_divisor, (int)mip_lower);
    float4 hi = sample_layer(img, pos.x / upper_divisor, pos.y / upper_divisor, (int)mip_upper);
    return mix(lo, hi, lower_weight);
}
static inline float circular_diff(float f1, float f2, float period) {  // cl.cl:3598-3605 (M_PI is double)
    f1 = (float)(f1 * (2 * 3.14159265358979323846 / period));
    f2 = (float)(f2 * (2 * 3.14159265358979323846 / period));
    return (float)(period * atan2f(sinf(f2 - f1), cosf(f2 - f1)) / (2 * 3.14159265358979323846));
}
static inline float energy_of(float3 v) { return v.x * 0.2125f + v.y * 0.7154f + v.z * 0.0721f; }  // cl.cl:5366-5369
static inline float3 redshift(float3 v, float z, const Features& dfg) {  // cl.cl:5371-5413
    float radiant_energy = energy_of(v);
    float3 red = f3(1 / 0.2125f, 0.f, 0.f);
    float3 green = f3(0, (float)(1 / 0.7154), 0.f);
    float3 blue = f3(0.f, 0.f, (float)(1 / 0.0721));
    float3 result;
    if (z > 0) result = mix(v, radiant_energy * red, tanhf(z));
    else {
        float iv1pz = (1 / (1 + z)) - 1;
        float interpolating_fraction = tanhf(iv1pz);
        float3 col = mix(v, radiant_energy * blue, interpolating_fraction);
        if (!dfg.use_old_redshift) {
            float final_energy = energy_of(clamp3(col, 0.f, 1.f));
            float real_energy = energy_of(col);
            float remaining_energy = real_energy - final_energy;
            col.x += remaining_energy * (red.x + green.x);
            col.y += remaining_energy * (red.y + green.y);
        }
        result = col;
    }
    return clamp3(result, 0.f, 1.f);
}

// one work-item of `render` (cl.cl:5453-5846); LINEAR_FRAMEBUFFER is defined in the app (metric_manager.hpp:78-81)
static inline float4 render_pixel(const render_data* rdata, int id, const MipImage& bg1, const MipImage& bg2, int width, int height, int maxProbes,
                                  const Features& dfg, int* out_sx, int* out_sy) {
    render_data rdat = rdata[id];
    int sx = rdat.sx, sy = rdat.sy, side = rdat.side;
    *out_sx = sx; *out_sy = sy;
    if (rdat.terminated != 1) return f4(0, 0, 0, 1);
    float sxf = rdat.tex_coord.x, syf = rdat.tex_coord.y;
    int dx = 1, dy = 1;
    if (sx == width - 1) dx = -1;
    if (sy == height - 1) dy = -1;
    float2 tl = rdata[sy * width + sx].tex_coord;
    float2 tr = rdata[sy * width + sx + dx].tex_coord;
    float2 bl = rdata[(sy + dy) * width + sx].tex_coord;
    float bias_frac = 1.3f;
    float2 dx_vtc = float2{circular_diff(tl.x, tr.x, 1.f) / bias_frac, circular_diff(tl.y, tr.y, 1.f) / bias_frac};
    float2 dy_vtc = float2{circular_diff(tl.x, bl.x, 1.f) / bias_frac, circular_diff(tl.y, bl.y, 1.f) / bias_frac};
    if (dx == -1) { dx_vtc.x = -dx_vtc.x; dx_vtc.y = -dx_vtc.y; }
    if (dy == -1) { dy_vtc.x = -dy_vtc.x; dy_vtc.y = -dy_vtc.y; }
    const MipImage& img = side > 0 ? bg1 : bg2;
    dx_vtc.x *= img.width; dy_vtc.x *= img.width;
    dx_vtc.y *= img.height; dy_vtc.y *= img.height;
    float dv_dx = dx_vtc.y, dv_dy = dy_vtc.y, du_dx = dx_vtc.x, du_dy = dy_vtc.x;
    float Ann = dv_dx * dv_dx + dv_dy * dv_dy + 1;
    float Bnn = -2 * (du_dx * dv_dx + du_dy * dv_dy);
    float Cnn = du_dx * du_dx + du_dy * du_dy + 1;
    float F = Ann * Cnn - Bnn * Bnn / 4;
    float A = Ann / F, B = Bnn / F, C = Cnn / F;
    float root = sqrtf((A - C) * (A - C) + B * B);
    float a_prime = (A + C - root) / 2, c_prime = (A + C + root) / 2;
    float majorRadius = 1.f / sqrtf(a_prime), minorRadius = 1.f / sqrtf(c_prime);
    float theta = atan2f(B, (A - C) / 2);
    majorRadius = fmaxf(majorRadius, 1.f);
    minorRadius = fmaxf(minorRadius, 1.f);
    majorRadius = fmaxf(majorRadius, minorRadius);
    float fProbes = 2 * (majorRadius / minorRadius) - 1;
    int iProbes = (int)floorf(fProbes + 0.5f);
    iProbes = std::min(iProbes, maxProbes);
    if (iProbes < fProbes) minorRadius = 2 * majorRadius / (iProbes + 1);
    float levelofdetail = log2f(minorRadius);
    int maxLod = img.levels - 1;
    if (levelofdetail > maxLod) { levelofdetail = (float)maxLod; iProbes = 1; }
    float4 end_result = f4(0, 0, 0, 0);
    if (iProbes <= 1) {
        if (iProbes < 1) levelofdetail = (float)maxLod;
        end_result = read_mipmap(bg1, bg2, side, float2{sxf, syf}, levelofdetail);
    } else {
        float lineLength = 2 * (majorRadius - minorRadius);
        float du = cosf(theta) * lineLength / (iProbes - 1);
        float dv = sinf(theta) * lineLength / (iProbes - 1);
        float4 totalWeight = f4(0, 0, 0, 0);
        float accumulatedProbes = 0;
        int startN = ((iProbes % 2) == 1) ? -2 * ((iProbes - 1) / 2) : -2 * (iProbes / 2) - 1;
        int currentN = startN;
        float alpha = 2;
        float sU = du / img.width, sV = dv / img.height;
        for (int cnt = 0; cnt < iProbes; cnt++) {
            float d_2 = (currentN * currentN / 4.f) * (du * du + dv * dv) / (majorRadius * majorRadius);
            float relativeWeight = expf(-alpha * d_2);
            float cu = sxf + (currentN / 2.f) * sU;
            float cv = syf + (currentN / 2.f) * sV;
            float4 fval = read_mipmap(bg1, bg2, side, float2{cu, cv}, levelofdetail);
            totalWeight = totalWeight + relativeWeight * fval;
            accumulatedProbes += relativeWeight;
            currentN += 2;
        }
        end_result = totalWeight / accumulatedProbes;
    }
    float3 rgb = f3(end_result.x, end_result.y, end_result.z);
    if (dfg.redshift) {
        float z_shift = rdat.z_shift;
        float3 lin_result = srgb_to_lin(rgb);
        float real_sol = 299792458;
        float test_wavelength = 555 / real_sol;
        float local_wavelength = test_wavelength / (z_shift + 1);
        float relative_luminance = 0.2126f * lin_result.x + 0.7152f * lin_result.y + 0.0722f * lin_result.z;
        float new_relative_luminance = powf(local_wavelength, 3) * relative_luminance / powf(test_wavelength, 3);
        new_relative_luminance = clampf(new_relative_luminance, 0.f, 1.f);
        if (relative_luminance > 0.00001f) {
            lin_result = (new_relative_luminance / relative_luminance) * lin_result;
            lin_result = clamp3(lin_result, 0.f, 1.f);
        }
        lin_result = redshift(lin_result, z_shift, dfg);
        lin_result = clamp3(lin_result, 0.f, 1.f);
        rgb = lin_result;  // LINEAR_FRAMEBUFFER
    } else rgb = srgb_to_lin(rgb);
    return f4(rgb.x, rgb.y, rgb.z, end_result.w);
}

// ---- camera on a geodesic (cl.cl:2163-2207, 2569-2620, 2738-2872, 3117-3141, 4735-4940) -----------
static inline float4 parallel_transport_get_velocity(float4 X, float4 geodesic_position, float4 geodesic_velocity, const grt_params& cfg) {
    float X_arr[4], Y_arr[4];
    to_arr(X, X_arr); to_arr(geodesic_velocity, Y_arr);
    float g[16], dg[64], christoffel[64];
    calculate_metric_big(geodesic_position, g, cfg);
    calculate_partial_derivatives_big(geodesic_position, dg, cfg);
    get_christoffel_generic(g, dg, christoffel);
    float vel[4];
    for (int a = 0; a < 4; a++) {
        float sum = 0;
        for (int b = 0; b < 4; b++)
            for (int s = 0; s < 4; s++) sum += christoffel[a * 16 + b * 4 + s] * X_arr[b] * Y_arr[s];
        vel[a] = -sum;
    }
    return from_arr(vel);
}
static inline void parallel_transport_quantity(const float4* geodesic_path, const float4* geodesic_velocity, const float* ds_in, float4 quantity,
                                               int cnt, float4* quantity_out, const grt_params& cfg) {  // cl.cl:2569-2620 (count = 1)
    if (cnt == 0) return;
    float4 current_quantity = quantity;
    quantity_out[0] = current_quantity;
    if (cnt == 1) return;
    for (int kk = 0; kk < cnt - 1; kk++) {
        float ds = ds_in[kk];
        float4 f_x = parallel_transport_get_velocity(current_quantity, geodesic_path[kk], geodesic_velocity[kk], cfg);
        float4 intermediate_next = current_quantity + f_x * ds;
        float4 next = current_quantity + 0.5f * ds * (f_x + parallel_transport_get_velocity(intermediate_next, geodesic_path[kk + 1], geodesic_velocity[kk + 1], cfg));
        quantity_out[kk] = current_quantity;
        current_quantity = next;
    }
    quantity_out[cnt - 1] = current_quantity;
}
static inline float periodic_component_diff(float in1, float in2, float period) {  // cl.cl:3618-3630
    if (period != 0) return circular_diff(in2, in1, period);
    return in1 - in2;
}
// cl.cl:4735-4940 for a single ray. Returns the number of points written.
static inline int get_geodesic_path(const lightray& ray_in, float4* positions_out, float4* velocities_out, float* ds_out, int max_path_length,
                                    const grt_params& cfg, const Features& dfg) {
    float4 position = ray_in.position, velocity = ray_in.velocity, acceleration = ray_in.acceleration;
    float f_in_x = fabsf(velocity.x);
    (void)f_in_x;
#if GRT_CONSTANT_THETA
    position.z = PI_F / 2; velocity.z = 0; acceleration.z = 0;
#endif
#if GRT_ADAPTIVE_PRECISION
    float max_accel = fminf(0.00001000f, dfg.max_acceleration_change);
#endif
    float next_ds = 0.00001f;
#if GRT_ADAPTIVE_PRECISION
    (void)acceleration_to_precision(acceleration, max_accel, &next_ds);
#endif
    float subambient_precision = 0.5f, ambient_precision = 0.2f;
    int bufc = 0;
    float4 periods = get_coordinate_period(cfg);
    float4 last_pos_generic = f4(0, 0, 0, 0);
    (void)periods; (void)last_pos_generic;
    float running_dlambda_dnew = 1;
    for (int i = 0; i < max_path_length; i++) {
#if GRT_CONSTANT_THETA
        position.z = PI_F / 2; velocity.z = 0; acceleration.z = 0;
#endif
        float new_max = dfg.max_precision_radius, new_min = 3;
        float4 polar_position = generic_to_spherical(position, cfg);
#if GRT_CONSTANT_THETA
        polar_position.z = PI_F / 2;
#endif
        float r_value = get_distance_to_object(polar_position, cfg);
        float ds = linear_val(fabsf(r_value), new_min, new_max, ambient_precision, subambient_precision);
#if GRT_ADAPTIVE_PRECISION
        ds = next_ds;
#endif
        if (fabsf(r_value) < new_max) ds = fminf(ds, ambient_precision);
        else ds = (float)(0.1 * (fabsf(r_value) - new_max) + ambient_precision);
        bool should_break = false;
#if !GRT_SINGULAR
        if (fabsf(polar_position.y) >= dfg.universe_size)
#else
        if (fabsf(polar_position.y) < GRT_SINGULAR_TERMINATOR || fabsf(polar_position.y) >= dfg.universe_size)
#endif
            should_break = true;
        float dLambda_dNew = 1;
        float4 next_position, next_velocity, next_acceleration;
        step_verlet(position, velocity, acceleration, ds, &next_position, &next_velocity, &next_acceleration, &dLambda_dNew, cfg, dfg);
        float old_dlambda = running_dlambda_dnew;
        running_dlambda_dnew *= dLambda_dNew;
#if GRT_ADAPTIVE_PRECISION
        if (fabsf(r_value) < new_max) {
            int res = calculate_ds_error(ds, next_acceleration, max_accel, &next_ds, dfg);
            if (res == DS_RETURN) should_break = true;
            if (res == DS_SKIP) continue;  // note: no i-- here (cl.cl:4858-4859)
        }
#endif
#if !GRT_UNCONDITIONALLY_NONSINGULAR
        if (fabsf(velocity.x / running_dlambda_dnew) > 1000 + f_in_x && fabsf(acceleration.x / running_dlambda_dnew) > 100) should_break = true;
#endif
        float4 generic_position_out = position;
        float4 generic_velocity_out = velocity / old_dlambda;
#if GRT_CONSTANT_THETA
        {
            float4 pos_spherical = generic_to_spherical(position, cfg);
            float4 vel_spherical = generic_velocity_to_spherical_velocity(position, velocity / old_dlambda, cfg);
            float fsign = signf(pos_spherical.y);
            pos_spherical.y = fabsf(pos_spherical.y);
            float3 pos_cart = polar_to_cartesian(yzw(pos_spherical));
            float3 vel_cart = spherical_velocity_to_cartesian_velocity(yzw(pos_spherical), yzw(vel_spherical));
            float4 quat = ray_in.initial_quat;
            pos_cart = rot_quat(pos_cart, quat);
            vel_cart = rot_quat(vel_cart, quat);
            float3 next_pos_spherical = cartesian_to_polar(pos_cart);
            float3 next_vel_spherical = cartesian_velocity_to_polar_velocity(pos_cart, vel_cart);
            if (fsign < 0) next_pos_spherical.x = -next_pos_spherical.x;
            float4 next_pos_generic = spherical_to_generic(f4(pos_spherical.x, next_pos_spherical), cfg);
            float4 next_vel_generic = spherical_velocity_to_generic_velocity(f4(pos_spherical.x, next_pos_spherical), f4(vel_spherical.x, next_vel_spherical), cfg);
            if (i != 0) {
                next_pos_generic = f4(periodic_component_diff(next_pos_generic.x, last_pos_generic.x, periods.x) + last_pos_generic.x,
                                      periodic_component_diff(next_pos_generic.y, last_pos_generic.y, periods.y) + last_pos_generic.y,
                                      periodic_component_diff(next_pos_generic.z, last_pos_generic.z, periods.z) + last_pos_generic.z,
                                      periodic_component_diff(next_pos_generic.w, last_pos_generic.w, periods.w) + last_pos_generic.w);
            }
            last_pos_generic = next_pos_generic;
            generic_position_out = next_pos_generic;
            generic_velocity_out = next_vel_generic;
        }
#endif
        if (any_degenerate(next_position) || any_degenerate(next_velocity) || any_degenerate(next_acceleration)) break;
        position = next_position;
        velocity = next_velocity;
        acceleration = next_acceleration;
        positions_out[bufc] = generic_position_out;
        velocities_out[bufc] = generic_velocity_out;
        ds_out[bufc] = ds * old_dlambda;
        bufc++;
        if (should_break) break;
    }
    return bufc;
}
// cl.cl:2738-2872
static inline void handle_interpolating_geodesic(const float4* geodesic_path, const float4* geodesic_velocity, const float* ds_in, float4* camera_generic_out,
                                                 const float4* const t_e[4], float4 e_out[4], float target_time, int cnt, int parallel_transport_observer,
                                                 float3 geodesic_basis_speed, float4* interpolated_geodesic_velocity, const grt_params& cfg) {
    if (cnt == 0) return;
    float4 start_generic = geodesic_path[0];
    if (!parallel_transport_observer) calculate_tetrads(start_generic, geodesic_basis_speed, e_out, cfg, 1);
    else for (int k = 0; k < 4; k++) e_out[k] = t_e[k][0];
    float current_proper_time = 0;
    *camera_generic_out = start_generic;
    *interpolated_geodesic_velocity = geodesic_velocity[0];
    if (cnt == 1) return;
    for (int i = 0; i < cnt - 1; i++) {
        float4 current_pos = geodesic_path[i], next_pos = geodesic_path[i + 1];
        float next_proper_time = current_proper_time + ds_in[i];
        if ((target_time >= current_proper_time && target_time < next_proper_time) || target_time < current_proper_time) {
            float dx = (target_time - current_proper_time) / (next_proper_time - current_proper_time);
            if (target_time < current_proper_time) dx = 0;
            float4 fin_generic = mix(current_pos, next_pos, dx);
            *camera_generic_out = fin_generic;
            float4 oe[4];
            for (int k = 0; k < 4; k++) oe[k] = mix(t_e[k][i], t_e[k][i + 1], dx);
            if (!parallel_transport_observer) calculate_tetrads(fin_generic, geodesic_basis_speed, oe, cfg, 1);
            *interpolated_geodesic_velocity = mix(geodesic_velocity[i], geodesic_velocity[i + 1], dx);
            for (int k = 0; k < 4; k++) e_out[k] = oe[k];
            return;
        }
        current_proper_time = next_proper_time;
    }
    *camera_generic_out = geodesic_path[cnt - 1];
    *interpolated_geodesic_velocity = geodesic_velocity[cnt - 1];
    if (!parallel_transport_observer) calculate_tetrads(geodesic_path[cnt - 1], geodesic_basis_speed, e_out, cfg, 1);
    else for (int k = 0; k < 4; k++) e_out[k] = t_e[k][cnt - 1];
}

}  // namespace oracle
