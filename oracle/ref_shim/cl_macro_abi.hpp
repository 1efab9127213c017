// =============================================================================================
// TEST INFRASTRUCTURE ONLY.  The generated-code ABI of the reference, i.e. the `-D` macro set that
// /root/reference/metric.hpp:725-959 (build_argument_string) hands to cl.cl, expressed over the
// functions this repository's emitter generates for a metric (GRT_METRIC_HEADER).  The reference
// produces these macros with its symbolic library deps/vec (20k/funvectormaths, an empty submodule
// here), so the *expressions* come from our emitter; everything that consumes them is the
// reference's own cl.cl.  Consumption sites: cl.cl:969-1199 (F*_I, F*_P, TEMPORARIES0),
// cl.cl:1202-1270 (TO_/FROM_[D]COORD), cl.cl:1338-1355 (COORDINATE_PERIODICITY), cl.cl:3253-3311
// (FIX_LIGHT, GEO_ACCEL), cl.cl:3377-3387 (DISTANCE_FUNC), cl.cl:940-964 (DYNVARS, features).
//
// Macro evaluation order relied on (checked against the cited lines): in every consumer the
// lowest-numbered macro of a family is evaluated first, so F1_I / F1_P / GEO_ACCEL0 fill a local
// array (declared through TEMPORARIES0, which every one of those sites declares as
// `float TEMPORARIES0;`) and the others read it.  GEO_ACCEL* use the reference formulation
// `grt_accel_ref` (symbolic Christoffel symbols with the symbolic inverse, metric.hpp:184-244).
// =============================================================================================
#pragma once
#include <cmath>
#define GRT_FN static inline __attribute__((always_inline))
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

namespace clref {
struct f4 { float a[4]; };
inline thread_local long long trial_steps = 0;   // one GEO_ACCEL0 evaluation = one trial step of step_verlet (cl.cl:3299)
#define CLREF_P const grt_params& P = *(const grt_params*)cfg; const float p[4] = {v1, v2, v3, v4}
GRT_FN float metric(float v1, float v2, float v3, float v4, const void* cfg, float* g) { CLREF_P; grt_metric(p, P, g); return g[0]; }
GRT_FN float partials(float v1, float v2, float v3, float v4, const void* cfg, float* dg, int first) { CLREF_P; grt_metric_partials(p, P, dg); return dg[first]; }
GRT_FN float accel(float v1, float v2, float v3, float v4, float iv1, float iv2, float iv3, float iv4, const void* cfg, float* a) {
    CLREF_P; const float v[4] = {iv1, iv2, iv3, iv4}; grt_accel_ref(p, v, P, a); trial_steps++; return a[0]; }
GRT_FN f4 to_polar(float v1, float v2, float v3, float v4, const void* cfg) { CLREF_P; f4 o; grt_to_polar(p, P, o.a); return o; }
GRT_FN f4 from_polar(float v1, float v2, float v3, float v4, const void* cfg) { CLREF_P; f4 o; grt_from_polar(p, P, o.a); return o; }
GRT_FN f4 to_polar_vel(float v1, float v2, float v3, float v4, float dv1, float dv2, float dv3, float dv4, const void* cfg) {
    CLREF_P; const float v[4] = {dv1, dv2, dv3, dv4}; f4 o; grt_to_polar_vel(p, v, P, o.a); return o; }
GRT_FN f4 from_polar_vel(float v1, float v2, float v3, float v4, float dv1, float dv2, float dv3, float dv4, const void* cfg) {
    CLREF_P; const float v[4] = {dv1, dv2, dv3, dv4}; f4 o; grt_from_polar_vel(p, v, P, o.a); return o; }
GRT_FN float distance(float v1, float v2, float v3, float v4, const void* cfg) { CLREF_P; float o; grt_distance(p, P, &o); return o; }
GRT_FN f4 period(const void* cfg) { const grt_params& P = *(const grt_params*)cfg; f4 o; grt_period(P, o.a); return o; }
#undef CLREF_P
}  // namespace clref

#define RS_IMPL 1
#define C_IMPL 1
#define GENERIC_METRIC
#define VERLET_INTEGRATION_GENERIC
#define TEMPORARIES0 clref_g[16], clref_dg[64], clref_acc[4]

#if GRT_BIG_METRIC
#define GENERIC_BIG_METRIC
#define F1_I clref::metric(v1, v2, v3, v4, cfg, clref_g)
#define F2_I clref_g[1]
#define F3_I clref_g[2]
#define F4_I clref_g[3]
#define F5_I clref_g[4]
#define F6_I clref_g[5]
#define F7_I clref_g[6]
#define F8_I clref_g[7]
#define F9_I clref_g[8]
#define F10_I clref_g[9]
#define F11_I clref_g[10]
#define F12_I clref_g[11]
#define F13_I clref_g[12]
#define F14_I clref_g[13]
#define F15_I clref_g[14]
#define F16_I clref_g[15]
#define F1_P clref::partials(v1, v2, v3, v4, cfg, clref_dg, 0)
#define F2_P clref_dg[1]
#define F3_P clref_dg[2]
#define F4_P clref_dg[3]
#define F5_P clref_dg[4]
#define F6_P clref_dg[5]
#define F7_P clref_dg[6]
#define F8_P clref_dg[7]
#define F9_P clref_dg[8]
#define F10_P clref_dg[9]
#define F11_P clref_dg[10]
#define F12_P clref_dg[11]
#define F13_P clref_dg[12]
#define F14_P clref_dg[13]
#define F15_P clref_dg[14]
#define F16_P clref_dg[15]
#define F17_P clref_dg[16]
#define F18_P clref_dg[17]
#define F19_P clref_dg[18]
#define F20_P clref_dg[19]
#define F21_P clref_dg[20]
#define F22_P clref_dg[21]
#define F23_P clref_dg[22]
#define F24_P clref_dg[23]
#define F25_P clref_dg[24]
#define F26_P clref_dg[25]
#define F27_P clref_dg[26]
#define F28_P clref_dg[27]
#define F29_P clref_dg[28]
#define F30_P clref_dg[29]
#define F31_P clref_dg[30]
#define F32_P clref_dg[31]
#define F33_P clref_dg[32]
#define F34_P clref_dg[33]
#define F35_P clref_dg[34]
#define F36_P clref_dg[35]
#define F37_P clref_dg[36]
#define F38_P clref_dg[37]
#define F39_P clref_dg[38]
#define F40_P clref_dg[39]
#define F41_P clref_dg[40]
#define F42_P clref_dg[41]
#define F43_P clref_dg[42]
#define F44_P clref_dg[43]
#define F45_P clref_dg[44]
#define F46_P clref_dg[45]
#define F47_P clref_dg[46]
#define F48_P clref_dg[47]
#define F49_P clref_dg[48]
#define F50_P clref_dg[49]
#define F51_P clref_dg[50]
#define F52_P clref_dg[51]
#define F53_P clref_dg[52]
#define F54_P clref_dg[53]
#define F55_P clref_dg[54]
#define F56_P clref_dg[55]
#define F57_P clref_dg[56]
#define F58_P clref_dg[57]
#define F59_P clref_dg[58]
#define F60_P clref_dg[59]
#define F61_P clref_dg[60]
#define F62_P clref_dg[61]
#define F63_P clref_dg[62]
#define F64_P clref_dg[63]
#else
#define F1_I clref::metric(v1, v2, v3, v4, cfg, clref_g)
#define F2_I clref_g[5]
#define F3_I clref_g[10]
#define F4_I clref_g[15]
#define F1_P clref::partials(v1, v2, v3, v4, cfg, clref_dg, 0)
#define F2_P clref_dg[16]
#define F3_P clref_dg[32]
#define F4_P clref_dg[48]
#define F5_P clref_dg[5]
#define F6_P clref_dg[21]
#define F7_P clref_dg[37]
#define F8_P clref_dg[53]
#define F9_P clref_dg[10]
#define F10_P clref_dg[26]
#define F11_P clref_dg[42]
#define F12_P clref_dg[58]
#define F13_P clref_dg[15]
#define F14_P clref_dg[31]
#define F15_P clref_dg[47]
#define F16_P clref_dg[63]
#endif
#define TO_COORD1 clref::to_polar(v1, v2, v3, v4, cfg).a[0]
#define TO_DCOORD1 clref::to_polar_vel(v1, v2, v3, v4, dv1, dv2, dv3, dv4, cfg).a[0]
#define FROM_COORD1 clref::from_polar(v1, v2, v3, v4, cfg).a[0]
#define FROM_DCOORD1 clref::from_polar_vel(v1, v2, v3, v4, dv1, dv2, dv3, dv4, cfg).a[0]
#define TO_COORD2 clref::to_polar(v1, v2, v3, v4, cfg).a[1]
#define TO_DCOORD2 clref::to_polar_vel(v1, v2, v3, v4, dv1, dv2, dv3, dv4, cfg).a[1]
#define FROM_COORD2 clref::from_polar(v1, v2, v3, v4, cfg).a[1]
#define FROM_DCOORD2 clref::from_polar_vel(v1, v2, v3, v4, dv1, dv2, dv3, dv4, cfg).a[1]
#define TO_COORD3 clref::to_polar(v1, v2, v3, v4, cfg).a[2]
#define TO_DCOORD3 clref::to_polar_vel(v1, v2, v3, v4, dv1, dv2, dv3, dv4, cfg).a[2]
#define FROM_COORD3 clref::from_polar(v1, v2, v3, v4, cfg).a[2]
#define FROM_DCOORD3 clref::from_polar_vel(v1, v2, v3, v4, dv1, dv2, dv3, dv4, cfg).a[2]
#define TO_COORD4 clref::to_polar(v1, v2, v3, v4, cfg).a[3]
#define TO_DCOORD4 clref::to_polar_vel(v1, v2, v3, v4, dv1, dv2, dv3, dv4, cfg).a[3]
#define FROM_COORD4 clref::from_polar(v1, v2, v3, v4, cfg).a[3]
#define FROM_DCOORD4 clref::from_polar_vel(v1, v2, v3, v4, dv1, dv2, dv3, dv4, cfg).a[3]
#if GRT_HAS_COORDINATE_PERIODICITY
#define HAS_COORDINATE_PERIODICITY
#define COORDINATE_PERIODICITY1 clref::period(cfg).a[0]
#define COORDINATE_PERIODICITY2 clref::period(cfg).a[1]
#define COORDINATE_PERIODICITY3 clref::period(cfg).a[2]
#define COORDINATE_PERIODICITY4 clref::period(cfg).a[3]
#endif
#define GEO_ACCEL0 clref::accel(v1, v2, v3, v4, iv1, iv2, iv3, iv4, cfg, clref_acc)
#define FIX_LIGHT0 0.0f   /* dead: the only caller is commented out, cl.cl:3315 */
#define CART_TO_POL0 0.0f   /* no live caller, cl.cl:1302-1335 */
#define CART_TO_POL_D0 0.0f
#define GEO_ACCEL1 clref_acc[1]
#define FIX_LIGHT1 0.0f   /* dead: the only caller is commented out, cl.cl:3315 */
#define CART_TO_POL1 0.0f   /* no live caller, cl.cl:1302-1335 */
#define CART_TO_POL_D1 0.0f
#define GEO_ACCEL2 clref_acc[2]
#define FIX_LIGHT2 0.0f   /* dead: the only caller is commented out, cl.cl:3315 */
#define CART_TO_POL2 0.0f   /* no live caller, cl.cl:1302-1335 */
#define CART_TO_POL_D2 0.0f
#define GEO_ACCEL3 clref_acc[3]
#define FIX_LIGHT3 0.0f   /* dead: the only caller is commented out, cl.cl:3315 */
#define CART_TO_POL3 0.0f   /* no live caller, cl.cl:1302-1335 */
#define CART_TO_POL_D3 0.0f
#define METRIC_TIME_G00 0.0f
#define DISTANCE_FUNC clref::distance(v1, v2, v3, v4, cfg)
#define DYNVARS clref_params[GRT_NUM_PARAMS > 0 ? GRT_NUM_PARAMS : 1]
#if GRT_CONSTANT_THETA
#define GENERIC_CONSTANT_THETA
#endif
#if GRT_SINGULAR
#define SINGULAR
#define SINGULAR_TERMINATOR GRT_SINGULAR_TERMINATOR
#if GRT_TRAVERSABLE_EVENT_HORIZON
#define TRAVERSABLE_EVENT_HORIZON
#endif
#endif
#if GRT_ADAPTIVE_PRECISION
#define ADAPTIVE_PRECISION
#if GRT_SINGULARITY_DETECTION
#define SINGULARITY_DETECTION
#endif
#endif
#define W_V1 GRT_W_V1
#define W_V2 GRT_W_V2
#define W_V3 GRT_W_V3
#define W_V4 GRT_W_V4
#if GRT_FORWARD_GEODESIC_PATH
#define FORWARD_GEODESIC_PATH
#endif
#if GRT_HAS_CYLINDRICAL_SINGULARITY
#define HAS_CYLINDRICAL_SINGULARITY
#define CYLINDRICAL_TERMINATOR GRT_CYLINDRICAL_TERMINATOR
#endif
#if GRT_UNCONDITIONALLY_NONSINGULAR
#define UNCONDITIONALLY_NONSINGULAR
#endif
// the dynamic program of dynamic_feature_config.cpp:122-180: floats first, then ints, each
// alphabetical (std::map order, dynamic_feature_config.cpp:190-232; features of main.cpp:1123-1158)
#define KERNEL_IS_DYNAMIC
#define DYNAMIC_FLOAT_FEATURES adaptive_sampling_threshold, field_of_view, max_acceleration_change, max_precision_radius, min_step, ray_skip, universe_size
#define DYNAMIC_BOOL_FEATURES adaptive_sampling, redshift, reparameterisation, use_old_redshift, use_triangle_rendering
// sett.is_srgb = true in the application (main.cpp:1019): the kernel writes linear RGB (cl.cl:5833-5843)
#define LINEAR_FRAMEBUFFER
