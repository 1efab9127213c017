// =============================================================================================
// TEST INFRASTRUCTURE ONLY.  A small OpenCL-C-1.2-as-C++ environment, written for this repository,
// that lets g++ compile the reference's device source /root/reference/cl.cl *where it lies* (it is
// piped through the token rewrites in oracle/ref_build.py, never copied into the repository) into
// oracle/_ref/libclref_<metric>.so, so that the reference's OWN kernels can be run on host cores:
//   * to pin the C++ restatement in geodesic_oracle.hpp and the CUDA kernels against them,
//   * as the `cpu_baseline.kind = "reference"` arm of bench.py.
// What is provided: the address-space / kernel qualifiers, float2/3/4 + int2/3/4 with the swizzles
// cl.cl uses, the OpenCL builtins cl.cl calls (IEEE single precision; `native_*` map to the libm
// functions), work-item ids (thread-local, set by the driver loop in ref_driver.cpp), atomics,
// and a normalised-REPEAT-LINEAR `read_imagef` on an RGBA8 2-D array image (OpenCL 1.2 spec 8.2).
// =============================================================================================
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#pragma GCC diagnostic ignored "-Wunused-variable"
#pragma GCC diagnostic ignored "-Wunused-parameter"
#pragma GCC diagnostic ignored "-Wreturn-type"
#pragma GCC diagnostic ignored "-Wnarrowing"
#pragma GCC diagnostic ignored "-Wunused-value"
#pragma GCC diagnostic ignored "-Wsign-compare"
#pragma GCC diagnostic ignored "-Wunused-but-set-variable"

// ---- qualifiers -------------------------------------------------------------------------------
#define __kernel extern "C" __attribute__((visibility("default")))
#define kernel __kernel
#define __global
#define global
#define __constant
#define constant
#define __local
#define local
#define __private
#define __read_only
#define __write_only
#define restrict __restrict__
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2
inline void barrier(int) {}
inline void mem_fence(int) {}
#define printf(...) ((void)0)

typedef unsigned int uint;
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned long ulong;

// ---- vectors ------------------------------------------------------------------------------------
namespace clshim {

template <typename T, int N> struct vec;

// a swizzle view onto the storage of a parent vector (same address, union member)
template <typename T, int M, int... I> struct swz {
    T d[M];
    static constexpr int n = sizeof...(I);
    operator vec<T, n>() const { return vec<T, n>{d[I]...}; }
    swz& operator=(const vec<T, n>& v) { int k = 0; T tmp[n]; for (int i = 0; i < n; i++) tmp[i] = v[i]; ((d[I] = tmp[k++]), ...); return *this; }
    swz& operator=(const swz& o) { return *this = (vec<T, n>)o; }
    swz& operator+=(const vec<T, n>& v) { return *this = (vec<T, n>)*this + v; }
    swz& operator-=(const vec<T, n>& v) { return *this = (vec<T, n>)*this - v; }
    swz& operator*=(const vec<T, n>& v) { return *this = (vec<T, n>)*this * v; }
    swz& operator/=(const vec<T, n>& v) { return *this = (vec<T, n>)*this / v; }
    swz& operator*=(T v) { return *this = (vec<T, n>)*this * v; }
    swz& operator/=(T v) { return *this = (vec<T, n>)*this / v; }
};

template <typename T> struct alignas(sizeof(T) * 2) vec<T, 2> {
    union {
        struct { T x, y; };
        T s[2];
        swz<T, 2, 0, 1> xy;
        swz<T, 2, 1, 0> yx;
    };
    vec() : s{0, 0} {}
    vec(T a) : s{a, a} {}
    vec(T a, T b) : s{a, b} {}
    vec(const vec& o) : s{o.s[0], o.s[1]} {}
    vec& operator=(const vec& o) { s[0] = o.s[0]; s[1] = o.s[1]; return *this; }
    T& operator[](int i) { return s[i]; }
    const T& operator[](int i) const { return s[i]; }
};

template <typename T> struct alignas(sizeof(T) * 4) vec<T, 3> {   // OpenCL: a 3-vector has the size and alignment of a 4-vector
    union {
        struct { T x, y, z; };
        T s[3];
        swz<T, 3, 0, 1> xy;
        swz<T, 3, 1, 0> yx;
        swz<T, 3, 1, 2> yz;
        swz<T, 3, 0, 1, 2> xyz;
    };
    vec() : s{0, 0, 0} {}
    vec(T a) : s{a, a, a} {}
    vec(T a, T b, T c) : s{a, b, c} {}
    vec(const vec<T, 2>& a, T c) : s{a.x, a.y, c} {}
    vec(T a, const vec<T, 2>& b) : s{a, b.x, b.y} {}
    vec(const vec& o) : s{o.s[0], o.s[1], o.s[2]} {}
    vec& operator=(const vec& o) { s[0] = o.s[0]; s[1] = o.s[1]; s[2] = o.s[2]; return *this; }
    T& operator[](int i) { return s[i]; }
    const T& operator[](int i) const { return s[i]; }
};

template <typename T> struct alignas(sizeof(T) * 4) vec<T, 4> {
    union {
        struct { T x, y, z, w; };
        T s[4];
        swz<T, 4, 0, 1> xy;
        swz<T, 4, 1, 0> yx;
        swz<T, 4, 1, 2> yz;
        swz<T, 4, 2, 3> zw;
        swz<T, 4, 0, 1, 2> xyz;
        swz<T, 4, 1, 2, 3> yzw;
    };
    vec() : s{0, 0, 0, 0} {}
    vec(T a) : s{a, a, a, a} {}
    vec(T a, T b, T c, T d) : s{a, b, c, d} {}
    vec(const vec<T, 3>& a, T d) : s{a.x, a.y, a.z, d} {}
    vec(T a, const vec<T, 3>& b) : s{a, b.x, b.y, b.z} {}
    vec(const vec<T, 2>& a, T c, T d) : s{a.x, a.y, c, d} {}
    vec(const vec<T, 2>& a, const vec<T, 2>& b) : s{a.x, a.y, b.x, b.y} {}
    vec(T a, T b, const vec<T, 2>& c) : s{a, b, c.x, c.y} {}
    vec(const vec& o) : s{o.s[0], o.s[1], o.s[2], o.s[3]} {}
    vec& operator=(const vec& o) { s[0] = o.s[0]; s[1] = o.s[1]; s[2] = o.s[2]; s[3] = o.s[3]; return *this; }
    T& operator[](int i) { return s[i]; }
    const T& operator[](int i) const { return s[i]; }
};

#define CLSHIM_BINOP(op)                                                                                                   \
    template <typename T, int N> inline vec<T, N> operator op(const vec<T, N>& a, const vec<T, N>& b) {                    \
        vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] op b[i]; return r; }                                          \
    template <typename T, int N> inline vec<T, N> operator op(const vec<T, N>& a, T b) {                                   \
        vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] op b; return r; }                                             \
    template <typename T, int N> inline vec<T, N> operator op(T a, const vec<T, N>& b) {                                   \
        vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a op b[i]; return r; }                                             \
    template <typename T, int N> inline vec<T, N>& operator op##=(vec<T, N>& a, const vec<T, N>& b) {                      \
        for (int i = 0; i < N; i++) a[i] op##= b[i]; return a; }                                                           \
    template <typename T, int N> inline vec<T, N>& operator op##=(vec<T, N>& a, T b) {                                     \
        for (int i = 0; i < N; i++) a[i] op##= b; return a; }
CLSHIM_BINOP(+)
CLSHIM_BINOP(-)
CLSHIM_BINOP(*)
CLSHIM_BINOP(/)
#undef CLSHIM_BINOP

// float vectors mixed with int / double scalars (OpenCL converts the scalar to the element type)
#define CLSHIM_MIXOP(op, S)                                                                                                \
    template <int N> inline vec<float, N> operator op(const vec<float, N>& a, S b) { return a op (float)b; }               \
    template <int N> inline vec<float, N> operator op(S a, const vec<float, N>& b) { return (float)a op b; }
CLSHIM_MIXOP(+, int) CLSHIM_MIXOP(-, int) CLSHIM_MIXOP(*, int) CLSHIM_MIXOP(/, int)
CLSHIM_MIXOP(+, double) CLSHIM_MIXOP(-, double) CLSHIM_MIXOP(*, double) CLSHIM_MIXOP(/, double)
#undef CLSHIM_MIXOP

template <typename T, int N> inline vec<T, N> operator-(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i < N; i++) r[i] = -a[i]; return r; }

// swizzles take part in arithmetic through their vector value
#define CLSHIM_SWZOP(op)                                                                                                   \
    template <typename T, int M, int... I, typename R> inline auto operator op(const swz<T, M, I...>& a, const R& b) {     \
        return (vec<T, sizeof...(I)>)a op b; }                                                                             \
    template <typename T, int M, int... I, int N> inline auto operator op(const vec<T, N>& a, const swz<T, M, I...>& b) {  \
        return a op (vec<T, sizeof...(I)>)b; }                                                                             \
    template <typename T, int M, int... I> inline auto operator op(T a, const swz<T, M, I...>& b) {                        \
        return a op (vec<T, sizeof...(I)>)b; }
CLSHIM_SWZOP(+)
CLSHIM_SWZOP(-)
CLSHIM_SWZOP(*)
CLSHIM_SWZOP(/)
#undef CLSHIM_SWZOP
template <typename T, int M, int... I> inline auto operator-(const swz<T, M, I...>& a) { return -(vec<T, sizeof...(I)>)a; }

// relational operators return int vectors with -1 for true (OpenCL 1.2 6.3.d)
#define CLSHIM_RELOP(op)                                                                                                   \
    template <typename T, int N> inline vec<int, N> operator op(const vec<T, N>& a, const vec<T, N>& b) {                  \
        vec<int, N> r; for (int i = 0; i < N; i++) r[i] = (a[i] op b[i]) ? -1 : 0; return r; }
CLSHIM_RELOP(==)
CLSHIM_RELOP(!=)
CLSHIM_RELOP(<)
CLSHIM_RELOP(>)
CLSHIM_RELOP(<=)
CLSHIM_RELOP(>=)
#undef CLSHIM_RELOP
template <int N> inline vec<int, N> operator!(const vec<int, N>& a) { vec<int, N> r; for (int i = 0; i < N; i++) r[i] = a[i] ? 0 : -1; return r; }
template <int N> inline vec<int, N> operator||(const vec<int, N>& a, const vec<int, N>& b) { vec<int, N> r; for (int i = 0; i < N; i++) r[i] = (a[i] || b[i]) ? -1 : 0; return r; }
template <int N> inline vec<int, N> operator&&(const vec<int, N>& a, const vec<int, N>& b) { vec<int, N> r; for (int i = 0; i < N; i++) r[i] = (a[i] && b[i]) ? -1 : 0; return r; }

}  // namespace clshim

typedef clshim::vec<float, 2> float2;
typedef clshim::vec<float, 3> float3;
typedef clshim::vec<float, 4> float4;
typedef clshim::vec<int, 2> int2;
typedef clshim::vec<int, 3> int3;
typedef clshim::vec<int, 4> int4;

// ---- builtins -------------------------------------------------------------------------------------
#define CLSHIM_MAP1(name, expr)                                                                                            \
    inline float name(float a) { return expr; }                                                                            \
    template <int N> inline clshim::vec<float, N> name(const clshim::vec<float, N>& v) {                                   \
        clshim::vec<float, N> r; for (int i = 0; i < N; i++) { float a = v[i]; r[i] = expr; } return r; }
#define CLSHIM_MAP1_SWZ(name)                                                                                              \
    template <int M, int... I> inline auto name(const clshim::swz<float, M, I...>& v) { return name((clshim::vec<float, sizeof...(I)>)v); }

// scalar double overloads (C++ <cmath>) remain visible; these are the float forms cl.cl means
CLSHIM_MAP1(native_sqrt, ::sqrtf(a))
CLSHIM_MAP1(native_rsqrt, 1.0f / ::sqrtf(a))
CLSHIM_MAP1(rsqrt, 1.0f / ::sqrtf(a))
CLSHIM_MAP1(native_sin, ::sinf(a))
CLSHIM_MAP1(native_cos, ::cosf(a))
CLSHIM_MAP1(native_tan, ::tanf(a))
CLSHIM_MAP1(native_exp, ::expf(a))
CLSHIM_MAP1(native_log, ::logf(a))
CLSHIM_MAP1(native_log2, ::log2f(a))
CLSHIM_MAP1(native_recip, 1.0f / a)
CLSHIM_MAP1(sign, (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : ((a == 0.0f) ? a : 0.0f)))
inline float native_divide(float a, float b) { return a / b; }
inline float native_powr(float a, float b) { return ::powf(a, b); }

// element-wise forms of the <cmath> functions for vectors
#define CLSHIM_VEC1(name, fn)                                                                                              \
    template <int N> inline clshim::vec<float, N> name(const clshim::vec<float, N>& v) {                                   \
        clshim::vec<float, N> r; for (int i = 0; i < N; i++) r[i] = fn(v[i]); return r; }                                  \
    CLSHIM_MAP1_SWZ(name)
CLSHIM_VEC1(sqrt, ::sqrtf) CLSHIM_VEC1(fabs, ::fabsf) CLSHIM_VEC1(floor, ::floorf) CLSHIM_VEC1(ceil, ::ceilf) CLSHIM_VEC1(round, ::roundf)
CLSHIM_VEC1(sin, ::sinf) CLSHIM_VEC1(cos, ::cosf) CLSHIM_VEC1(tan, ::tanf) CLSHIM_VEC1(exp, ::expf) CLSHIM_VEC1(log, ::logf)
CLSHIM_VEC1(tanh, ::tanhf) CLSHIM_VEC1(acos, ::acosf) CLSHIM_VEC1(asin, ::asinf) CLSHIM_VEC1(atan, ::atanf)

// float scalars: route the unsuffixed names to the single-precision libm entry points (C++ overloads
// for float exist in <cmath>; integer / mixed arguments would otherwise promote to double)
using std::acos; using std::asin; using std::atan; using std::atan2; using std::ceil; using std::cos; using std::cosh; using std::exp;
using std::fabs; using std::floor; using std::fmod; using std::log; using std::log2; using std::pow; using std::round; using std::sin;
using std::sinh; using std::sqrt; using std::tan; using std::tanh; using std::exp2; using std::log10; using std::trunc;
inline float fmod(float a, int b) { return ::fmodf(a, (float)b); }
inline float pow(float a, int b) { return ::powf(a, (float)b); }
inline float atan2(float a, int b) { return ::atan2f(a, (float)b); }
inline float atan2(int a, float b) { return ::atan2f((float)a, b); }
inline float fmod(float a, double b) { return ::fmodf(a, (float)b); }
inline float pow(float a, double b) { return ::powf(a, (float)b); }
inline float pow(double a, float b) { return ::powf((float)a, b); }
inline float atan2(float a, double b) { return ::atan2f(a, (float)b); }
inline float atan2(double a, float b) { return ::atan2f((float)a, b); }

template <int N> inline clshim::vec<float, N> fmod(const clshim::vec<float, N>& a, float b) { clshim::vec<float, N> r; for (int i = 0; i < N; i++) r[i] = ::fmodf(a[i], b); return r; }
template <int N> inline clshim::vec<float, N> fmod(const clshim::vec<float, N>& a, const clshim::vec<float, N>& b) { clshim::vec<float, N> r; for (int i = 0; i < N; i++) r[i] = ::fmodf(a[i], b[i]); return r; }
template <int N> inline clshim::vec<float, N> pow(const clshim::vec<float, N>& a, float b) { clshim::vec<float, N> r; for (int i = 0; i < N; i++) r[i] = ::powf(a[i], b); return r; }
template <int N> inline clshim::vec<float, N> pow(const clshim::vec<float, N>& a, const clshim::vec<float, N>& b) { clshim::vec<float, N> r; for (int i = 0; i < N; i++) r[i] = ::powf(a[i], b[i]); return r; }

// min / max / clamp / mix.  Scalar float min/max come from /root/reference/common.cl itself under
// __cplusplus (it defines them with std::min / std::max); the mixed forms are added here.
constexpr float min(float v1, float v2);
constexpr float max(float v1, float v2);
constexpr int min(int a, int b) { return a < b ? a : b; }
constexpr int max(int a, int b) { return a > b ? a : b; }
constexpr float min(float a, int b) { return a < (float)b ? a : (float)b; }
constexpr float max(float a, int b) { return a > (float)b ? a : (float)b; }
constexpr float min(int a, float b) { return (float)a < b ? (float)a : b; }
constexpr float max(int a, float b) { return (float)a > b ? (float)a : b; }
constexpr float min(float a, double b) { return a < (float)b ? a : (float)b; }
constexpr float max(float a, double b) { return a > (float)b ? a : (float)b; }
constexpr float min(double a, float b) { return (float)a < b ? (float)a : b; }
constexpr float max(double a, float b) { return (float)a > b ? (float)a : b; }
template <typename T, int N> inline clshim::vec<T, N> min(const clshim::vec<T, N>& a, const clshim::vec<T, N>& b) { clshim::vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] < b[i] ? a[i] : b[i]; return r; }
template <typename T, int N> inline clshim::vec<T, N> max(const clshim::vec<T, N>& a, const clshim::vec<T, N>& b) { clshim::vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] > b[i] ? a[i] : b[i]; return r; }
template <typename T, int N> inline clshim::vec<T, N> min(const clshim::vec<T, N>& a, T b) { clshim::vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] < b ? a[i] : b; return r; }
template <typename T, int N> inline clshim::vec<T, N> max(const clshim::vec<T, N>& a, T b) { clshim::vec<T, N> r; for (int i = 0; i < N; i++) r[i] = a[i] > b ? a[i] : b; return r; }

inline float clamp(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline float clamp(float x, int lo, int hi) { return clamp(x, (float)lo, (float)hi); }
inline float clamp(float x, double lo, double hi) { return clamp(x, (float)lo, (float)hi); }
inline float clamp(float x, int lo, float hi) { return clamp(x, (float)lo, hi); }
inline float clamp(float x, float lo, int hi) { return clamp(x, lo, (float)hi); }
inline int clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
template <int N> inline clshim::vec<float, N> clamp(const clshim::vec<float, N>& x, float lo, float hi) { clshim::vec<float, N> r; for (int i = 0; i < N; i++) r[i] = clamp(x[i], lo, hi); return r; }
template <int N> inline clshim::vec<float, N> clamp(const clshim::vec<float, N>& x, const clshim::vec<float, N>& lo, const clshim::vec<float, N>& hi) { clshim::vec<float, N> r; for (int i = 0; i < N; i++) r[i] = clamp(x[i], lo[i], hi[i]); return r; }
template <int N> inline clshim::vec<int, N> clamp(const clshim::vec<int, N>& x, int lo, int hi) { clshim::vec<int, N> r; for (int i = 0; i < N; i++) r[i] = clamp(x[i], lo, hi); return r; }
template <int N> inline clshim::vec<int, N> clamp(const clshim::vec<int, N>& x, const clshim::vec<int, N>& lo, const clshim::vec<int, N>& hi) { clshim::vec<int, N> r; for (int i = 0; i < N; i++) r[i] = clamp(x[i], lo[i], hi[i]); return r; }

inline float mix(float a, float b, float t) { return a + (b - a) * t; }
template <int N> inline clshim::vec<float, N> mix(const clshim::vec<float, N>& a, const clshim::vec<float, N>& b, float t) { return a + (b - a) * t; }
template <int N> inline clshim::vec<float, N> mix(const clshim::vec<float, N>& a, const clshim::vec<float, N>& b, const clshim::vec<float, N>& t) { return a + (b - a) * t; }
template <int M, int... I, int M2, int... J, typename W> inline auto mix(const clshim::swz<float, M, I...>& a, const clshim::swz<float, M2, J...>& b, const W& t) {
    return mix((clshim::vec<float, sizeof...(I)>)a, (clshim::vec<float, sizeof...(J)>)b, t); }
inline float mad(float a, float b, float c) { return a * b + c; }
inline float fma(float a, float b, float c) { return ::fmaf(a, b, c); }

// geometric
template <int N> inline float dot(const clshim::vec<float, N>& a, const clshim::vec<float, N>& b) { float s = 0; for (int i = 0; i < N; i++) s += a[i] * b[i]; return s; }
inline float dot(float a, float b) { return a * b; }
template <int N> inline float length(const clshim::vec<float, N>& a) { return ::sqrtf(dot(a, a)); }
inline float length(float a) { return ::fabsf(a); }
template <int N> inline float fast_length(const clshim::vec<float, N>& a) { return ::sqrtf(dot(a, a)); }
template <int N> inline float distance(const clshim::vec<float, N>& a, const clshim::vec<float, N>& b) { return length(a - b); }
template <int N> inline clshim::vec<float, N> normalize(const clshim::vec<float, N>& a) { return a / length(a); }
template <int N> inline clshim::vec<float, N> fast_normalize(const clshim::vec<float, N>& a) { return a / length(a); }
inline float3 cross(const float3& a, const float3& b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float4 cross(const float4& a, const float4& b) { return float4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f); }
#define CLSHIM_SWZ_GEOM(name)                                                                                              \
    template <int M, int... I> inline auto name(const clshim::swz<float, M, I...>& v) { return name((clshim::vec<float, sizeof...(I)>)v); }
CLSHIM_SWZ_GEOM(length) CLSHIM_SWZ_GEOM(fast_length) CLSHIM_SWZ_GEOM(normalize) CLSHIM_SWZ_GEOM(fast_normalize)
template <int M, int... I, int M2, int... J> inline float dot(const clshim::swz<float, M, I...>& a, const clshim::swz<float, M2, J...>& b) { return dot((clshim::vec<float, sizeof...(I)>)a, (clshim::vec<float, sizeof...(J)>)b); }
template <int M, int... I> inline float dot(const clshim::swz<float, M, I...>& a, const clshim::vec<float, sizeof...(I)>& b) { return dot((clshim::vec<float, sizeof...(I)>)a, b); }
template <int M, int... I> inline float dot(const clshim::vec<float, sizeof...(I)>& a, const clshim::swz<float, M, I...>& b) { return dot(a, (clshim::vec<float, sizeof...(I)>)b); }

// relational
using std::isnan; using std::isfinite; using std::isinf;
template <int N> inline clshim::vec<int, N> isnan(const clshim::vec<float, N>& a) { clshim::vec<int, N> r; for (int i = 0; i < N; i++) r[i] = std::isnan(a[i]) ? -1 : 0; return r; }
template <int N> inline clshim::vec<int, N> isfinite(const clshim::vec<float, N>& a) { clshim::vec<int, N> r; for (int i = 0; i < N; i++) r[i] = std::isfinite(a[i]) ? -1 : 0; return r; }
template <int N> inline clshim::vec<int, N> isinf(const clshim::vec<float, N>& a) { clshim::vec<int, N> r; for (int i = 0; i < N; i++) r[i] = std::isinf(a[i]) ? -1 : 0; return r; }
template <int N> inline int any(const clshim::vec<int, N>& a) { for (int i = 0; i < N; i++) if (a[i] < 0) return 1; return 0; }   // MSB set
template <int N> inline int all(const clshim::vec<int, N>& a) { for (int i = 0; i < N; i++) if (!(a[i] < 0)) return 0; return 1; }
inline int any(int a) { return a < 0; }
inline int all(int a) { return a < 0; }
template <int N> inline clshim::vec<float, N> select(const clshim::vec<float, N>& a, const clshim::vec<float, N>& b, const clshim::vec<int, N>& c) { clshim::vec<float, N> r; for (int i = 0; i < N; i++) r[i] = c[i] < 0 ? b[i] : a[i]; return r; }
inline float select(float a, float b, int c) { return c ? b : a; }

// conversions
inline int convert_int(float a) { return (int)a; }
inline float convert_float(int a) { return (float)a; }
template <int N> inline clshim::vec<int, N> convert_int_n(const clshim::vec<float, N>& a) { clshim::vec<int, N> r; for (int i = 0; i < N; i++) r[i] = (int)a[i]; return r; }
inline int2 convert_int2(const float2& a) { return convert_int_n(a); }
inline int4 convert_int4(const float4& a) { return convert_int_n(a); }
inline float2 convert_float2(const int2& a) { return float2((float)a.x, (float)a.y); }
inline float4 convert_float4(const int4& a) { return float4((float)a.x, (float)a.y, (float)a.z, (float)a.w); }

// ---- work items -------------------------------------------------------------------------------------
namespace clshim {
struct WorkItem { size_t gid[3], lid[3], grp[3], gsz[3], lsz[3]; };
inline thread_local WorkItem work_item = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {1, 1, 1}, {1, 1, 1}};
}  // namespace clshim
inline size_t get_global_id(int d) { return clshim::work_item.gid[d]; }
inline size_t get_local_id(int d) { return clshim::work_item.lid[d]; }
inline size_t get_group_id(int d) { return clshim::work_item.grp[d]; }
inline size_t get_global_size(int d) { return clshim::work_item.gsz[d]; }
inline size_t get_local_size(int d) { return clshim::work_item.lsz[d]; }
inline size_t get_num_groups(int d) { return clshim::work_item.gsz[d] / clshim::work_item.lsz[d]; }

// ---- atomics (32-bit, relaxed) ------------------------------------------------------------------------
inline int atomic_add(volatile int* p, int v) { return __atomic_fetch_add((int*)p, v, __ATOMIC_RELAXED); }
inline int atomic_sub(volatile int* p, int v) { return __atomic_fetch_sub((int*)p, v, __ATOMIC_RELAXED); }
inline int atomic_inc(volatile int* p) { return __atomic_fetch_add((int*)p, 1, __ATOMIC_RELAXED); }
inline int atomic_dec(volatile int* p) { return __atomic_fetch_sub((int*)p, 1, __ATOMIC_RELAXED); }
inline int atomic_xchg(volatile int* p, int v) { return __atomic_exchange_n((int*)p, v, __ATOMIC_RELAXED); }
inline int atomic_min(volatile int* p, int v) { int o = __atomic_load_n((int*)p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n((int*)p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline int atomic_max(volatile int* p, int v) { int o = __atomic_load_n((int*)p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n((int*)p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }

// ---- images -------------------------------------------------------------------------------------------
typedef unsigned int sampler_t;
#define CLK_NORMALIZED_COORDS_FALSE 0u
#define CLK_NORMALIZED_COORDS_TRUE 1u
#define CLK_ADDRESS_NONE 0u
#define CLK_ADDRESS_CLAMP_TO_EDGE 2u
#define CLK_ADDRESS_CLAMP 4u
#define CLK_ADDRESS_REPEAT 6u
#define CLK_FILTER_NEAREST 0x10u
#define CLK_FILTER_LINEAR 0x20u

struct image2d_t {        // write-only float4 RGBA target (the reference's GL texture is RGBA32F)
    float* data; int width, height;
};
struct image2d_array_t {  // read-only CL_RGBA / CL_UNORM_INT8 array (graphics_settings.cpp:205-211)
    const unsigned char* data; int width, height, layers;
};
inline int get_image_width(const image2d_t& i) { return i.width; }
inline int get_image_height(const image2d_t& i) { return i.height; }
inline int get_image_width(const image2d_array_t& i) { return i.width; }
inline int get_image_height(const image2d_array_t& i) { return i.height; }
inline int2 get_image_dim(const image2d_array_t& i) { return int2(i.width, i.height); }
inline int2 get_image_dim(const image2d_t& i) { return int2(i.width, i.height); }
inline size_t get_image_array_size(const image2d_array_t& i) { return (size_t)i.layers; }
inline void write_imagef(image2d_t img, int2 c, float4 v) {
    if (c.x < 0 || c.y < 0 || c.x >= img.width || c.y >= img.height) return;
    float* p = img.data + ((size_t)c.y * img.width + c.x) * 4;
    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
}
namespace clshim {
inline float4 texel(const image2d_array_t& img, int x, int y, int layer) {
    const unsigned char* p = img.data + (((size_t)layer * img.height + y) * img.width + x) * 4;
    return float4(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f);
}
}  // namespace clshim
// OpenCL 1.2 spec 8.2 (normalised coordinates, CLK_ADDRESS_REPEAT, CLK_FILTER_LINEAR) and 8.4
// (array layer = clamp(rint(w), 0, layers - 1)).  The only sampler cl.cl's live path uses (cl.cl:5425).
inline float4 read_imagef(const image2d_array_t& img, sampler_t sam, float4 c) {
    int layer = (int)std::fmin(std::fmax(::rintf(c.z), 0.0f), (float)(img.layers - 1));
    if (!(sam & CLK_NORMALIZED_COORDS_TRUE)) { c.x /= img.width; c.y /= img.height; }
    if (sam & CLK_FILTER_LINEAR) {
        float u = (c.x - ::floorf(c.x)) * img.width, v = (c.y - ::floorf(c.y)) * img.height;
        int i0 = (int)::floorf(u - 0.5f), j0 = (int)::floorf(v - 0.5f);
        int i1 = i0 + 1, j1 = j0 + 1;
        if (i0 < 0) i0 += img.width;
        if (i1 > img.width - 1) i1 -= img.width;
        if (j0 < 0) j0 += img.height;
        if (j1 > img.height - 1) j1 -= img.height;
        float a = (u - 0.5f) - ::floorf(u - 0.5f), b = (v - 0.5f) - ::floorf(v - 0.5f);
        return (1 - a) * (1 - b) * clshim::texel(img, i0, j0, layer) + a * (1 - b) * clshim::texel(img, i1, j0, layer) +
               (1 - a) * b * clshim::texel(img, i0, j1, layer) + a * b * clshim::texel(img, i1, j1, layer);
    }
    float u = (c.x - ::floorf(c.x)) * img.width, v = (c.y - ::floorf(c.y)) * img.height;
    int i = (int)::floorf(u), j = (int)::floorf(v);
    if (i > img.width - 1) i -= img.width;
    if (j > img.height - 1) j -= img.height;
    return clshim::texel(img, i, j, layer);
}
