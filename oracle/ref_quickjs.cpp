// =============================================================================================
// TEST INFRASTRUCTURE ONLY.  Runs the reference's metric / coordinate / origin scripts in the
// REFERENCE'S OWN JavaScript engine - the vendored QuickJS 2020-04-12 (bignum operator-overloading
// extension), compiled from /root/reference/deps/quickjs where it lies (oracle/ref_build.py) -
// after the reference's own number.js, to pin the product's script engine (csrc/minijs.cpp) and
// symbolic layer (csrc/sym.cpp) against it.
//
// The reference binds CShim / CMath to *symbolic* dual numbers from its deps/vec library (an empty
// submodule here).  This binding follows /root/reference/js_interop.cpp line by line - storage
// {real | complex} hidden on each JS object (:129-149, 299-350, 394-446), CShim.{add,mul,sub,div,
// neg,construct} (:448-478), the CMath table (:481-616, 688-736), comparisons with the `gte` is `>`
// quirk (:251-297), $pin (:352-380), the $cfg proxy with `$default` (:762-846), the 4 -> 16 expansion
// taking real parts (:382-392, 855-901) - but over *numeric* dual numbers: a complex value and its
// four complex partial derivatives in double precision.  One run therefore gives g_ij and all
// d_k g_ij at a point (the reference makes four runs, one per variable, metric.hpp:38-80).
// =============================================================================================
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
#include "quickjs.h"
}

namespace {

typedef std::complex<double> cd;

struct Storage {
    int which = 0;   // 0 real, 1 complex (js_interop.cpp:129-149)
    cd v = 0;
    cd d[4] = {0, 0, 0, 0};
};

struct Session {
    std::vector<std::string> names;
    std::vector<double> defaults, values;
    std::vector<int> has_value;
    std::string error;
    int find(const std::string& n) const { for (size_t i = 0; i < names.size(); i++) if (names[i] == n) return (int)i; return -1; }
};

JSClassID box_class;
thread_local Session* g_session = nullptr;

void box_finalizer(JSRuntime*, JSValue val) { delete (Storage*)JS_GetOpaque(val, box_class); }

JSValue throw_msg(JSContext* ctx, const char* msg) { return JS_ThrowTypeError(ctx, "%s", msg); }

// get (js_interop.cpp:324-350): undefined -> 0, number / boolean -> real constant, object -> hidden storage (default 0)
Storage get(JSContext* ctx, JSValueConst v) {
    Storage s;
    if (JS_IsUndefined(v)) return s;
    if (JS_IsBool(v)) { s.v = JS_ToBool(ctx, v) ? 1.0 : 0.0; return s; }
    if (JS_IsNumber(v)) { double d = 0; JS_ToFloat64(ctx, &d, v); s.v = d; return s; }
    if (JS_IsObject(v)) {
        if (Storage* p = (Storage*)JS_GetOpaque(v, box_class)) return *p;
        JSValue h = JS_GetPropertyStr(ctx, v, "__v");
        if (Storage* p = (Storage*)JS_GetOpaque(h, box_class)) s = *p;
        JS_FreeValue(ctx, h);
        return s;
    }
    double d = 0;
    if (JS_ToFloat64(ctx, &d, v) == 0) s.v = d;   // BigFloat etc. produced under "use math"
    return s;
}

JSValue make_box(JSContext* ctx, const Storage& s) {
    JSValue o = JS_NewObjectClass(ctx, (int)box_class);
    JS_SetOpaque(o, new Storage(s));
    return o;
}

// to_value (js_interop.cpp:394-446): box -> make_class(box) -> new number(box) -> CShim.construct(this, box)
JSValue to_value(JSContext* ctx, const Storage& s) {
    JSValue global = JS_GetGlobalObject(ctx);
    JSValue fn = JS_GetPropertyStr(ctx, global, "make_class");
    JSValue box = make_box(ctx, s);
    JSValue r = JS_Call(ctx, fn, global, 1, &box);
    JS_FreeValue(ctx, box);
    JS_FreeValue(ctx, fn);
    JS_FreeValue(ctx, global);
    return r;
}

// ---- numeric dual arithmetic -----------------------------------------------------------------------
Storage promote(const Storage& a, const Storage& b, Storage r) { r.which = (a.which || b.which) ? 1 : 0; return r; }
Storage s_add(const Storage& a, const Storage& b) { Storage r; r.v = a.v + b.v; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] + b.d[i]; return promote(a, b, r); }
Storage s_sub(const Storage& a, const Storage& b) { Storage r; r.v = a.v - b.v; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] - b.d[i]; return promote(a, b, r); }
Storage s_mul(const Storage& a, const Storage& b) { Storage r; r.v = a.v * b.v; for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return promote(a, b, r); }
Storage s_div(const Storage& a, const Storage& b) {
    Storage r; r.v = a.v / b.v;
    for (int i = 0; i < 4; i++) r.d[i] = (a.d[i] * b.v - a.v * b.d[i]) / (b.v * b.v);
    return promote(a, b, r);
}
Storage s_neg(const Storage& a) { Storage r = a; r.v = -a.v; for (int i = 0; i < 4; i++) r.d[i] = -a.d[i]; return r; }
// f(a) with derivative factor df
Storage chain(const Storage& a, cd f, cd df) { Storage r; r.which = a.which; r.v = f; for (int i = 0; i < 4; i++) r.d[i] = df * a.d[i]; return r; }
Storage real_const(double x) { Storage r; r.v = x; return r; }
double re(const Storage& a) { return a.v.real(); }

#define ARG(i) ((i) < argc ? argv[i] : JS_UNDEFINED)
#define REAL_ONLY(s) if ((s).which != 0) return throw_msg(ctx, "Invalid type in dual/complex internals, this is a real-only function")

JSValue js_binop(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv, int magic) {
    Storage a = get(ctx, ARG(0)), b = get(ctx, ARG(1));
    switch (magic) {
        case 0: return to_value(ctx, s_add(a, b));
        case 1: return to_value(ctx, s_mul(a, b));
        case 2: return to_value(ctx, s_sub(a, b));
        case 3: return to_value(ctx, s_div(a, b));
        default: break;
    }
    if (a.which || b.which) return throw_msg(ctx, "Can only use comparisons on real values");
    bool r = false;
    switch (magic) {
        case 4: r = re(a) < re(b); break;
        case 5: r = re(a) <= re(b); break;
        case 6: r = re(a) == re(b); break;
        case 7: r = re(a) > re(b); break;
        case 8: r = re(a) > re(b); break;    // s_gte uses `>` (js_interop.cpp:291-297)
    }
    return to_value(ctx, real_const(r ? 1.0 : 0.0));
}
JSValue js_neg(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) { return to_value(ctx, s_neg(get(ctx, ARG(0)))); }
JSValue js_construct(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) {   // js_interop.cpp:448-456
    Storage s = get(ctx, ARG(1));
    JSValue box = make_box(ctx, s);
    JS_DefinePropertyValueStr(ctx, ARG(0), "__v", box, JS_PROP_CONFIGURABLE | JS_PROP_WRITABLE);
    return JS_UNDEFINED;
}

enum Unary { U_SIN, U_COS, U_TAN, U_ASIN, U_ACOS, U_ATAN, U_FABS, U_LOG, U_SQRT, U_PSQRT, U_EXP, U_SINH, U_COSH, U_TANH, U_CONJ, U_SELFCONJ, U_REAL, U_IMAG, U_CSQRT };
JSValue js_unary(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv, int magic) {
    Storage a = get(ctx, ARG(0));
    cd x = a.v;
    switch (magic) {
        case U_SIN: return to_value(ctx, chain(a, std::sin(x), std::cos(x)));
        case U_COS: return to_value(ctx, chain(a, std::cos(x), -std::sin(x)));
        case U_TAN: REAL_ONLY(a); return to_value(ctx, chain(a, std::tan(re(a)), 1.0 / (std::cos(re(a)) * std::cos(re(a)))));
        case U_ASIN: REAL_ONLY(a); return to_value(ctx, chain(a, std::asin(re(a)), 1.0 / std::sqrt(1.0 - re(a) * re(a))));
        case U_ACOS: REAL_ONLY(a); return to_value(ctx, chain(a, std::acos(re(a)), -1.0 / std::sqrt(1.0 - re(a) * re(a))));
        case U_ATAN: REAL_ONLY(a); return to_value(ctx, chain(a, std::atan(re(a)), 1.0 / (1.0 + re(a) * re(a))));
        case U_FABS:
            if (a.which == 0) return to_value(ctx, chain(a, std::fabs(re(a)), re(a) > 0 ? 1.0 : (re(a) < 0 ? -1.0 : 0.0)));
            { Storage r = chain(a, std::abs(x), 0.0); r.which = 0; for (int i = 0; i < 4; i++) r.d[i] = (x.real() * a.d[i].real() + x.imag() * a.d[i].imag()) / std::abs(x); return to_value(ctx, r); }
        case U_LOG: REAL_ONLY(a); return to_value(ctx, chain(a, std::log(re(a)), 1.0 / re(a)));
        case U_SQRT:
        case U_PSQRT:
            if (a.which == 0) { double s = std::sqrt(re(a)); return to_value(ctx, chain(a, s, 0.5 / s)); }
            { cd s = std::sqrt(x); return to_value(ctx, chain(a, s, 0.5 / s)); }
        case U_EXP: REAL_ONLY(a); return to_value(ctx, chain(a, std::exp(re(a)), std::exp(re(a))));
        case U_SINH: REAL_ONLY(a); return to_value(ctx, chain(a, std::sinh(re(a)), std::cosh(re(a))));
        case U_COSH: REAL_ONLY(a); return to_value(ctx, chain(a, std::cosh(re(a)), std::sinh(re(a))));
        case U_TANH: REAL_ONLY(a); { double t = std::tanh(re(a)); return to_value(ctx, chain(a, t, 1.0 - t * t)); }
        case U_CONJ: { Storage r = a; r.v = std::conj(a.v); for (int i = 0; i < 4; i++) r.d[i] = std::conj(a.d[i]); return to_value(ctx, r); }
        case U_SELFCONJ: {   // z * conj(z), a real quantity
            Storage c = a; c.v = std::conj(a.v); for (int i = 0; i < 4; i++) c.d[i] = std::conj(a.d[i]);
            Storage r = s_mul(a, c); r.which = a.which; return to_value(ctx, r); }
        case U_REAL: { Storage r; r.v = a.v.real(); for (int i = 0; i < 4; i++) r.d[i] = a.d[i].real(); return to_value(ctx, r); }
        case U_IMAG: { Storage r; r.v = a.v.imag(); for (int i = 0; i < 4; i++) r.d[i] = a.d[i].imag(); return to_value(ctx, r); }
        case U_CSQRT: {
            if (a.which != 0) return throw_msg(ctx, "csqrt must be used with purely real arguments");
            cd s = std::sqrt(cd(re(a), 0.0)); Storage r = chain(a, s, 0.5 / s); r.which = 1; return to_value(ctx, r); }
    }
    return JS_UNDEFINED;
}
JSValue js_atan2(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) {
    Storage y = get(ctx, ARG(0)), x = get(ctx, ARG(1));
    Storage r; r.v = std::atan2(re(y), re(x));
    double den = re(x) * re(x) + re(y) * re(y);
    for (int i = 0; i < 4; i++) r.d[i] = (re(x) * y.d[i].real() - re(y) * x.d[i].real()) / den;
    return to_value(ctx, r);
}
JSValue js_fast_length(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) {
    Storage a = get(ctx, ARG(0)), b = get(ctx, ARG(1)), c = get(ctx, ARG(2));
    double l = std::sqrt(re(a) * re(a) + re(b) * re(b) + re(c) * re(c));
    Storage r; r.v = l;
    for (int i = 0; i < 4; i++) r.d[i] = (re(a) * a.d[i].real() + re(b) * b.d[i].real() + re(c) * c.d[i].real()) / l;
    return to_value(ctx, r);
}
JSValue js_smooth_fmod(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) {   // only b's real part is used (js_interop.cpp:530-536)
    Storage a = get(ctx, ARG(0)), b = get(ctx, ARG(1));
    Storage r; r.v = std::fmod(re(a), re(b));
    for (int i = 0; i < 4; i++) r.d[i] = a.d[i].real();
    return to_value(ctx, r);
}
JSValue js_select(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) {   // dual_if on the real condition (js_interop.cpp:538-547)
    Storage c = get(ctx, ARG(0)), t = get(ctx, ARG(1)), f = get(ctx, ARG(2));
    Storage r = (re(c) != 0.0) ? t : f;
    r.which = 0;
    return to_value(ctx, r);
}
JSValue js_pow(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) {   // js_interop.cpp:585-616
    Storage a = get(ctx, ARG(0)), b = get(ctx, ARG(1));
    if (b.which != 0) return throw_msg(ctx, "Pow cannot be used with a complex second argument");
    bool b_const = true;
    for (int i = 0; i < 4; i++) if (b.d[i] != cd(0)) b_const = false;
    if (a.which == 0) {
        double x = re(a), e = re(b), p = std::pow(x, e);
        Storage r; r.v = p;
        for (int i = 0; i < 4; i++) {
            double d = e * std::pow(x, e - 1.0) * a.d[i].real();
            if (!b_const) d += p * std::log(x) * b.d[i].real();
            r.d[i] = d;
        }
        return to_value(ctx, r);
    }
    if (!b_const) return throw_msg(ctx, "Pow's second argument must be constant if arg1 is complex");
    double e = re(b);
    if (e != (int)e) return throw_msg(ctx, "With a complex first argument, the exponent must be integral");
    cd p = std::pow(a.v, (int)e);
    Storage r = chain(a, p, e * std::pow(a.v, (int)e - 1));
    r.which = 1;
    return to_value(ctx, r);
}
JSValue js_get_i(JSContext* ctx, JSValueConst, int, JSValueConst*) { Storage s; s.which = 1; s.v = cd(0, 1); return to_value(ctx, s); }
JSValue js_pin(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) { return JS_DupValue(ctx, ARG(0)); }   // a CSE hint only
JSValue js_debug(JSContext*, JSValueConst, int, JSValueConst*) { return JS_UNDEFINED; }

// $cfg proxy (js_interop.cpp:753-846): identifier check, first-access ordering, `$default`
bool valid_name(const char* s) { for (; *s; s++) if (!(isalnum((unsigned char)*s) || *s == '_')) return false; return true; }
JSValue js_cfg_value(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) {
    const char* key = JS_ToCString(ctx, ARG(0));
    if (!key) return JS_EXCEPTION;
    std::string k = key;
    JS_FreeCString(ctx, key);
    if (!valid_name(k.c_str())) return throw_msg(ctx, "Value must be alphanumeric or _");
    Session* S = g_session;
    int i = S->find(k);
    if (i < 0) { S->names.push_back(k); S->defaults.push_back(0.0); S->values.push_back(0.0); S->has_value.push_back(0); i = (int)S->names.size() - 1; }
    return to_value(ctx, real_const(S->has_value[i] ? S->values[i] : S->defaults[i]));
}
JSValue js_cfg_default(JSContext* ctx, JSValueConst, int argc, JSValueConst* argv) {
    const char* key = JS_ToCString(ctx, ARG(0));
    if (!key) return JS_EXCEPTION;
    double d = 0;
    JS_ToFloat64(ctx, &d, ARG(1));
    int i = g_session->find(key);
    JS_FreeCString(ctx, key);
    if (i >= 0) g_session->defaults[i] = (double)(float)d;   // `float valf = (double)value` (js_interop.cpp:787)
    return JS_UNDEFINED;
}

const char* GLUE_BEFORE =   // installed before number.js: the globals extract_function creates (js_interop.cpp:665-743)
    "globalThis.CShim = {add: __b0, mul: __b1, sub: __b2, div: __b3, neg: __neg, construct: __construct};\n"
    "globalThis.CMath = {sin: __u0, cos: __u1, tan: __u2, asin: __u3, acos: __u4, atan: __u5, atan2: __atan2, fabs: __u6, log: __u7,\n"
    "  select: __select, pow: __pow, sqrt: __u8, psqrt: __u9, csqrt: __u18, exp: __u10, fast_length: __fast_length, length: __fast_length,\n"
    "  smooth_fmod: __smooth_fmod, sinh: __u11, cosh: __u12, tanh: __u13, M_PI: Math.PI, PI: Math.PI, conjugate: __u14,\n"
    "  self_conjugate_multiply: __u15, Real: __u16, Imaginary: __u17, debug: __debug, get_i: __get_i, lt: __b4, lte: __b5, eq: __b6, gt: __b7, gte: __b8};\n"
    "globalThis.M_PI = Math.PI;\n"
    "globalThis.$pin = __pin;\n";
const char* GLUE_AFTER =    // inject_config (js_interop.cpp:826-846): global getter $cfg -> Proxy {get, set}
    "Object.defineProperty(globalThis, '$cfg', {get: function() { return new Proxy(function(){}, {\n"
    "  get: function(target, prop, receiver) { var n = __cfg_value(String(prop));\n"
    "    Object.defineProperty(n, '$default', {get: function(){}, set: function(v) { __cfg_default(String(prop), Number(v)); }}); return n; },\n"
    "  set: function(target, prop, val, receiver) { return true; } }); }, set: function(){} });\n";

std::string read_file(const char* path) {
    std::string s;
    FILE* f = fopen(path, "rb");
    if (!f) return s;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

std::string exception_text(JSContext* ctx) {
    JSValue e = JS_GetException(ctx);
    const char* c = JS_ToCString(ctx, e);
    std::string s = c ? c : "unknown exception";
    if (c) JS_FreeCString(ctx, c);
    JS_FreeValue(ctx, e);
    return s;
}

}  // namespace

extern "C" {

// One script, one point.  kind: 0 metric (4 or 16 entries -> out[16], dout[64] = d_k out_ij at [k*16 + ij]),
// 1 coordinate function (4 entries -> out[4], dout[16] at [k*4 + i]), 2 origin distance (1 entry -> out[1], dout[4]).
// Parameters: names/values in (n_in of them; has_value all 1), discovered names/defaults out (first-access order).
// Returns 0, or -1 with the QuickJS error message in err.
int qjsref_run(const char* number_js_path, const char* script_path, int kind, const double* pos, int n_in, const char* const* names_in, const double* values_in,
               double* out, double* dout, int* out_len, int* n_names_out, char* names_out /* 32 x 64 */, double* defaults_out, char* err, int err_len) {
    Session S;
    for (int i = 0; i < n_in; i++) { S.names.push_back(names_in[i]); S.defaults.push_back(0.0); S.values.push_back(values_in[i]); S.has_value.push_back(1); }
    g_session = &S;
    JSRuntime* rt = JS_NewRuntime();
    JS_SetMemoryLimit(rt, 4 * 1024 * 1024);   // quickjs_cpp.cpp:261
    JSContext* ctx = JS_NewContext(rt);
    JS_AddIntrinsicBigFloat(ctx);
    JS_AddIntrinsicBigDecimal(ctx);
    JS_AddIntrinsicOperators(ctx);
    JS_EnableBignumExt(ctx, 1);
    if (box_class == 0) JS_NewClassID(&box_class);
    JSClassDef def = {"storage_box", box_finalizer, nullptr, nullptr, nullptr};
    JS_NewClass(rt, box_class, &def);

    JSValue global = JS_GetGlobalObject(ctx);
    char nm[32];
    for (int i = 0; i <= 8; i++) { snprintf(nm, sizeof nm, "__b%d", i); JS_SetPropertyStr(ctx, global, nm, JS_NewCFunctionMagic(ctx, js_binop, nm, 2, JS_CFUNC_generic_magic, i)); }
    for (int i = 0; i <= 18; i++) { snprintf(nm, sizeof nm, "__u%d", i); JS_SetPropertyStr(ctx, global, nm, JS_NewCFunctionMagic(ctx, js_unary, nm, 1, JS_CFUNC_generic_magic, i)); }
    struct { const char* n; JSCFunction* f; int argc; } fns[] = {
        {"__neg", js_neg, 1}, {"__construct", js_construct, 2}, {"__atan2", js_atan2, 2}, {"__fast_length", js_fast_length, 3}, {"__smooth_fmod", js_smooth_fmod, 2},
        {"__select", js_select, 3}, {"__pow", js_pow, 2}, {"__get_i", js_get_i, 0}, {"__pin", js_pin, 1}, {"__debug", js_debug, 1}, {"__cfg_value", js_cfg_value, 1},
        {"__cfg_default", js_cfg_default, 2}};
    for (auto& f : fns) JS_SetPropertyStr(ctx, global, f.n, JS_NewCFunction(ctx, f.f, f.n, f.argc));

    int rc = 0;
    std::string message;
    auto eval = [&](const std::string& text, const char* name, JSValue* result) {
        JSValue r = JS_Eval(ctx, text.c_str(), text.size(), name, JS_EVAL_TYPE_GLOBAL);
        if (JS_IsException(r)) { message = std::string(name) + ": " + exception_text(ctx); rc = -1; JS_FreeValue(ctx, r); return false; }
        if (result) *result = r; else JS_FreeValue(ctx, r);
        return true;
    };
    JSValue func = JS_UNDEFINED, result = JS_UNDEFINED;
    std::string number_js = read_file(number_js_path), script = read_file(script_path);
    if (number_js.empty() || script.empty()) { message = "cannot read number.js or the script"; rc = -1; }
    if (rc == 0 && eval(GLUE_BEFORE, "<glue>", nullptr) && eval(number_js, "number.js", nullptr) && eval(script, script_path, &func) && eval(GLUE_AFTER, "<glue2>", nullptr)) {
        if (!JS_IsFunction(ctx, func)) { message = "Expected function in eval of script"; rc = -1; }
        else {
            JSValue args[4];
            for (int i = 0; i < 4; i++) { Storage s; s.v = pos[i]; s.d[i] = 1.0; args[i] = to_value(ctx, s); }
            result = JS_Call(ctx, func, global, 4, args);
            for (int i = 0; i < 4; i++) JS_FreeValue(ctx, args[i]);
            if (JS_IsException(result)) { message = "Error in script exec " + exception_text(ctx); rc = -1; }
            else {
                std::vector<Storage> vals;
                if (kind == 2) vals.push_back(get(ctx, result));
                else if (!JS_IsArray(ctx, result)) { message = "Must return array"; rc = -1; }
                else {
                    JSValue lenv = JS_GetPropertyStr(ctx, result, "length");
                    int len = 0;
                    JS_ToInt32(ctx, &len, lenv);
                    JS_FreeValue(ctx, lenv);
                    for (int i = 0; i < len; i++) { JSValue e = JS_GetPropertyUint32(ctx, result, (uint32_t)i); vals.push_back(get(ctx, e)); JS_FreeValue(ctx, e); }
                }
                if (rc == 0) {
                    int n = (int)vals.size();
                    if (kind == 0 && n != 4 && n != 16) { message = "Must return array length of 4 or 16"; rc = -1; }
                    if (kind == 1 && n != 4) { message = "Must return array size of 4"; rc = -1; }
                    if (rc == 0) {
                        int total = kind == 0 ? 16 : (kind == 1 ? 4 : 1);
                        for (int i = 0; i < total; i++) out[i] = 0;
                        for (int i = 0; i < total * 4; i++) dout[i] = 0;
                        for (int i = 0; i < n; i++) {
                            int dst = (kind == 0 && n == 4) ? i * 5 : i;     // 4 -> diagonal of 16 (js_interop.cpp:879-886)
                            out[dst] = vals[i].v.real();                      // getr: the real part (js_interop.cpp:382-392)
                            for (int k = 0; k < 4; k++) dout[k * total + dst] = vals[i].d[k].real();
                        }
                        *out_len = n;
                    }
                }
            }
        }
    }
    JS_FreeValue(ctx, result);
    JS_FreeValue(ctx, func);
    JS_FreeValue(ctx, global);
    *n_names_out = (int)S.names.size();
    for (size_t i = 0; i < S.names.size() && i < 32; i++) { snprintf(names_out + i * 64, 64, "%s", S.names[i].c_str()); defaults_out[i] = S.defaults[i]; }
    if (rc != 0 && err) snprintf(err, (size_t)err_len, "%s", message.c_str());
    JS_FreeContext(ctx);
    JS_FreeRuntime(rt);
    g_session = nullptr;
    return rc;
}

}  // extern "C"
