// Metric plugin loader. See metric_loader.hpp for the reference files this mirrors.
#include "metric_loader.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <functional>
#include <set>
#include <sstream>
#include <stdexcept>

#include "mini_json.hpp"
#include "minijs.hpp"

namespace fs = std::filesystem;
using sym::Expr;
using sym::Graph;

namespace grt {

static std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot read file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// ------------------------------------------------------------------------------------------
// Config (metric.hpp:359-434)
// ------------------------------------------------------------------------------------------
void MetricConfig::load_json(const std::string& text) {
    mini_json::Value js = mini_json::parse(text);
    if (js.type != mini_json::Value::Object) throw std::runtime_error("metric config must be a JSON object");
    auto as_bool = [](const mini_json::Value& v) {
        if (v.type == mini_json::Value::Bool) return v.b;
        if (v.type == mini_json::Value::Number) return v.num != 0;
        throw std::runtime_error("expected boolean");
    };
    auto as_num = [](const mini_json::Value& v) {
        if (v.type == mini_json::Value::Number) return (float)v.num;
        if (v.type == mini_json::Value::Bool) return v.b ? 1.f : 0.f;
        throw std::runtime_error("expected number");
    };
    auto as_str = [](const mini_json::Value& v) {
        if (v.type != mini_json::Value::String) throw std::runtime_error("expected string");
        return v.str;
    };
    for (auto& [key, value] : js.obj) {
        if (key == "name") name = as_str(value);
        else if (key == "description") description = as_str(value);
        else if (key == "use_prepass") use_prepass = as_bool(value);
        else if (key == "max_acceleration_change") max_acceleration_change = as_num(value);
        else if (key == "singular") singular = as_bool(value);
        else if (key == "traversable_event_horizon") traversable_event_horizon = as_bool(value);
        else if (key == "singular_terminator") singular_terminator = as_num(value);
        else if (key == "adaptive_precision") adaptive_precision = as_bool(value);
        else if (key == "detect_singularities") detect_singularities = as_bool(value);
        else if (key == "follow_geodesics_forward") follow_geodesics_forward = as_bool(value);
        else if (key == "coordinate_system") {
            std::string s = as_str(value);
            if (s == "X_Y_THETA_PHI") system = CoordinateSystem::X_Y_THETA_PHI;
            else if (s == "CARTESIAN") system = CoordinateSystem::CARTESIAN;
            else if (s == "CYLINDRICAL") system = CoordinateSystem::CYLINDRICAL;
            else system = CoordinateSystem::OTHER;
        } else if (key == "to_polar") to_polar = as_str(value);
        else if (key == "from_polar") from_polar = as_str(value);
        else if (key == "origin_distance") origin_distance = as_str(value);
        else if (key == "coordinate_periodicity") coordinate_periodicity = as_str(value);
        else if (key == "inherit_settings") inherit_settings = as_str(value);
        else if (key == "has_cylindrical_singularity") has_cylindrical_singularity = as_bool(value);
        else if (key == "cylindrical_terminator") cylindrical_terminator = as_num(value);
        else if (key == "unconditionally_nonsingular") unconditionally_nonsingular = as_bool(value);
        else warnings.push_back("unknown key name " + key);
    }
}

void ConfigVariables::add(const std::string& name, float val) {
    for (auto& e : names) if (e == name) return;
    names.push_back(name);
    default_values.push_back(val);
    current_values.push_back(val);
}
void ConfigVariables::set_default(const std::string& name, float val) {
    for (size_t i = 0; i < names.size(); i++) {
        if (names[i] == name) { default_values[i] = val; current_values[i] = val; return; }
    }
    names.push_back(name);
    default_values.push_back(val);
    current_values.push_back(val);
}

// ------------------------------------------------------------------------------------------
// Content folders (content_manager.cpp:153-379)
// ------------------------------------------------------------------------------------------
static std::vector<std::string> files_with_extension(const fs::path& folder, const std::string& ext) {
    std::vector<std::string> ret;
    std::error_code ec;
    if (!fs::is_directory(folder, ec)) return ret;
    for (auto& entry : fs::directory_iterator(folder, ec)) {
        std::string s = entry.path().string();
        if (s.size() >= ext.size() && s.compare(s.size() - ext.size(), ext.size(), ext) == 0)
            ret.push_back(fs::absolute(entry.path()).string());
    }
    std::sort(ret.begin(), ret.end());  // directory order is unspecified; make it deterministic
    return ret;
}

MetricConfig load_config(const ContentManager& cm, const std::string& filename, bool inherit) {
    MetricConfig cfg;
    try {
        std::string text = read_file(filename);
        mini_json::Value js = mini_json::parse(text);
        const mini_json::Value* inh = js.find("inherit_settings");
        if (inherit && inh && inh->type == mini_json::Value::String) {
            std::string parent_file;
            if (cm.lookup_config_file(inh->str, parent_file)) {
                MetricConfig parent;
                parent.load_json(read_file(parent_file));
                cfg = parent;
            } else cfg.warnings.push_back("could not look up inherited config " + inh->str);
        }
        cfg.load_json(text);
    } catch (std::exception& e) {
        cfg.name = fs::path(filename).filename().string() + " (broken)";
        cfg.warnings.push_back(std::string("error loading config: ") + e.what());
    }
    return cfg;
}

void ContentManager::add_content_folder(const std::string& folder_in) {
    std::string folder = fs::absolute(folder_in).lexically_normal().string();
    for (auto& c : dirs) if (c.folder == folder) return;
    ContentFolder con;
    con.folder = folder;
    con.metrics = files_with_extension(folder, ".js");
    con.configs = files_with_extension(folder, ".json");
    con.coordinates = files_with_extension(fs::path(folder) / "coordinates", ".js");
    con.origins = files_with_extension(fs::path(folder) / "origins", ".js");
    for (size_t i = 0; i < con.configs.size(); i++) {
        if (fs::path(con.configs[i]).filename() == "sorting.json") {
            try {
                mini_json::Value js = mini_json::parse(read_file(con.configs[i]));
                for (auto& v : js.arr) if (v.type == mini_json::Value::String) con.sorting.push_back(v.str);
            } catch (...) {}
            con.configs.erase(con.configs.begin() + i);
            break;
        }
    }
    if (!con.sorting.empty()) {
        std::set<std::string> named(con.sorting.begin(), con.sorting.end());
        std::vector<std::string> sorted, rest;
        for (auto& nm : con.sorting) sorted.push_back((fs::path(folder) / nm).string());
        for (auto& m : con.metrics) if (!named.count(fs::path(m).filename().string())) rest.push_back(m);
        sorted.insert(sorted.end(), rest.begin(), rest.end());
        con.metrics = sorted;
    }
    dirs.push_back(con);
    ContentFolder& stored = dirs.back();
    for (auto& c : stored.configs) stored.base_configs.push_back(load_config(*this, c, false));
}

bool ContentManager::lookup_config_file(const std::string& name, std::string& out) const {
    for (auto& c : dirs)
        for (size_t i = 0; i < c.base_configs.size(); i++)
            if (c.base_configs[i].name == name) { out = c.configs[i]; return true; }
    return false;
}
bool ContentManager::lookup_metric_file(const std::string& name, std::string& out) const {
    std::string cfg;
    if (!lookup_config_file(name, cfg)) return false;
    out = fs::path(cfg).replace_extension(".js").string();
    return true;
}
static bool lookup_stem(const std::vector<std::string>& files, const std::string& name, std::string& out) {
    for (auto& s : files)
        if (fs::path(s).filename().string() == name + ".js") { out = s; return true; }
    return false;
}
bool ContentManager::lookup_coordinates_file(const std::string& name, std::string& out) const {
    for (auto& c : dirs) if (lookup_stem(c.coordinates, name, out)) return true;
    return false;
}
bool ContentManager::lookup_origins_file(const std::string& name, std::string& out) const {
    for (auto& c : dirs) if (lookup_stem(c.origins, name, out)) return true;
    return false;
}
std::vector<std::string> ContentManager::metric_names() const {
    std::vector<std::string> out;
    for (auto& c : dirs)
        for (auto& m : c.metrics) {
            std::string cfgfile = fs::path(m).replace_extension(".json").string();
            for (size_t i = 0; i < c.configs.size(); i++)
                if (c.configs[i] == cfgfile) out.push_back(c.base_configs[i].name);
        }
    return out;
}

// ------------------------------------------------------------------------------------------
// Complex symbolic values (stands in for deps/vec dual / dual_complex; js_interop.cpp:129-297)
// ------------------------------------------------------------------------------------------
struct CVal {
    Expr re, im;
    bool cx = false;
    CVal() {}
    CVal(Expr r) : re(r), im(0.0), cx(false) {}
    CVal(Expr r, Expr i) : re(r), im(i), cx(true) {}
};

static CVal c_add(const CVal& a, const CVal& b) { return (a.cx || b.cx) ? CVal(a.re + b.re, a.im + b.im) : CVal(a.re + b.re); }
static CVal c_sub(const CVal& a, const CVal& b) { return (a.cx || b.cx) ? CVal(a.re - b.re, a.im - b.im) : CVal(a.re - b.re); }
static CVal c_mul(const CVal& a, const CVal& b) {
    if (!a.cx && !b.cx) return CVal(a.re * b.re);
    return CVal(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
static CVal c_div(const CVal& a, const CVal& b) {
    if (!a.cx && !b.cx) return CVal(a.re / b.re);
    if (!b.cx || b.im.is_const(0.0)) return CVal(a.re / b.re, a.im / b.re);
    Expr d = b.re * b.re + b.im * b.im;
    return CVal((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
static CVal c_neg(const CVal& a) { return a.cx ? CVal(-a.re, -a.im) : CVal(-a.re); }
static CVal c_conj(const CVal& a) { return a.cx ? CVal(a.re, -a.im) : a; }
static Expr c_abs2(const CVal& a) { return a.cx ? a.re * a.re + a.im * a.im : a.re * a.re; }
static CVal c_sqrt(const CVal& a) {
    if (!a.cx) return CVal(sym::sqrt(a.re));
    // principal complex square root as one differentiable primitive pair: ds = dz / (2 s) stays finite
    // on the real axis, where the (|z| +- re)/2 formulas have 0/0 derivatives
    return CVal(sym::csqrt_re(a.re, a.im), sym::csqrt_im(a.re, a.im));
}
static CVal c_pow_int(const CVal& a, int n) {
    if (n == 0) return CVal(Expr(1.0), Expr(0.0));
    CVal result, base = a;
    bool have = false;
    int m = std::abs(n);
    while (m) {
        if (m & 1) { result = have ? c_mul(result, base) : base; have = true; }
        m >>= 1;
        if (m) base = c_mul(base, base);
    }
    if (n < 0) return c_div(CVal(Expr(1.0), Expr(0.0)), result);
    return result;
}

// ------------------------------------------------------------------------------------------
// Script binding (js_interop.cpp:299-846)
// ------------------------------------------------------------------------------------------
static const int TAG_NUMBER = 1, TAG_CFG = 2;

struct Sandbox {
    ConfigVariables* cfg = nullptr;
};

static void validate_ident(const std::string& s) {
    if (s.empty()) throw minijs::JsError("Value must be alphanumeric or _");
    for (char c : s)
        if (!(isalnum((unsigned char)c) || c == '_')) throw minijs::JsError("Value must be alphanumeric or _");
}

struct ScriptFunction {
    minijs::Interp in;
    minijs::Value fn;
    Sandbox* sand;
    std::string path;

    static minijs::Value wrap(const CVal& v) {
        auto payload = std::make_shared<CVal>(v);
        return minijs::Value(minijs::Interp::make_host(TAG_NUMBER, payload));
    }
    // js_interop.cpp:322-350 `get`
    static CVal unwrap(const minijs::Value& v, const char* what = "value") {
        switch (v.type) {
            case minijs::Value::Undefined: return CVal(Expr(0.0));
            case minijs::Value::Number: return CVal(Expr(v.num));
            case minijs::Value::Bool: return CVal(Expr(v.b ? 1.0 : 0.0));
            case minijs::Value::Obj:
                if (v.obj->kind == minijs::Object::Host && v.obj->host_tag == TAG_NUMBER) return *std::static_pointer_cast<CVal>(v.obj->host);
                if (v.obj->props.count("__number")) return unwrap(v.obj->props["__number"], what);
                break;
            default: break;
        }
        throw minijs::JsError(std::string("TypeError: ") + what + " is not a number");
    }
    static Expr real_arg(const minijs::Value& v, const char* fn_name) {
        CVal c = unwrap(v, fn_name);
        if (c.cx) throw minijs::JsError(std::string("Invalid type in dual/complex internals, ") + fn_name + " is a real-only function");
        return c.re;
    }

    ScriptFunction(Sandbox* s, const std::string& file) : sand(s), path(file) {
        install();
        std::string src = read_file(file);
        fn = in.eval(src, fs::path(file).filename().string());
    }

    void install() {
        using minijs::Interp;
        using minijs::Value;
        using Args = std::vector<Value>;
        auto arg = [](Args& a, size_t i) -> Value { return i < a.size() ? a[i] : Value(); };

        in.host_operator = [](Interp&, const std::string& op, const Value& a, const Value& b) -> Value {
            if (op == "neg") return wrap(c_neg(unwrap(a)));
            if (op == "pos") return a;
            CVal x = unwrap(a, "left operand"), y = unwrap(b, "right operand");
            if (op == "+") return wrap(c_add(x, y));
            if (op == "-") return wrap(c_sub(x, y));
            if (op == "*") return wrap(c_mul(x, y));
            if (op == "/") return wrap(c_div(x, y));
            throw minijs::JsError("TypeError: operator '" + op + "' is not overloaded for number objects (use CMath.lt/lte/eq/gt/gte/select)");
        };

        auto cmath = Interp::make_object();
        auto def = [&](const char* name, minijs::NativeFn f) { cmath->props[name] = Value(Interp::make_native(std::move(f))); };
        // functions defined for real and complex arguments (UNARY_JS)
        def("sin", [arg](Interp&, const Value&, Args& a) {
            CVal v = unwrap(arg(a, 0), "sin");
            if (!v.cx) return wrap(CVal(sym::sin(v.re)));
            return wrap(CVal(sym::sin(v.re) * sym::cosh(v.im), sym::cos(v.re) * sym::sinh(v.im)));
        });
        def("cos", [arg](Interp&, const Value&, Args& a) {
            CVal v = unwrap(arg(a, 0), "cos");
            if (!v.cx) return wrap(CVal(sym::cos(v.re)));
            return wrap(CVal(sym::cos(v.re) * sym::cosh(v.im), -(sym::sin(v.re) * sym::sinh(v.im))));
        });
        def("fabs", [arg](Interp&, const Value&, Args& a) {
            CVal v = unwrap(arg(a, 0), "fabs");
            return wrap(v.cx ? CVal(sym::sqrt(c_abs2(v))) : CVal(sym::fabs(v.re)));
        });
        def("sqrt", [arg](Interp&, const Value&, Args& a) { return wrap(c_sqrt(unwrap(arg(a, 0), "sqrt"))); });
        def("psqrt", [arg](Interp&, const Value&, Args& a) { return wrap(c_sqrt(unwrap(arg(a, 0), "psqrt"))); });
        def("conjugate", [arg](Interp&, const Value&, Args& a) { return wrap(c_conj(unwrap(arg(a, 0), "conjugate"))); });
        def("self_conjugate_multiply", [arg](Interp&, const Value&, Args& a) { return wrap(CVal(c_abs2(unwrap(arg(a, 0), "self_conjugate_multiply")))); });
        // real-only (UNARY_JS_REAL)
#define REAL1(nm, fnx) def(nm, [arg](Interp&, const Value&, Args& a) { return wrap(CVal(fnx(real_arg(arg(a, 0), nm)))); })
        REAL1("tan", sym::tan); REAL1("asin", sym::asin); REAL1("acos", sym::acos); REAL1("atan", sym::atan);
        REAL1("log", sym::log); REAL1("exp", sym::exp); REAL1("sinh", sym::sinh); REAL1("cosh", sym::cosh); REAL1("tanh", sym::tanh);
#undef REAL1
        def("atan2", [arg](Interp&, const Value&, Args& a) { return wrap(CVal(sym::atan2(real_arg(arg(a, 0), "atan2"), real_arg(arg(a, 1), "atan2")))); });
        auto fast_length = [arg](Interp&, const Value&, Args& a) {
            Expr x = real_arg(arg(a, 0), "fast_length"), y = real_arg(arg(a, 1), "fast_length"), z = real_arg(arg(a, 2), "fast_length");
            return wrap(CVal(sym::sqrt(x * x + y * y + z * z)));
        };
        def("fast_length", fast_length);
        def("length", fast_length);
        def("smooth_fmod", [arg](Interp&, const Value&, Args& a) { return wrap(CVal(sym::fmod(real_arg(arg(a, 0), "smooth_fmod"), real_arg(arg(a, 1), "smooth_fmod")))); });
        def("select", [arg](Interp&, const Value&, Args& a) {
            return wrap(CVal(sym::select(real_arg(arg(a, 0), "select"), real_arg(arg(a, 1), "select"), real_arg(arg(a, 2), "select"))));
        });
        def("Real", [arg](Interp&, const Value&, Args& a) { return wrap(CVal(unwrap(arg(a, 0), "Real").re)); });
        def("Imaginary", [arg](Interp&, const Value&, Args& a) { CVal v = unwrap(arg(a, 0), "Imaginary"); return wrap(CVal(v.cx ? v.im : Expr(0.0))); });
        def("csqrt", [arg](Interp&, const Value&, Args& a) {
            CVal v = unwrap(arg(a, 0), "csqrt");
            if (v.cx) throw minijs::JsError("csqrt must be used with purely real arguments");
            return wrap(c_sqrt(CVal(v.re, Expr(0.0))));
        });
        def("pow", [arg](Interp&, const Value&, Args& a) {
            CVal s1 = unwrap(arg(a, 0), "pow"), s2 = unwrap(arg(a, 1), "pow");
            if (s2.cx) throw minijs::JsError("Pow cannot be used with a complex second argument");
            if (!s1.cx) return wrap(CVal(sym::pow(s1.re, s2.re)));
            if (!s2.re.is_const()) throw minijs::JsError("Pow's second argument must be constant if arg1 is complex");
            double e = s2.re.cval();
            if (e != (int)e) throw minijs::JsError("With a complex first argument, the exponent must be integral");
            return wrap(c_pow_int(s1, (int)e));
        });
        def("debug", [](Interp&, const Value&, Args&) { return Value(); });
        def("get_i", [](Interp&, const Value&, Args&) { return wrap(CVal(Expr(0.0), Expr(1.0))); });
        // comparisons: real only; the reference's gte is implemented with '>' (js_interop.cpp:291-297)
#define CMP(nm, fnx) def(nm, [arg](Interp&, const Value&, Args& a) { \
            CVal x = unwrap(arg(a, 0), nm), y = unwrap(arg(a, 1), nm); \
            if (x.cx || y.cx) throw minijs::JsError(std::string("Can only use ") + nm + " on real values"); \
            return wrap(CVal(fnx(x.re, y.re))); })
        CMP("lt", sym::lt); CMP("lte", sym::lte); CMP("eq", sym::eq); CMP("gt", sym::gt); CMP("gte", sym::gt);
#undef CMP
        cmath->props["M_PI"] = Value(3.14159265358979323846);
        cmath->props["PI"] = Value(3.14159265358979323846);
        // number.js:4-10,73-80: `class number`, CMath.number, CMath.i, make_class
        auto number_ctor = Interp::make_native([arg](Interp&, const Value&, Args& a) { return wrap(unwrap(arg(a, 0), "number")); });
        cmath->props["number"] = Value(number_ctor);
        cmath->props["i"] = wrap(CVal(Expr(0.0), Expr(1.0)));
        in.set_global("CMath", Value(cmath));
        in.set_global("number", Value(number_ctor));
        in.set_global("make_class", Value(number_ctor));
        in.set_global("M_PI", Value(3.14159265358979323846));

        auto cshim = Interp::make_object();
        auto bin = [&](const char* name, const char* op) {
            std::string ops = op;
            cshim->props[name] = Value(Interp::make_native([arg, ops](Interp& ip, const Value&, Args& a) { return ip.host_operator(ip, ops, wrap(unwrap(arg(a, 0))), arg(a, 1)); }));
        };
        bin("add", "+"); bin("mul", "*"); bin("sub", "-"); bin("div", "/");
        cshim->props["neg"] = Value(Interp::make_native([arg](Interp&, const Value&, Args& a) { return wrap(c_neg(unwrap(arg(a, 0)))); }));
        cshim->props["construct"] = Value(Interp::make_native([arg](Interp&, const Value&, Args& a) {
            Value self = arg(a, 0);
            if (self.type == Value::Obj) self.obj->props["__number"] = wrap(unwrap(arg(a, 1)));
            return Value();
        }));
        in.set_global("CShim", Value(cshim));

        // $pin: CSE hint. The DAG is hash-consed, so sharing is automatic; returns its argument
        // (js_interop.cpp:352-380; no-op on non-number values such as arrays, :354-355).
        in.set_global("$pin", Value(Interp::make_native([arg](Interp&, const Value&, Args& a) { return arg(a, 0); })));

        // $cfg: proxy whose property reads declare a runtime parameter (js_interop.cpp:795-846)
        Sandbox* sb = sand;
        in.set_global_getter("$cfg", [sb](Interp&) -> Value {
            auto proxy = Interp::make_host(TAG_CFG, nullptr);
            proxy->get_hook = [sb](Interp&, const std::string& key, Value& out) -> bool {
                validate_ident(key);
                sb->cfg->add(key, 0.f);
                auto payload = std::make_shared<CVal>(CVal(Expr::param(key)));
                auto num = Interp::make_host(TAG_NUMBER, payload);
                num->set_hook = [sb, key](Interp&, const std::string& prop, const Value& v) -> bool {
                    if (prop != "$default") return false;
                    sb->cfg->set_default(key, (float)Interp::to_number(v));
                    return true;
                };
                out = Value(num);
                return true;
            };
            proxy->set_hook = [](Interp&, const std::string&, const Value&) -> bool { return true; };  // "Warning, setting a config from js"
            return Value(proxy);
        });
    }

    std::vector<CVal> call4(const Expr v[4]) {
        if (fn.type != minijs::Value::Obj || (fn.obj->kind != minijs::Object::Function && fn.obj->kind != minijs::Object::Native))
            throw std::runtime_error("Expected function in eval of script " + path);
        std::vector<minijs::Value> args;
        for (int i = 0; i < 4; i++) args.push_back(wrap(CVal(v[i])));
        minijs::Value r = in.call(fn, minijs::Value(), args);
        std::vector<CVal> out;
        if (r.type == minijs::Value::Obj && r.obj->kind == minijs::Object::Array) {
            for (auto& e : r.obj->arr) out.push_back(unwrap(e, "returned array element"));
        } else out.push_back(unwrap(r, "returned value"));
        return out;
    }
    bool returned_array = false;
};

// ------------------------------------------------------------------------------------------
// Descriptor (metric.hpp:82-274, 557-708)
// ------------------------------------------------------------------------------------------
static long tree_occurrences(const Graph& g, int id, int leaf, std::map<int, long>& memo) {
    if (id < 0) return 0;
    if (id == leaf) return 1;
    auto it = memo.find(id);
    if (it != memo.end()) return it->second;
    const sym::Node& n = g.node(id);
    long c = tree_occurrences(g, n.a, leaf, memo) + tree_occurrences(g, n.b, leaf, memo) + tree_occurrences(g, n.c, leaf, memo);
    memo[id] = c;
    return c;
}

// symmetric 4x4 inverse by cofactors, the arrangement of cl.cl:469-558 (upper triangle only)
static void symbolic_metric_inverse(const Expr m[16], Expr inv_out[16]) {
    Expr inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[11] - m[6] * m[6] * m[15] + m[6] * m[7] * m[11] + m[7] * m[6] * m[11] - m[7] * m[7] * m[10];
    inv[1] = -(m[1] * m[10] * m[15]) + m[1] * m[11] * m[11] + m[6] * m[2] * m[15] - m[6] * m[3] * m[11] - m[7] * m[2] * m[11] + m[7] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[11] - m[2] * m[2] * m[15] + m[2] * m[3] * m[11] + m[3] * m[2] * m[11] - m[3] * m[3] * m[10];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[11] - m[5] * m[2] * m[15] + m[5] * m[3] * m[11] + m[7] * m[2] * m[7] - m[7] * m[3] * m[6];
    inv[6] = -(m[0] * m[6] * m[15]) + m[0] * m[7] * m[11] + m[1] * m[2] * m[15] - m[1] * m[3] * m[11] - m[3] * m[2] * m[7] + m[3] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[7] - m[1] * m[1] * m[15] + m[1] * m[3] * m[7] + m[3] * m[1] * m[7] - m[3] * m[3] * m[5];
    inv[3] = -(m[1] * m[6] * m[11]) + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[6] * m[2] * m[7] + m[6] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[1] * m[2] * m[11] + m[1] * m[3] * m[10] + m[2] * m[2] * m[7] - m[2] * m[3] * m[6];
    inv[11] = -(m[0] * m[5] * m[11]) + m[0] * m[7] * m[6] + m[1] * m[1] * m[11] - m[1] * m[3] * m[6] - m[2] * m[1] * m[7] + m[2] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[6] - m[1] * m[1] * m[10] + m[1] * m[2] * m[6] + m[2] * m[1] * m[6] - m[2] * m[2] * m[5];
    inv[4] = inv[1]; inv[8] = inv[2]; inv[12] = inv[3]; inv[9] = inv[6]; inv[13] = inv[7]; inv[14] = inv[11];
    Expr det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    Expr idet = Expr(1.0) / det;
    for (int i = 0; i < 16; i++) inv_out[i] = inv[i] * idet;
}

// a = -g^{-1} w for a symmetric 4x4 g through 2x2 minors (Laplace expansion); never forms the
// full inverse.  g is indexed row-major, only the upper triangle is read.
static void symbolic_solve_sym4(const Expr g[16], const Expr w[4], Expr out[4]) {
    auto G = [&](int i, int j) { return i <= j ? g[i * 4 + j] : g[j * 4 + i]; };
    // minors of rows 0,1 (s) and rows 2,3 (c)
    Expr s0 = G(0, 0) * G(1, 1) - G(1, 0) * G(0, 1);
    Expr s1 = G(0, 0) * G(1, 2) - G(1, 0) * G(0, 2);
    Expr s2 = G(0, 0) * G(1, 3) - G(1, 0) * G(0, 3);
    Expr s3 = G(0, 1) * G(1, 2) - G(1, 1) * G(0, 2);
    Expr s4 = G(0, 1) * G(1, 3) - G(1, 1) * G(0, 3);
    Expr s5 = G(0, 2) * G(1, 3) - G(1, 2) * G(0, 3);
    Expr c5 = G(2, 2) * G(3, 3) - G(3, 2) * G(2, 3);
    Expr c4 = G(2, 1) * G(3, 3) - G(3, 1) * G(2, 3);
    Expr c3 = G(2, 1) * G(3, 2) - G(3, 1) * G(2, 2);
    Expr c2 = G(2, 0) * G(3, 3) - G(3, 0) * G(2, 3);
    Expr c1 = G(2, 0) * G(3, 2) - G(3, 0) * G(2, 2);
    Expr c0 = G(2, 0) * G(3, 1) - G(3, 0) * G(2, 1);
    Expr det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    Expr adj[16];
    adj[0] = G(1, 1) * c5 - G(1, 2) * c4 + G(1, 3) * c3;
    adj[1] = -(G(0, 1) * c5) + G(0, 2) * c4 - G(0, 3) * c3;
    adj[2] = G(3, 1) * s5 - G(3, 2) * s4 + G(3, 3) * s3;
    adj[3] = -(G(2, 1) * s5) + G(2, 2) * s4 - G(2, 3) * s3;
    adj[5] = G(0, 0) * c5 - G(0, 2) * c2 + G(0, 3) * c1;
    adj[6] = -(G(3, 0) * s5) + G(3, 2) * s2 - G(3, 3) * s1;
    adj[7] = G(2, 0) * s5 - G(2, 2) * s2 + G(2, 3) * s1;
    adj[10] = G(3, 0) * s4 - G(3, 1) * s2 + G(3, 3) * s0;
    adj[11] = -(G(2, 0) * s4) + G(2, 1) * s2 - G(2, 3) * s0;
    adj[15] = G(2, 0) * s3 - G(2, 1) * s1 + G(2, 2) * s0;
    adj[4] = adj[1]; adj[8] = adj[2]; adj[12] = adj[3]; adj[9] = adj[6]; adj[13] = adj[7]; adj[14] = adj[11];
    Expr nidet = Expr(-1.0) / det;
    for (int i = 0; i < 4; i++) {
        Expr sum = adj[i * 4 + 0] * w[0] + adj[i * 4 + 1] * w[1] + adj[i * 4 + 2] * w[2] + adj[i * 4 + 3] * w[3];
        out[i] = sum * nidet;
    }
}

// The same solve with the cofactors and the determinant multiplied out as polynomials over the atoms of the metric
// entries (Graph::to_poly), so that everything that cancels in them cancels symbolically: for a Kerr-Schild form
// eta + H l l the ten cofactors collapse to a few monomials in (H, l_i) each and the determinant to -(1 + H(l.l - 1)).
// Returns false when the expansion would be too large; the caller keeps whichever arrangement has fewer operations.
static bool symbolic_solve_sym4_poly(Graph& gr, const Expr g[16], const Expr w[4], Expr out[4]) {
    typedef Graph::Poly Poly;
    const size_t cap = 6000;
    Poly P[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = i; j < 4; j++) {
            if (!gr.to_poly(g[i * 4 + j].id, P[i][j], cap)) return false;
            P[j][i] = P[i][j];
        }
    bool ok = true;
    auto mul = [&](const Poly& a, const Poly& b) { Poly r; if (!Graph::poly_mul(a, b, r, cap)) ok = false; return r; };
    auto minor2 = [&](int r0, int r1, int c0, int c1) {   // P[r0][c0] P[r1][c1] - P[r1][c0] P[r0][c1]
        Poly r = mul(P[r0][c0], P[r1][c1]);
        Graph::poly_axpy(r, -1.0, mul(P[r1][c0], P[r0][c1]));
        return r;
    };
    Poly s0 = minor2(0, 1, 0, 1), s1 = minor2(0, 1, 0, 2), s2 = minor2(0, 1, 0, 3), s3 = minor2(0, 1, 1, 2), s4 = minor2(0, 1, 1, 3), s5 = minor2(0, 1, 2, 3);
    Poly c5 = minor2(2, 3, 2, 3), c4 = minor2(2, 3, 1, 3), c3 = minor2(2, 3, 1, 2), c2 = minor2(2, 3, 0, 3), c1 = minor2(2, 3, 0, 2), c0 = minor2(2, 3, 0, 1);
    if (!ok) return false;
    auto comb3 = [&](double sa, const Poly& a, const Poly& ma, double sb, const Poly& b, const Poly& mb, double sc, const Poly& c, const Poly& mc) {
        Poly r;
        Graph::poly_axpy(r, sa, mul(a, ma));
        Graph::poly_axpy(r, sb, mul(b, mb));
        Graph::poly_axpy(r, sc, mul(c, mc));
        return r;
    };
    Poly det;
    Graph::poly_axpy(det, 1.0, mul(s0, c5)); Graph::poly_axpy(det, -1.0, mul(s1, c4)); Graph::poly_axpy(det, 1.0, mul(s2, c3));
    Graph::poly_axpy(det, 1.0, mul(s3, c2)); Graph::poly_axpy(det, -1.0, mul(s4, c1)); Graph::poly_axpy(det, 1.0, mul(s5, c0));
    Poly adj[16];
    adj[0] = comb3(1, P[1][1], c5, -1, P[1][2], c4, 1, P[1][3], c3);
    adj[1] = comb3(-1, P[0][1], c5, 1, P[0][2], c4, -1, P[0][3], c3);
    adj[2] = comb3(1, P[3][1], s5, -1, P[3][2], s4, 1, P[3][3], s3);
    adj[3] = comb3(-1, P[2][1], s5, 1, P[2][2], s4, -1, P[2][3], s3);
    adj[5] = comb3(1, P[0][0], c5, -1, P[0][2], c2, 1, P[0][3], c1);
    adj[6] = comb3(-1, P[3][0], s5, 1, P[3][2], s2, -1, P[3][3], s1);
    adj[7] = comb3(1, P[2][0], s5, -1, P[2][2], s2, 1, P[2][3], s1);
    adj[10] = comb3(1, P[3][0], s4, -1, P[3][1], s2, 1, P[3][3], s0);
    adj[11] = comb3(-1, P[2][0], s4, 1, P[2][1], s2, -1, P[2][3], s0);
    adj[15] = comb3(1, P[2][0], s3, -1, P[2][1], s1, 1, P[2][2], s0);
    if (!ok) return false;
    Expr A[16];
    const int upper[10] = {0, 1, 2, 3, 5, 6, 7, 10, 11, 15};
    for (int k : upper) A[k] = Expr::from_id(gr.from_poly(adj[k]));
    A[4] = A[1]; A[8] = A[2]; A[12] = A[3]; A[9] = A[6]; A[13] = A[7]; A[14] = A[11];
    Expr nidet = Expr(-1.0) / Expr::from_id(gr.from_poly(det));
    for (int i = 0; i < 4; i++) {
        Expr sum = A[i * 4 + 0] * w[0] + A[i * 4 + 1] * w[1] + A[i * 4 + 2] * w[2] + A[i * 4 + 3] * w[3];
        out[i] = sum * nidet;
    }
    return true;
}

static void build_descriptor(Metric& m, ScriptFunction& fmetric, ScriptFunction& fto, ScriptFunction& ffrom,
                             ScriptFunction& forigin, ScriptFunction* fperiod) {
    MetricDescriptor& d = m.desc;
    Graph& G = d.graph;
    Expr v[4] = {Expr::var("v1"), Expr::var("v2"), Expr::var("v3"), Expr::var("v4")};
    Expr iv[4] = {Expr::var("iv1"), Expr::var("iv2"), Expr::var("iv3"), Expr::var("iv4")};
    Expr dv[4] = {Expr::var("dv1"), Expr::var("dv2"), Expr::var("dv3"), Expr::var("dv4")};

    // js_metric::operator() (js_interop.cpp:855-901): 4 -> diagonal, 16 -> row-major, real parts
    std::vector<CVal> met = fmetric.call4(v);
    Expr g[16];
    if (met.size() == 4) {
        for (int i = 0; i < 4; i++) g[i * 4 + i] = met[i].re;
    } else if (met.size() == 16) {
        for (int i = 0; i < 16; i++) g[i] = met[i].re;
    } else throw std::runtime_error("Must return array length of 4 or 16");

    // debiggen (metric.hpp:665-708): diagonal if every off-diagonal is literally zero
    d.big = false;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (i != j && !g[i * 4 + j].is_const(0.0)) d.big = true;
    // only the upper triangle is read by the kernels (cl.cl:1035-1050); mirror it
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < i; j++) g[i * 4 + j] = g[j * 4 + i];

    Expr dg[64];
    for (int k = 0; k < 4; k++)
        for (int ij = 0; ij < 16; ij++) dg[k * 16 + ij] = g[ij].diff(v[k]);
    for (int i = 0; i < 16; i++) d.g[i] = g[i].id;
    for (int i = 0; i < 64; i++) d.dg[i] = dg[i].id;

    // coordinate maps + total differentials (metric.hpp:247-274)
    auto coord_map = [&](ScriptFunction& f, int out[4], int dout[4], const char* what) {
        std::vector<CVal> r = f.call4(v);
        if (r.size() != 4) throw std::runtime_error(std::string("Must return array size of 4 (") + what + ")");
        for (int i = 0; i < 4; i++) {
            out[i] = r[i].re.id;
            Expr accum(0.0);
            for (int j = 0; j < 4; j++) accum += r[i].re.diff(v[j]) * dv[j];
            dout[i] = accum.id;
        }
    };
    coord_map(fto, d.to_polar, d.dto_polar, "to_polar");
    coord_map(ffrom, d.from_polar, d.dfrom_polar, "from_polar");
    {
        std::vector<CVal> r = forigin.call4(v);
        if (r.size() != 1) throw std::runtime_error("origin_distance must return a single value");
        d.distance = r[0].re.id;
    }
    d.has_period = false;
    for (int i = 0; i < 4; i++) d.period[i] = Expr(0.0).id;
    if (fperiod) {
        std::vector<CVal> r = fperiod->call4(v);
        if (r.size() != 4) throw std::runtime_error("coordinate_periodicity must return array size of 4");
        for (int i = 0; i < 4; i++) d.period[i] = r[i].re.id;
        d.has_period = true;
    }

    // is_polar_spherically_symmetric (metric.hpp:557-622)
    {
        bool ok = true;
        if (d.big) {
            const int bad[4] = {0 * 4 + 2, 0 * 4 + 3, 1 * 4 + 2, 1 * 4 + 3};
            for (int idx : bad) if (!g[idx].is_const(0.0)) ok = false;
        }
        if (ok) {
            Expr theta = g[2 * 4 + 2], phi = g[3 * 4 + 3];
            Expr sin_theta = sym::sin(v[2]);
            std::map<int, long> memo;
            long count = tree_occurrences(G, phi.id, sin_theta.id, memo);
            if (count != 2) ok = false;
            else {
                std::map<int, int> repl{{sin_theta.id, Expr(1.0).id}};
                ok = G.substitute(phi.id, repl) == theta.id;
            }
        }
        d.is_spherical = ok;
    }

    // reference formulation (metric.hpp:184-244): Gamma^i_kl = 1/2 g^im (d_l g_mk + d_k g_ml - d_m g_kl)
    {
        Expr inv[16];
        symbolic_metric_inverse(g, inv);
        Expr accel[4];
        auto P = [&](int wrt, int i, int j) { return dg[wrt * 16 + i * 4 + j]; };
        Expr christ[64];
        for (int i = 0; i < 4; i++)
            for (int k = 0; k < 4; k++)
                for (int l = 0; l < 4; l++) {
                    Expr sum(0.0);
                    for (int mm = 0; mm < 4; mm++) {
                        sum += inv[i * 4 + mm] * P(l, mm, k);
                        sum += inv[i * 4 + mm] * P(k, mm, l);
                        sum += -(inv[i * 4 + mm] * P(mm, k, l));
                    }
                    christ[i * 16 + k * 4 + l] = Expr(0.5) * sum;
                }
        for (int uu = 0; uu < 4; uu++) {
            Expr sum(0.0);
            for (int aa = 0; aa < 4; aa++)
                for (int bb = 0; bb < 4; bb++) sum += iv[aa] * iv[bb] * christ[uu * 16 + aa * 4 + bb];
            accel[uu] = -sum;
            d.accel_ref[uu] = accel[uu].id;
        }
    }
    // lowered formulation: w_m = sum_kl (d_k g_ml - 1/2 d_m g_kl) v^k v^l ; a = -g^{-1} w
    {
        Expr w[4];
        Expr u[4][4];  // u[k][m] = sum_l d_k g_ml v^l
        for (int k = 0; k < 4; k++)
            for (int mm = 0; mm < 4; mm++) {
                Expr s(0.0);
                for (int l = 0; l < 4; l++) s += dg[k * 16 + mm * 4 + l] * iv[l];
                u[k][mm] = s;
            }
        for (int mm = 0; mm < 4; mm++) {
            Expr first(0.0), second(0.0);
            for (int k = 0; k < 4; k++) first += iv[k] * u[k][mm];
            for (int k = 0; k < 4; k++) second += iv[k] * u[mm][k];
            w[mm] = first - Expr(0.5) * second;
        }
        Expr a[4];
        symbolic_solve_sym4(g, w, a);
        for (int i = 0; i < 4; i++) d.accel[i] = a[i].id;
        // Second arrangement of the same w_m: with the covector P_m = g_ml v^l and the scalar
        // S = g_kl v^k v^l (v held constant), w_m = (v . grad) P_m - 1/2 d_m S.  P and S are cleaned up
        // (common factors pulled out) *before* they are differentiated, so metrics with product
        // structure - Kerr-Schild forms g = eta + H l l, conformally flat ones - never expand the
        // 40 separate partials.  The cheaper of the two arrangements is emitted.
        // every arrangement goes through the value-based merge (Graph::unify) before it is costed: mathematically equal
        // sub-expressions that differ in form - the same l.v summed in three orders, a cofactor equal to a minor, |l|^2 = 1 -
        // become one node or a constant.  GRT_UNIFY=0 switches the pass off (tools/compile_catalogue.py compares both).
        std::map<std::string, double> param_env;
        for (size_t i = 0; i < m.params.names.size(); i++) param_env[m.params.names[i]] = m.params.current_values[i];
        const bool use_unify = !(getenv("GRT_UNIFY") && atoi(getenv("GRT_UNIFY")) == 0);
        auto merged = [&](const int cand[4], int out[4]) {
            std::vector<int> u(cand, cand + 4);
            if (use_unify) u = G.unify(u, param_env);
            for (int i = 0; i < 4; i++) out[i] = u[i];
        };
        size_t best_n;
        {
            int first[4];
            merged(d.accel, first);
            for (int i = 0; i < 4; i++) d.accel[i] = first[i];
            best_n = G.instr_count(std::vector<int>(d.accel, d.accel + 4));
        }
        auto consider = [&](const int raw[4]) {
            int cand[4];
            merged(raw, cand);
            size_t n = G.instr_count(std::vector<int>(cand, cand + 4));
            if (getenv("GRT_DEBUG_EMIT"))
                fprintf(stderr, "arrangement: %zu ops, %zu instructions (before the value-based merge %zu)\n", G.op_count(std::vector<int>(cand, cand + 4)), n,
                        G.instr_count(std::vector<int>(raw, raw + 4)));
            if (n < best_n) { best_n = n; for (int i = 0; i < 4; i++) d.accel[i] = cand[i]; }
        };
        const int base[4] = {d.accel[0], d.accel[1], d.accel[2], d.accel[3]};
        for (int dist = 0; dist < 2; dist++) {
            int cur[4];
            for (int i = 0; i < 4; i++) cur[i] = G.factor(base[i], dist != 0);
            consider(cur);
            Expr P[4], S(0.0);
            for (int mm = 0; mm < 4; mm++) {
                Expr s(0.0);
                for (int l = 0; l < 4; l++) s += g[mm * 4 + l] * iv[l];
                P[mm] = Expr::from_id(G.factor(s.id, dist != 0));
            }
            for (int mm = 0; mm < 4; mm++) S += P[mm] * iv[mm];
            S = Expr::from_id(G.factor(S.id, dist != 0));
            Expr w2[4], a2[4];
            for (int mm = 0; mm < 4; mm++) {
                Expr first(0.0);
                for (int k = 0; k < 4; k++) first += P[mm].diff(v[k]) * iv[k];
                w2[mm] = Expr::from_id(G.factor((first - Expr(0.5) * S.diff(v[mm])).id, dist != 0));
            }
            symbolic_solve_sym4(g, w2, a2);
            int alt[4], alt_raw[4];
            for (int i = 0; i < 4; i++) { alt_raw[i] = a2[i].id; alt[i] = G.factor(a2[i].id, dist != 0); }
            consider(alt_raw);
            consider(alt);
            // Third arrangement: the same P and S, with (v . grad) P_m taken as ONE forward-mode pass along v
            // (Graph::diff_along) and grad S as ONE reverse-mode sweep (Graph::gradient): the tangents and
            // adjoints of shared sub-expressions (r, H, l_i of a Kerr-Schild form) are computed once.
            Expr w3[4], a4[4];
            {
                std::map<int, int> along;
                for (int k = 0; k < 4; k++) along[v[k].id] = iv[k].id;
                std::vector<int> gS = G.gradient(S.id, {v[0].id, v[1].id, v[2].id, v[3].id});
                for (int mm = 0; mm < 4; mm++) w3[mm] = Expr::from_id(G.diff_along(P[mm].id, along)) - Expr(0.5) * Expr::from_id(gS[mm]);
            }
            symbolic_solve_sym4(g, w3, a4);
            { int c_raw[4], c_fac[4]; for (int i = 0; i < 4; i++) { c_raw[i] = a4[i].id; c_fac[i] = G.factor(a4[i].id, dist != 0); } consider(c_raw); consider(c_fac); }
            // every arrangement of w with the multiplied-out cofactors
            for (const Expr* ww : {(const Expr*)w, (const Expr*)w2, (const Expr*)w3}) {
                Expr a3[4];
                if (!symbolic_solve_sym4_poly(G, g, ww, a3)) continue;
                int p_raw[4], p_fac[4];
                for (int i = 0; i < 4; i++) { p_raw[i] = a3[i].id; p_fac[i] = G.factor(a3[i].id, dist != 0); }
                consider(p_raw);
                consider(p_fac);
            }
        }
    }
    // constant-theta specialisation (cl.cl:3294-3306): v3 = (float)(pi/2), iv3 = 0, a.z = 0
    if (m.constant_theta()) {
        double half_pi_f = (double)(float)(3.14159265358979323846 / 2);
        std::map<int, int> repl{{v[2].id, Expr(half_pi_f).id}, {iv[2].id, Expr(0.0).id}};
        for (int i = 0; i < 4; i++) {
            d.accel[i] = (i == 2) ? Expr(0.0).id : G.substitute(d.accel[i], repl);
            d.accel_ref[i] = (i == 2) ? Expr(0.0).id : G.substitute(d.accel_ref[i], repl);
        }
    }
    // radii: polar r and distance-to-object straight from native coordinates (cl.cl:4062-4068)
    {
        std::map<int, int> repl;
        for (int i = 0; i < 4; i++) repl[v[i].id] = d.to_polar[i];
        if (m.constant_theta()) repl[v[2].id] = Expr((double)(float)(3.14159265358979323846 / 2)).id;  // polar_position.z = pi/2
        d.radii[0] = d.to_polar[1];
        d.radii[1] = G.substitute(d.distance, repl);
        // the integration loop evaluates the radii of a point together with its acceleration (grt_accel_radii): merge the two
        // DAGs by value so that r, rho^2 ... are computed once
        if (!(getenv("GRT_UNIFY") && atoi(getenv("GRT_UNIFY")) == 0)) {
            std::map<std::string, double> param_env;
            for (size_t i = 0; i < m.params.names.size(); i++) param_env[m.params.names[i]] = m.params.current_values[i];
            std::vector<int> all = {d.accel[0], d.accel[1], d.accel[2], d.accel[3], d.radii[0], d.radii[1]};
            all = G.unify(all, param_env);
            for (int i = 0; i < 4; i++) d.accel[i] = all[i];
            d.radii[0] = all[4];
            d.radii[1] = all[5];
        }
    }
    // derived parameters (see MetricDescriptor::derived)
    {
        std::vector<int> all(d.g, d.g + 16);
        all.insert(all.end(), d.dg, d.dg + 64);
        for (const int* arr : {d.to_polar, d.dto_polar, d.from_polar, d.dfrom_polar, d.period, d.accel, d.accel_ref}) all.insert(all.end(), arr, arr + 4);
        all.push_back(d.distance);
        all.push_back(d.radii[0]);
        all.push_back(d.radii[1]);
        std::map<int, int> kind;   // 0: constants only, 1: parameters (no variables), 2: depends on a variable
        std::function<int(int)> classify = [&](int x) -> int {
            if (x < 0) return 0;
            auto it = kind.find(x);
            if (it != kind.end()) return it->second;
            const sym::Node n = G.node(x);
            int k = n.op == sym::Op::Var ? 2 : (n.op == sym::Op::Param ? 1 : std::max(classify(n.a), std::max(classify(n.b), classify(n.c))));
            kind[x] = k;
            return k;
        };
        for (int o : all) classify(o);
        auto is_op = [&](int x) { const sym::Node& n = G.node(x); return n.op != sym::Op::Const && n.op != sym::Op::Var && n.op != sym::Op::Param; };
        std::set<int> maximal;
        size_t param_only_ops = 0;
        for (auto& kv : kind) {
            const int x = kv.first;
            if (kv.second == 1 && is_op(x)) param_only_ops++;
            if (kv.second != 2) continue;
            const sym::Node n = G.node(x);
            for (int ch : {n.a, n.b, n.c})
                if (ch >= 0 && kind[ch] == 1 && is_op(ch)) maximal.insert(ch);
        }
        for (int o : all)
            if (kind[o] == 1 && is_op(o)) maximal.insert(o);
        if (param_only_ops >= 1 && maximal.size() <= 160) d.derived.assign(maximal.begin(), maximal.end());
    }
    std::vector<int> outs(d.g, d.g + 16);
    outs.insert(outs.end(), d.dg, d.dg + 64);
    // per-evaluation operation counts: derived parameters are inputs, not work
    d.flops_expr = G.op_count(outs, &d.derived);
    d.flops_accel = G.op_count(std::vector<int>(d.accel, d.accel + 4), &d.derived);
    d.flops_accel_ref = G.op_count(std::vector<int>(d.accel_ref, d.accel_ref + 4), &d.derived);
}

std::unique_ptr<Metric> load_metric_from_script(const ContentManager& cm, const std::string& js_path) {
    auto met = std::make_unique<Metric>();
    met->js_path = js_path;
    std::string cfg_path = fs::path(js_path).replace_extension(".json").string();
    met->cfg = load_config(cm, cfg_path, true);

    Graph::set_cur(&met->desc.graph);
    Sandbox sand;
    sand.cfg = &met->params;
    try {
        ScriptFunction fmetric(&sand, js_path);
        std::string to_polar_file, from_polar_file, origin_file, period_file;
        if (!cm.lookup_coordinates_file(met->cfg.to_polar, to_polar_file)) throw std::runtime_error("No to polar file " + met->cfg.to_polar);
        if (!cm.lookup_coordinates_file(met->cfg.from_polar, from_polar_file)) throw std::runtime_error("No from polar file " + met->cfg.from_polar);
        if (!cm.lookup_origins_file(met->cfg.origin_distance, origin_file)) throw std::runtime_error("No origin coordinate system file " + met->cfg.origin_distance);
        ScriptFunction fto(&sand, to_polar_file);
        ScriptFunction ffrom(&sand, from_polar_file);
        ScriptFunction forigin(&sand, origin_file);
        std::unique_ptr<ScriptFunction> fperiod;
        if (cm.lookup_coordinates_file(met->cfg.coordinate_periodicity, period_file)) fperiod = std::make_unique<ScriptFunction>(&sand, period_file);
        build_descriptor(*met, fmetric, fto, ffrom, forigin, fperiod.get());
    } catch (minijs::JsError& e) {
        Graph::set_cur(nullptr);
        throw std::runtime_error(std::string("Error in script exec: ") + e.what());
    } catch (...) {
        Graph::set_cur(nullptr);
        throw;
    }
    Graph::set_cur(nullptr);
    return met;
}

// ------------------------------------------------------------------------------------------
// Emission (replaces build_argument_string, metric.hpp:725-959)
// ------------------------------------------------------------------------------------------
// W_V1..4 (metric.hpp:849-869): weights of the acceleration components in the step-size control
void acceleration_weights(const Metric& m, double w[4]) {
    w[0] = w[1] = w[2] = w[3] = 1;
    if (m.cfg.system == CoordinateSystem::X_Y_THETA_PHI) { w[2] = 8; w[3] = m.constant_theta() ? 8 : 32; }
    else if (m.cfg.system == CoordinateSystem::CYLINDRICAL) { w[2] = 8; }
}

std::string emit_metric_source(const Metric& m, bool is_static) {
    const MetricDescriptor& d = m.desc;
    const Graph& G = d.graph;
    sym::EmitOptions opt;
    if (is_static)
        for (size_t i = 0; i < m.params.names.size(); i++) opt.param_values[m.params.names[i]] = m.params.current_values[i];

    std::ostringstream o;
    auto defflag = [&](const char* nm, bool v) { o << "#define " << nm << " " << (v ? 1 : 0) << "\n"; };
    auto deffloat = [&](const char* nm, double v) { o << "#define " << nm << " " << sym::float_literal(v) << "\n"; };
    o << "// generated by geodesic_raytracing_b200 from " << fs::path(m.js_path).filename().string()
      << " (" << (is_static ? "static" : "dynamic") << " program)\n";
    o << "// F_expr(g+dg) = " << d.flops_expr << " ops, accel(lowered) = " << d.flops_accel
      << " ops, accel(reference formulation) = " << d.flops_accel_ref << " ops\n";
    bool const_theta = m.constant_theta();
    defflag("GRT_BIG_METRIC", d.big);
    defflag("GRT_CONSTANT_THETA", const_theta);
    defflag("GRT_SINGULAR", m.cfg.singular);
    deffloat("GRT_SINGULAR_TERMINATOR", m.cfg.singular_terminator);
    defflag("GRT_TRAVERSABLE_EVENT_HORIZON", m.cfg.singular && m.cfg.traversable_event_horizon);
    defflag("GRT_ADAPTIVE_PRECISION", m.cfg.adaptive_precision);
    defflag("GRT_SINGULARITY_DETECTION", m.cfg.adaptive_precision && m.cfg.detect_singularities);
    // W_V1..4 (metric.hpp:849-869)
    double w[4];
    acceleration_weights(m, w);
    deffloat("GRT_W_V1", w[0]); deffloat("GRT_W_V2", w[1]); deffloat("GRT_W_V3", w[2]); deffloat("GRT_W_V4", w[3]);
    defflag("GRT_FORWARD_GEODESIC_PATH", m.cfg.follow_geodesics_forward);
    defflag("GRT_HAS_CYLINDRICAL_SINGULARITY", m.cfg.has_cylindrical_singularity);
    deffloat("GRT_CYLINDRICAL_TERMINATOR", m.cfg.cylindrical_terminator);
    defflag("GRT_UNCONDITIONALLY_NONSINGULAR", m.cfg.unconditionally_nonsingular);
    defflag("GRT_HAS_COORDINATE_PERIODICITY", d.has_period);
    defflag("GRT_USE_PREPASS", m.cfg.use_prepass);
    defflag("GRT_STATIC_PARAMS", is_static);
    // named parameters, then the derived ones (MetricDescriptor::derived); the host fills the whole block (pack_params)
    o << "#define GRT_NUM_PARAMS " << (m.params.names.size() + (is_static ? 0 : d.derived.size())) << "\n";
    if (!d.derived.empty()) o << "#define GRT_NUM_NAMED_PARAMS " << m.params.names.size() << "\n";
    if (!d.derived.empty()) {
        std::map<std::string, double> env;
        for (size_t i = 0; i < m.params.names.size(); i++) env[m.params.names[i]] = m.params.current_values[i];
        for (size_t k = 0; k < d.derived.size(); k++)
            opt.leaf_names[d.derived[k]] = is_static ? sym::float_literal((double)(float)G.eval(d.derived[k], env)) : "cfg.grt_d" + std::to_string(k);
    }
    // principal complex square root (stable on both half planes): re >= 0 -> (t, im / 2t), else (|im| / 2t, +-t)
    o << "GRT_FN float grt_csqrt_re(float re, float im) { float t = GRT_SQRT(0.5f * (GRT_SQRT(re * re + im * im) + fabsf(re))); "
         "return re >= 0.0f ? t : (t > 0.0f ? GRT_DIV(fabsf(im), 2.0f * t) : 0.0f); }\n";
    o << "GRT_FN float grt_csqrt_im(float re, float im) { float t = GRT_SQRT(0.5f * (GRT_SQRT(re * re + im * im) + fabsf(re))); "
         "return re >= 0.0f ? (t > 0.0f ? GRT_DIV(im, 2.0f * t) : 0.0f) : (im < 0.0f ? -t : t); }\n";
    // struct dynamic_config { float DYNVARS; } (cl.cl:940-945): first-access order
    o << "struct grt_params {\n";
    if (m.params.names.empty()) o << "    float _unused;\n";
    for (auto& n : m.params.names) o << "    float " << n << ";\n";
    if (!is_static)
        for (size_t k = 0; k < d.derived.size(); k++) o << "    float grt_d" << k << ";\n";
    o << "};\n\n";

    auto fn = [&](const std::string& sig, const std::vector<std::pair<std::string, int>>& outs) {
        o << "GRT_FN " << sig << "\n{\n    (void)cfg;\n" << sym::emit_block(G, outs, opt) << "}\n\n";
    };
    const std::string pos_decl = "    const float v1 = p[0], v2 = p[1], v3 = p[2], v4 = p[3]; (void)v1; (void)v2; (void)v3; (void)v4;\n";
    const std::string vel_decl = "    const float iv1 = v[0], iv2 = v[1], iv3 = v[2], iv4 = v[3]; (void)iv1; (void)iv2; (void)iv3; (void)iv4;\n";
    const std::string dvel_decl = "    const float dv1 = v[0], dv2 = v[1], dv3 = v[2], dv4 = v[3]; (void)dv1; (void)dv2; (void)dv3; (void)dv4;\n";
    auto fn_pos = [&](const std::string& sig, const std::string& decls, const std::vector<std::pair<std::string, int>>& outs) {
        o << "GRT_FN " << sig << "\n{\n    (void)cfg;\n" << decls << sym::emit_block(G, outs, opt) << "}\n\n";
    };
    (void)fn;

    std::vector<std::pair<std::string, int>> outs;
    for (int i = 0; i < 16; i++) outs.push_back({"g[" + std::to_string(i) + "]", d.g[i]});
    fn_pos("void grt_metric(const float* p, const grt_params& cfg, float* g)", pos_decl, outs);
    std::vector<std::pair<std::string, int>> douts;
    for (int i = 0; i < 64; i++) douts.push_back({"dg[" + std::to_string(i) + "]", d.dg[i]});
    fn_pos("void grt_metric_partials(const float* p, const grt_params& cfg, float* dg)", pos_decl, douts);
    std::vector<std::pair<std::string, int>> both = outs;
    both.insert(both.end(), douts.begin(), douts.end());
    fn_pos("void grt_metric_and_partials(const float* p, const grt_params& cfg, float* g, float* dg)", pos_decl, both);

    auto vec4 = [&](const int ids[4]) {
        std::vector<std::pair<std::string, int>> r;
        for (int i = 0; i < 4; i++) r.push_back({"o[" + std::to_string(i) + "]", ids[i]});
        return r;
    };
    fn_pos("void grt_to_polar(const float* p, const grt_params& cfg, float* o)", pos_decl, vec4(d.to_polar));
    fn_pos("void grt_to_polar_vel(const float* p, const float* v, const grt_params& cfg, float* o)", pos_decl + dvel_decl, vec4(d.dto_polar));
    fn_pos("void grt_from_polar(const float* p, const grt_params& cfg, float* o)", pos_decl, vec4(d.from_polar));
    fn_pos("void grt_from_polar_vel(const float* p, const float* v, const grt_params& cfg, float* o)", pos_decl + dvel_decl, vec4(d.dfrom_polar));
    fn_pos("void grt_distance(const float* p, const grt_params& cfg, float* o)", pos_decl, {{"o[0]", d.distance}});
    fn_pos("void grt_radii(const float* p, const grt_params& cfg, float* o)", pos_decl, {{"o[0]", d.radii[0]}, {"o[1]", d.radii[1]}});
    {
        // get_coordinate_period evaluates with v1..v4 = 0 (cl.cl:1338-1355)
        std::string zero_decl = "    const float v1 = 0.0f, v2 = 0.0f, v3 = 0.0f, v4 = 0.0f; (void)v1; (void)v2; (void)v3; (void)v4;\n";
        fn_pos("void grt_period(const grt_params& cfg, float* o)", zero_decl, vec4(d.period));
    }
    fn_pos("void grt_accel(const float* p, const float* v, const grt_params& cfg, float* o)", pos_decl + vel_decl, vec4(d.accel));
    {
        // acceleration and radii of the same point in one straight-line block (shared sub-expressions computed once): what a step
        // of the integration loop evaluates at the trial position
        std::vector<std::pair<std::string, int>> both_outs = vec4(d.accel);
        both_outs.push_back({"rr[0]", d.radii[0]});
        both_outs.push_back({"rr[1]", d.radii[1]});
        fn_pos("void grt_accel_radii(const float* p, const float* v, const grt_params& cfg, float* o, float* rr)", pos_decl + vel_decl, both_outs);
    }
    fn_pos("void grt_accel_ref(const float* p, const float* v, const grt_params& cfg, float* o)", pos_decl + vel_decl, vec4(d.accel_ref));
    return o.str();
}

std::vector<float> pack_params(const Metric& m) {
    std::vector<float> out(m.params.current_values.begin(), m.params.current_values.end());
    if (!m.desc.derived.empty()) {
        std::map<std::string, double> env;
        for (size_t i = 0; i < m.params.names.size(); i++) env[m.params.names[i]] = m.params.current_values[i];
        for (int id : m.desc.derived) out.push_back((float)m.desc.graph.eval(id, env));
    }
    return out;
}

}  // namespace grt
