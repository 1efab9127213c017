// Symbolic expression DAG: hash-consing, simplification, differentiation, emission.
// See sym.hpp for what this replaces in the reference.
#include "sym.hpp"

#include <array>
#include <cstring>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <functional>
#include <set>
#include <stdexcept>

namespace sym {

static thread_local Graph* g_cur = nullptr;
Graph* Graph::cur() {
    if (!g_cur) throw std::runtime_error("sym: no current graph");
    return g_cur;
}
void Graph::set_cur(Graph* g) { g_cur = g; }

Graph::Graph() { nodes_.reserve(4096); }

int Graph::intern(const Node& n) {
    char buf[96];
    std::string key;
    if (n.op == Op::Const) {
        uint64_t bits;
        double v = n.val;
        if (v == 0.0) v = 0.0;  // merge -0 and +0
        memcpy(&bits, &v, 8);
        snprintf(buf, sizeof buf, "c%llx", (unsigned long long)bits);
        key = buf;
    } else if (n.op == Op::Var || n.op == Op::Param) {
        key = (n.op == Op::Var ? "v:" : "p:") + n.name;
    } else {
        snprintf(buf, sizeof buf, "o%d:%d:%d:%d", (int)n.op, n.a, n.b, n.c);
        key = buf;
    }
    auto it = table_.find(key);
    if (it != table_.end()) return it->second;
    int id = (int)nodes_.size();
    nodes_.push_back(n);
    table_.emplace(std::move(key), id);
    return id;
}

int Graph::constant(double v) {
    Node n;
    n.op = Op::Const;
    n.val = (v == 0.0) ? 0.0 : v;
    return intern(n);
}
int Graph::var(const std::string& name) {
    Node n;
    n.op = Op::Var;
    n.name = name;
    return intern(n);
}
int Graph::param(const std::string& name) {
    Node n;
    n.op = Op::Param;
    n.name = name;
    return intern(n);
}

static bool is_commutative(Op op) {
    return op == Op::Add || op == Op::Mul || op == Op::Min || op == Op::Max || op == Op::Eq;
}

static double fold1(Op op, double a) {
    switch (op) {
        case Op::Neg: return -a;
        case Op::Sin: return std::sin(a);
        case Op::Cos: return std::cos(a);
        case Op::Tan: return std::tan(a);
        case Op::Asin: return std::asin(a);
        case Op::Acos: return std::acos(a);
        case Op::Atan: return std::atan(a);
        case Op::Fabs: return std::fabs(a);
        case Op::Log: return std::log(a);
        case Op::Sqrt: return std::sqrt(a);
        case Op::Exp: return std::exp(a);
        case Op::Sinh: return std::sinh(a);
        case Op::Cosh: return std::cosh(a);
        case Op::Tanh: return std::tanh(a);
        case Op::Sign: return a > 0 ? 1.0 : (a < 0 ? -1.0 : 0.0);
        default: throw std::runtime_error("sym: fold1 bad op");
    }
}
static double fold2(Op op, double a, double b) {
    switch (op) {
        case Op::Add: return a + b;
        case Op::Sub: return a - b;
        case Op::Mul: return a * b;
        case Op::Div: return a / b;
        case Op::Atan2: return std::atan2(a, b);
        case Op::Pow: return std::pow(a, b);
        case Op::Lt: return a < b ? 1.0 : 0.0;
        case Op::Lte: return a <= b ? 1.0 : 0.0;
        case Op::Eq: return a == b ? 1.0 : 0.0;
        case Op::Gt: return a > b ? 1.0 : 0.0;
        case Op::Gte: return a >= b ? 1.0 : 0.0;
        case Op::Fmod: return std::fmod(a, b);
        case Op::Min: return std::min(a, b);
        case Op::Max: return std::max(a, b);
        default: throw std::runtime_error("sym: fold2 bad op");
    }
}

// Returns a node id when the expression collapses (done=true), else done=false and the caller interns.
int Graph::simplify(Op op, int a, int b, int c, bool& done) {
    done = true;
    const bool ca = a >= 0 && is_const(a), cb = b >= 0 && is_const(b);
    auto isneg = [&](int id) { return nodes_[id].op == Op::Neg; };
    switch (op) {
        case Op::Add:
            if (ca && cb) return constant(cval(a) + cval(b));
            if (is_const(a, 0.0)) return b;
            if (is_const(b, 0.0)) return a;
            if (isneg(b)) return make(Op::Sub, a, nodes_[b].a);
            if (isneg(a)) return make(Op::Sub, b, nodes_[a].a);
            break;
        case Op::Sub:
            if (ca && cb) return constant(cval(a) - cval(b));
            if (is_const(b, 0.0)) return a;
            if (is_const(a, 0.0)) return make(Op::Neg, b);
            if (a == b) return constant(0.0);
            if (isneg(b)) return make(Op::Add, a, nodes_[b].a);
            if (isneg(a)) return make(Op::Neg, make(Op::Add, nodes_[a].a, b));
            break;
        case Op::Mul:
            if (ca && cb) return constant(cval(a) * cval(b));
            if (is_const(a, 0.0) || is_const(b, 0.0)) return constant(0.0);
            if (is_const(a, 1.0)) return b;
            if (is_const(b, 1.0)) return a;
            if (is_const(a, -1.0)) return make(Op::Neg, b);
            if (is_const(b, -1.0)) return make(Op::Neg, a);
            if (isneg(a) && isneg(b)) return make(Op::Mul, nodes_[a].a, nodes_[b].a);
            if (isneg(a)) return make(Op::Neg, make(Op::Mul, nodes_[a].a, b));
            if (isneg(b)) return make(Op::Neg, make(Op::Mul, a, nodes_[b].a));
            // fold constants through nested products: c1 * (c2 * x) -> (c1*c2) * x
            if (ca && nodes_[b].op == Op::Mul && is_const(nodes_[b].a))
                return make(Op::Mul, constant(cval(a) * cval(nodes_[b].a)), nodes_[b].b);
            if (cb && nodes_[a].op == Op::Mul && is_const(nodes_[a].a))
                return make(Op::Mul, constant(cval(b) * cval(nodes_[a].a)), nodes_[a].b);
            if (ca && cval(a) < 0) return make(Op::Neg, make(Op::Mul, constant(-cval(a)), b));
            if (cb && cval(b) < 0) return make(Op::Neg, make(Op::Mul, a, constant(-cval(b))));
            break;
        case Op::Div:
            if (ca && cb) return constant(cval(a) / cval(b));
            if (is_const(a, 0.0)) return constant(0.0);
            if (is_const(b, 1.0)) return a;
            if (is_const(b, -1.0)) return make(Op::Neg, a);
            if (a == b) return constant(1.0);
            if (isneg(a) && isneg(b)) return make(Op::Div, nodes_[a].a, nodes_[b].a);
            if (isneg(a)) return make(Op::Neg, make(Op::Div, nodes_[a].a, b));
            if (isneg(b)) return make(Op::Neg, make(Op::Div, a, nodes_[b].a));
            if (cb) return make(Op::Mul, a, constant(1.0 / cval(b)));
            break;
        case Op::Neg:
            if (ca) return constant(-cval(a));
            if (isneg(a)) return nodes_[a].a;
            if (nodes_[a].op == Op::Sub) return make(Op::Sub, nodes_[a].b, nodes_[a].a);
            break;
        case Op::Sin:
            if (ca) return constant(fold1(op, cval(a)));
            if (isneg(a)) return make(Op::Neg, make(Op::Sin, nodes_[a].a));
            if (nodes_[a].op == Op::Atan2) {  // sin(atan2(y,x)) = y / hypot
                int y = nodes_[a].a, x = nodes_[a].b;
                int h = make(Op::Sqrt, make(Op::Add, make(Op::Mul, y, y), make(Op::Mul, x, x)));
                return make(Op::Div, y, h);
            }
            break;
        case Op::Cos:
            if (ca) return constant(fold1(op, cval(a)));
            if (isneg(a)) return make(Op::Cos, nodes_[a].a);
            if (nodes_[a].op == Op::Atan2) {
                int y = nodes_[a].a, x = nodes_[a].b;
                int h = make(Op::Sqrt, make(Op::Add, make(Op::Mul, y, y), make(Op::Mul, x, x)));
                return make(Op::Div, x, h);
            }
            break;
        case Op::Fabs:
            if (ca) return constant(std::fabs(cval(a)));
            if (isneg(a)) return make(Op::Fabs, nodes_[a].a);
            if (nodes_[a].op == Op::Fabs || nodes_[a].op == Op::Sqrt || nodes_[a].op == Op::Exp ||
                nodes_[a].op == Op::Cosh)
                return a;
            if (nodes_[a].op == Op::Mul && nodes_[a].a == nodes_[a].b) return a;  // |x*x|
            break;
        case Op::Tan: case Op::Asin: case Op::Acos: case Op::Atan: case Op::Log:
        case Op::Exp: case Op::Sinh: case Op::Cosh: case Op::Tanh: case Op::Sign:
            if (ca) return constant(fold1(op, cval(a)));
            break;
        case Op::Sqrt:
            if (ca) return constant(std::sqrt(cval(a)));
            break;
        case Op::Pow:
            if (ca && cb) return constant(std::pow(cval(a), cval(b)));
            if (cb) {
                double e = cval(b);
                if (e == 0.0) return constant(1.0);
                if (e == 1.0) return a;
                if (e == 0.5) return make(Op::Sqrt, a);
                if (e == -0.5) return make(Op::Div, constant(1.0), make(Op::Sqrt, a));
                if (e == std::floor(e) && std::fabs(e) <= 16) {
                    int n = (int)std::fabs(e);
                    int result = -1, base = a;
                    while (n) {  // square and multiply
                        if (n & 1) result = result < 0 ? base : make(Op::Mul, result, base);
                        n >>= 1;
                        if (n) base = make(Op::Mul, base, base);
                    }
                    return e < 0 ? make(Op::Div, constant(1.0), result) : result;
                }
            }
            if (is_const(a, 1.0)) return constant(1.0);
            break;
        case Op::Atan2: case Op::Fmod:
            if (ca && cb) return constant(fold2(op, cval(a), cval(b)));
            break;
        case Op::Min: case Op::Max:
            if (ca && cb) return constant(fold2(op, cval(a), cval(b)));
            if (a == b) return a;
            break;
        case Op::CsqrtRe: case Op::CsqrtIm:
            if (ca && cb) {
                std::complex<double> r = std::sqrt(std::complex<double>(cval(a), cval(b) == 0.0 ? 0.0 : cval(b)));
                return constant(op == Op::CsqrtRe ? r.real() : r.imag());
            }
            break;
        case Op::Lt: case Op::Lte: case Op::Eq: case Op::Gt: case Op::Gte:
            if (ca && cb) return constant(fold2(op, cval(a), cval(b)));
            break;
        case Op::Select:
            if (ca) return cval(a) != 0.0 ? b : c;
            if (b == c) return b;
            break;
        default: break;
    }
    done = false;
    return -1;
}

int Graph::make(Op op, int a, int b, int c) {
    bool done = false;
    int r = simplify(op, a, b, c, done);
    if (done) return r;
    if (is_commutative(op) && a > b) std::swap(a, b);
    Node n;
    n.op = op;
    n.a = a;
    n.b = b;
    n.c = c;
    return intern(n);
}

int Graph::diff(int id, int wrt) { return diff_key(id, wrt); }

int Graph::diff_along(int id, const std::map<int, int>& tangents) {
    int key = -1;
    for (size_t i = 0; i < tangent_sets_.size(); i++)
        if (tangent_sets_[i] == tangents) key = -((int)i + 2);
    if (key == -1) { tangent_sets_.push_back(tangents); key = -((int)tangent_sets_.size() + 1); }
    return diff_key(id, key);
}

// wrt >= 0: partial derivative with respect to that variable node; wrt <= -2: tangent set -(wrt + 2)
int Graph::diff_key(int id, int wrt) {
    auto key = std::make_pair(id, wrt);
    auto it = diff_memo_.find(key);
    if (it != diff_memo_.end()) return it->second;
    const Node n = nodes_[id];  // copy: nodes_ may reallocate
    int r;
    auto D = [&](int x) { return diff_key(x, wrt); };
    auto C = [&](double v) { return constant(v); };
    auto M = [&](Op op, int x, int y = -1, int z = -1) { return make(op, x, y, z); };
    switch (n.op) {
        case Op::Const: case Op::Param: r = C(0.0); break;
        case Op::Var:
            if (wrt >= 0) r = C(id == wrt ? 1.0 : 0.0);
            else { const auto& ts = tangent_sets_[-(wrt + 2)]; auto f = ts.find(id); r = f == ts.end() ? C(0.0) : f->second; }
            break;
        case Op::Add: r = M(Op::Add, D(n.a), D(n.b)); break;
        case Op::Sub: r = M(Op::Sub, D(n.a), D(n.b)); break;
        case Op::Mul: r = M(Op::Add, M(Op::Mul, D(n.a), n.b), M(Op::Mul, n.a, D(n.b))); break;
        case Op::Div: {
            // q = a/b: q' = (a' - q b')/b reuses the quotient (three ops instead of five)
            int da = D(n.a), db = D(n.b);
            if (is_const(db, 0.0)) r = M(Op::Div, da, n.b);
            else r = M(Op::Div, M(Op::Sub, da, M(Op::Mul, id, db)), n.b);
            break;
        }
        case Op::Neg: r = M(Op::Neg, D(n.a)); break;
        case Op::Sin: r = M(Op::Mul, M(Op::Cos, n.a), D(n.a)); break;
        case Op::Cos: r = M(Op::Neg, M(Op::Mul, M(Op::Sin, n.a), D(n.a))); break;
        case Op::Tan: { int cs = M(Op::Cos, n.a); r = M(Op::Div, D(n.a), M(Op::Mul, cs, cs)); break; }
        case Op::Asin: r = M(Op::Div, D(n.a), M(Op::Sqrt, M(Op::Sub, C(1.0), M(Op::Mul, n.a, n.a)))); break;
        case Op::Acos: r = M(Op::Neg, M(Op::Div, D(n.a), M(Op::Sqrt, M(Op::Sub, C(1.0), M(Op::Mul, n.a, n.a))))); break;
        case Op::Atan: r = M(Op::Div, D(n.a), M(Op::Add, C(1.0), M(Op::Mul, n.a, n.a))); break;
        case Op::Atan2: {
            // d atan2(y,x) = (x dy - y dx)/(x^2+y^2)
            int y = n.a, x = n.b;
            r = M(Op::Div, M(Op::Sub, M(Op::Mul, x, D(y)), M(Op::Mul, y, D(x))),
                  M(Op::Add, M(Op::Mul, x, x), M(Op::Mul, y, y)));
            break;
        }
        case Op::Fabs: r = M(Op::Mul, M(Op::Sign, n.a), D(n.a)); break;
        case Op::Log: r = M(Op::Div, D(n.a), n.a); break;
        case Op::Sqrt: r = M(Op::Div, D(n.a), M(Op::Mul, C(2.0), id)); break;
        case Op::Exp: r = M(Op::Mul, id, D(n.a)); break;
        case Op::Sinh: r = M(Op::Mul, M(Op::Cosh, n.a), D(n.a)); break;
        case Op::Cosh: r = M(Op::Mul, M(Op::Sinh, n.a), D(n.a)); break;
        case Op::Tanh: r = M(Op::Mul, M(Op::Sub, C(1.0), M(Op::Mul, id, id)), D(n.a)); break;
        case Op::Pow: {
            int da = D(n.a), db = D(n.b);
            if (is_const(n.b)) {
                r = M(Op::Mul, M(Op::Mul, n.b, M(Op::Pow, n.a, C(cval(n.b) - 1.0))), da);
            } else {
                // a^b (b' ln a + b a'/a)
                r = M(Op::Mul, id, M(Op::Add, M(Op::Mul, db, M(Op::Log, n.a)), M(Op::Div, M(Op::Mul, n.b, da), n.a)));
            }
            break;
        }
        case Op::Select: r = M(Op::Select, n.a, D(n.b), D(n.c)); break;
        case Op::Lt: case Op::Lte: case Op::Eq: case Op::Gt: case Op::Gte: case Op::Sign:
            r = C(0.0); break;
        case Op::Fmod: r = D(n.a); break;
        case Op::Min: r = M(Op::Select, M(Op::Lt, n.a, n.b), D(n.a), D(n.b)); break;
        case Op::Max: r = M(Op::Select, M(Op::Gt, n.a, n.b), D(n.a), D(n.b)); break;
        case Op::CsqrtRe: case Op::CsqrtIm: {
            // s = sqrt(z): ds = dz / (2 s) = dz conj(s) / (2 |s|^2)
            int sr = M(Op::CsqrtRe, n.a, n.b), si = M(Op::CsqrtIm, n.a, n.b);
            int da = D(n.a), db = D(n.b);
            int den = M(Op::Mul, C(2.0), M(Op::Add, M(Op::Mul, sr, sr), M(Op::Mul, si, si)));
            if (n.op == Op::CsqrtRe) r = M(Op::Div, M(Op::Add, M(Op::Mul, da, sr), M(Op::Mul, db, si)), den);
            else r = M(Op::Div, M(Op::Sub, M(Op::Mul, db, sr), M(Op::Mul, da, si)), den);
            break;
        }
        default: throw std::runtime_error("sym: diff bad op");
    }
    diff_memo_[key] = r;
    return r;
}

std::vector<int> Graph::gradient(int id, const std::vector<int>& wrt) {
    // reachable nodes, processed from the root down: children are created before their parents, so ids decrease
    std::set<int> reach;
    std::vector<int> stack{id};
    while (!stack.empty()) {
        int x = stack.back();
        stack.pop_back();
        if (!reach.insert(x).second) continue;
        const Node& n = nodes_[x];
        if (n.a >= 0) stack.push_back(n.a);
        if (n.b >= 0) stack.push_back(n.b);
        if (n.c >= 0) stack.push_back(n.c);
    }
    std::map<int, int> adj;
    adj[id] = constant(1.0);
    auto C = [&](double v) { return constant(v); };
    auto M = [&](Op op, int x, int y = -1, int z = -1) { return make(op, x, y, z); };
    auto push = [&](int child, int contribution) {
        if (is_const(contribution, 0.0)) return;
        auto it = adj.find(child);
        if (it == adj.end()) adj[child] = contribution;
        else it->second = M(Op::Add, it->second, contribution);
    };
    for (auto rit = reach.rbegin(); rit != reach.rend(); ++rit) {
        const int x = *rit;
        auto ai = adj.find(x);
        if (ai == adj.end()) continue;
        const int A = ai->second;
        const Node n = nodes_[x];
        switch (n.op) {
            case Op::Const: case Op::Param: case Op::Var: break;
            case Op::Add: push(n.a, A); push(n.b, A); break;
            case Op::Sub: push(n.a, A); push(n.b, M(Op::Neg, A)); break;
            case Op::Neg: push(n.a, M(Op::Neg, A)); break;
            case Op::Mul: push(n.a, M(Op::Mul, A, n.b)); push(n.b, M(Op::Mul, A, n.a)); break;
            case Op::Div: { int t = M(Op::Div, A, n.b); push(n.a, t); push(n.b, M(Op::Neg, M(Op::Mul, t, x))); break; }
            case Op::Sin: push(n.a, M(Op::Mul, A, M(Op::Cos, n.a))); break;
            case Op::Cos: push(n.a, M(Op::Neg, M(Op::Mul, A, M(Op::Sin, n.a)))); break;
            case Op::Tan: { int cs = M(Op::Cos, n.a); push(n.a, M(Op::Div, A, M(Op::Mul, cs, cs))); break; }
            case Op::Asin: push(n.a, M(Op::Div, A, M(Op::Sqrt, M(Op::Sub, C(1.0), M(Op::Mul, n.a, n.a))))); break;
            case Op::Acos: push(n.a, M(Op::Neg, M(Op::Div, A, M(Op::Sqrt, M(Op::Sub, C(1.0), M(Op::Mul, n.a, n.a)))))); break;
            case Op::Atan: push(n.a, M(Op::Div, A, M(Op::Add, C(1.0), M(Op::Mul, n.a, n.a)))); break;
            case Op::Atan2: {
                int y = n.a, xx = n.b;
                int t = M(Op::Div, A, M(Op::Add, M(Op::Mul, xx, xx), M(Op::Mul, y, y)));
                push(y, M(Op::Mul, t, xx));
                push(xx, M(Op::Neg, M(Op::Mul, t, y)));
                break;
            }
            case Op::Fabs: push(n.a, M(Op::Mul, A, M(Op::Sign, n.a))); break;
            case Op::Log: push(n.a, M(Op::Div, A, n.a)); break;
            case Op::Sqrt: push(n.a, M(Op::Div, A, M(Op::Mul, C(2.0), x))); break;
            case Op::Exp: push(n.a, M(Op::Mul, A, x)); break;
            case Op::Sinh: push(n.a, M(Op::Mul, A, M(Op::Cosh, n.a))); break;
            case Op::Cosh: push(n.a, M(Op::Mul, A, M(Op::Sinh, n.a))); break;
            case Op::Tanh: push(n.a, M(Op::Mul, A, M(Op::Sub, C(1.0), M(Op::Mul, x, x)))); break;
            case Op::Pow:
                if (is_const(n.b)) push(n.a, M(Op::Mul, A, M(Op::Mul, n.b, M(Op::Pow, n.a, C(cval(n.b) - 1.0)))));
                else {
                    int t = M(Op::Mul, A, x);
                    push(n.a, M(Op::Div, M(Op::Mul, t, n.b), n.a));
                    push(n.b, M(Op::Mul, t, M(Op::Log, n.a)));
                }
                break;
            case Op::Select: push(n.b, M(Op::Select, n.a, A, C(0.0))); push(n.c, M(Op::Select, n.a, C(0.0), A)); break;
            case Op::Lt: case Op::Lte: case Op::Eq: case Op::Gt: case Op::Gte: case Op::Sign: break;
            case Op::Fmod: push(n.a, A); break;
            case Op::Min: { int c = M(Op::Lt, n.a, n.b); push(n.a, M(Op::Select, c, A, C(0.0))); push(n.b, M(Op::Select, c, C(0.0), A)); break; }
            case Op::Max: { int c = M(Op::Gt, n.a, n.b); push(n.a, M(Op::Select, c, A, C(0.0))); push(n.b, M(Op::Select, c, C(0.0), A)); break; }
            case Op::CsqrtRe: case Op::CsqrtIm: {
                int sr = M(Op::CsqrtRe, n.a, n.b), si = M(Op::CsqrtIm, n.a, n.b);
                int t = M(Op::Div, A, M(Op::Mul, C(2.0), M(Op::Add, M(Op::Mul, sr, sr), M(Op::Mul, si, si))));
                if (n.op == Op::CsqrtRe) { push(n.a, M(Op::Mul, t, sr)); push(n.b, M(Op::Mul, t, si)); }
                else { push(n.a, M(Op::Neg, M(Op::Mul, t, si))); push(n.b, M(Op::Mul, t, sr)); }
                break;
            }
            default: throw std::runtime_error("sym: gradient bad op");
        }
    }
    std::vector<int> out;
    for (int w : wrt) { auto it = adj.find(w); out.push_back(it == adj.end() ? constant(0.0) : it->second); }
    return out;
}

int Graph::substitute(int id, const std::map<int, int>& repl) {
    std::unordered_map<int, int> memo;
    std::function<int(int)> go = [&](int x) -> int {
        auto r = repl.find(x);
        if (r != repl.end()) return r->second;
        auto m = memo.find(x);
        if (m != memo.end()) return m->second;
        const Node n = nodes_[x];
        int out;
        if (n.op == Op::Const || n.op == Op::Var || n.op == Op::Param) out = x;
        else {
            int a = n.a >= 0 ? go(n.a) : -1, b = n.b >= 0 ? go(n.b) : -1, c = n.c >= 0 ? go(n.c) : -1;
            out = (a == n.a && b == n.b && c == n.c) ? x : make(n.op, a, b, c);
        }
        memo[x] = out;
        return out;
    };
    return go(id);
}

namespace {
struct Term {
    double coef = 1.0;
    std::vector<int> factors;   // sorted node ids, repeats allowed
};
// x*y - x*y, 2ab + ab: terms with the same factor list are one term
void merge_like_terms(std::vector<Term>& ts) {
    std::map<std::vector<int>, double> acc;
    std::vector<std::vector<int>> order;
    for (auto& t : ts) {
        auto it = acc.find(t.factors);
        if (it == acc.end()) { acc[t.factors] = t.coef; order.push_back(t.factors); }
        else it->second += t.coef;
    }
    if (order.size() == ts.size()) return;
    ts.clear();
    for (auto& f : order) {
        double c = acc[f];
        if (c != 0.0) { Term t; t.coef = c; t.factors = f; ts.push_back(t); }
    }
}
}  // namespace

void Graph::poly_axpy(Poly& acc, double alpha, const Poly& x) {
    for (auto& kv : x.terms) {
        double& c = acc.terms[kv.first];
        c += alpha * kv.second;
        if (c == 0.0) acc.terms.erase(kv.first);
    }
}

bool Graph::poly_mul(const Poly& a, const Poly& b, Poly& out, size_t max_terms) {
    Poly r;
    for (auto& ka : a.terms)
        for (auto& kb : b.terms) {
            std::vector<int> m(ka.first.size() + kb.first.size());
            std::merge(ka.first.begin(), ka.first.end(), kb.first.begin(), kb.first.end(), m.begin());
            double& c = r.terms[m];
            c += ka.second * kb.second;
            if (r.terms.size() > max_terms) return false;
        }
    for (auto it = r.terms.begin(); it != r.terms.end();) it = (it->second == 0.0) ? r.terms.erase(it) : std::next(it);
    out = r;
    return true;
}

bool Graph::to_poly(int id, Poly& out, size_t max_terms) {
    auto memo = poly_memo_.find(id);
    if (memo != poly_memo_.end()) { out = memo->second; return true; }
    const Node n = nodes_[id];
    Poly r;
    switch (n.op) {
        case Op::Const:
            if (n.val != 0.0) r.terms[{}] = n.val;
            break;
        case Op::Add: case Op::Sub: {
            Poly a, b;
            if (!to_poly(n.a, a, max_terms) || !to_poly(n.b, b, max_terms)) return false;
            r = a;
            poly_axpy(r, n.op == Op::Add ? 1.0 : -1.0, b);
            break;
        }
        case Op::Neg: {
            Poly a;
            if (!to_poly(n.a, a, max_terms)) return false;
            poly_axpy(r, -1.0, a);
            break;
        }
        case Op::Mul: {
            Poly a, b;
            if (!to_poly(n.a, a, max_terms) || !to_poly(n.b, b, max_terms)) return false;
            if (!poly_mul(a, b, r, max_terms)) return false;
            break;
        }
        case Op::Pow: {
            const Node e = nodes_[n.b];
            if (e.op == Op::Const && e.val == std::floor(e.val) && e.val >= 1.0 && e.val <= 4.0) {
                Poly a;
                if (!to_poly(n.a, a, max_terms)) return false;
                r = a;
                for (int k = 1; k < (int)e.val; k++) { Poly t; if (!poly_mul(r, a, t, max_terms)) return false; r = t; }
                break;
            }
            r.terms[{id}] = 1.0;
            break;
        }
        default:
            r.terms[{id}] = 1.0;   // an atom: variable, parameter, quotient, function value
            break;
    }
    if (r.terms.size() > max_terms) return false;
    poly_memo_[id] = r;
    out = r;
    return true;
}

int Graph::from_poly(const Poly& p) {
    // one sum of plain products, built in a canonical order; `factor` then pulls the common factors out
    int acc = -1;
    std::vector<std::pair<std::vector<int>, double>> ts(p.terms.begin(), p.terms.end());
    std::stable_sort(ts.begin(), ts.end(), [](const auto& a, const auto& b) { return (a.second > 0) > (b.second > 0); });
    for (auto& t : ts) {
        int prod = -1;
        for (int f : t.first) prod = prod < 0 ? f : make(Op::Mul, prod, f);
        double c = t.second;
        bool neg = c < 0 && acc >= 0;
        if (neg) c = -c;
        int term = prod < 0 ? constant(c) : (c == 1.0 ? prod : make(Op::Mul, constant(c), prod));
        acc = acc < 0 ? term : make(neg ? Op::Sub : Op::Add, acc, term);
    }
    if (acc < 0) return constant(0.0);
    return factor(acc, false);
}

int Graph::factor(int id, bool distribute) {
    auto& memo_table = factor_memo_[distribute ? 1 : 0];
    auto memo = memo_table.find(id);
    if (memo != memo_table.end()) return memo->second;
    const Node n = nodes_[id];
    int result;
    auto is_sum = [&](Op op) { return op == Op::Add || op == Op::Sub; };
    if (n.op == Op::Const || n.op == Op::Var || n.op == Op::Param) {
        result = id;
    } else if (is_sum(n.op) || n.op == Op::Neg || n.op == Op::Mul) {
        // flatten the sum rooted here into signed products of (recursively cleaned) atomic factors
        std::vector<Term> terms;
        std::function<void(int, Term&)> flatten_mul = [&](int x, Term& t) {
            const Node m = nodes_[x];
            if (m.op == Op::Mul) { flatten_mul(m.a, t); flatten_mul(m.b, t); }
            else if (m.op == Op::Neg) { t.coef = -t.coef; flatten_mul(m.a, t); }
            else if (m.op == Op::Const) t.coef *= m.val;
            else t.factors.push_back(factor(x, distribute));
        };
        std::function<void(int, double)> collect = [&](int x, double sign) {
            const Node m = nodes_[x];
            if (m.op == Op::Add) { collect(m.a, sign); collect(m.b, sign); }
            else if (m.op == Op::Sub) { collect(m.a, sign); collect(m.b, -sign); }
            else if (m.op == Op::Neg) collect(m.a, -sign);
            else {
                Term t;
                t.coef = sign;
                flatten_mul(x, t);
                // a cleaned factor may itself have become a product or a constant: absorb it
                std::vector<int> flat;
                for (size_t i = 0; i < t.factors.size(); i++) {
                    const Node f = nodes_[t.factors[i]];
                    if (f.op == Op::Const) t.coef *= f.val;
                    else if (f.op == Op::Neg) { t.coef = -t.coef; t.factors.push_back(f.a); }
                    else if (f.op == Op::Mul) { t.factors.push_back(f.a); t.factors.push_back(f.b); }
                    else flat.push_back(t.factors[i]);
                }
                std::sort(flat.begin(), flat.end());
                t.factors = flat;
                if (t.coef == 0.0) return;
                if (distribute) {
                    // multiply out short sums sitting inside the product (bounded, so nothing blows up)
                    std::vector<Term> work{t}, done;
                    bool gave_up = false;
                    while (!work.empty() && !gave_up) {
                        Term cur = work.back();
                        work.pop_back();
                        size_t at = cur.factors.size();
                        for (size_t i = 0; i < cur.factors.size(); i++)
                            if (is_sum(nodes_[cur.factors[i]].op)) { at = i; break; }
                        if (at == cur.factors.size()) { done.push_back(cur); continue; }
                        int sum_node = cur.factors[at];
                        cur.factors.erase(cur.factors.begin() + at);
                        std::vector<std::pair<double, int>> parts;
                        std::function<void(int, double)> split = [&](int y, double sg) {
                            const Node q = nodes_[y];
                            if (q.op == Op::Add) { split(q.a, sg); split(q.b, sg); }
                            else if (q.op == Op::Sub) { split(q.a, sg); split(q.b, -sg); }
                            else if (q.op == Op::Neg) split(q.a, -sg);
                            else parts.push_back({sg, y});
                        };
                        split(sum_node, 1.0);
                        if (parts.size() > 4 || work.size() + done.size() + parts.size() > 12) { gave_up = true; break; }
                        for (auto& pr : parts) {
                            Term nt = cur;
                            nt.coef *= pr.first;
                            std::vector<int> stack{pr.second};
                            while (!stack.empty()) {
                                int y = stack.back();
                                stack.pop_back();
                                const Node q = nodes_[y];
                                if (q.op == Op::Mul) { stack.push_back(q.a); stack.push_back(q.b); }
                                else if (q.op == Op::Neg) { nt.coef = -nt.coef; stack.push_back(q.a); }
                                else if (q.op == Op::Const) nt.coef *= q.val;
                                else nt.factors.push_back(y);
                            }
                            std::sort(nt.factors.begin(), nt.factors.end());
                            if (nt.coef != 0.0) work.push_back(nt);
                        }
                    }
                    if (!gave_up) { for (auto& d : done) terms.push_back(d); return; }
                }
                terms.push_back(t);
            }
        };
        collect(id, 1.0);
        merge_like_terms(terms);
        std::function<int(std::vector<Term>&)> build = [&](std::vector<Term>& ts) -> int {
            if (ts.empty()) return constant(0.0);
            // most frequent factor (counted once per term)
            std::map<int, int> freq;
            for (auto& t : ts) {
                int prev = -1;
                for (int f : t.factors) { if (f != prev) freq[f]++; prev = f; }
            }
            int best = -1, best_n = 1;
            for (auto& kv : freq) if (kv.second > best_n) { best = kv.first; best_n = kv.second; }
            auto product = [&](const Term& t) {
                int acc = -1;
                for (int f : t.factors) acc = acc < 0 ? f : make(Op::Mul, acc, f);
                if (acc < 0) return constant(t.coef);
                if (t.coef != 1.0) acc = make(Op::Mul, constant(t.coef), acc);
                return acc;
            };
            if (best < 0 || ts.size() == 1) {
                // no shared factor: plain sum, positive terms first so that it emits as a + b - c
                std::stable_sort(ts.begin(), ts.end(), [](const Term& a, const Term& b) { return (a.coef > 0) > (b.coef > 0); });
                int acc = -1;
                for (auto& t : ts) {
                    Term p = t;
                    bool neg = p.coef < 0 && acc >= 0;
                    if (neg) p.coef = -p.coef;
                    int v = product(p);
                    acc = acc < 0 ? v : make(neg ? Op::Sub : Op::Add, acc, v);
                }
                return acc;
            }
            std::vector<Term> with, without;
            for (auto& t : ts) {
                auto it = std::find(t.factors.begin(), t.factors.end(), best);
                if (it == t.factors.end()) without.push_back(t);
                else { Term r = t; r.factors.erase(r.factors.begin() + (it - t.factors.begin())); with.push_back(r); }
            }
            int inner = build(with);
            int lhs = make(Op::Mul, best, inner);
            if (without.empty()) return lhs;
            int rest = build(without);
            return make(Op::Add, lhs, rest);
        };
        result = build(terms);
    } else {
        int a = n.a >= 0 ? factor(n.a, distribute) : -1, b = n.b >= 0 ? factor(n.b, distribute) : -1, c = n.c >= 0 ? factor(n.c, distribute) : -1;
        result = (a == n.a && b == n.b && c == n.c) ? id : make(n.op, a, b, c);
    }
    memo_table[id] = result;
    memo_table[result] = result;
    return result;
}

bool Graph::depends_on(int id, int leaf) const {
    std::set<int> seen;
    std::function<bool(int)> go = [&](int x) -> bool {
        if (x == leaf) return true;
        if (x < 0 || !seen.insert(x).second) return false;
        const Node& n = nodes_[x];
        return (n.a >= 0 && go(n.a)) || (n.b >= 0 && go(n.b)) || (n.c >= 0 && go(n.c));
    };
    return go(id);
}

size_t Graph::op_count(const std::vector<int>& outs, const std::vector<int>* leaves) const {
    std::set<int> seen;
    if (leaves) seen.insert(leaves->begin(), leaves->end());
    size_t cnt = 0;
    std::function<void(int)> go = [&](int x) {
        if (x < 0 || !seen.insert(x).second) return;
        const Node& n = nodes_[x];
        if (n.op == Op::Const || n.op == Op::Var || n.op == Op::Param) return;
        cnt++;
        go(n.a); go(n.b); go(n.c);
    };
    for (int o : outs) go(o);
    return cnt;
}

size_t Graph::instr_count(const std::vector<int>& outs) const {
    std::set<int> seen;
    std::map<int, int> uses;
    std::function<void(int)> go = [&](int x) {
        if (x < 0) return;
        uses[x]++;
        if (!seen.insert(x).second) return;
        const Node& n = nodes_[x];
        go(n.a); go(n.b); go(n.c);
    };
    for (int o : outs) go(o);
    size_t cnt = 0;
    std::set<int> fused;
    for (int x : seen) {
        const Node& n = nodes_[x];
        switch (n.op) {
            case Op::Const: case Op::Var: case Op::Param: case Op::Neg: break;
            case Op::Div: cnt += 2; break;
            case Op::Add: case Op::Sub: {
                cnt++;
                for (int ch : {n.a, n.b}) {
                    int y = ch;
                    if (nodes_[y].op == Op::Neg) y = nodes_[y].a;
                    if (nodes_[y].op == Op::Mul && uses[y] == 1 && !fused.count(y)) { fused.insert(y); cnt--; break; }   // FMA: the product is free
                }
                break;
            }
            default: cnt++; break;
        }
    }
    return cnt;
}

// ---- value-based merging --------------------------------------------------------------------
std::vector<int> Graph::unify(const std::vector<int>& outs, const std::map<std::string, double>& params) {
    constexpr int K = 8;
    constexpr double kTol = 1e-10;
    // reachable nodes, ascending = topological
    std::set<int> reach;
    {
        std::vector<int> stack(outs.begin(), outs.end());
        while (!stack.empty()) {
            int x = stack.back();
            stack.pop_back();
            if (x < 0 || !reach.insert(x).second) continue;
            const Node& n = nodes_[x];
            stack.push_back(n.a); stack.push_back(n.b); stack.push_back(n.c);
        }
    }
    // deterministic sample generator (splitmix64)
    uint64_t state = 0x9e3779b97f4a7c15ull;
    auto rnd = [&]() {
        uint64_t z = (state += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    };
    using Vals = std::array<double, K>;
    std::unordered_map<int, Vals> val;          // old id -> sample values
    auto is_atom_op = [](Op op) {
        switch (op) {
            case Op::Select: case Op::Lt: case Op::Lte: case Op::Eq: case Op::Gt: case Op::Gte:
            case Op::Fabs: case Op::Sign: case Op::Min: case Op::Max: case Op::Fmod: case Op::Atan2:
            case Op::CsqrtRe: case Op::CsqrtIm:
                return true;
            default: return false;
        }
    };
    auto finite = [](const Vals& v) { for (double x : v) if (!std::isfinite(x)) return false; return true; };
    // representatives: new id + values, bucketed by the float rounding of |first sample|
    struct Rep { int id; Vals v; };
    std::unordered_map<uint32_t, std::vector<Rep>> buckets;
    auto key_of = [](double x) { float f = (float)std::fabs(x); uint32_t u; memcpy(&u, &f, 4); return u; };
    auto close = [&](const Vals& a, const Vals& b, double sign) {
        for (int k = 0; k < K; k++) {
            double scale = std::max(std::fabs(a[k]), std::fabs(b[k]));
            if (std::fabs(a[k] - sign * b[k]) > kTol * scale) return false;
        }
        return true;
    };
    auto find_rep = [&](const Vals& v, double* sign) -> int {
        uint32_t k0 = key_of(v[0]);
        for (uint32_t k : {k0, k0 + 1, k0 - 1}) {
            auto it = buckets.find(k);
            if (it == buckets.end()) continue;
            for (const Rep& r : it->second) {
                if (close(v, r.v, 1.0)) { *sign = 1.0; return r.id; }
                if (close(v, r.v, -1.0)) { *sign = -1.0; return r.id; }
            }
        }
        return -1;
    };
    std::unordered_map<int, int> repl;          // old id -> new id
    for (int x : reach) {
        const Node n = nodes_[x];
        Vals v{};
        if (n.op == Op::Const) {
            v.fill(n.val);
            val[x] = v;
            repl[x] = x;
            continue;                            // constants are never representatives: make() folds them
        }
        if (n.op == Op::Var || n.op == Op::Param) {
            auto pit = params.find(n.name);
            for (int k = 0; k < K; k++) {
                if (n.op == Op::Param) {
                    double base = pit == params.end() ? 0.0 : pit->second;
                    v[k] = base == 0.0 ? 0.0 : base * (0.8 + 0.45 * rnd());      // a parameter that is exactly zero stays zero
                } else {
                    double mag = 0.6 + 1.9 * rnd();
                    v[k] = (k >= K / 2 && rnd() < 0.5) ? -mag : mag;              // the first half all positive, the second with both signs
                }
            }
            val[x] = v;
            repl[x] = x;
            buckets[key_of(v[0])].push_back({x, v});
            continue;
        }
        const int a = n.a >= 0 ? repl.at(n.a) : -1, b = n.b >= 0 ? repl.at(n.b) : -1, c = n.c >= 0 ? repl.at(n.c) : -1;
        int nx = (a == n.a && b == n.b && c == n.c) ? x : make(n.op, a, b, c);
        if (is_atom_op(n.op)) {
            // not analytic: an independent random atom (one per rebuilt node, so identical rebuilt nodes share it)
            auto known = val.find(nx);
            if (known != val.end() && nx != x) v = known->second;
            else for (int k = 0; k < K; k++) v[k] = 0.5 + 1.5 * rnd();
            val[x] = v;
            val[nx] = v;
            repl[x] = nx;
            continue;
        }
        const Vals* va = n.a >= 0 ? &val.at(n.a) : nullptr;
        const Vals* vb = n.b >= 0 ? &val.at(n.b) : nullptr;
        for (int k = 0; k < K; k++) {
            switch (n.op) {
                case Op::Add: v[k] = (*va)[k] + (*vb)[k]; break;
                case Op::Sub: v[k] = (*va)[k] - (*vb)[k]; break;
                case Op::Mul: v[k] = (*va)[k] * (*vb)[k]; break;
                case Op::Div: v[k] = (*va)[k] / (*vb)[k]; break;
                case Op::Pow: v[k] = fold2(n.op, (*va)[k], (*vb)[k]); break;
                default: v[k] = fold1(n.op, (*va)[k]); break;
            }
        }
        val[x] = v;
        if (!finite(v)) { repl[x] = nx; val[nx] = v; continue; }     // outside the sampled domain somewhere: leave it alone
        if (nodes_[nx].op == Op::Const || nodes_[nx].op == Op::Var || nodes_[nx].op == Op::Param) { repl[x] = nx; continue; }
        // a sum that cancels to rounding everywhere is zero
        if (n.op == Op::Add || n.op == Op::Sub) {
            bool zero = true;
            for (int k = 0; k < K; k++)
                if (std::fabs(v[k]) > 1e-11 * (std::fabs((*va)[k]) + std::fabs((*vb)[k]))) { zero = false; break; }
            if (zero) { repl[x] = constant(0.0); continue; }
        }
        // the same number at every sample: a constant
        {
            double lo = v[0], hi = v[0];
            for (double t : v) { lo = std::min(lo, t); hi = std::max(hi, t); }
            double mean = 0.5 * (lo + hi);
            if (hi - lo <= kTol * std::fabs(mean) && mean != 0.0) {
                char buf[64];
                snprintf(buf, sizeof buf, "%.11g", mean);        // snap: 0.99999999999999978 -> 1
                repl[x] = constant(strtod(buf, nullptr));
                continue;
            }
        }
        double sign = 1.0;
        int rep = find_rep(v, &sign);
        if (rep >= 0 && rep != nx) {
            repl[x] = sign > 0 ? rep : make(Op::Neg, rep);
            continue;
        }
        repl[x] = nx;
        val[nx] = v;
        if (rep < 0) buckets[key_of(v[0])].push_back({nx, v});
    }
    std::vector<int> res;
    for (int o : outs) res.push_back(repl.at(o));
    return res;
}

double Graph::eval(int id, const std::map<std::string, double>& env) const {
    std::map<int, double> memo;
    std::function<double(int)> go = [&](int x) -> double {
        auto it = memo.find(x);
        if (it != memo.end()) return it->second;
        const Node& n = nodes_[x];
        double r = 0.0;
        switch (n.op) {
            case Op::Const: r = n.val; break;
            case Op::Var: case Op::Param: { auto e = env.find(n.name); r = e == env.end() ? 0.0 : e->second; break; }
            case Op::Add: r = go(n.a) + go(n.b); break;
            case Op::Sub: r = go(n.a) - go(n.b); break;
            case Op::Mul: r = go(n.a) * go(n.b); break;
            case Op::Div: r = go(n.a) / go(n.b); break;
            case Op::Neg: r = -go(n.a); break;
            case Op::Select: r = go(n.a) != 0.0 ? go(n.b) : go(n.c); break;
            case Op::CsqrtRe: case Op::CsqrtIm: {
                std::complex<double> s = std::sqrt(std::complex<double>(go(n.a), go(n.b)));
                r = n.op == Op::CsqrtRe ? s.real() : s.imag();
                break;
            }
            case Op::Atan2: case Op::Pow: case Op::Fmod: case Op::Min: case Op::Max:
            case Op::Lt: case Op::Lte: case Op::Eq: case Op::Gt: case Op::Gte:
                r = fold2(n.op, go(n.a), go(n.b)); break;
            default: r = fold1(n.op, go(n.a)); break;
        }
        memo[x] = r;
        return r;
    };
    return go(id);
}

static const char* fn_name(Op op) {
    switch (op) {
        case Op::Sin: return "sin"; case Op::Cos: return "cos"; case Op::Tan: return "tan";
        case Op::Asin: return "asin"; case Op::Acos: return "acos"; case Op::Atan: return "atan";
        case Op::Atan2: return "atan2"; case Op::Fabs: return "fabs"; case Op::Log: return "log";
        case Op::Sqrt: return "sqrt"; case Op::Exp: return "exp"; case Op::Sinh: return "sinh";
        case Op::Cosh: return "cosh"; case Op::Tanh: return "tanh"; case Op::Pow: return "pow";
        case Op::Fmod: return "fmod"; case Op::Sign: return "sign"; case Op::Min: return "min";
        case Op::Max: return "max"; case Op::Select: return "select";
        case Op::CsqrtRe: return "csqrt_re"; case Op::CsqrtIm: return "csqrt_im";
        default: return "?";
    }
}

std::string Graph::to_string(int id) const {
    const Node& n = nodes_[id];
    auto S = [&](int x) { return to_string(x); };
    char buf[64];
    switch (n.op) {
        case Op::Const: snprintf(buf, sizeof buf, "%.9g", n.val); return buf;
        case Op::Var: return n.name;
        case Op::Param: return "cfg->" + n.name;
        case Op::Add: return "(" + S(n.a) + "+" + S(n.b) + ")";
        case Op::Sub: return "(" + S(n.a) + "-" + S(n.b) + ")";
        case Op::Mul: return "(" + S(n.a) + "*" + S(n.b) + ")";
        case Op::Div: return "(" + S(n.a) + "/" + S(n.b) + ")";
        case Op::Neg: return "(-" + S(n.a) + ")";
        case Op::Lt: return "(" + S(n.a) + "<" + S(n.b) + ")";
        case Op::Lte: return "(" + S(n.a) + "<=" + S(n.b) + ")";
        case Op::Eq: return "(" + S(n.a) + "==" + S(n.b) + ")";
        case Op::Gt: return "(" + S(n.a) + ">" + S(n.b) + ")";
        case Op::Gte: return "(" + S(n.a) + ">=" + S(n.b) + ")";
        default: {
            std::string s = std::string(fn_name(n.op)) + "(" + S(n.a);
            if (n.b >= 0) s += "," + S(n.b);
            if (n.c >= 0) s += "," + S(n.c);
            return s + ")";
        }
    }
}

std::string float_literal(double v) {
    if (std::isnan(v)) return "GRT_NAN";
    if (std::isinf(v)) return v > 0 ? "GRT_INF" : "(-GRT_INF)";
    char buf[64];
    snprintf(buf, sizeof buf, "%.9g", (double)(float)v);
    std::string s = buf;
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos && s.find("inf") == std::string::npos)
        s += ".0";
    return s + "f";
}

std::string emit_block(const Graph& g, const std::vector<std::pair<std::string, int>>& outs,
                       const EmitOptions& opt) {
    // topological order of reachable op nodes (ids are already topological: children < parents)
    std::set<int> reach;
    std::function<void(int)> go = [&](int x) {
        if (x < 0 || !reach.insert(x).second) return;
        if (opt.leaf_names.count(x)) return;   // a named leaf: nothing below it is needed
        const Node& n = g.node(x);
        go(n.a); go(n.b); go(n.c);
    };
    for (auto& o : outs) go(o.second);

    auto ref = [&](int x) -> std::string {
        auto leaf = opt.leaf_names.find(x);
        if (leaf != opt.leaf_names.end()) return leaf->second;
        const Node& n = g.node(x);
        switch (n.op) {
            case Op::Const: return float_literal(n.val);
            case Op::Var: return n.name;
            case Op::Param: {
                auto it = opt.param_values.find(n.name);
                if (it != opt.param_values.end()) return float_literal(it->second);
                return opt.param_prefix + n.name;
            }
            default: return opt.tmp_prefix + std::to_string(x);
        }
    };

    std::string out;
    for (int x : reach) {  // std::set iterates ascending = topological
        const Node& n = g.node(x);
        if (n.op == Op::Const || n.op == Op::Var || n.op == Op::Param || opt.leaf_names.count(x)) continue;
        std::string rhs;
        auto A = [&]() { return ref(n.a); };
        auto B = [&]() { return ref(n.b); };
        auto Cc = [&]() { return ref(n.c); };
        switch (n.op) {
            case Op::Add: rhs = A() + " + " + B(); break;
            case Op::Sub: rhs = A() + " - " + B(); break;
            case Op::Mul: rhs = A() + " * " + B(); break;
            case Op::Div: rhs = "GRT_DIV(" + A() + ", " + B() + ")"; break;
            case Op::Neg: rhs = "-" + A(); break;
            case Op::Sin: rhs = "GRT_SIN(" + A() + ")"; break;
            case Op::Cos: rhs = "GRT_COS(" + A() + ")"; break;
            case Op::Tan: rhs = "GRT_TAN(" + A() + ")"; break;
            case Op::Asin: rhs = "asinf(" + A() + ")"; break;
            case Op::Acos: rhs = "acosf(" + A() + ")"; break;
            case Op::Atan: rhs = "atanf(" + A() + ")"; break;
            case Op::Atan2: rhs = "atan2f(" + A() + ", " + B() + ")"; break;
            case Op::Fabs: rhs = "fabsf(" + A() + ")"; break;
            case Op::Log: rhs = "GRT_LOG(" + A() + ")"; break;
            case Op::Sqrt: rhs = "GRT_SQRT(" + A() + ")"; break;
            case Op::Exp: rhs = "GRT_EXP(" + A() + ")"; break;
            case Op::Sinh: rhs = "sinhf(" + A() + ")"; break;
            case Op::Cosh: rhs = "coshf(" + A() + ")"; break;
            case Op::Tanh: rhs = "GRT_TANH(" + A() + ")"; break;
            case Op::Pow: rhs = "GRT_POW(" + A() + ", " + B() + ")"; break;
            case Op::Select: rhs = "((" + A() + " != 0.0f) ? " + B() + " : " + Cc() + ")"; break;
            case Op::Lt: rhs = "((" + A() + " < " + B() + ") ? 1.0f : 0.0f)"; break;
            case Op::Lte: rhs = "((" + A() + " <= " + B() + ") ? 1.0f : 0.0f)"; break;
            case Op::Eq: rhs = "((" + A() + " == " + B() + ") ? 1.0f : 0.0f)"; break;
            case Op::Gt: rhs = "((" + A() + " > " + B() + ") ? 1.0f : 0.0f)"; break;
            case Op::Gte: rhs = "((" + A() + " >= " + B() + ") ? 1.0f : 0.0f)"; break;
            case Op::Fmod: rhs = "fmodf(" + A() + ", " + B() + ")"; break;
            case Op::Sign: rhs = "GRT_SIGN(" + A() + ")"; break;
            case Op::Min: rhs = "fminf(" + A() + ", " + B() + ")"; break;
            case Op::Max: rhs = "fmaxf(" + A() + ", " + B() + ")"; break;
            case Op::CsqrtRe: rhs = "grt_csqrt_re(" + A() + ", " + B() + ")"; break;
            case Op::CsqrtIm: rhs = "grt_csqrt_im(" + A() + ", " + B() + ")"; break;
            default: throw std::runtime_error("sym: emit bad op");
        }
        out += opt.indent + "const float " + opt.tmp_prefix + std::to_string(x) + " = " + rhs + ";\n";
    }
    for (auto& o : outs) out += opt.indent + o.first + " = " + ref(o.second) + ";\n";
    return out;
}

}  // namespace sym
