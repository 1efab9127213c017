// Symbolic expression DAG used by the metric plugin loader.
//
// Replaces the reference's missing `deps/vec` layer (20k/funvectormaths @ 7a8b639:
// `value`, `dual`, `dual_complex`, `equivalent()`, `type_to_string()`), which
// /root/reference/js_interop.cpp:219-297,481-616 and /root/reference/metric.hpp:38-274
// drive.  Design differences (deliberate, B200-first):
//   * hash-consed DAG: structurally equal sub-expressions are the same node, so
//     `equivalent()` is an id compare and common sub-expression elimination is free
//     (the reference needs script authors to call `$pin` for that,
//     /root/reference/equation_context.hpp:16-43);
//   * derivatives are taken symbolically on the DAG (memoised reverse walk) instead of
//     re-running the script once per variable with forward-mode duals
//     (/root/reference/metric.hpp:38-80) - same partials, one script run;
//   * code is emitted as straight-line SSA `float` statements for a CUDA `__device__`
//     function (and the identical text for the CPU oracle), not as `-D` macro strings.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace sym {

enum class Op : uint8_t {
    Const, Var, Param,
    Add, Sub, Mul, Div, Neg,
    Sin, Cos, Tan, Asin, Acos, Atan, Atan2,
    Fabs, Log, Sqrt, Exp, Sinh, Cosh, Tanh, Pow,
    Select, Lt, Lte, Eq, Gt, Gte,
    Fmod, Sign, Min, Max,
    CsqrtRe, CsqrtIm   // real / imaginary part of the principal complex square root of (a + i b)
};

struct Node {
    Op op = Op::Const;
    int a = -1, b = -1, c = -1;
    double val = 0.0;     // Const
    std::string name;     // Var / Param
};

class Graph {
public:
    Graph();
    // the graph new Expr constants are created in (one per metric load, single threaded)
    static Graph* cur();
    static void set_cur(Graph* g);

    int constant(double v);
    int var(const std::string& name);
    int param(const std::string& name);
    int make(Op op, int a, int b = -1, int c = -1);  // simplifying constructor

    const Node& node(int id) const { return nodes_[id]; }
    size_t size() const { return nodes_.size(); }
    bool is_const(int id) const { return nodes_[id].op == Op::Const; }
    bool is_const(int id, double v) const { return nodes_[id].op == Op::Const && nodes_[id].val == v; }
    double cval(int id) const { return nodes_[id].val; }

    // d(id)/d(variable node `wrt`)
    int diff(int id, int wrt);
    // Forward-mode derivative along a direction: every variable node in `tangents` carries the given tangent
    // expression (others 0), so with {v_k -> iv_k} this is the directional derivative sum_k iv_k d/dv_k in ONE
    // pass over the DAG (shared sub-expressions get one tangent each) instead of four gradients summed.
    int diff_along(int id, const std::map<int, int>& tangents);
    // Reverse-mode gradient: d(id)/d(wrt[i]) for all i in one adjoint sweep (about three times the cost of the
    // expression itself, whatever the number of variables).
    std::vector<int> gradient(int id, const std::vector<int>& wrt);
    // replace nodes by other nodes (bottom-up rebuild with re-simplification)
    int substitute(int id, const std::map<int, int>& repl);
    // does `id` depend on node `leaf`?
    bool depends_on(int id, int leaf) const;
    // number of arithmetic nodes reachable from the outputs (flop-ish count, FMA not fused)
    size_t op_count(const std::vector<int>& outs, const std::vector<int>* leaves = nullptr) const;   // `leaves`: nodes counted as inputs
    // Estimated machine instructions: like op_count, but a single-use product feeding a sum is one FMA with it, a
    // negation folds into the operand modifier of its consumer, and a quotient costs two (reciprocal + multiply).
    size_t instr_count(const std::vector<int>& outs) const;
    // Algebraic clean-up of an expression: every maximal sum is flattened into signed product terms and
    // common factors are pulled out greedily (a*b*x + a*c*x -> a*x*(b + c)), recursively, bottom-up and
    // memoised so shared sub-expressions stay shared.  Same value up to floating-point reassociation.
    // With `distribute`, a product that contains a short sum is multiplied out first ((c + H l l) v ->
    // c v + H l l v) so that factors hidden inside it can be shared with the neighbouring terms.
    int factor(int id, bool distribute = false);
    // Multivariate polynomial view of an expression.  `to_poly` multiplies sums and products out completely over
    // "atoms" (every node that is not + - * neg, a constant or a small integer power), merging like terms, so that
    // cancellations which only show after expansion - the cofactors and determinant of a metric with product
    // structure, e.g. Kerr-Schild eta + H l l - happen symbolically.  Returns false if more than `max_terms`
    // monomials appear.  `from_poly` rebuilds an expression with common factors pulled out (same greedy scheme as
    // `factor`).  Same value as the input up to floating-point reassociation.
    struct Poly { std::map<std::vector<int>, double> terms; };   // sorted atom ids (repeats = powers) -> coefficient
    bool to_poly(int id, Poly& out, size_t max_terms = 20000);
    int from_poly(const Poly& p);
    static bool poly_mul(const Poly& a, const Poly& b, Poly& out, size_t max_terms = 20000);
    static void poly_axpy(Poly& acc, double alpha, const Poly& x);   // acc += alpha * x
    // Value-based merging of sub-expressions (probabilistic identity testing).  The DAG reachable from `outs` is evaluated
    // in double precision at a handful of random sample points (variables with both signs, parameters perturbed around their
    // current values, every non-analytic node - select, comparisons, fabs, sign, min, max, fmod - replaced by an independent
    // random atom so that what remains is analytic and equality on the samples means equality everywhere).  Then, bottom-up:
    //   * a node whose value equals that of an earlier node at every sample IS that node (differently associated copies of
    //     l.v, a cofactor that equals a minor computed elsewhere ...), a node equal to minus an earlier node is its negation;
    //   * a node that is the same number at every sample is that constant (|l|^2 = 1 for the null vector of a Kerr-Schild form);
    //   * a sum that cancels to rounding at every sample is zero.
    // Hash-consing finds structurally equal sub-expressions; this finds the mathematically equal ones that no local rewrite
    // rule reaches.  Same value as the input up to floating-point reassociation; returns the new ids of `outs`.
    // `params`: current parameter values (name -> value).  Deterministic (fixed seed).
    std::vector<int> unify(const std::vector<int>& outs, const std::map<std::string, double>& params);
    // numeric value of an expression in double precision; variables / parameters by name (missing -> 0)
    double eval(int id, const std::map<std::string, double>& env) const;
    // human readable infix (tests, debugging, classification decisions)
    std::string to_string(int id) const;

private:
    int intern(const Node& n);
    int simplify(Op op, int a, int b, int c, bool& done);
    std::vector<Node> nodes_;
    std::unordered_map<std::string, int> table_;
    int diff_key(int id, int key);
    std::map<std::pair<int, int>, int> diff_memo_;
    std::vector<std::map<int, int>> tangent_sets_;   // diff_along directions; key = -(index + 2)
    std::unordered_map<int, int> factor_memo_[2];
    std::map<int, Poly> poly_memo_;
};

// value-semantics handle
struct Expr {
    int id = 0;
    Expr() : id(Graph::cur()->constant(0.0)) {}
    Expr(double v) : id(Graph::cur()->constant(v)) {}
    Expr(int v) : id(Graph::cur()->constant((double)v)) {}
    static Expr from_id(int i) { Expr e(0.0); e.id = i; return e; }
    static Expr var(const std::string& n) { return from_id(Graph::cur()->var(n)); }
    static Expr param(const std::string& n) { return from_id(Graph::cur()->param(n)); }
    bool is_const() const { return Graph::cur()->is_const(id); }
    bool is_const(double v) const { return Graph::cur()->is_const(id, v); }
    double cval() const { return Graph::cur()->cval(id); }
    Expr diff(const Expr& wrt) const { return from_id(Graph::cur()->diff(id, wrt.id)); }
    std::string str() const { return Graph::cur()->to_string(id); }
    bool same(const Expr& o) const { return id == o.id; }
};

inline Expr mk(Op op, const Expr& a) { return Expr::from_id(Graph::cur()->make(op, a.id)); }
inline Expr mk(Op op, const Expr& a, const Expr& b) { return Expr::from_id(Graph::cur()->make(op, a.id, b.id)); }
inline Expr mk(Op op, const Expr& a, const Expr& b, const Expr& c) { return Expr::from_id(Graph::cur()->make(op, a.id, b.id, c.id)); }

inline Expr operator+(const Expr& a, const Expr& b) { return mk(Op::Add, a, b); }
inline Expr operator-(const Expr& a, const Expr& b) { return mk(Op::Sub, a, b); }
inline Expr operator*(const Expr& a, const Expr& b) { return mk(Op::Mul, a, b); }
inline Expr operator/(const Expr& a, const Expr& b) { return mk(Op::Div, a, b); }
inline Expr operator-(const Expr& a) { return mk(Op::Neg, a); }
inline Expr& operator+=(Expr& a, const Expr& b) { a = a + b; return a; }
inline Expr& operator-=(Expr& a, const Expr& b) { a = a - b; return a; }

inline Expr sin(const Expr& a) { return mk(Op::Sin, a); }
inline Expr cos(const Expr& a) { return mk(Op::Cos, a); }
inline Expr tan(const Expr& a) { return mk(Op::Tan, a); }
inline Expr asin(const Expr& a) { return mk(Op::Asin, a); }
inline Expr acos(const Expr& a) { return mk(Op::Acos, a); }
inline Expr atan(const Expr& a) { return mk(Op::Atan, a); }
inline Expr atan2(const Expr& y, const Expr& x) { return mk(Op::Atan2, y, x); }
inline Expr fabs(const Expr& a) { return mk(Op::Fabs, a); }
inline Expr log(const Expr& a) { return mk(Op::Log, a); }
inline Expr sqrt(const Expr& a) { return mk(Op::Sqrt, a); }
inline Expr exp(const Expr& a) { return mk(Op::Exp, a); }
inline Expr sinh(const Expr& a) { return mk(Op::Sinh, a); }
inline Expr cosh(const Expr& a) { return mk(Op::Cosh, a); }
inline Expr tanh(const Expr& a) { return mk(Op::Tanh, a); }
inline Expr pow(const Expr& a, const Expr& b) { return mk(Op::Pow, a, b); }
inline Expr select(const Expr& c, const Expr& t, const Expr& f) { return mk(Op::Select, c, t, f); }
inline Expr fmod(const Expr& a, const Expr& b) { return mk(Op::Fmod, a, b); }
inline Expr sign(const Expr& a) { return mk(Op::Sign, a); }
inline Expr min(const Expr& a, const Expr& b) { return mk(Op::Min, a, b); }
inline Expr max(const Expr& a, const Expr& b) { return mk(Op::Max, a, b); }
inline Expr csqrt_re(const Expr& re, const Expr& im) { return mk(Op::CsqrtRe, re, im); }
inline Expr csqrt_im(const Expr& re, const Expr& im) { return mk(Op::CsqrtIm, re, im); }
inline Expr lt(const Expr& a, const Expr& b) { return mk(Op::Lt, a, b); }
inline Expr lte(const Expr& a, const Expr& b) { return mk(Op::Lte, a, b); }
inline Expr eq(const Expr& a, const Expr& b) { return mk(Op::Eq, a, b); }
inline Expr gt(const Expr& a, const Expr& b) { return mk(Op::Gt, a, b); }
inline Expr gte(const Expr& a, const Expr& b) { return mk(Op::Gte, a, b); }

// ---- code emission -------------------------------------------------------------------------
struct EmitOptions {
    // Param nodes are written as `<param_prefix><name>` (dynamic program) unless a literal is
    // given in `param_values` (static program: /root/reference/metric.hpp:494-547 substitutes
    // `cfg->x` by a literal).
    std::string param_prefix = "cfg.";
    std::map<std::string, double> param_values;
    // nodes emitted as a given text instead of being computed (derived parameters: sub-expressions of the parameters only,
    // evaluated once on the host and handed to the kernels in the parameter block)
    std::map<int, std::string> leaf_names;
    std::string tmp_prefix = "t";
    std::string indent = "    ";
};

// Emit straight-line SSA statements computing every `outs[i].second` into the lvalue
// `outs[i].first`.  Shared sub-expressions are computed once.
std::string emit_block(const Graph& g, const std::vector<std::pair<std::string, int>>& outs,
                       const EmitOptions& opt);
std::string float_literal(double v);

}  // namespace sym
