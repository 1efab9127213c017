"""The script engine (csrc/minijs.cpp) through the plugin boundary: the JavaScript the reference's scripts rely on
(number.js:4-81 operator overloading, function-scoped `var`, sparse arrays with `.length`, closures, Math.*)."""
import os
import textwrap

import numpy as np
import pytest

from test_loader import write_folder


def m_path(tmp_path):
    """The probe script metric_diag wrote (test_loader.write_folder puts it into <tmp>/content)."""
    for root, _, files in os.walk(str(tmp_path)):
        if "probe.js" in files:
            return os.path.join(root, "probe.js")
    raise FileNotFoundError("probe.js")


def metric_diag(grt, oracle_mod, tmp_path, js, pos=(0.3, 1.5, 0.7, 0.4)):
    folder = write_folder(tmp_path, js)
    m = grt.Content(folder).load(os.path.join(folder, "probe.js"))
    g, _ = oracle_mod.for_metric(m).metric_eval(pos)
    return g.reshape(4, 4), m


def test_var_hoisting_shadowing_and_loops(grt, oracle_mod, tmp_path):
    # kerr_schild.js re-declares `var a` as a loop counter after using it as a parameter (function scope, hoisted)
    g, _ = metric_diag(grt, oracle_mod, tmp_path, """
        function f(t, x, y, z) {
            var a = x * 2;
            var keep = a;
            var acc = 0;
            for (var a = 0; a < 4; a++) { for (var b = 0; b < 3; b++) { if (b == 1) continue; acc += a * b; } }
            var i = 0; while (true) { i++; if (i >= 5) break; }
            do { i += 10; } while (i < 30);
            return [-(1 + 0 * x), keep, acc + 0 * x, i + a + 0 * x];
        }
        f
        """)
    assert g[0, 0] == -1 and np.isclose(g[1, 1], 3.0) and g[2, 2] == 12 and g[3, 3] == 39


def test_arrays_closures_and_arrows(grt, oracle_mod, tmp_path):
    g, _ = metric_diag(grt, oracle_mod, tmp_path, """
        const scale = (k) => (v) => v * k;
        function f(t, x, y, z) {
            let xs = [x, y, z].map(scale(2));
            let arr = []; arr.length = 16; arr[5] = xs[0]; arr.push(7);
            const total = [1, 2, 3, 4].reduce((s, v) => s + v, 0);
            const obj = {p: xs[1], q: {r: xs[2]}};
            const pick = arr[3] === undefined ? 1 : 0;
            return [-(total + pick - 10), arr[5], obj.p + (arr.length - 17), obj.q.r * [4, 5, 6].indexOf(5) + Math.max(1, 2) - Math.sqrt(4)];
        }
        f
        """)
    assert g[0, 0] == -1 and np.isclose(g[1, 1], 3.0) and np.isclose(g[2, 2], 1.4) and np.isclose(g[3, 3], 0.8)


def test_operator_overloading_both_ways_and_math(grt, oracle_mod, tmp_path):
    g, _ = metric_diag(grt, oracle_mod, tmp_path, """
        function f(t, x, y, z) {
            var a = 2 - x;           // Number on the left
            var b = x - 2;           // number on the left
            var c = -x / 3 + 3 / x;  // unary minus, division both ways
            var d = x * x * x - CMath.pow(x, 3) + Math.PI - M_PI + (1 + 2) * 3 % 4;
            var e = (Math.abs(-2) > 1 ? 1 : 0);
            return [-(a + b + 1), c, 1 + d, 1 + e];
        }
        f
        """)
    x = 1.5
    assert g[0, 0] == -1 and np.isclose(g[1, 1], -x / 3 + 3 / x, rtol=1e-6) and np.isclose(g[2, 2], 2.0, atol=1e-5)


def test_comparison_of_number_objects_is_refused(grt, tmp_path):
    folder = write_folder(tmp_path, "function f(t, x, y, z) { if (x < y) { return [-1, 1, 1, 1]; } return [-1, 1, 1, 2]; }\nf")
    with pytest.raises(grt.GrtError) as e:
        grt.Content(folder).load(os.path.join(folder, "probe.js"))
    assert "select" in e.value.message     # the error points at CMath.lt/select, the differentiable branch of the plugin API


def test_runaway_script_is_stopped(grt, tmp_path):
    folder = write_folder(tmp_path, "function f(t, x, y, z) { while (true) {} }\nf")
    with pytest.raises(grt.GrtError):
        grt.Content(folder).load(os.path.join(folder, "probe.js"))


# ---- the ES2015+ surface QuickJS 2020-04-12 accepts and plugin authors may use -------------------------------------------------
# every script returns a diagonal metric whose entries are built with the construct under test; expected values are computed by hand,
# and where the reference is mounted the same script runs in the reference's own QuickJS (oracle/ref_quickjs.cpp) and must agree.
MODERN_SCRIPTS = {
    "template literals": ("""
        function f(t, x, y, z) {
            const n = 3, label = `k${n}-${1 + 1}`;
            const tbl = {"k3-2": 2.5};
            const multi = `a
b`;
            return [-1, tbl[label] * x, multi.length + 0 * y, `${n}${n}` * 1 + 0 * z];
        }
        f""", [-1, 2.5 * 1.5, 3, 33]),
    "default and rest parameters": ("""
        function scale(v, k = 2, m = k + 1) { return v * k + m; }
        const total = (first, ...rest) => rest.reduce((s, q) => s + q, first);
        function f(t, x, y, z) {
            return [-1, scale(x), scale(x, 3), total(1, 2, 3, 4) + scale(0, undefined, 5) + 0 * z];
        }
        f""", [-1, 1.5 * 2 + 3, 1.5 * 3 + 4, 15]),
    "destructuring declarations": ("""
        function f(t, x, y, z) {
            const [a, b = 7, , c, ...tail] = [x, undefined, 99, y, 1, 2, 3];
            const {p, q: {r, s = 4}, w = 10} = {p: 2, q: {r: 3}};
            let [[m, n], {k}] = [[1, 2], {k: 5}];
            return [-1, a + b, c * p + r + s + w, tail.length + m + n + k + 0 * z];
        }
        f""", [-1, 1.5 + 7, 0.7 * 2 + 3 + 4 + 10, 3 + 1 + 2 + 5]),
    "destructuring parameters and for-of": ("""
        function norm2([u, v], {weight = 2} = {}) { return (u * u + v * v) * weight; }
        function f(t, x, y, z) {
            let acc = 0;
            for (const [k, v] of [[1, x], [2, y]]) acc = acc + k * v;
            let keys = 0;
            for (const {id} of [{id: 3}, {id: 4}]) keys += id;
            return [-1, norm2([x, y]), acc, keys + norm2([1, 1], {weight: 0.5}) + 0 * z];
        }
        f""", [-1, (1.5 ** 2 + 0.7 ** 2) * 2, 1.5 + 2 * 0.7, 7 + 1]),
    "spread": ("""
        function add3(a, b, c) { return a + b + c; }
        function f(t, x, y, z) {
            const base = [x, y];
            const more = [...base, z, ...[1, 2]];
            const o = {...{a: 1, b: 2}, b: 3, ...{c: 4}};
            return [-1, add3(...base, 1), more.length + more[2], o.a + o.b + o.c + Math.max(...[1, 5, 2]) + 0 * z];
        }
        f""", [-1, 1.5 + 0.7 + 1, 5 + 0.4, 1 + 3 + 4 + 5]),
    "switch": ("""
        function kind(n) {
            let r = 0;
            switch (n) {
                case 1: r += 1;
                case 2: r += 10; break;
                case "2": r += 100; break;
                default: r += 1000;
                case 3: r += 5;
            }
            return r;
        }
        function f(t, x, y, z) { return [-1, kind(1) + 0 * x, kind(2) + kind("2") + 0 * y, kind(7) + kind(3) + 0 * z]; }
        f""", [-1, 11, 110, 1005 + 5]),
    "classes": ("""
        class Vec {
            constructor(a, b) { this.a = a; this.b = b; }
            dot(o) { return this.a * o.a + this.b * o.b; }
            get norm2() { return this.dot(this); }
            set first(v) { this.a = v * 2; }
            static unit() { return new Vec(1, 0); }
        }
        class Weighted extends Vec {
            constructor(a, b, w) { super(a, b); this.w = w; }
            dot(o) { return super.dot(o) * this.w; }
        }
        class Plain extends Vec {}
        function f(t, x, y, z) {
            const v = new Vec(x, y), u = new Weighted(x, y, 2), q = new Plain(3, 4);
            const e = Vec.unit();
            e.first = 4;
            const kinds = (u instanceof Vec ? 1 : 0) + (u instanceof Weighted ? 10 : 0) + (v instanceof Weighted ? 100 : 0) + (q instanceof Vec ? 1000 : 0);
            return [-1, v.norm2, u.dot(v), q.norm2 + e.a + kinds + 0 * z];
        }
        f""", [-1, 1.5 ** 2 + 0.7 ** 2, (1.5 ** 2 + 0.7 ** 2) * 2, 25 + 8 + 1011]),
    "class expression and closures": ("""
        const make = (k) => class { constructor(v) { this.v = v; } times() { return this.v * k; } };
        const Twice = make(2), Thrice = make(3);
        function f(t, x, y, z) { return [-1, new Twice(x).times(), new Thrice(y).times(), (typeof Twice === "function" ? 1 : 0) + 0 * z]; }
        f""", [-1, 3.0, 2.1, 1]),
    "property enumeration order": ("""
        function f(t, x, y, z) {
            const o = {r: 1, theta: 2, phi: 3, 10: 4, 2: 5};
            o.alpha = 6;
            let code = 0, m = 1;
            for (const k in o) { code += m * o[k]; m *= 10; }
            let code2 = 0; m = 1;
            for (const k of Object.keys({...o, beta: 7})) { code2 += m * (o[k] === undefined ? 7 : o[k]); m *= 10; }
            const {theta, ...others} = o;
            return [-1, code + 0 * x, code2 + 0 * y, Object.keys(others).join("") === "210rphialpha" ? 1 : 0];
        }
        f""", [-1, 632145, 7632145, 1]),
    "computed property names": ("""
        function f(t, x, y, z) {
            const k = "ab", i = 2;
            const o = {[k]: x, ["c" + i]: y, [i * 2]: 7, [`m${i}`](v) { return v * 3; }, plain: 1};
            return [-1, o.ab + o.c2, o[4] + o.plain + 0 * y, o.m2(z)];
        }
        f""", [-1, 1.5 + 0.7, 8, 1.2]),
    "class fields": ("""
        class Base { scale = 2; offset; static count = 3; static twice = this.count * 2; constructor(v) { this.v = v * this.scale; } }
        class Child extends Base { extra = this.v + 1; constructor(v) { super(v); this.sum = this.extra + this.scale; } }
        class Quiet extends Base { tag = 5; }
        class Bare { w = 4; }
        function f(t, x, y, z) {
            const c = new Child(x), q = new Quiet(y), b = new Bare();
            return [-1, c.v, c.sum, q.tag + q.v + b.w + Base.count + Base.twice + (c.offset === undefined ? 1 : 0) + 0 * z];
        }
        f""", [-1, 3.0, 3.0 + 1 + 2, 5 + 1.4 + 4 + 3 + 6 + 1]),
    "labelled statements": ("""
        function f(t, x, y, z) {
            let hits = 0, rows = 0;
            outer: for (let i = 0; i < 4; i++) {
                rows++;
                for (let j = 0; j < 4; j++) {
                    if (j == 2) continue outer;
                    if (i == 3) break outer;
                    hits++;
                }
            }
            let w = 0;
            blk: { w = 1; if (w > 0) break blk; w = 100; }
            let n = 0, k = 0;
            scan: while (true) { k++; switch (k) { case 3: break scan; default: n += k; } }
            return [-1, hits + 0 * x, rows + w + 0 * y, n + k + 0 * z];
        }
        f""", [-1, 6, 5, 6]),
    "destructuring assignment": ("""
        function f(t, x, y, z) {
            let a = x, b = y, c, d, rest, o = {}, arr = [0, 0];
            [a, b] = [b, a];
            ({c, d = 7, ...rest} = {c: 2, e: 5, g: 6});
            [o.p, arr[1], [c] = [9]] = [3, 4];
            const r = ([a, b] = [a + 1, b + 1]);
            let m, n2;
            ({m, q: {n2}} = {m: 1, q: {n2: 2}});
            return [-1, a, b + d + rest.e + rest.g + r.length, o.p + arr[1] + c + m + n2 + 0 * z];
        }
        f""", [-1, 0.7 + 1, 1.5 + 1 + 7 + 5 + 6 + 2, 3 + 4 + 9 + 1 + 2]),
    "object rest in declarations": ("""
        function pick({keep, ...others}) { return keep + Object.keys(others).length; }
        function f(t, x, y, z) {
            const {a, b: {c, ...inner}, ...outer} = {a: x, b: {c: 1, d: 2, e: 3}, f: 4, g: 5};
            return [-1, a + c, inner.d + inner.e + outer.f + outer.g + 0 * y, pick({keep: z, u: 1, v: 2})];
        }
        f""", [-1, 2.5, 14, 2.4]),
}


@pytest.mark.parametrize("what", sorted(MODERN_SCRIPTS))
def test_modern_javascript_constructs(grt, oracle_mod, have_reference, tmp_path, what):
    js, want = MODERN_SCRIPTS[what]
    pos = (0.3, 1.5, 0.7, 0.4)
    g, m = metric_diag(grt, oracle_mod, tmp_path, js, pos)
    assert np.allclose(np.diag(g), want, rtol=1e-6), (what, np.diag(g), want)
    if have_reference:
        try:
            q = oracle_mod.QuickJSReference()
        except FileNotFoundError:
            return
        qg, qdg, _, _ = q.metric(m_path(tmp_path), np.array(pos, np.float64))
        sym = m.eval_symbolic(pos)
        assert np.abs(sym["g"] - qg).max() <= 1e-9 * max(1.0, np.abs(qg).max()), what
        assert np.abs(sym["dg"] - qdg).max() <= 1e-8 * max(1.0, np.abs(qdg).max()), what


REFUSED = {
    "generator function": ("function* gen() { yield 1; }\nfunction f(t,x,y,z){ return [-1,1,1,1]; }\nf", "generator"),
    "async function": ("async function g() { return 1; }\nfunction f(t,x,y,z){ return [-1,1,1,1]; }\nf", "async"),
    "tagged template": ("function tag(s) { return 1; }\nfunction f(t,x,y,z){ return [-1, tag`a`, 1, 1]; }\nf", "tagged template"),
    "private class member": ("class A { #n = 1; }\nfunction f(t,x,y,z){ return [-1,1,1,1]; }\nf", "unexpected character '#'"),
    "shorthand default outside a pattern": ("function f(t,x,y,z){ var o = {q = 1}; return [-1,1,1,1]; }\nf", "shorthand property"),
    "field of a derived class without a top-level super": ("class A {}\nclass B extends A { n = 1; constructor() { if (true) { super(); } } }\nfunction f(t,x,y,z){ return [-1,1,1,1]; }\nf", "top-level statement"),
    "unary plus on a number object": ("function f(t,x,y,z){ return [-1, +x, 1, 1]; }\nf", "operator pos"),
    "let redeclares a parameter": ("function f(t,x,y,z){ let x = 1; return [-1,1,1,1]; }\nf", "redefinition of parameter"),
    "regular expression literal": ("function f(t,x,y,z){ var r = /ab+c/.test('abc'); return [-1,1,1,1]; }\nf", "regular expression"),
    "with statement": ("function f(t,x,y,z){ var o = {q: 1}; with (o) { var w = q; } return [-1,1,1,1]; }\nf", "'with'"),
    "new.target": ("function F(){ return new.target; }\nfunction f(t,x,y,z){ return [-1,1,1,1]; }\nf", "new.target"),
    "duplicate let": ("function f(t,x,y,z){ let a = 1; let a = 2; return [-1,1,1,1]; }\nf", "redefinition of lexical identifier"),
    "getter with a parameter": ("function f(t,x,y,z){ var o = { get v(a) { return 1; } }; return [-1,1,1,1]; }\nf", "getter or setter"),
    "unary minus under **": ("function f(t,x,y,z){ return [-1, -2 ** 2, 1, 1]; }\nf", "left-hand side of '**'"),
    "class called without new": ("class A {}\nfunction f(t,x,y,z){ A(); return [-1,1,1,1]; }\nf", "'new'"),
}


@pytest.mark.parametrize("what", sorted(REFUSED))
def test_unsupported_javascript_is_refused_loudly(grt, tmp_path, what):
    """What the interpreter does not implement is a SyntaxError (or TypeError) naming the script and the line - never skipped."""
    js, needle = REFUSED[what]
    folder = write_folder(tmp_path, js)
    with pytest.raises(grt.GrtError) as e:
        grt.Content(folder).load(os.path.join(folder, "probe.js"))
    msg = e.value.message
    assert needle in msg and "probe.js:" in msg, msg


def _probes():
    import json
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "script_engine_probes.json")))


@pytest.mark.parametrize("what", sorted(_probes()))
def test_probe_scripts_agree_with_quickjs(grt, oracle_mod, have_reference, tmp_path, what):
    """Corners of JavaScript a metric script may rely on (tests/golden/make_script_engine_probes.py): the interpreter gives the metric
    the reference's QuickJS gave when the table was recorded - and gives live, where the reference is mounted."""
    body, want = _probes()[what]
    js = "function f(t, x, y, z) {\n" + body + "\n}\nf"
    pos = (0.3, 1.5, 0.7, 0.4)
    folder = write_folder(tmp_path, js)
    m = grt.Content(folder).load(os.path.join(folder, "probe.js"))
    got = np.diag(np.asarray(m.eval_symbolic(pos)["g"]).reshape(4, 4))
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12), (what, got, want)
    if have_reference:
        try:
            q = oracle_mod.QuickJSReference()
        except FileNotFoundError:
            return
        qg, _, _, _ = q.metric(os.path.join(folder, "probe.js"), np.array(pos, np.float64))
        assert np.allclose(got, np.diag(np.asarray(qg).reshape(4, 4)), rtol=1e-9, atol=1e-12), what
