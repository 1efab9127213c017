"""Differential test of the script engine: randomly generated metric scripts run in the product's interpreter + symbolic layer
(csrc/minijs.cpp, csrc/sym.cpp) and in the REFERENCE'S OWN engine (vendored QuickJS + number.js with numeric duals,
oracle/ref_quickjs.cpp); values and all first derivatives must agree to 1e-9 in float64.  Needs /root/reference."""
import os

import numpy as np
import pytest

from test_loader import write_folder


class Gen:
    """Seeded generator of JavaScript in the dialect the reference's scripts use (ES5 functions and `var`, and the ES2015+ constructs the interpreter implements)."""

    UNARY = ["CMath.sin", "CMath.cos", "CMath.exp", "CMath.tanh", "CMath.sinh", "CMath.cosh", "CMath.atan", "CMath.fabs"]

    def __init__(self, seed, force_body=None):
        self.rng = np.random.default_rng(seed)
        self.tmp = 0
        self.force_body = force_body     # index of the statement block every script must contain (None: random)

    def pick(self, xs):
        return xs[int(self.rng.integers(0, len(xs)))]

    def num(self):
        return self.pick(["2", "0.5", "3", "1.25", "Math.PI", "M_PI", "(1 + 2)", "Math.sqrt(2)", "Math.pow(2, 3)", "(7 % 4)", "Math.max(1.5, 0.25)", "1e-1"])

    def expr(self, names, depth):
        r = self.rng.random()
        if depth <= 0 or r < 0.18:
            return self.pick(names) if self.rng.random() < 0.75 else self.num()
        a, b = self.expr(names, depth - 1), self.expr(names, depth - 1)
        if r < 0.34:
            return f"({a} + {b})"
        if r < 0.46:
            return f"({a} - {b})"
        if r < 0.62:
            return f"({a} * {b})"
        if r < 0.70:
            return f"({a} / (2 + {b} * {b}))"            # denominators bounded away from zero
        if r < 0.76:
            return f"(-{a})"
        if r < 0.88:
            return f"{self.pick(self.UNARY)}({a})"
        if r < 0.91:
            return f"CMath.sqrt(1 + {a} * {a})"
        if r < 0.94:
            return f"CMath.pow({self.pick(names)} * {self.pick(names)} + 1, {self.pick(['2', '3', '0.5', '1.5', '-1'])})"
        if r < 0.96:
            return f"CMath.atan2({a}, 1 + {b} * {b})"
        if r < 0.98:
            return f"CMath.select(CMath.lt({self.pick(names)}, {self.num()}), {a}, {b})"
        return f"CMath.fast_length({a}, {b}, {self.num()})"

    def statement_block(self, names):
        """An immediately-invoked function exercising statements: compound assignment, while, ternaries on plain numbers,
        nested functions and closures, object and array literals, let/const."""
        a, b, c = self.expr(names, 2), self.expr(names, 2), self.expr(names, 1)
        n = int(self.rng.integers(2, 5))
        bodies = [
            f"var acc = {a}; var k = 0; while (k < {n}) {{ acc += {b} * (k % 2 == 0 ? 1 : -0.5); k++; }} acc *= 0.5; acc -= {c}; return acc;",
            f"function inner(u) {{ return u * u + {c}; }} var o = {{p: {a}, q: [{b}, {c}, 3]}}; return inner(o.p) + o.q[0] * o.q.length - o.q[2];",
            f"let total = 0; const parts = [{a}, {b}, {c}]; for (let i = 0; i < parts.length; i++) {{ if (i == 1) continue; total = total + parts[i] / (i + 1); }} return total;",
            f"var mk = function(s) {{ return function(v) {{ return v * s + {c}; }}; }}; var f2 = mk(2), f3 = mk(-3); return f2({a}) + f3({b});",
            f"var r = {a}; var flag = (Math.floor(2.7) == 2 && !(1 > 2)) || false; if (flag) {{ r = r + {b}; }} else {{ r = r - {b}; }} do {{ r = r * 0.5; }} while (false); return r;",
            f"var xs = [{a}, {b}].map(function(v) {{ return v * 2; }}); var s = xs.reduce((p, v) => p + v, 0); return s + [1, 2, 3].indexOf(3) + (typeof s === 'object' ? 1 : 0);",
            # ES2015+: destructuring, defaults, rest / spread, template literals, switch, classes
            f"const [p0, p1 = {c}, ...more] = [{a}, undefined, {b}, 2]; const {{u, v: {{w = 3}}}} = {{u: p0, v: {{}}}}; return u * w + p1 + more.length * more[0];",
            f"function lin(v, k = {n}, ...extra) {{ return extra.reduce((s, e) => s + e, v * k); }} const args = [{a}, undefined, {b}]; return lin(...args) + lin({c}, 2) + [...args, 1].length;",
            f"const tag = `n${{{n}}}`; let r; switch (tag) {{ case 'n2': r = {a}; break; case 'n3': r = {b}; break; default: r = {c}; }} const tbl = {{n2: 2, n3: 3, n4: 4}}; return r * tbl[tag] + `${{1}}${{2}}` * 1;",
            f"class Acc {{ constructor(s) {{ this.s = s; }} add(v) {{ this.s = this.s + v; return this; }} get doubled() {{ return this.s * 2; }} static of(v) {{ return new Acc(v); }} }} "
            f"class Neg extends Acc {{ add(v) {{ return super.add(-v); }} }} return Acc.of({a}).add({b}).doubled + new Neg({c}).add({b}).s;",
            # computed keys, class fields, labels, destructuring assignment, object rest
            f"const key = 'k' + {n}; const o = {{[key]: {a}, ['w' + 1]: {b}, [{n} * 2]: {c}}}; let p, q, others; ({{[key]: p, ...others}} = o); [p, q] = [others.w1, p]; return p * 2 + q + others[{n} * 2];",
            f"class Cell {{ gain = {n}; bias = {c}; static made = 1; constructor(v) {{ this.v = v * this.gain + this.bias; }} }} class Wide extends Cell {{ extra = this.v * 0.5; }} return new Cell({a}).v + new Wide({b}).extra + Cell.made;",
            f"let acc = {c}; outer: for (let i = 0; i < {n}; i++) {{ for (let j = 0; j < 3; j++) {{ if (j == 1) continue outer; if (i == {n} - 1) break outer; acc = acc + {a} * (i + 1); }} }} done: {{ acc = acc * 0.5; if (acc !== 0) break done; acc = 0; }} return acc + {b};",
        ]
        body = bodies[self.force_body] if self.force_body is not None else self.pick(bodies)
        return f"(function() {{ {body} }})()"

    def script(self):
        names = ["t", "x", "y", "z"]
        lines = []
        helper = self.rng.random() < 0.6
        if helper:
            lines.append(f"function helper(p, q) {{ var s = {self.expr(['p', 'q'], 2)}; return s * {self.num()} + q; }}")
        if self.rng.random() < 0.4:
            lines.append(f"const twice = (v) => v * 2 + {self.num()};")
        lines.append("function metric(t, x, y, z) {")
        if self.rng.random() < 0.6:
            lines.append(f"    $cfg.alpha.$default = {self.pick(['0.75', '1.5', '-0.25'])};")
            lines.append("    var alpha = $cfg.alpha;")
            names = names + ["alpha"]
        if self.rng.random() < 0.4:
            lines.append(f"    var beta = $cfg.beta; $cfg.beta.$default = {self.pick(['2', '0.5'])};")
            names = names + ["beta"]
        for count in range(int(self.rng.integers(2, 6))):
            v = f"w{self.tmp}"
            self.tmp += 1
            form = self.rng.random() if not (self.force_body is not None and count == 1) else 0.95
            if form < 0.55:
                lines.append(f"    var {v} = {self.expr(names, 3)};")
            elif form < 0.7 and helper:
                lines.append(f"    var {v} = helper({self.expr(names, 2)}, {self.pick(names)});")
            elif form < 0.85:
                lines.append(f"    var {v} = 0; for (var i = 0; i < {int(self.rng.integers(1, 4))}; i++) {{ {v} = {v} + {self.expr(names, 2)} * (i + 1); }}")
            elif form < 0.9:
                lines.append(f"    var {v} = $pin({self.expr(names, 2)});")
            else:
                lines.append(f"    var {v} = {self.statement_block(names)};")
            names = names + [v]
        if self.rng.random() < 0.3 and "twice" in "\n".join(lines):
            lines.append(f"    var tw = twice({self.pick(names)});")
            names = names + ["tw"]
        if self.rng.random() < 0.5:   # diagonal
            ents = [f"-(1 + {self.expr(names, 2)} * {self.expr(names, 1)} * 0.1)"] + [f"1 + 0.1 * {self.expr(names, 3)}" for _ in range(3)]
            lines.append("    return [" + ", ".join(ents) + "];")
        else:                         # full 16, symmetric by construction, with holes and plain numbers
            lines.append("    var g = []; g.length = 16;")
            lines.append(f"    g[0] = -(1 + 0.1 * {self.expr(names, 2)}); g[5] = 1 + 0.1 * {self.expr(names, 3)}; g[10] = 1; g[15] = 1 + 0.1 * {self.expr(names, 2)};")
            lines.append(f"    g[1] = 0.05 * {self.expr(names, 3)}; g[4] = g[1];")
            lines.append(f"    g[7] = 0.05 * {self.expr(names, 2)}; g[13] = g[7];")
            if self.rng.random() < 0.5:
                lines.append("    g[2] = 0; g[8] = 0;")
            lines.append("    return g;")
        lines.append("}")
        lines.append("metric")
        return "\n".join(lines) + "\n"


@pytest.mark.parametrize("block", range(8))
def test_random_scripts_agree_with_quickjs(grt, oracle_mod, have_reference, tmp_path, block):
    if not have_reference:
        pytest.skip("reference not mounted")
    try:
        q = oracle_mod.QuickJSReference()
    except FileNotFoundError:
        pytest.skip("reference not mounted")
    checked = 0
    for k in range(13):
        seed = 1000 * block + k
        js = Gen(seed, force_body=(k % 13) if block % 2 else None).script()
        assert block % 2 == 0 or "(function() {" in js
        folder = write_folder(tmp_path / f"s{seed}", js)
        path = os.path.join(folder, "probe.js")
        m = grt.Content(folder).load(path)
        rng = np.random.default_rng(seed)
        for _ in range(2):
            pos = rng.uniform(0.3, 1.6, 4)
            g, dg, n, found = q.metric(path, pos)
            assert list(found.keys()) == m.param_names[:len(found)], js
            assert np.allclose(list(found.values()), m.param_defaults[:len(found)]), js
            sym = m.eval_symbolic(pos)
            for got, want, what in [(sym["g"], g, "g"), (sym["dg"], dg, "dg")]:
                assert not (np.isfinite(want) & ~np.isfinite(got)).any(), (what, js)
                ok = np.isfinite(want)
                scale = max(1.0, np.abs(want[ok]).max()) if ok.any() else 1.0
                assert not ok.any() or np.abs(got[ok] - want[ok]).max() <= 1e-9 * scale, (what, seed, np.abs(got[ok] - want[ok]).max(), js)
            checked += 1
    assert checked == 26


def complex_script(seed):
    """Scripts that go through complex values the way the double-Kerr family does (csqrt of a real, CMath.i, products, integer
    powers, conjugates, Real / Imaginary / self_conjugate_multiply back to reals)."""
    g = Gen(seed)
    names = ["t", "x", "y", "z"]
    e = lambda d=2: g.expr(names, d)   # noqa: E731
    lines = ["function metric(t, x, y, z) {"]
    lines.append(f"    var za = CMath.csqrt({e()} - {g.pick(['0.5', '2', '3.5'])});")                      # real or imaginary depending on the point
    lines.append(f"    var zb = {e()} + CMath.i * ({e()});")
    lines.append(f"    var zc = za * zb + {g.pick(['zb / (za * za + 3)', 'CMath.conjugate(zb) * 0.5', 'CMath.pow(zb, 2)', 'CMath.pow(za + 2, 3)', 'zb - za * 2'])};")
    lines.append(f"    var zd = CMath.sqrt(zc * zc + {g.pick(['4', '9'])});")                                # complex square root away from the cut most of the time
    lines.append("    var r0 = CMath.Real(zc), i0 = CMath.Imaginary(zc), m0 = CMath.self_conjugate_multiply(zd);")
    lines.append(f"    var mix = CMath.Real(zd * CMath.conjugate(zb)) + CMath.Imaginary(za * {g.pick(['zb', 'zd', 'zc'])});")
    lines.append("    return [-(1 + 0.01 * m0), 1 + 0.05 * r0, 1 + 0.05 * i0, 1 + 0.02 * mix];")
    lines.append("}")
    lines.append("metric")
    return "\n".join(lines) + "\n"


def test_complex_valued_scripts_agree_with_quickjs(grt, oracle_mod, have_reference, tmp_path):
    if not have_reference:
        pytest.skip("reference not mounted")
    try:
        q = oracle_mod.QuickJSReference()
    except FileNotFoundError:
        pytest.skip("reference not mounted")
    compared = 0
    for seed in range(40):
        js = complex_script(seed)
        folder = write_folder(tmp_path / f"c{seed}", js)
        path = os.path.join(folder, "probe.js")
        m = grt.Content(folder).load(path)
        rng = np.random.default_rng(seed)
        for _ in range(3):
            pos = rng.uniform(0.3, 1.6, 4)
            g, dg, n, _ = q.metric(path, pos)
            sym = m.eval_symbolic(pos)
            assert np.isfinite(g).all() and np.isfinite(sym["g"]).all(), js
            assert np.abs(sym["g"] - g).max() <= 1e-9 * max(1.0, np.abs(g).max()), (seed, js, sym["g"] - g)
            # numeric duals give 0 * inf = NaN where a constant sits on a branch point (csqrt(0.5 - 0.5)) and the symbolic derivative is an
            # exact 0: whatever we leave undefined must be undefined there too, the defined entries must agree
            assert not (np.isfinite(dg) & ~np.isfinite(sym["dg"])).any(), js
            ok = np.isfinite(dg)
            assert ok.mean() >= 0.5
            assert np.abs(sym["dg"][ok] - dg[ok]).max() <= 1e-8 * max(1.0, np.abs(dg[ok]).max()), (seed, js)
            compared += 1
    assert compared == 120


BAD_SCRIPTS = {
    "tan of a complex value": "function f(t,x,y,z){ return [-1, CMath.tan(CMath.csqrt(x - 5)), 1, 1]; }\nf\n",
    "log of a complex value": "function f(t,x,y,z){ return [-1, CMath.log(CMath.i * x), 1, 1]; }\nf\n",
    "complex power with a variable exponent": "function f(t,x,y,z){ return [-1, CMath.Real(CMath.pow(CMath.i * x, y)), 1, 1]; }\nf\n",
    "complex power with a fractional exponent": "function f(t,x,y,z){ return [-1, CMath.Real(CMath.pow(CMath.i * x, 1.5)), 1, 1]; }\nf\n",
    "csqrt of a complex value": "function f(t,x,y,z){ return [-1, CMath.Real(CMath.csqrt(CMath.i * x)), 1, 1]; }\nf\n",
    "comparison of complex values": "function f(t,x,y,z){ return [-1, CMath.lt(CMath.i * x, 1), 1, 1]; }\nf\n",
    "three entries": "function f(t,x,y,z){ return [-1, 1, 1]; }\nf\n",
    "not an array": "function f(t,x,y,z){ return x; }\nf\n",
    "not a function": "var f = 3;\nf\n",
    "syntax error": "function f(t,x,y,z){ return [-1, 1, 1, ; }\nf\n",
    "undefined identifier": "function f(t,x,y,z){ return [-1, nothing_here * x, 1, 1]; }\nf\n",
    "bad $cfg name": "function f(t,x,y,z){ var q = $cfg['a-b']; return [-1, q, 1, 1]; }\nf\n",
}


@pytest.mark.parametrize("what", sorted(BAD_SCRIPTS))
def test_scripts_the_reference_engine_rejects_are_rejected(grt, oracle_mod, have_reference, tmp_path, what):
    """Error behaviour at the plugin boundary (js_interop.cpp:171-184, 585-616, 753-760, 862-900): a script QuickJS + the reference's
    bindings refuse must be refused by the product's loader too (GRT_E_SCRIPT with a message), never silently evaluated."""
    if not have_reference:
        pytest.skip("reference not mounted")
    try:
        q = oracle_mod.QuickJSReference()
    except FileNotFoundError:
        pytest.skip("reference not mounted")
    folder = write_folder(tmp_path, BAD_SCRIPTS[what])
    path = os.path.join(folder, "probe.js")
    with pytest.raises(RuntimeError):
        q.metric(path, [0.1, 1.2, 0.7, 0.4])
    with pytest.raises(grt.GrtError) as err:
        grt.Content(folder).load(path)
    assert str(err.value)
