"""The script engine (csrc/minijs.cpp) through the plugin boundary: the JavaScript the reference's scripts rely on
(number.js:4-81 operator overloading, function-scoped `var`, sparse arrays with `.length`, closures, Math.*)."""
import os
import textwrap

import numpy as np
import pytest

from test_loader import write_folder


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
