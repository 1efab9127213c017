"""Plugin boundary: content folders, config inheritance, the script API, generated code."""
import os
import textwrap

import numpy as np
import pytest


def test_bundled_content_lists_in_sorting_order(content):
    assert content.metric_names() == ["minkowski", "schwarzschild_fast", "schwarzschild", "schwarzschild_skewed", "schwarzschild_ingoing_ef", "kerr_schild", "kerr_boyer", "wormhole (morris-thorne)", "alcubierre", "double_schwarzschild", "double_unequal_kerr", "kerr_newman_schild", "double_kerr", "godel_cylindrical",
                                      "godel_cylindrical_forward", "misner_4d"]


def test_config_inheritance_and_flags(content):
    ks = content.load_by_name("kerr_schild").info
    assert ks.is_big == 1 and ks.is_constant_theta == 0 and ks.use_prepass == 1 and ks.adaptive_precision == 1 and ks.detect_singularities == 1
    assert ks.coordinate_system == 1 and abs(ks.max_acceleration_change - 1e-6) < 1e-12
    assert list(ks.w_v) == [1, 1, 1, 1]
    sw = content.load_by_name("schwarzschild_fast").info
    assert sw.is_big == 0 and sw.is_constant_theta == 1 and sw.singular == 1 and abs(sw.singular_terminator - 1.05) < 1e-6 and sw.adaptive_precision == 0
    assert list(sw.w_v) == [1, 1, 8, 8] and sw.has_periodicity == 1
    mk = content.load_by_name("minkowski").info
    assert mk.is_big == 0 and mk.flops_expr == 0 and mk.flops_accel == 0 and abs(mk.max_acceleration_change - 1e-4) < 1e-10
    al = content.load_by_name("alcubierre").info
    assert al.unconditionally_nonsingular == 1 and al.is_big == 1
    ds = content.load_by_name("double_schwarzschild").info
    assert ds.coordinate_system == 2 and list(ds.w_v) == [1, 1, 8, 1]


def test_params_first_access_order_and_defaults(content):
    m = content.load_by_name("kerr_schild")
    assert m.param_names == ["a", "rs"] and m.param_defaults == [-0.5, 1.0]
    m = content.load_by_name("alcubierre")
    assert m.param_names == ["velocity", "sigma", "R"] and m.param_defaults == [2.0, 1.0, 2.0]


def test_static_source_substitutes_literals(content):
    m = content.load_by_name("kerr_schild")
    dyn, sta = m.source(False), m.source(True)
    assert "cfg.a" in dyn and "cfg.rs" in dyn
    assert "cfg.a" not in sta and "-0.5f" in sta
    m.set_param("a", 0.25)
    assert "0.25f" in m.source(True)
    for fn in ["grt_metric", "grt_metric_partials", "grt_metric_and_partials", "grt_to_polar", "grt_to_polar_vel", "grt_from_polar", "grt_from_polar_vel", "grt_distance",
               "grt_radii", "grt_period", "grt_accel", "grt_accel_ref"]:
        assert f"GRT_FN void {fn}(" in dyn


def write_folder(tmp_path, js, cfg='{"name": "probe", "inherit_settings": "flat_cartesian"}'):
    import shutil

    import geodesic_raytracing_b200 as g

    folder = tmp_path / "content"
    shutil.copytree(g.CONTENT_DIR, folder)
    (folder / "probe.js").write_text(textwrap.dedent(js))
    (folder / "probe.json").write_text(cfg)
    return str(folder)


def test_script_api_surface(grt, tmp_path, oracle_mod):
    """Every CMath / CShim entry the reference exposes (js_interop.cpp:665-743) is callable; mixed Number/number arithmetic works both ways."""
    folder = write_folder(tmp_path, """
        function f(t, x, y, z) {
            $cfg.k.$default = 2;
            var k = $cfg.k;
            var a = CMath.sin(x) + CMath.cos(y) * CMath.tan(z) - CMath.asin(x / 10) + CMath.acos(y / 10) + CMath.atan(z) + CMath.atan2(y, x);
            var b = CMath.fabs(x) + CMath.log(x * x + 1) + CMath.sqrt(y * y + 1) + CMath.psqrt(z * z + 1) + CMath.exp(-x) + CMath.sinh(y / 4) + CMath.cosh(z / 4) + CMath.tanh(x);
            var c = CMath.pow(x * x + 1, 1.5) + CMath.fast_length(x, y, z) + CMath.length(x, y, z) + CMath.smooth_fmod(x + 9, 2) + CMath.select(CMath.lt(x, y), x, y);
            var d = CMath.lte(x, 1) + CMath.eq(x, x) + CMath.gt(y, 0) + CMath.gte(z, 0) + 2 * k - k / 2 + (1 - k) + (-k) + M_PI + CMath.M_PI + CMath.PI + Math.PI;
            var zc = (x + CMath.i * y) * CMath.conjugate(x + CMath.i * y);
            var e = CMath.Real(zc) + CMath.Imaginary(zc) + CMath.self_conjugate_multiply(x + CMath.get_i() * z) + CMath.Real(CMath.pow(x + CMath.i, 2)) + CMath.Real(CMath.csqrt(-(x * x) - 1) / CMath.i);
            var n = new CMath.number(3);
            var s = CShim.add(n, 1) ;
            $pin(a); $pin([a, b]);
            CMath.debug(a);
            return [-(1 + a * a), 1 + b * b, 1 + c * c + d * d + e * e, CShim.mul(s, CShim.div(CShim.sub(n, 1), CShim.neg(n))) + 5];
        }
        f
        """)
    c = grt.Content(folder)
    m = c.load(os.path.join(folder, "probe.js"))
    assert m.param_names == ["k"] and m.params == [2.0]
    o = oracle_mod.for_metric(m)
    x, y, z, k = 0.7, 1.3, 0.4, 2.0
    g, dg = o.metric_eval([0.1, x, y, z])
    a = np.sin(x) + np.cos(y) * np.tan(z) - np.arcsin(x / 10) + np.arccos(y / 10) + np.arctan(z) + np.arctan2(y, x)
    b = abs(x) + np.log(x * x + 1) + 2 * np.sqrt(y * y + 1) - np.sqrt(y * y + 1) + np.sqrt(z * z + 1) + np.exp(-x) + np.sinh(y / 4) + np.cosh(z / 4) + np.tanh(x)
    L = np.sqrt(x * x + y * y + z * z)
    cc = (x * x + 1) ** 1.5 + 2 * L + np.fmod(x + 9, 2) + min(x, y)
    d = 1 + 1 + 1 + 1 + 2 * k - k / 2 + (1 - k) - k + 4 * np.pi   # gte(z, 0) with z > 0 is 1 under the reference's '>' too
    e = (x * x + y * y) + 0 + (x * x + z * z) + (x * x - 1) + np.sqrt(x * x + 1)
    assert np.isclose(g[0], -(1 + a * a), rtol=2e-5)
    assert np.isclose(g[5], 1 + b * b, rtol=2e-5)
    assert np.isclose(g[10], 1 + cc * cc + d * d + e * e, rtol=2e-5)
    assert np.isclose(g[15], 4 * (2 / -3) + 5, rtol=2e-5)
    # partials against central differences of the same generated g
    for k_ in range(1, 4):
        h = 1e-3
        p0, p1 = np.array([0.1, x, y, z]), np.array([0.1, x, y, z])
        p0[k_] -= h
        p1[k_] += h
        fd = (o.metric_eval(p1)[0].astype(np.float64) - o.metric_eval(p0)[0]) / (2 * h)
        assert np.allclose(dg[k_ * 16:(k_ + 1) * 16], fd, rtol=5e-2, atol=5e-2)


@pytest.mark.parametrize("js, message", [
    ("function f(t, x, y, z) { return [1, 2, 3]; }\nf", "Must return array length of 4 or 16"),
    ("42", "Expected function"),
    ("function f(t, x, y, z) { return [CMath.tan(x + CMath.i), 1, 1, 1]; }\nf", "real-only"),
    ("function f(t, x, y, z) { return [CMath.lt(x + CMath.i, 1), 1, 1, 1]; }\nf", "real values"),
    ("function f(t, x, y, z) { return [CMath.pow(x + CMath.i, 1.5), 1, 1, 1]; }\nf", "integral"),
    ("function f(t, x, y, z) { var q = $cfg['bad name']; return [-1, 1, 1, 1]; }\nf", "alphanumeric"),
    ("function f(t, x, y, z) { return [-1, 1, 1, 1 }\nf", "Error in script exec"),
    ("function f(t, x, y, z) { return undefined_thing(x); }\nf", "undefined_thing"),
])
def test_script_errors_surface_as_script_errors(grt, tmp_path, js, message):
    folder = write_folder(tmp_path, js)
    c = grt.Content(folder)
    with pytest.raises(grt.GrtError) as e:
        c.load(os.path.join(folder, "probe.js"))
    assert e.value.code == -2 and message in e.value.message


def test_missing_coordinate_script_is_reported(grt, tmp_path):
    folder = write_folder(tmp_path, "(t, x, y, z) => [-1, 1, 1, 1]", '{"name": "probe", "inherit_settings": "flat_cartesian", "to_polar": "nope"}')
    c = grt.Content(folder)
    with pytest.raises(grt.GrtError) as e:
        c.load(os.path.join(folder, "probe.js"))
    assert "No to polar file nope" in e.value.message


def test_sixteen_entry_returns_allow_holes_and_plain_numbers(grt, tmp_path, oracle_mod):
    folder = write_folder(tmp_path, """
        function f(t, x, y, z) { var g = []; g.length = 16; g[0] = -1; g[5] = 1; g[10] = true; g[15] = 1 + 0 * x; g[1] = 0.5 * x; return g; }
        f
        """)
    m = grt.Content(folder).load(os.path.join(folder, "probe.js"))
    assert m.info.is_big == 1
    g, _ = oracle_mod.for_metric(m).metric_eval([0, 2.0, 0, 0])
    assert g[0] == -1 and g[1] == 1.0 and g[4] == 1.0 and g[10] == 1.0 and g[2] == 0


def test_load_by_unknown_name_fails(grt, content):
    with pytest.raises(grt.GrtError):
        content.load_by_name("no_such_metric")


# ---- the emitted arrangement of the acceleration against the reference formulation, float64 on the host ----------------------
# (grt_metric_eval_symbolic: plugin-layer diagnostic; the arrangement is chosen per metric among explicit partials /
#  P-S form / forward-mode + reverse-mode form, each with the 2x2-minor or the multiplied-out cofactors - DESIGN.md 2)
def _check_arrangement(metric, seed):
    rng = np.random.default_rng(seed)
    worst = 0.0
    for _ in range(6):
        pos = rng.uniform(0.7, 2.6, 4)
        vel = rng.uniform(-1, 1, 4)
        vel[0] = -1.0
        if metric.info.is_constant_theta:
            pos[2] = float(np.float32(np.pi / 2))
            vel[2] = 0.0
        e = metric.eval_symbolic(pos, vel)
        a, r = e["accel"], e["accel_ref"]
        if not (np.isfinite(a).all() and np.isfinite(r).all()):
            assert np.array_equal(np.isfinite(a), np.isfinite(r))
            continue
        scale = max(1.0, np.abs(r).max())
        worst = max(worst, np.abs(a - r).max() / scale)
    return worst


@pytest.mark.parametrize("name", ["minkowski", "schwarzschild", "schwarzschild_skewed", "schwarzschild_ingoing_ef", "kerr_schild", "kerr_boyer", "wormhole (morris-thorne)", "alcubierre",
                                  "double_schwarzschild"])
def test_emitted_acceleration_equals_reference_formulation_in_float64(content, name):
    assert _check_arrangement(content.load_by_name(name), 3) < 1e-9


def test_emitted_acceleration_of_every_reference_metric(grt, have_reference):
    if not have_reference:
        pytest.skip("reference not mounted")
    c = grt.Content("/root/reference/scripts")
    for k, name in enumerate(c.metric_names()):
        assert _check_arrangement(c.load_by_name(name), 40 + k) < 1e-8, name


# ---- derived parameters (MetricDescriptor::derived): sub-expressions of the $cfg parameters alone travel in the parameter block --------
PARAM_HEAVY_JS = """
function f(t, r, theta, phi) {
    $cfg.m.$default = 0.4; $cfg.q.$default = 0.25; $cfg.k.$default = 1.5;
    var m = $cfg.m, q = $cfg.q, k = $cfg.k;
    // a tower of parameter-only algebra, as the double-Kerr scripts have (roots of a cubic in the masses and spins)
    var p1 = CMath.sqrt(m * m + q * q + k) * (m + 2 * q) / (1 + k * k);
    var p2 = CMath.pow(p1 * p1 + m * q * k + 3, 1 / 3) - q / (2 + m * m);
    var p3 = CMath.exp(-p2 * 0.1) * (p1 + p2) + CMath.sin(k) * CMath.cos(m + q);
    var p4 = (p1 * p2 + p2 * p3 + p3 * p1) / (1 + p1 * p1 + p2 * p2 + p3 * p3);
    var rs = 2 * m * (1 + 0.05 * p4);
    var f0 = 1 - rs / r + p3 * 0.01 / (r * r);
    return [-f0, 1 / f0, r * r * (1 + 0.001 * p2), r * r * CMath.sin(theta) * CMath.sin(theta)];
}
f
"""


def test_derived_parameters_travel_in_the_parameter_block(grt, oracle_mod, tmp_path):
    folder = write_folder(tmp_path, PARAM_HEAVY_JS, cfg='{"name": "probe", "inherit_settings": "polar_base"}' if os.path.exists(os.path.join(grt.CONTENT_DIR, "polar_base.json")) else
                          '{"name": "probe", "inherit_settings": "spherical"}')
    m = grt.Content(folder).load(os.path.join(folder, "probe.js"))
    assert m.param_names == ["m", "q", "k"]
    packed = m.packed_params
    assert len(packed) > 3 and packed[:3] == m.params                       # named values first, derived ones behind them
    src = m.source(False)
    assert f"#define GRT_NUM_PARAMS {len(packed)}" in src and "#define GRT_NUM_NAMED_PARAMS 3" in src and "float grt_d0;" in src
    body = src[src.index("void grt_accel("):src.index("void grt_accel_ref(")]
    assert "GRT_EXP" not in body and "GRT_POW" not in body                   # the parameter-only tower is not recomputed per evaluation
    static = m.source(True)
    assert "cfg." not in static[static.index("void grt_metric("):]           # static program: everything folded into literals
    rng = np.random.default_rng(4)
    for trial in range(2):
        orc = oracle_mod.for_metric(m)                                       # the emitted fp32 functions with the packed block ...
        orc_static = oracle_mod.for_metric(m, is_static=True)
        for _ in range(4):
            pos = rng.uniform(1.5, 3.0, 4)
            vel = rng.uniform(-1, 1, 4)
            e = m.eval_symbolic(pos, vel)                                    # ... against the float64 symbolic evaluation
            for o in (orc, orc_static):
                g, dg = o.metric_eval(pos.astype(np.float32))
                a = o.accel(pos.astype(np.float32), vel.astype(np.float32))
                assert np.allclose(g, e["g"], rtol=2e-5, atol=2e-5) and np.allclose(dg, e["dg"], rtol=2e-4, atol=2e-4)
                assert np.allclose(a["lowered"], e["accel"], rtol=5e-4, atol=5e-4) and np.allclose(a["ref"], e["accel_ref"], rtol=5e-4, atol=5e-4)
        before = list(packed)
        m.set_param("k", 0.7 + 0.2 * trial)                                  # a parameter change re-evaluates the derived values
        m.set_param("m", 0.55 - 0.1 * trial)
        packed = m.packed_params
        assert packed[:3] == m.params and packed[3:] != before[3:]


def test_parameter_only_products_travel_in_the_parameter_block(content):
    """Even a single parameter-only operation (the a * a of Kerr-Schild) is evaluated on the host: a parameter lives in a uniform
    register, so a * a in the loop costs a move and a multiply per use.  A metric whose parameters only meet variables has no block."""
    m = content.load_by_name("kerr_schild")
    named = m.params
    a = named[m.param_names.index("a")]
    assert m.packed_params[:len(named)] == named and len(m.packed_params) > len(named)
    assert any(abs(d - a * a) < 1e-6 for d in m.packed_params[len(named):])
    src = m.source(False)
    body = src[src.index("void grt_accel_radii("):src.index("void grt_accel_ref(")]
    assert "cfg.a * cfg.a" not in body and "cfg.grt_d0" in src
    for name in ["minkowski", "schwarzschild_fast"]:
        m = content.load_by_name(name)
        assert m.packed_params == m.params and "grt_d0" not in m.source(False)
