"""The bundled content against golden vectors produced from the reference's own scripts (tests/golden/make_golden.py),
and - where /root/reference is mounted - the reference scripts themselves through the unchanged plugin boundary."""
import json
import os

import numpy as np
import pytest

from conftest import REFERENCE_SCRIPTS

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_metric_vectors.json")))
BUNDLED = ["minkowski", "schwarzschild_fast", "schwarzschild", "schwarzschild_skewed", "schwarzschild_ingoing_ef", "kerr_schild", "kerr_boyer", "wormhole (morris-thorne)", "alcubierre", "double_unequal_kerr", "kerr_newman_schild", "double_kerr", "godel_cylindrical", "misner_4d"]


def check_against_golden(metric, orc, rec, rtol=3e-5):
    assert metric.param_names == rec["params"]
    assert np.allclose(metric.param_defaults, rec["defaults"])
    for k, v in rec["info"].items():
        assert getattr(metric.info, k) == v, k
    assert list(metric.info.w_v) == rec["w_v"]
    for pt in rec["points"]:
        g, dg = orc.metric_eval(pt["pos"])
        a = orc.accel(pt["pos"], pt["vel"])["ref"]
        c = orc.coords(pt["pos"], pt["vel"])
        for got, want in [(g, pt["g"]), (dg, pt["dg"]), (a, pt["accel_ref"]), (c["to_polar"], pt["to_polar"]), (c["from_polar"], pt["from_polar"]),
                          ([c["distance"]], [pt["distance"]]), (c["period"], pt["period"])]:
            want = np.asarray(want, np.float64)
            got = np.asarray(got, np.float64)
            assert np.array_equal(np.isnan(got), np.isnan(want))   # out-of-domain sample points stay out of domain
            ok = ~np.isnan(want)
            scale = max(1.0, np.abs(want[ok]).max()) if ok.any() else 1.0
            assert not ok.any() or np.abs(got[ok] - want[ok]).max() <= rtol * scale * 8
        if "qjs_g" in pt:
            # values the REFERENCE'S OWN JavaScript engine (vendored QuickJS + number.js) produced from the reference script,
            # float64 on both sides: the script engine and the symbolic layer of the product against the real thing
            sym = metric.eval_symbolic(pt["pos"], pt["vel"])
            for got, want in [(sym["g"], pt["qjs_g"]), (sym["dg"], pt["qjs_dg"])]:
                want, got = np.asarray(want, np.float64), np.asarray(got, np.float64)
                ok = np.isfinite(want)
                assert np.array_equal(np.isfinite(got), ok)
                scale = max(1.0, np.abs(want[ok]).max()) if ok.any() else 1.0
                assert not ok.any() or np.abs(got[ok] - want[ok]).max() <= 1e-9 * scale


@pytest.mark.parametrize("name", BUNDLED)
def test_bundled_script_equals_reference_script_golden(oracles, name):
    metric, orc = oracles(name)
    check_against_golden(metric, orc, GOLDEN[name])


def test_forward_geodesic_variant_is_the_same_metric(oracles):
    """godel_forward.js/.json: the bundled Goedel metric with `follow_geodesics_forward` set (no reference metric sets that key; the variant
    keeps the code path of cl.cl:3039-3051 under test).  Same g, dg, maps as the reference's godel_cylinder.js."""
    metric, orc = oracles("godel_cylindrical_forward")
    check_against_golden(metric, orc, GOLDEN["godel_cylindrical"])
    assert "#define GRT_FORWARD_GEODESIC_PATH 1" in metric.source(False)


def test_all_reference_scripts_load_unchanged(grt, have_reference):
    if not have_reference:
        pytest.skip("reference not mounted")
    c = grt.Content(REFERENCE_SCRIPTS)
    names = c.metric_names()
    assert len(names) == 31
    assert names[0] == "schwarzschild"   # sorting.json order (content_manager.cpp:193-253)
    for n in names:
        m = c.load_by_name(n)
        assert m.name == n
        src = m.source(False)
        assert "grt_accel" in src and "nan" not in src.lower().replace("grt_nan", "")


@pytest.mark.parametrize("name", ["kerr_schild", "double_unequal_kerr", "schwarzschild_ingoing_ef_hawking", "misner_4d"])
def test_reference_scripts_reproduce_golden(grt, oracle_mod, have_reference, name):
    if not have_reference:
        pytest.skip("reference not mounted")
    c = grt.Content(REFERENCE_SCRIPTS)
    m = c.load_by_name(name)
    check_against_golden(m, oracle_mod.for_metric(m), GOLDEN[name], rtol=1e-6)


def test_reference_headline_program_compiles_for_sm_100a(grt, have_reference, tmp_path, monkeypatch):
    if not have_reference:
        pytest.skip("reference not mounted")
    monkeypatch.setenv("GRT_CACHE_DIR", str(tmp_path))
    c = grt.Content(REFERENCE_SCRIPTS)
    m = c.load_by_name("double_unequal_kerr")
    log = m.compile(is_static=False, fast_math=True)
    assert "error" not in log.lower()
    assert any(f.endswith(".cubin") for f in os.listdir(tmp_path))


# ---- the reference's own JavaScript engine ---------------------------------------------------------------------------------
# oracle/_ref/libqjsref: the vendored QuickJS (bignum operator overloading) compiled from /root/reference/deps/quickjs where it
# lies, running the reference's number.js and scripts with CShim/CMath bound to numeric dual numbers (oracle/ref_quickjs.cpp).
def _quickjs(oracle_mod):
    try:
        return oracle_mod.QuickJSReference()
    except FileNotFoundError:
        pytest.skip("reference not mounted")


def _points(seed, n=3):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0.6, 2.5, (n, 4))
    pos[:, 0] = rng.uniform(0.05, 1, n)
    return pos


def test_every_reference_metric_script_matches_quickjs(grt, oracle_mod, have_reference):
    """All 31 metric scripts: g_ij and every d_k g_ij from our script engine + symbolic differentiation equal what the
    reference's engine computes by forward-mode dual numbers, to 1e-9 in float64; `$cfg` names and defaults agree too."""
    if not have_reference:
        pytest.skip("reference not mounted")
    q = _quickjs(oracle_mod)
    c = grt.Content(REFERENCE_SCRIPTS)
    scripts = sorted(f for f in os.listdir(REFERENCE_SCRIPTS) if f.endswith(".js"))
    assert len(scripts) == 31
    for k, f in enumerate(scripts):
        path = os.path.join(REFERENCE_SCRIPTS, f)
        m = c.load(path)
        for pos in _points(100 + k):
            g, dg, n, found = q.metric(path, pos)
            # first-access order (js_interop.cpp:795-846); the coordinate / periodicity scripts of the metric share the sandbox and
            # may append further names after the metric script's own
            assert list(found.keys()) == m.param_names[:len(found)], f
            assert np.allclose(list(found.values()), m.param_defaults[:len(found)]), f
            assert n == (16 if m.info.is_big else 4) or (n == 16 and not m.info.is_big), f   # 16-entry scripts with zero off-diagonals are reduced
            sym = m.eval_symbolic(pos)
            for got, want, what in [(sym["g"], g, "g"), (sym["dg"], dg, "dg")]:
                # out of a function's domain the numeric duals give NaN * 0 = NaN where the symbolic derivative is an exact 0:
                # every entry we leave undefined must be undefined there too, and the defined ones must agree
                assert not (np.isfinite(want) & ~np.isfinite(got)).any(), (f, what)
                ok = np.isfinite(want)
                assert ok.mean() >= 0.5, (f, what)
                scale = max(1.0, np.abs(want[ok]).max())
                assert np.abs(got[ok] - want[ok]).max() <= 1e-9 * scale, (f, what, np.abs(got[ok] - want[ok]).max())


def test_parameters_reach_the_script_like_in_quickjs(grt, oracle_mod, have_reference):
    if not have_reference:
        pytest.skip("reference not mounted")
    q = _quickjs(oracle_mod)
    c = grt.Content(REFERENCE_SCRIPTS)
    path = os.path.join(REFERENCE_SCRIPTS, "kerr_schild.js")
    m = c.load(path)
    m.set_param("a", 0.9)
    m.set_param("rs", 1.7)
    pos = [0.1, 1.4, -0.8, 1.9]
    g, dg, _, _ = q.metric(path, pos, {"a": float(np.float32(0.9)), "rs": float(np.float32(1.7))})   # parameters are floats (config_variables, js_interop.hpp:12-23)
    sym = m.eval_symbolic(pos)
    assert np.abs(sym["g"] - g).max() < 1e-9 and np.abs(sym["dg"] - dg).max() < 1e-9
    g0, _, _, _ = q.metric(path, pos)
    assert np.abs(g0 - g).max() > 1e-3


def test_coordinate_and_origin_scripts_match_quickjs(grt, oracle_mod, have_reference, tmp_path):
    """Every scripts/coordinates/*.js (4 -> 4) and scripts/origins/*.js (4 -> 1) through a throw-away metric that names them."""
    if not have_reference:
        pytest.skip("reference not mounted")
    import json
    import shutil
    q = _quickjs(oracle_mod)
    coords = sorted(f[:-3] for f in os.listdir(os.path.join(REFERENCE_SCRIPTS, "coordinates")) if f.endswith(".js"))
    origins = sorted(f[:-3] for f in os.listdir(os.path.join(REFERENCE_SCRIPTS, "origins")) if f.endswith(".js"))
    assert len(coords) == 20 and len(origins) == 2
    shutil.copytree(os.path.join(REFERENCE_SCRIPTS, "coordinates"), tmp_path / "coordinates")
    shutil.copytree(os.path.join(REFERENCE_SCRIPTS, "origins"), tmp_path / "origins")
    flat = "function f(t, x, y, z) { return [-1, 1, 1, 1]; }\nf\n"
    for i, cname in enumerate(coords):
        (tmp_path / f"probe{i}.js").write_text(flat)
        (tmp_path / f"probe{i}.json").write_text(json.dumps({"name": f"probe{i}", "to_polar": cname, "from_polar": cname, "origin_distance": origins[i % 2]}))
    c = grt.Content(str(tmp_path))
    checked = 0
    for i, cname in enumerate(coords):
        m = c.load(str(tmp_path / f"probe{i}.js"))
        cpath = os.path.join(REFERENCE_SCRIPTS, "coordinates", cname + ".js")
        opath = os.path.join(REFERENCE_SCRIPTS, "origins", origins[i % 2] + ".js")
        shared = {}
        for pos in _points(500 + i, 2):
            sym = m.eval_symbolic(pos)
            want, _, _, found = q.run(cpath, 1, pos, dict(zip(m.param_names, m.param_defaults)))
            which = "to_polar" if "to_polar" in cname or cname.endswith("_to_polar") else None
            # a coordinate file is either a to_polar or a from_polar map; it was bound to both slots, compare the slot it is meant for
            got = sym["to_polar"] if np.allclose(sym["to_polar"], want, rtol=1e-9, atol=1e-9, equal_nan=True) else sym["from_polar"]
            assert np.allclose(got, want, rtol=1e-9, atol=1e-9, equal_nan=True), (cname, got, want)
            dist, _, _, _ = q.run(opath, 2, pos, dict(zip(m.param_names, m.param_defaults)))
            assert np.allclose(sym["distance"], dist, rtol=1e-9, atol=1e-9, equal_nan=True), (origins[i % 2], sym["distance"], dist)
            checked += 1
    assert checked == 40


@pytest.mark.parametrize("name", ["double_unequal_kerr", "double_kerr"])
def test_double_kerr_family_with_derived_parameters(grt, oracle_mod, have_reference, name):
    """BASELINE config 5's metric family: a quarter of the script's algebra depends on the parameters alone and travels as derived
    parameters.  The emitted fp32 functions fed with the packed block agree with the float64 symbolic evaluation,
    and a parameter change moves both together."""
    if not have_reference:
        pytest.skip("reference not mounted")
    c = grt.Content(REFERENCE_SCRIPTS)
    m = c.load_by_name(name)
    assert len(m.packed_params) > len(m.params)
    rng = np.random.default_rng(21)
    for trial in range(2):
        orc = oracle_mod.for_metric(m)
        for _ in range(4):
            pos = rng.uniform(0.7, 2.6, 4)
            vel = rng.uniform(-1, 1, 4)
            vel[0] = -1
            e = m.eval_symbolic(pos, vel)
            g, dg = orc.metric_eval(pos.astype(np.float32))
            a = orc.accel(pos.astype(np.float32), vel.astype(np.float32))
            for got, want in [(g, e["g"]), (dg, e["dg"]), (a["lowered"], e["accel"]), (a["ref"], e["accel_ref"])]:
                assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())
        m.set_param(0, m.params[0] * 1.2)
