"""The bundled content against golden vectors produced from the reference's own scripts (tests/golden/make_golden.py),
and - where /root/reference is mounted - the reference scripts themselves through the unchanged plugin boundary."""
import json
import os

import numpy as np
import pytest

from conftest import REFERENCE_SCRIPTS

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_metric_vectors.json")))
BUNDLED = ["minkowski", "schwarzschild_fast", "schwarzschild", "schwarzschild_skewed", "schwarzschild_ingoing_ef", "kerr_schild", "kerr_boyer", "wormhole (morris-thorne)", "alcubierre"]


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


@pytest.mark.parametrize("name", BUNDLED)
def test_bundled_script_equals_reference_script_golden(oracles, name):
    metric, orc = oracles(name)
    check_against_golden(metric, orc, GOLDEN[name])


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
