"""The C-ABI library loads without a GPU and exports every symbol include/grt.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "grt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(grt_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_boundary():
    names = declared_functions()
    for required in ["grt_content_add_folder", "grt_metric_load", "grt_metric_load_by_name", "grt_metric_param_set", "grt_metric_source", "grt_create", "grt_set_metric",
                     "grt_set_features", "grt_set_background", "grt_set_camera", "grt_snapshot_geodesic", "grt_set_camera_on_geodesic", "grt_render", "grt_render_rows",
                     "grt_render_stripes", "grt_last_error"]:
        assert required in names


def test_library_exports_every_declared_symbol(grt):
    lib = ctypes.CDLL(grt.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol(grt):
    L = grt.lib()
    for n in declared_functions():
        assert getattr(L, n).argtypes is not None or n in ("grt_last_error", "grt_version"), n


def test_no_gpu_fails_loudly(grt):
    """No CPU fallback: without a usable B200 the device half refuses to start."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(grt.GrtError) as e:
        grt.Context(0)
    assert e.value.code == -4


def test_features_default(grt):
    f = grt.default_features()
    assert (f.universe_size, f.max_precision_radius, f.field_of_view, f.adaptive_sampling_threshold) == (20.0, 10.0, 90.0, 32.0)
    assert abs(f.min_step - 1e-6) < 1e-12 and f.adaptive_sampling == 1 and f.redshift == 0 and f.max_probes == 8 and f.use_prepass == -1


def test_cpp_host_fails_loudly_without_a_gpu(grt):
    """grt_headless (the C++ host over the C ABI) must not fall back to anything when no sm_100 GPU is usable."""
    import os
    import subprocess
    try:
        import torch
        if torch.cuda.is_available():
            import pytest
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    exe = os.path.join(os.path.dirname(grt.__file__), "grt_headless")
    assert os.path.exists(exe), "make -C geodesic_raytracing_b200 builds grt_headless"
    r = subprocess.run([exe, "-content", grt.CONTENT_DIR, "-bench", "kerr_schild", "-size", "64", "48", "-frames", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr
    usage = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert usage.returncode == 2 and "usage" in usage.stderr
