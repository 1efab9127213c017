import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_SCRIPTS = "/root/reference/scripts"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def grt():
    import geodesic_raytracing_b200 as g

    g.build_library()
    return g


@pytest.fixture(scope="session")
def content(grt):
    return grt.Content(grt.CONTENT_DIR)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    return oracle


@pytest.fixture(scope="session")
def oracles(content, oracle_mod):
    """name -> (Metric, Oracle), built lazily and cached for the session."""
    cache = {}

    def get(name):
        if name not in cache:
            m = content.load_by_name(name)
            cache[name] = (m, oracle_mod.for_metric(m))
        return cache[name]

    return get


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir(REFERENCE_SCRIPTS)


@pytest.fixture(scope="session")
def gpu_ctx(grt):
    ctx = grt.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def background(grt):
    return grt.synthetic_background(1024, 512)
