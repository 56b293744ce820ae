import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the in-tree libraries.  nvcc cross-compiles without a GPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200zk_build", os.path.join(ROOT, "zk-apps_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(os.path.join(ROOT, "zk-apps_b200", "libb200zk.so")) or os.environ.get("B200ZK_REBUILD"):
        mod.build_cuda()
    mod.build_hostcheck()
    mod.build_oracle()
    return mod


@pytest.fixture(scope="session")
def ctx(built):
    """GPU context.  No fallback: on a box without a GPU this raises (gpu tests are not run there)."""
    import zk_apps_b200 as z
    c = z.Context(0)
    yield c
    c.close()
