import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_package():
    """The package directory is `cp-cals_b200` (hyphen): import it under the module name cp_cals_b200."""
    if "cp_cals_b200" in sys.modules:
        return sys.modules["cp_cals_b200"]
    path = os.path.join(ROOT, "cp-cals_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("cp_cals_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["cp_cals_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def engine(pkg):
    eng = pkg.Engine(0)  # raises (no fallback) when there is no B200 or the library is not built
    yield eng
    eng.close()
