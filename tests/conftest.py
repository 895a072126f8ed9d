import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def amdg():
    """the product package (its directory name has hyphens, so it is imported through importlib)"""
    return importlib.import_module("adaptive-multiresolution-dg_b200")


GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_names():
    return sorted(f[:-len(".dump.xz")] for f in os.listdir(GOLDEN) if f.endswith(".dump.xz"))


_cache = {}


def load_golden(name):
    import refdump
    if name not in _cache:
        _cache[name] = refdump.load(os.path.join(GOLDEN, name + ".dump.xz"))
    return _cache[name]
