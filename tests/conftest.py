import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def wg():
    """The product package (ctypes over libwinograd_b200.so)."""
    import wg_loader
    return wg_loader.load()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def seeded_data(tmp_path_factory):
    """The reference's data/ directory written by the seeded restatement of data_generator.py (seed 0)."""
    import datagen
    out = tmp_path_factory.mktemp("wgdata") / "data"
    tensors = datagen.generate_all(str(out), seed=0, f4x4=True, goldens=True)
    return str(out), tensors
