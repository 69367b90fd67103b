import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    d = ROOT / "tests" / "golden"
    return {name: np.load(d / (name + ".npz")) for name in ("ref_scene0", "ref_scene1", "ref_scene2", "ref_kat", "ref_stratified_kat")}
