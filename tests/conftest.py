import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def seq_c1():
    """5 frames of the C1 (640x480, TUM3 intrinsics) synthetic sequence, seed 20241108."""
    from sindslam_b200 import synth
    scene, frames = synth.make_sequence(5, synth.TUM3, seq=0, kind="box", start=8)
    return scene, frames


@pytest.fixture(scope="session")
def lib_built():
    from sindslam_b200 import build
    return build.build()
