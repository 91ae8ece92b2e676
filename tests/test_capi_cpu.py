"""CPU-only checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol
include/sindyn.h declares (no compute calls without a GPU); without a device the API fails loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sindyn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sindyn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_bound(lib_built):
    from sindslam_b200 import capi
    lib = capi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sindyn.h but not exported"
    assert set(declared) == set(capi.SIGNATURES), set(declared) ^ set(capi.SIGNATURES)
    assert b"sm_100a" in lib.sindyn_version()


def test_library_is_sm100a_only(lib_built):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", lib_built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_device_fails_loudly(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sindslam_b200 import capi
    with pytest.raises(capi.SindynError):
        capi.SinDyn(640, 480)


def test_config_struct_layout_matches_header(lib_built):
    from sindslam_b200 import capi
    cfg = capi.Config()
    capi.load_library().sindyn_default_config(ctypes.byref(cfg), 640, 480)
    assert (cfg.width, cfg.height) == (640, 480)
    assert abs(cfg.brox_alpha - 0.197) < 1e-6 and cfg.brox_gamma == 50.0 and abs(cfg.brox_pyr_scale - 0.8) < 1e-6
    assert (cfg.brox_inner, cfg.brox_outer, cfg.brox_solver) == (10, 77, 10)   # DynaDetect.cc:1029
    assert (cfg.n_row_cluster, cfg.n_col_cluster) == (3, 4) and cfg.depth_weight == 1.5  # DynaDetect.cc:46-48
    assert abs(cfg.flow_scale - 0.6) < 1e-6                                      # DynaDetect.cc:1033
