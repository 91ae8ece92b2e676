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


def test_create_rejects_unusable_configs(lib_built):
    """sindyn_create validates the configuration BEFORE it looks for a device: the cluster grid must give the 12 clusters the
    kernels are compiled for (DynaDetect.cc:46-47), the flow scale and the solver parameters must be usable."""
    from sindslam_b200 import capi
    lib = capi.load_library()

    def status(**over):
        cfg = capi.Config()
        lib.sindyn_default_config(ctypes.byref(cfg), 640, 480)
        for k, v in over.items():
            setattr(cfg, k, v)
        h = ctypes.c_void_p()
        st = lib.sindyn_create(ctypes.byref(cfg), ctypes.byref(h))
        if h:
            lib.sindyn_destroy(h)
        return st

    for bad in (dict(n_row_cluster=4, n_col_cluster=4), dict(n_row_cluster=2, n_col_cluster=4), dict(flow_scale=0.0), dict(flow_scale=1.5),
                dict(flow_scale=0.01), dict(brox_inner=0), dict(brox_pyr_scale=1.0), dict(brox_omega=2.5), dict(fx=0.0), dict(depth_scale=-1.0),
                dict(width=32)):
        assert status(**bad) == 1, bad                      # SINDYN_ERR_INVALID
    assert status() in (0, 3)                               # OK with a device, NO_DEVICE without: never INVALID
