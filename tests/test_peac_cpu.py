"""The two statements of the PEAC oracle -- the readable Python loops (oracle/peac_oracle.py) and the C file the long
tests and the CPU baseline run (oracle/peac_cpu.c) -- restate the same reference code (include/PEAC/*.hpp,
DynaDetect.cc:558-593) and must agree exactly: coarse planes, block map, grown membership, final plane ids."""
import numpy as np
import pytest

from oracle import peac_oracle as po
from sindslam_b200 import synth


@pytest.mark.parametrize("cam_name,kind,hole", [("TUM3", "box", 0.0005), ("D455_848", "humanoid", 0.002)])
def test_c_and_python_peac_oracles_agree(cam_name, kind, hole):
    cam = getattr(synth, cam_name)
    _, frames = synth.make_sequence(1, cam, seq=5, kind=kind, start=11, hole_rate=hole)
    f = frames[0]
    pts = po.organized_cloud(f.depth, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    m1, n1, d1 = po.plane_fit(pts)
    m2, n2, d2 = po.plane_fit_c(pts)
    assert n1 == n2 and n1 >= 3
    assert d1["coarse"] == d2["coarse"]
    assert np.array_equal(d1["blk_map"], d2["blk_map"])
    assert np.array_equal(d1["grown"], d2["grown"])
    assert np.array_equal(m1, m2)
    assert d2["stats"]["levels"] > 10 and d2["stats"]["queue_entries"] > d2["stats"]["seeds"]
