"""Edge cases of the full path on the device: invalid depth everywhere, identical frames (zero flow, zero residual),
a small non-VGA resolution, wrong call order and bad arguments (error behaviour of the C ABI)."""
import cv2
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu


def test_all_invalid_depth(seq_c1):
    from sindslam_b200.capi import SinDyn
    _, frames = seq_c1
    cam = synth.TUM3
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    s.set_prev_frames(frames[0].bgr, frames[0].bgr)
    o = orc.DynaDetectOracle(frames[0].bgr, frames[0].bgr, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=True)
    zero = np.zeros((cam.height, cam.width), np.uint16)
    far = np.full((cam.height, cam.width), 40000, np.uint16)          # 8 m: beyond the 6 m validity limit
    for k, depth in ((1, zero), (2, far)):
        mask, label = s.detect(frames[k].bgr, depth, k)
        fr = s.flow_results()
        r = o.detect(frames[k].bgr, depth, inject_masks=(fr["low"], fr["high"]))
        assert np.array_equal(mask, r["mask"]) and np.array_equal(label, r["label"])
        assert not (mask == 255).any() and int(label.max()) == 0       # nothing can be dynamic without valid depth
    s.close()


def test_identical_frames_do_not_crash(seq_c1):
    """Zero flow -> zero residual -> the reference divides by maxError = 0 (DynaDetect.cc:1281); whatever comes out must be
    a well-formed mask, and the handle must stay usable."""
    from sindslam_b200.capi import SinDyn
    _, frames = seq_c1
    cam = synth.TUM3
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    s.set_prev_frames(frames[0].bgr, frames[0].bgr)
    mask, label = s.detect(frames[0].bgr, frames[0].depth, 1)
    assert set(np.unique(mask)) <= {0, 125, 255}
    mask2, _ = s.detect(frames[1].bgr, frames[1].depth, 2)
    assert set(np.unique(mask2)) <= {0, 125, 255}
    s.close()


def test_small_resolution_320x240():
    from sindslam_b200.capi import Orb, SinDyn
    from oracle import orb_oracle as oo
    cam = synth.CameraConfig(320, 240, 267.7, 269.6, 160.0, 123.8, 5000.0, "TUM3-half")
    _, frames = synth.make_sequence(4, cam, seq=5, kind="box", start=8, hole_rate=0.0003)
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    s.set_prev_frames(frames[0].bgr, frames[0].bgr)
    o = orc.DynaDetectOracle(frames[0].bgr, frames[0].bgr, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=False)
    s2 = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=0)
    s2.set_prev_frames(frames[0].bgr, frames[0].bgr)
    for k in range(1, 4):
        mask, label = s2.detect(frames[k].bgr, frames[k].depth, k)
        fr = s2.flow_results()
        r = o.detect(frames[k].bgr, frames[k].depth, inject_masks=(fr["low"], fr["high"]))
        assert np.array_equal(label, r["label"]) and np.array_equal(mask, r["mask"])
        m1, _ = s.detect(frames[k].bgr, frames[k].depth, k)            # with the plane fitter on: runs, well-formed output
        assert set(np.unique(m1)) <= {0, 125, 255}
    orb = Orb(500, 1.2, 6, 20, 7, cam.width, cam.height)
    gray = cv2.cvtColor(frames[1].bgr, cv2.COLOR_BGR2GRAY)
    kps, desc = orb.extract(gray, None)
    rk, rd = oo.OrbOracle(500, 1.2, 6, 20, 7).extract(gray, None)
    assert len(kps) == len(rk) and np.array_equal(desc, rd)
    s.close(); s2.close(); orb.close()


def test_error_behaviour():
    from sindslam_b200.capi import SinDyn, SindynError
    cam = synth.TUM3
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    img = np.zeros((cam.height, cam.width, 3), np.uint8)
    with pytest.raises(SindynError, match="STATE"):                    # detect before the constructor frames were given
        s.detect(img, np.zeros((cam.height, cam.width), np.uint16), 1)
    with pytest.raises(SindynError, match="INVALID"):                  # flow grid size mismatch
        s.flow_brox(np.zeros((100, 100), np.float32), np.zeros((100, 100), np.float32))
    s.close()
    with pytest.raises(SindynError):                                   # width / height must be multiples of 8
        SinDyn(636, 476, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
