"""End-to-end, UN-INJECTED parity on the configurations the metric names (BASELINE.json configs[2] and configs[3]):

sindyn_detect streams the whole synthetic sequence with PEAC plane edges on and its own free-running state recurrence
(DynaDetect.cc:1377-1666, driver loop rgbd_tum_noros.cc:113-139), and is compared per frame with the CPU oracle streamed
over the same frames with ITS OWN flow (oracle/brox_cpu.c), the real cv2.VariationalRefinement, the real
cv2.findHomography(RHO) (DynaDetect.cc:1235), PEAC on and its own free-running state.  Nothing the GPU computes is fed to
the oracle: its run is independent of the device, so it is generated offline by tools/make_e2e_golden.py (committed,
~3 s per frame on the CPU) and stored in tests/golden/e2e_*.npz together with a checksum of every input frame.

Gate (SURVEY.md 8d "Parity gates", north_star): dynamic-mask IoU(255 class) >= 0.99 per frame.  Frames on which the
oracle's own 255 region is tiny (< 400 px: IoU of two near-empty sets is meaningless) are gated on the absolute number of
differing pixels instead.  Label agreement and the large-motion decisions are reported and gated as well.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from sindslam_b200 import synth

pytestmark = pytest.mark.gpu

IOU_MIN = 0.99            # north_star: dynamic mask IoU >= 0.99 against the reference's mask
SMALL_REGION_PX = 400     # below this many oracle 255-pixels the frame is gated on |A xor B| <= SMALL_DIFF_PX instead
SMALL_DIFF_PX = 40


def _iou(a, b):
    u = int((a | b).sum())
    return 1.0 if u == 0 else float((a & b).sum()) / float(u)


def _run(name):
    import make_e2e_golden as g
    from sindslam_b200.capi import SinDyn
    path = os.path.join(ROOT, "tests", "golden", "e2e_%s.npz" % name)
    assert os.path.exists(path), "golden oracle run missing: python tools/make_e2e_golden.py " + name
    z = np.load(path)
    cam_name, kind, seq, n_frames, hole = g.CONFIGS[name]
    cam = getattr(synth, cam_name)
    n = len(z["mask"]) + 1
    _, frames = synth.make_sequence_parallel(n, cam, seq=seq, kind=kind, start=0, hole_rate=hole)
    for k in (0, 1, n // 2, n - 1):     # the frames the oracle saw are the frames rendered here
        assert g.frame_crc(frames[k]) == int(z["crc"][k]), "synthetic frame %d differs from the one the golden oracle run used" % k
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1)
    s.set_prev_frames(frames[0].bgr, frames[0].bgr)          # rgbd_tum_noros.cc:103-107
    ious, bad, lab_agree, lm_diff, worst = [], [], [], 0, (1.0, -1)
    for k in range(1, n):
        mask, label = s.detect(frames[k].bgr, frames[k].depth, k)
        fr = s.flow_results()
        om, ol = z["mask"][k - 1], z["label"][k - 1]
        a, b = mask == 255, om == 255
        iou = _iou(a, b)
        nd = int((a ^ b).sum())
        ok = (nd <= SMALL_DIFF_PX) if int(b.sum()) < SMALL_REGION_PX else (iou >= IOU_MIN)
        ious.append(iou)
        lab_agree.append(float((label == ol).mean()))
        lm_diff += int(bool(fr["large_motion"]) != bool(z["large_motion"][k - 1]))
        if iou < worst[0]:
            worst = (iou, k)
        if not ok:
            bad.append((k, round(iou, 4), nd, int(b.sum())))
        if not ok or k % 25 == 0:
            dH = float(np.abs(fr["H"] / fr["H"][2, 2] - z["H"][k - 1] / z["H"][k - 1][2, 2]).max())
            print("%s frame %3d: IoU(255) %.4f  xor %5d px (oracle %6d px)  label agreement %.4f  thr gpu %s oracle %s  max|dH| %.2e" % (
                name, k, iou, nd, int(b.sum()), lab_agree[-1], np.round(fr["thr"], 2), np.round(z["thr"][k - 1], 2), dH))
    s.close()
    ious = np.array(ious)
    print("%s: %d frames, IoU(255) min %.4f (frame %d) p1 %.4f median %.4f; frames below the gate: %d; exact-mask frames %d; "
          "label agreement median %.4f min %.4f; large-motion decisions differing: %d" % (
              name, n - 1, ious.min(), worst[1], float(np.quantile(ious, 0.01)), float(np.median(ious)), len(bad), int((ious == 1.0).sum()),
              float(np.median(lab_agree)), float(np.min(lab_agree)), lm_diff))
    return bad, lm_diff


def test_e2e_c3_300_frames_uninjected():
    """configs[2]: 300-frame walking_xyz-shaped 640x480 sequence."""
    bad, lm_diff = _run("c3")
    assert lm_diff == 0
    assert not bad, bad


def test_e2e_c4_848x480_humanoid_uninjected():
    """configs[3]: 848x480 D455-shaped sequence with a humanoid-sized dynamic region."""
    bad, lm_diff = _run("c4")
    assert lm_diff == 0
    assert not bad, bad
