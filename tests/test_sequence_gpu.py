"""BASELINE configs[2] / configs[3] shaped parity runs: the full sindyn_detect streamed over longer synthetic
sequences -- 640x480 walking_xyz-shaped with the box object (including a frame jump that triggers the large-motion
fallback) and 848x480 D455-shaped with the humanoid-sized dynamic region.  The oracle gets the GPU's own low/high masks
(identical flow), so merged labels, the dynamic mask and the state recurrence must be bit-exact; the masked ORB
keypoint sets are compared with the oracle extractor on the dilated GPU mask."""
import cv2
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from oracle import orb_oracle as oo
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu


def _iou(a, b):
    u = (a | b).sum()
    return 1.0 if u == 0 else float((a & b).sum()) / float(u)


def _stream(cam, frames, order, orb_every=3):
    from sindslam_b200.capi import Orb, SinDyn
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1)
    orb = Orb(1000, 1.2, 8, 20, 7, cam.width, cam.height)
    oorb = oo.OrbOracle(1000, 1.2, 8, 20, 7)
    s.set_prev_frames(frames[order[0]].bgr, frames[order[0]].bgr)     # the driver primes with frame 0 twice (rgbd_tum_noros.cc:103-107)
    o = orc.DynaDetectOracle(frames[order[0]].bgr, frames[order[0]].bgr, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=True)
    n_large, ious = 0, []
    for n, k in enumerate(order[1:], 1):
        f = frames[k]
        mask, label = s.detect(f.bgr, f.depth, n)
        fr = s.flow_results()
        n_large += int(fr["large_motion"])
        r = o.detect(f.bgr, f.depth, inject_masks=(fr["low"], fr["high"]))
        assert np.array_equal(label, r["label"]), (n, int((label != r["label"]).sum()))
        assert np.array_equal(mask, r["mask"]), (n, int((mask != r["mask"]).sum()))
        gt = cv2.dilate(f.dyn_mask.astype(np.uint8), orc.ellipse(9)) > 0
        ious.append(_iou(mask == 255, gt))
        if n % orb_every == 0:
            dil = s.morph_ellipse(mask, 15, 0)
            assert np.array_equal(dil, cv2.dilate(mask, orc.ellipse(15)))
            gray = cv2.cvtColor(f.bgr, cv2.COLOR_BGR2GRAY)
            kps, desc = orb.extract(gray, dil)
            rk, rd = oorb.extract(gray, dil)
            assert len(kps) == len(rk) and np.array_equal(desc, rd)
            assert np.array_equal(np.stack([kps["x"], kps["y"]], 1), rk[:, :2].astype(np.float32))
            # erased-keypoint set (ORBextractor.cc:1063-1092): unless the < 250 fallback restored everything, no key point
            # survives on mask == 255 at the reference's lookup position int(pt_level * s), s = (float)pow(scaleFactor, octave);
            # the output point is pt_level * mvScaleFactor[octave] (:1156-1160), same float product
            if len(kps) >= 250:
                msf = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2))]).astype(np.float32), dtype=np.float32)   # mvScaleFactor
                lx, ly = np.rint(kps["x"] / msf[kps["octave"]]), np.rint(kps["y"] / msf[kps["octave"]])    # FAST positions are integers in level coordinates
                sc = (np.float64(np.float32(1.2)) ** kps["octave"]).astype(np.float32)                      # (float)pow(scaleFactor, octave)
                px, py = (lx.astype(np.float32) * sc).astype(int), (ly.astype(np.float32) * sc).astype(int)
                assert not (dil[py, px] == 255).any()
    s.close()
    orb.close()
    return n_large, ious


def test_c1_stream_with_large_motion_jump():
    cam = synth.TUM3
    _, frames = synth.make_sequence(24, cam, seq=1, kind="box", start=4, hole_rate=0.0005)    # sensor-like holes: the PEAC stage has planes to fit
    order = list(range(0, 8)) + [23] + list(range(9, 12))     # 7 -> 23 -> 9: 14-frame jumps, large-motion fallback expected
    n_large, ious = _stream(cam, frames, order)
    print("large-motion frames:", n_large, "IoU vs rendered truth:", np.round(ious, 3))
    assert n_large >= 1
    assert max(ious) > 0.9


def test_c2_848x480_humanoid_stream():
    cam = synth.D455_848
    _, frames = synth.make_sequence(7, cam, seq=2, kind="humanoid", start=6, hole_rate=0.0005)
    n_large, ious = _stream(cam, frames, list(range(7)))
    print("848x480 humanoid: IoU vs rendered truth:", np.round(ious, 3))


def test_long_stream_invariants():
    """BASELINE configs[2] shape (long walking_xyz-like sequence), at full size through size-independent properties: 120
    frames through sindyn_detect + dilation + masked ORB (graph replay, state recurrence, fixed-capacity lists) must keep
    every invariant of the reference outputs; every 30th frame is additionally checked bit-exactly against the oracle
    fed with the device's state and masks."""
    from sindslam_b200.capi import Orb, SinDyn
    cam = synth.TUM3
    n = 120
    _, frames = synth.make_sequence(n, cam, seq=4, kind="box", start=0, hole_rate=0.002)
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    orb = Orb(1500, 1.2, 8, 15, 5, cam.width, cam.height)
    s.set_prev_frames(frames[0].bgr, frames[0].bgr)
    ious, n_kp = [], []
    for k in range(1, n):
        check = k % 30 == 0
        if check:
            st = [s.get_state(i) for i in range(5)]
        mask, label = s.detect(frames[k].bgr, frames[k].depth, k)
        assert set(np.unique(mask)) <= {0, 125, 255}
        assert int(label.max()) < 128
        if check:
            fr = s.flow_results()
            o = orc.DynaDetectOracle(st[3], st[4], cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=True)
            o.dyna_last, o.high_last, o.label_last = st[0], st[1], st[2]
            s2 = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1)
            s2.set_state(3, st[3]); s2.set_state(4, st[4]); s2.set_state(0, st[0]); s2.set_state(1, st[1]); s2.set_state(2, st[2])
            m2, l2 = s2.detect(frames[k].bgr, frames[k].depth, k)
            f2 = s2.flow_results()
            r = o.detect(frames[k].bgr, frames[k].depth, inject_masks=(f2["low"], f2["high"]))
            assert np.array_equal(m2, r["mask"]) and np.array_equal(l2, r["label"]), k
            s2.close()
        dil = s.morph_ellipse(mask, 15, 0)
        kps, desc = orb.extract(cv2.cvtColor(frames[k].bgr, cv2.COLOR_RGB2GRAY), dil)
        assert 250 <= len(kps) <= 1500 + 8 and desc.shape == (len(kps), 32)
        n_kp.append(len(kps))
        if k >= 3:
            gt = cv2.dilate(frames[k].dyn_mask.astype(np.uint8), orc.ellipse(9)) > 0
            ious.append(_iou(mask == 255, gt))
    print("120-frame stream: median IoU vs rendered truth %.3f (p10 %.3f), keypoints %d..%d" % (
        float(np.median(ious)), float(np.quantile(ious, 0.1)), min(n_kp), max(n_kp)))
    assert np.median(ious) > 0.8
    s.close()
    orb.close()


def test_fuzz_streams_bit_exact():
    """A reduced tools/fuzz_parity.py under the gpu marker: 12 random short sequences (both camera models, box / humanoid
    objects, depth-hole rates 0.03 % - 5 %, every 4th with frame jumps) streamed through sindyn_detect with the PEAC stage ON,
    free-running state on both sides, the oracle fed with the device's low / high masks: merged labels, dynamic masks and the
    masked ORB key points + descriptors must be bit-exact on every frame."""
    n = 0
    for seq in range(2000, 2012):
        kind = ("box", "humanoid")[seq % 2]
        cam = (synth.TUM3, synth.D455_848)[(seq // 2) % 2]
        hole = (0.015, 0.0003, 0.05, 0.001)[seq % 4]
        _, frames = synth.make_sequence(5, cam, seq=seq, kind=kind, start=3 + seq % 5, hole_rate=hole)
        order = list(range(5)) if seq % 4 else [0, 1, 4, 2, 3]
        _stream(cam, frames, order, orb_every=2)
        n += 1
    assert n == 12
