"""GPU parity of the masked ORB extractor (sindyn_orb_* = ORB_SLAM2::ORBextractor, ORBextractor.cc) against
oracle/orb_oracle.py: pyramid, FAST candidates, distributed keypoints, angles, erased-keypoint sets and descriptors
must be bit-exact."""
import cv2
import numpy as np
import pytest

from oracle import orb_oracle as oo
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu


def _check(orb, o, gray, mask):
    kps, desc = orb.extract(gray, mask)
    dbg = {}
    rk, rd = o.extract(gray, mask, debug=dbg)
    for lvl in range(o.nlevels):
        assert np.array_equal(orb.pyramid_level(lvl), o.level_image(lvl)), ("pyramid", lvl)
    for lvl in range(o.nlevels):
        got = orb.candidates(lvl)
        ref = np.array([[int(x), int(y), int(r)] for x, y, r in dbg["candidates"][lvl]], np.int32).reshape(-1, 3)
        assert got.shape == ref.shape, ("candidate count", lvl, got.shape, ref.shape)
        assert np.array_equal(got, ref), ("candidates", lvl)
    print("keypoints gpu %d oracle %d" % (len(kps), len(rk)))
    assert len(kps) == len(rk)
    g = np.stack([kps["x"], kps["y"], kps["size"], kps["angle"], kps["response"], kps["octave"].astype(np.float32)], 1).astype(np.float64)
    assert np.array_equal(g[:, [0, 1, 2, 4, 5]], rk[:, [0, 1, 2, 4, 5]].astype(np.float32).astype(np.float64)), "keypoint set / order"
    assert np.array_equal(g[:, 3], rk[:, 3].astype(np.float32).astype(np.float64)), "angles"
    assert np.array_equal(desc, rd), int((desc != rd).sum())
    return kps


def test_orb_bit_exact_c1(seq_c1):
    from sindslam_b200.capi import Orb
    _, frames = seq_c1
    o = oo.OrbOracle(1500, 1.2, 8, 15, 5)          # TUM3.yaml
    orb = Orb(1500, 1.2, 8, 15, 5, 640, 480)
    for f in frames[:2]:
        gray = cv2.cvtColor(f.bgr, cv2.COLOR_BGR2GRAY)
        k0 = _check(orb, o, gray, None)
        # dynamic mask as the driver passes it: DetectDynaArea output dilated 15x15 (rgbd_tum_noros.cc:136-139)
        mask = np.where(f.dyn_mask, 255, 125).astype(np.uint8)
        mask = cv2.dilate(mask, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (15, 15)))
        k1 = _check(orb, o, gray, mask)
        assert len(k1) < len(k0)
        # everything masked -> fewer than 250 survivors -> mask ignored
        k2 = _check(orb, o, gray, np.full((480, 640), 255, np.uint8))
        assert len(k2) == len(k0)
    orb.close()


def test_orb_other_parameters_and_flat_image():
    from sindslam_b200.capi import Orb
    rng = np.random.default_rng(3)
    noise = (rng.random((480, 848)) * 255).astype(np.uint8)
    img = cv2.GaussianBlur(noise, (0, 0), 2.5)
    img = np.clip((img.astype(np.int32) - 128) * 5 + 128, 0, 255).astype(np.uint8)
    o = oo.OrbOracle(1000, 1.2, 8, 20, 7)          # TUM1.yaml values on an 848x480 (D455-shaped) frame
    orb = Orb(1000, 1.2, 8, 20, 7, 848, 480)
    _check(orb, o, img, None)
    # white noise: more FAST corners than the fixed-capacity candidate list holds -> a clean CAPACITY error, no crash
    from sindslam_b200.capi import SindynError
    with pytest.raises(SindynError, match="CAPACITY"):
        orb.extract(noise, None)
    flat = np.full((480, 848), 90, np.uint8)        # no corners at all
    k, d = orb.extract(flat, None)
    assert len(k) == 0 and d.shape == (0, 32)
    orb.close()


@pytest.mark.parametrize("dist", [(0.0, 0.0, 0.0, 0.0, 0.0), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)])   # TUM3.yaml / TUM1.yaml
def test_frame_features_after_orb(seq_c1, dist):
    """Frame.cc:143-170 (undistortion, RGB-D stereo coordinates, 64 x 48 grid) on the device-resident
    keypoints, bit-exact against the oracle (real cv2.undistortPoints)."""
    from oracle import frame_oracle as fo
    from sindslam_b200.capi import Orb
    _, frames = seq_c1
    cam = synth.TUM3
    orb = Orb(1000, 1.2, 8, 20, 7, 640, 480)
    gray = cv2.cvtColor(frames[1].bgr, cv2.COLOR_BGR2GRAY)
    kps, _ = orb.extract(gray, None)
    bf, dmf = 40.0, 1.0 / 5000.0
    un, dep, ur, b, off, idx = orb.frame_features(frames[1].depth, cam.fx, cam.fy, cam.cx, cam.cy, dist, bf, dmf)
    r = fo.frame_features(np.stack([kps["x"], kps["y"]], 1), frames[1].depth, 640, 480, cam.fx, cam.fy, cam.cx, cam.cy, dist, bf, dmf)
    assert np.array_equal(b, r["bounds"])
    assert np.array_equal(un, r["keys_un"])
    assert np.array_equal(dep, r["depth"]) and np.array_equal(ur, r["u_right"])
    for i in range(64):
        for j in range(48):
            c = i * 48 + j
            assert list(idx[off[c]:off[c + 1]]) == r["grid"][i][j], (i, j)
    assert (dep > 0).mean() > 0.9 and off[-1] >= len(kps) - 5
    orb.close()
