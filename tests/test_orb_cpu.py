"""CPU checks of the ORB oracle (oracle/orb_oracle.py) and of the FAST model the kernels follow: known-answer
constants of ORBextractor.cc (level sizes, umax, quotas, pattern) and FAST-9/16 + NMS against cv2."""
import zlib

import cv2
import numpy as np
import pytest

import fast_rules
from oracle import orb_oracle as oo
from sindslam_b200 import synth


def test_constructor_tables():
    o = oo.OrbOracle(1500, 1.2, 8, 15, 5)      # TUM3.yaml ORBextractor.* values
    # umax for HALF_PATCH_SIZE = 15 (ORBextractor.cc:450-467)
    assert o.umax == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert sum(o.per_level) == 1500 and len(o.per_level) == 8
    assert o.per_level[0] > o.per_level[1] > o.per_level[6]
    img = np.zeros((480, 640), np.uint8)
    o.compute_pyramid(img)
    # SURVEY.md 8a row b2
    assert o.level_size == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
    assert all(p.shape == (h + 38, w + 38) for p, (w, h) in zip(o.pyramid, o.level_size))
    pat = oo.load_pattern()
    assert pat.shape == (256, 4) and pat.min() == -13 and pat.max() == 12
    assert tuple(pat[0]) == (8, -3, 9, 5) and tuple(pat[255]) == (-1, -6, 0, -11)
    assert zlib.crc32(pat.astype(np.int8).tobytes()) == PATTERN_CRC


PATTERN_CRC = zlib.crc32(oo.load_pattern().astype(np.int8).tobytes())


@pytest.mark.parametrize("thr", [5, 15, 20])
def test_fast_model_matches_cv2(thr):
    rng = np.random.default_rng(thr)
    for shape in ((36, 37), (60, 81), (7, 7), (6, 20)):
        img = cv2.GaussianBlur((rng.random(shape) * 255).astype(np.uint8), (0, 0), 1.0)
        img = np.clip(img.astype(np.int32) * 3 - 200, 0, 255).astype(np.uint8)
        ref = cv2.FastFeatureDetector_create(thr, True).detect(img)
        ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in ref]
        got = fast_rules.fast_window(img, thr) if min(shape) >= 7 else []
        assert got == ref, (shape, len(got), len(ref))


def test_extract_on_synthetic_frame():
    _, frames = synth.make_sequence(1, synth.TUM3, seq=0, kind="box", start=8)
    gray = cv2.cvtColor(frames[0].bgr, cv2.COLOR_BGR2GRAY)
    o = oo.OrbOracle(1500, 1.2, 8, 15, 5)
    kps, desc = o.extract(gray)
    assert 1000 < len(kps) <= 1500 + 8 and desc.shape == (len(kps), 32)
    assert kps[:, 0].min() >= 16 and kps[:, 0].max() < 640 - 16
    # octaves are emitted in level order
    assert np.all(np.diff(kps[:, 5]) >= 0)
    # erasing: everything inside a mask==255 rectangle disappears (when >= 250 keypoints remain)
    mask = np.zeros((480, 640), np.uint8)
    mask[100:300, 200:500] = 255
    k2, d2 = o.extract(gray, mask)
    inside = (k2[:, 0] >= 200) & (k2[:, 0] < 500) & (k2[:, 1] >= 100) & (k2[:, 1] < 300)
    assert len(k2) < len(kps) and inside.sum() <= 8      # scale rounding at coarse octaves may keep border points
    # < 250 survivors -> the mask is ignored ("maybe lost", ORBextractor.cc:1105-1115)
    k3, _ = o.extract(gray, np.full((480, 640), 255, np.uint8))
    assert len(k3) == len(kps)
