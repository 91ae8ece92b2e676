"""GPU parity: k-means, cluster ordering, morphology, depth edges through the C ABI vs the oracle."""
import cv2
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    from sindslam_b200.capi import SinDyn
    cam = synth.TUM3
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    yield s
    s.close()


@pytest.mark.parametrize("k", [3, 4, 5, 7, 9, 10, 15])
def test_morph_ellipse_bit_exact(sd, k):
    rng = np.random.default_rng(k)
    img = (cv2.GaussianBlur(rng.random((480, 640)).astype(np.float32), (0, 0), 3) > 0.5).astype(np.uint8) * 255
    img[100:140, 200:260] = 128  # non-binary values (the low-error mask holds 128)
    se = orc.ellipse(k)
    for op, cvop in ((0, cv2.MORPH_DILATE), (1, cv2.MORPH_ERODE), (2, cv2.MORPH_OPEN), (3, cv2.MORPH_CLOSE)):
        got = sd.morph_ellipse(img, k, op)
        ref = cv2.morphologyEx(img, cvop, se)
        assert np.array_equal(got, ref), (k, op, int((got != ref).sum()))


def test_morph_ragged_sizes(sd):
    rng = np.random.default_rng(0)
    for (h, w) in ((1, 1), (7, 5), (33, 131), (479, 637)):
        img = (rng.random((h, w)) > 0.6).astype(np.uint8) * 255
        for k in (4, 9):
            got = sd.morph_ellipse(img, k, 2)
            assert np.array_equal(got, cv2.morphologyEx(img, cv2.MORPH_OPEN, orc.ellipse(k)))


def test_kmeans_first_frame_bit_exact_vs_fx(sd, seq_c1):
    _, frames = seq_c1
    cam = synth.TUM3
    sd.set_state(2, np.zeros((480, 640), np.uint8))
    for f in frames[:2]:
        lab, pts, ctr = sd.kmeans(f.depth)
        olab, opts, octr = orc.seg_by_kmeans(f.depth, np.zeros((480, 640), np.uint8), cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, "fx")
        assert np.array_equal(pts, opts)
        print("label mismatches vs oracle-fx:", int((lab != olab).sum()), "centre maxdiff", float(np.abs(ctr - octr).max()))
        assert np.array_equal(lab, olab)
        assert np.array_equal(ctr, octr)
        cvlab, _, _ = orc.seg_by_kmeans(f.depth, np.zeros((480, 640), np.uint8), cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, "cv2")
        agree = (lab == cvlab).mean()
        print("agreement with cv2.kmeans (sequential float32 sums): %.6f" % agree)
        assert agree > 0.999
        img, order = sd.cluster_order()
        okept, oseg, _ = orc.cluster_order(olab, octr)
        assert list(order) == okept
        assert np.array_equal(img, oseg)


def test_kmeans_with_previous_labels_and_empty_cluster_repair(sd, seq_c1):
    _, frames = seq_c1
    cam = synth.TUM3
    lab0, _, _ = orc.seg_by_kmeans(frames[0].depth, np.zeros((480, 640), np.uint8), cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, "fx")
    prev = (lab0 % 5).astype(np.uint8)  # merged-label-like image with ids 0..4: seven clusters start empty
    sd.set_state(2, prev)
    lab, pts, ctr = sd.kmeans(frames[1].depth)
    olab, _, octr = orc.seg_by_kmeans(frames[1].depth, prev, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, "fx")
    print("repair case label mismatches:", int((lab != olab).sum()))
    assert np.array_equal(lab, olab)
    assert np.array_equal(ctr, octr)
    sd.set_state(2, np.zeros((480, 640), np.uint8))


def test_depth_edges_bit_exact(sd, seq_c1):
    _, frames = seq_c1
    cam = synth.TUM3
    for f in frames[:3]:
        ta, ge, ep = sd.depth_edges(f.depth)
        ota, oge, oep = orc.depth_edges(f.depth, cam.depth_factor)
        print("edges px", int((oge > 0).sum()), "endpoints", len(oep), len(ep))
        assert np.array_equal(ta, ota)
        assert np.array_equal(ge, oge)
        assert np.array_equal(ep, oep)
