"""GPU parity: sample list (bit-exact) and robust homography: k_rho restates cv::findHomography(..., RHO) (DynaDetect.cc:1235,
OpenCV calib3d rho.cpp) operation by operation and must return the library's H BIT FOR BIT on the same ordered sample list
(and the same inlier mask)."""
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu

import cv2


@pytest.fixture(scope="module")
def sd():
    from sindslam_b200.capi import SinDyn
    cam = synth.TUM3
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=0)
    yield s
    s.close()


def _states(seq_c1):
    _, frames = seq_c1
    H, W = 480, 640
    rng = np.random.default_rng(5)
    lab = (rng.integers(0, 7, (H // 40, W // 40)).repeat(40, 0).repeat(40, 1)).astype(np.uint8)
    dyn = np.full((H, W), 125, np.uint8)
    dyn[:40] = 0
    dyn[frames[1].dyn_mask] = 255
    return lab, dyn


def _h_flow(Hm, W=640, H=480):
    col, row = np.meshgrid(np.arange(0, W, 8, dtype=np.float64), np.arange(0, H, 8, dtype=np.float64))
    den = Hm[2, 0] * col + Hm[2, 1] * row + Hm[2, 2]
    return np.stack([(Hm[0, 0] * col + Hm[0, 1] * row + Hm[0, 2]) / den, (Hm[1, 0] * col + Hm[1, 1] * row + Hm[1, 2]) / den], -1)


def test_sample_pairs_bit_exact(sd, seq_c1):
    scene, frames = seq_c1
    flow = -synth.gt_flow(scene, synth.TUM3, 10, 8, frames[2])
    flow[:30, :30] = 200.0   # push some samples out of the image (exercises the inBorder filter)
    flow[-40:, -40:] = -35.0
    for lab, dyn in (_states(seq_c1), (np.zeros((480, 640), np.uint8), np.zeros((480, 640), np.uint8))):
        sd.set_state(2, lab)
        sd.set_state(0, dyn)
        p, q = sd.sample_pairs(flow)
        op, oq = orc.sample_pairs(flow, dyn, lab)
        assert len(p) == len(op) and len(p) < 2961
        assert np.array_equal(p, op)
        assert np.array_equal(q, oq)


def test_homography_equals_cv2_rho_on_sample_lists(sd, seq_c1):
    scene, frames = seq_c1
    lab, dyn = _states(seq_c1)
    sd.set_state(2, lab)
    sd.set_state(0, dyn)
    # (a) pure homography flow + gross outliers ; (b) rendered scene flow (parallax + moving box) ; (c) the same with noise
    Ht = np.array([[1.01, 0.004, -3.0], [-0.003, 0.995, 2.0], [1e-5, -2e-5, 1.0]])
    col, row = np.meshgrid(np.arange(640, dtype=np.float64), np.arange(480, dtype=np.float64))
    den = Ht[2, 0] * col + Ht[2, 1] * row + Ht[2, 2]
    fa = np.stack([col - (Ht[0, 0] * col + Ht[0, 1] * row + Ht[0, 2]) / den, row - (Ht[1, 0] * col + Ht[1, 1] * row + Ht[1, 2]) / den], -1).astype(np.float32)
    fa[100:300, 200:330] += np.float32(9.0)
    fb = -synth.gt_flow(scene, synth.TUM3, 10, 8, frames[2])
    fc = fb + np.random.default_rng(3).normal(0, 0.3, fb.shape).astype(np.float32)
    for name, flow in (("synthetic-H", fa), ("scene", fb), ("scene + noise", fc)):
        Hg, n = sd.estimate_homography(flow)
        p, q = orc.sample_pairs(flow, dyn, lab)
        assert n == len(p)
        Hc = orc.estimate_homography(p, q)          # the real cv2.findHomography(p, q, cv2.RHO)
        print(name, "max |H_gpu - H_cv2| = %.3g" % np.abs(Hg - Hc).max())
        assert np.array_equal(Hg, Hc)
    assert np.abs(_h_flow(sd.estimate_homography(fa)[0]) - _h_flow(Ht)).max() < 0.02


def test_rho_equals_cv2_on_random_correspondences(sd):
    """sindyn_find_homography_rho vs cv2.findHomography(..., cv2.RHO): random ordered lists (grid and scattered sources, 5 to
    2961 points, 0-60 % outliers, 0-0.5 px noise): same inlier mask, bit-identical H; degenerate inputs fail the same way."""
    rng = np.random.default_rng(11)
    n_checked = 0
    for t in range(120):
        N = int(rng.choice([5, 6, 8, 12, 30, 50, 200, 300, 1000, 1500, 2961]))
        outl = float(rng.choice([0, 0.1, 0.3, 0.5, 0.6]))
        noise = float(rng.choice([0, 0.02, 0.1, 0.5]))
        if t % 2 == 0:
            g = np.stack(np.meshgrid(np.arange(10, 640, 10), np.arange(10, 480, 10)), -1).reshape(-1, 2).astype(np.float32)
            g = g[rng.permutation(len(g))[:N]]
        else:
            g = (rng.random((N, 2)) * [640, 480]).astype(np.float32)
        Ht = np.eye(3) + rng.normal(0, 1, (3, 3)) * [[0.01, 0.01, 3], [0.01, 0.01, 3], [1e-5, 1e-5, 0]]
        ph = np.concatenate([g, np.ones((N, 1))], 1) @ Ht.T
        d = (ph[:, :2] / ph[:, 2:]).astype(np.float32) + rng.normal(0, noise, (N, 2)).astype(np.float32)
        k = int(outl * N)
        idx = rng.permutation(N)[:k]
        d[idx] += rng.normal(0, 30, (k, 2)).astype(np.float32)
        g, d = np.ascontiguousarray(g, np.float32), np.ascontiguousarray(d, np.float32)
        Hc, mc = cv2.findHomography(g, d, cv2.RHO)
        Hg, mg, info = sd.find_homography_rho(g, d)
        if Hc is None:
            assert Hg is None, (t, N, outl, noise)
            continue
        assert Hg is not None, (t, N, outl, noise)
        assert np.array_equal(mg, mc.ravel()), (t, N, outl, noise, int(mg.sum()), int(mc.sum()))
        assert np.array_equal(Hg, Hc), (t, N, outl, noise, float(np.abs(Hg - Hc).max()), info.tolist())
        n_checked += 1
    assert n_checked >= 100
    # all points on one line: no model survives the degeneracy tests -> failure, like the library
    line = np.stack([np.arange(40, dtype=np.float32) * 7, np.arange(40, dtype=np.float32) * 3], 1)
    Hc, _ = cv2.findHomography(line, line + 1, cv2.RHO)
    Hg, _, _ = sd.find_homography_rho(line, line + 1)
    assert (Hc is None) == (Hg is None)
