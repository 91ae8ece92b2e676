"""GPU parity: sample list (bit-exact) and robust homography (tolerance vs cv2.findHomography RHO)."""
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu

# Maximum difference between the flow induced by our H and by cv2's RHO H, anywhere in the image (px).
# RHO is a randomised PROSAC estimator; both land on the same consensus set and differ by the refinement.
H_FLOW_TOL = 0.15


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


def test_homography_close_to_rho(sd, seq_c1):
    scene, frames = seq_c1
    lab, dyn = _states(seq_c1)
    sd.set_state(2, lab)
    sd.set_state(0, dyn)
    # (a) pure homography flow + 20 % gross outliers ; (b) rendered scene flow (parallax + moving box)
    Ht = np.array([[1.01, 0.004, -3.0], [-0.003, 0.995, 2.0], [1e-5, -2e-5, 1.0]])
    col, row = np.meshgrid(np.arange(640, dtype=np.float64), np.arange(480, dtype=np.float64))
    den = Ht[2, 0] * col + Ht[2, 1] * row + Ht[2, 2]
    fa = np.stack([col - (Ht[0, 0] * col + Ht[0, 1] * row + Ht[0, 2]) / den, row - (Ht[1, 0] * col + Ht[1, 1] * row + Ht[1, 2]) / den], -1).astype(np.float32)
    fa[100:300, 200:330] += np.float32(9.0)
    fb = -synth.gt_flow(scene, synth.TUM3, 10, 8, frames[2])
    for name, flow in (("synthetic-H", fa), ("scene", fb)):
        Hg, n = sd.estimate_homography(flow)
        p, q = orc.sample_pairs(flow, dyn, lab)
        assert n == len(p)
        Hc = orc.estimate_homography(p, q)
        d = np.abs(_h_flow(Hg) - _h_flow(Hc)).max()
        # both should explain the consensus set equally well
        def inl(Hm):
            ph = np.concatenate([p, np.ones((len(p), 1))], 1) @ Hm.T
            e = np.linalg.norm(ph[:, :2] / ph[:, 2:3] - q, axis=1)
            return int((e <= 3.0).sum()), float(np.median(e))
        print(name, "max |Hx_gpu - Hx_rho| = %.4f px; inliers/median err gpu %s rho %s" % (d, inl(Hg), inl(Hc)))
        assert inl(Hg)[0] >= inl(Hc)[0] * 0.98
        assert d <= H_FLOW_TOL
    assert np.abs(_h_flow(sd.estimate_homography(fa)[0]) - _h_flow(Ht)).max() < 0.02
