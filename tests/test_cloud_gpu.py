"""Row f3: point-cloud generation of the dense-map consumer (octomap_pub/src/pubPointCloud.cc
generatePointCloud, both overloads) -- CUDA path through the C ABI against oracle/cloud_oracle.py."""
import numpy as np
import pytest

from sindslam_b200 import synth
from oracle import cloud_oracle as oc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    from sindslam_b200.capi import SinDyn
    cam = synth.TUM3
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    yield s
    s.close()


@pytest.fixture(scope="module")
def frames():
    _, fr = synth.make_sequence(6, synth.TUM3, seq=3, kind="box", start=4)
    return fr


INTR = (535.4, 539.2, 320.1, 247.6, 5000.0)   # the node reads doubles from its parameter file


def _labels(rng, H, W, with_invalid=True):
    """12 cluster labels in 3 x 4 cells with ragged borders, plus a stripe of label values >= 12 (skipped by the reference)."""
    yy, xx = np.mgrid[0:H, 0:W]
    lab = ((yy + (20 * np.sin(xx / 37.0)).astype(int)) * 3 // (H + 40)).clip(0, 2) * 4 + ((xx + (15 * np.cos(yy / 23.0)).astype(int)) * 4 // (W + 30)).clip(0, 3)
    lab = lab.astype(np.uint8)
    if with_invalid:
        lab[:, 300:306] = 12
        lab[100:104, :] = 200
    return lab


def _cmp_points(got, xyz, col):
    g = np.stack([got["x"], got["y"], got["z"]], -1)
    assert g.shape == xyz.shape
    assert np.array_equal(np.isnan(g), np.isnan(xyz))
    fin = ~np.isnan(xyz)
    assert np.array_equal(g[fin], xyz[fin])        # same operations in the same order: bit-exact
    assert np.array_equal(np.stack([got["b"], got["g"], got["r"]], -1), col)


def test_cloud_single_frame(sd, frames):
    f = frames[2]
    mask = np.where(f.dyn_mask, 255, 0).astype(np.uint8)
    mask[50:60, 50:90] = 240    # boundary of the >= 240 test
    mask[60:70, 50:90] = 239
    depth = f.depth.copy()
    depth[200:210, 100:140] = 49       # 0.0098 m: below the 0.01 m bound
    depth[210:220, 100:140] = 50001    # above 10 m
    for Twc in (np.eye(4), f.T_wc):
        got = sd.cloud_single(f.bgr, depth, mask, Twc, INTR)
        xyz, col = oc.generate_single(f.bgr, depth, mask, Twc, *INTR)
        assert len(got) == 160 * 214
        _cmp_points(got, xyz, col)
    assert np.isnan(xyz).any() and (~np.isnan(xyz)).any()
    # default intrinsics = the handle's configuration (floats widened to double)
    cam = synth.TUM3
    got = sd.cloud_single(f.bgr, depth, mask, f.T_wc)
    k = [float(np.float32(v)) for v in (cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)]
    xyz, col = oc.generate_single(f.bgr, depth, mask, f.T_wc, *k)
    _cmp_points(got, xyz, col)


@pytest.mark.parametrize("case", ["static_scene", "moving_box", "all_rejected", "big_motion"])
def test_cloud_cross_frame_consistency(sd, frames, case):
    rng = np.random.default_rng(5)
    cur, last = frames[4], frames[2]
    H, W = cur.depth.shape
    label = _labels(rng, H, W)
    mask = np.where(cur.dyn_mask, 255, 0).astype(np.uint8)
    mask_last = np.where(last.dyn_mask, 255, 0).astype(np.uint8)
    T_rel = np.linalg.inv(last.T_wc) @ cur.T_wc       # poseRelative: current camera -> previous key frame
    depth, depth_last = cur.depth, last.depth
    if case == "static_scene":
        mask[:] = 0; mask_last[:] = 0
        depth_last = depth.copy(); T_rel = np.eye(4)
    elif case == "all_rejected":
        mask_last[:] = 255                             # every re-projected pixel votes "dynamic"
    elif case == "big_motion":
        T_rel = T_rel.copy(); T_rel[:3, 3] += (0.4, -0.2, 0.3)   # many pixels leave the previous view
    got = sd.cloud_consistent(cur.bgr, depth, depth_last, mask, mask_last, label, T_rel, cur.T_wc, INTR)
    ref = oc.generate_consistent(cur.bgr, depth, depth_last, mask, mask_last, label, T_rel, cur.T_wc, *INTR)
    print(case, "occlusion", got["occlusion"], "kept", got["kept"].astype(int), "points", len(got["points"]))
    assert np.array_equal(got["occlusion"], ref["occlusion"])
    assert np.array_equal(got["label_count"], ref["label_count"])
    assert np.array_equal(got["kept"], ref["kept"])
    assert np.array_equal(got["mask_new"], ref["mask_new"])
    assert np.array_equal(got["depth_new"], ref["depth_new"])
    _cmp_points(got["points"], ref["xyz"], ref["bgr"])
    if case == "static_scene":
        assert got["kept"].all() and got["occlusion"].sum() < 600   # only row-0 / column-0 samples that re-project to -1e-8
    if case == "all_rejected":
        assert got["kept"][0] and not got["kept"][1:].any()
        assert (got["mask_new"][label < 12][label[label < 12] >= 1] == 255).all()
