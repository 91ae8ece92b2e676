"""Oracle of row f3 (oracle/cloud_oracle.py): known answers on tiny hand-checkable inputs, no GPU."""
import numpy as np

from oracle import cloud_oracle as oc

K = (500.0, 500.0, 4.0, 3.0, 1000.0)


def test_single_known_answers():
    H, W = 6, 9
    depth = np.full((H, W), 2000, np.uint16)          # 2 m
    bgr = np.zeros((H, W, 3), np.uint8); bgr[..., 0] = np.arange(W); bgr[..., 1] = np.arange(H)[:, None]
    mask = np.zeros((H, W), np.uint8)
    mask[0, 3] = 240; depth[3, 0] = 5; depth[3, 6] = 20000
    xyz, col = oc.generate_single(bgr, depth, mask, np.eye(4), *K)
    assert xyz.shape == (2 * 3, 3)                    # rows 0,3 x cols 0,3,6
    # (m, n) = (0, 0): x = (0 - 4) * 2 / 500, y = (0 - 3) * 2 / 500
    assert np.allclose(xyz[0], [-0.016, -0.012, 2.0], atol=1e-7)
    assert np.isnan(xyz[1]).all()                     # mask >= 240
    assert np.isnan(xyz[3]).all()                     # 0.005 m < 0.01 m
    assert np.isnan(xyz[5]).all()                     # 20 m > 10 m
    assert np.array_equal(col[4], [3, 3, 0])
    # rigid transform: translation only moves finite points
    T = np.eye(4); T[:3, 3] = (1, 2, 3)
    xyz2, _ = oc.generate_single(bgr, depth, mask, T, *K)
    assert np.allclose(xyz2[0], xyz[0] + (1, 2, 3), atol=1e-6) and np.isnan(xyz2[1]).all()


def test_consistent_static_scene_keeps_everything():
    H, W = 8, 10
    depth = np.full((H, W), 1500, np.uint16)
    bgr = np.zeros((H, W, 3), np.uint8)
    mask = np.zeros((H, W), np.uint8)
    label = (np.arange(W)[None, :] // 4 + np.zeros((H, 1), int)).astype(np.uint8)   # labels 0, 1, 2
    r = oc.generate_consistent(bgr, depth, depth, mask, mask, label, np.eye(4), np.eye(4), *K)
    # (samples of row 0 / column 0 can re-project to -1e-8 after the float back-projection and count as "left the view",
    #  exactly as in the reference)
    assert r["occlusion"][1:].sum() == 0 and r["occlusion"][0] <= 4 + 5 and r["kept"].all()
    assert len(r["xyz"]) == 4 * 5 and np.array_equal(r["mask_new"], mask)
    assert np.array_equal(r["label_count"][:3], [32, 32, 16])
    # cluster order then raster order: the first points all belong to label 0 (columns 0 and 2)
    assert np.allclose(r["xyz"][0, :2], [(0 - 4) * 1.5 / 500, (0 - 3) * 1.5 / 500])
    assert np.array_equal(r["depth_new"][2::2, 2::2], depth[2::2, 2::2])


def test_consistent_rejects_moved_cluster_and_skips_invalid_labels():
    H, W = 8, 10
    depth = np.full((H, W), 1500, np.uint16)
    depth_last = depth.copy(); depth_last[:, 4:8] = 3000      # cluster 1 was 1.5 m further away in the previous key frame
    bgr = np.zeros((H, W, 3), np.uint8)
    mask = np.zeros((H, W), np.uint8)
    label = (np.arange(W)[None, :] // 4 + np.zeros((H, 1), int)).astype(np.uint8)
    label[:, 9] = 12                                           # skipped: no vote, no point
    r = oc.generate_consistent(bgr, depth, depth_last, mask, mask, label, np.eye(4), np.eye(4), *K)
    assert r["occlusion"][1] == 8 and not r["kept"][1] and r["kept"][0] and r["kept"][2]
    assert (r["mask_new"][:, 4:8] == 255).all() and (r["mask_new"][:, :4] == 0).all()
    assert len(r["xyz"]) == 4 * 2 + 4 * 1                      # label 0: cols 0, 2; label 2: col 8
    # a dynamic previous mask alone is enough to vote
    ml = mask.copy(); ml[:, 0:4] = 255
    r2 = oc.generate_consistent(bgr, depth, depth, mask, ml, label, np.eye(4), np.eye(4), *K)
    assert r2["occlusion"][0] >= 6 and r2["kept"][0]           # cluster 0 is always kept
