"""CPU oracle of the dense-map consumer's per-keyframe point-cloud generation.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product path is sindslam_b200/csrc/cloud.cu.
Parity unpinned: the reference node (octomap_pub/src/pubPointCloud.cc) needs ROS + PCL + octomap and cannot be built
here, and it ships no tests; this file restates its two `generatePointCloud` overloads line by line.

  generate_single      pubPointCloud.cc:392-470   every 3rd pixel, mask >= 240 or depth outside [0.01, 10] m -> NaN point
  generate_consistent  pubPointCloud.cc:471-678   every 2nd pixel, per-cluster occlusion vote against the previous
                                                  key frame's depth / mask, rejected clusters are painted 255 into the mask

pcl::transformPointCloud (un-vendored PCL 1.10, common/impl/transforms.hpp) is restated as: non-finite points are copied,
finite ones become float(R p + t) evaluated in double (the node passes Eigen::Isometry3d::matrix(), a double matrix).
"""
from __future__ import annotations

import numpy as np

NAN = np.float32(np.nan)


def _transform(pts: np.ndarray, Twc: np.ndarray) -> np.ndarray:
    out = pts.copy()
    fin = np.isfinite(pts).all(axis=1)
    p = pts[fin].astype(np.float64)
    R, t = Twc[:3, :3], Twc[:3, 3]
    q = np.empty_like(p)
    for i in range(3):
        q[:, i] = ((R[i, 0] * p[:, 0] + R[i, 1] * p[:, 1]) + R[i, 2] * p[:, 2]) + t[i]
    out[fin] = q.astype(np.float32)
    return out


def generate_single(bgr, depth, mask, Twc, fx, fy, cx, cy, depth_scale):
    """pubPointCloud.cc:392-470.  Returns (xyz float32 [K,3] in world coordinates, bgr uint8 [K,3]); K = ceil(H/3)*ceil(W/3)."""
    H, W = depth.shape
    m, n = np.meshgrid(np.arange(0, H, 3), np.arange(0, W, 3), indexing="ij")
    d = (depth[m, n].astype(np.float64) * (1.0 / depth_scale)).astype(np.float32)             # :413
    bad = (mask[m, n].astype(np.int32) >= 240) | (d.astype(np.float64) < 0.01) | (d.astype(np.float64) > 10)   # :417-424
    z = d
    x = (n.astype(np.float32) - np.float32(cx)) * z * np.float32(1.0 / fx)                    # :428 (float arithmetic)
    y = (m.astype(np.float32) - np.float32(cy)) * z * np.float32(1.0 / fy)
    xyz = np.stack([x, y, z], -1).astype(np.float32)
    xyz[bad] = NAN
    xyz = xyz.reshape(-1, 3)
    col = bgr[m, n].reshape(-1, 3).copy()
    return _transform(xyz, np.asarray(Twc, np.float64)), col


def generate_consistent(bgr, depth, depth_last, mask, mask_last, label, T_rel, Twc, fx, fy, cx, cy, depth_scale):
    """pubPointCloud.cc:471-678.  Returns dict(xyz, bgr, mask_new, occlusion[12], label_count[12], kept[12], depth_new)."""
    H, W = depth.shape
    T = np.asarray(T_rel, np.float64)
    R, t = T[:3, :3], T[:3, 3]
    m, n = np.meshgrid(np.arange(0, H, 2), np.arange(0, W, 2), indexing="ij")
    inv = 1.0 / depth_scale
    dcur = (depth[m, n].astype(np.float64) * inv).astype(np.float32)                           # :558
    lab = label[m, n].astype(np.int32)
    use = lab < 12                                                                             # :561-565
    # back-projection in float, widened to double by the Eigen::Vector3d constructor (:569)
    px = ((n.astype(np.float32) - np.float32(cx)) * dcur / np.float32(fx)).astype(np.float64)
    py = ((m.astype(np.float32) - np.float32(cy)) * dcur / np.float32(fy)).astype(np.float64)
    pz = dcur.astype(np.float64)
    q = [((R[i, 0] * px + R[i, 1] * py) + R[i, 2] * pz) + t[i] for i in range(3)]
    u = fx * q[0] + cx * q[2]                                                                  # K * (...), :570
    v = fy * q[1] + cy * q[2]
    zz = q[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        xt = (u / zz).astype(np.float32)
        yt = (v / zz).astype(np.float32)
    inside = (yt >= 0.0) & (yt < np.float32(H)) & (xt >= 0.0) & (xt < np.float32(W))          # :582
    iy = np.where(inside, yt, 0).astype(np.int32)
    ix = np.where(inside, xt, 0).astype(np.int32)
    dlast = np.where(inside, (depth_last[iy, ix].astype(np.float64) * inv).astype(np.float32), np.float32(0))
    dyn_last = inside & (mask_last[iy, ix] > 240)
    depth_new = np.zeros((H, W), np.uint16)
    dn = (dlast.astype(np.float64) * depth_scale).astype(np.int64).astype(np.uint16)           # (ushort)(dLast * depthScale), :585
    sel = inside & use
    depth_new[m[sel], n[sel]] = dn[sel]
    rng_ok = (dcur >= 0) & (dcur < 10) & (dlast >= 0) & (dlast < 10)                           # :588
    diff = dcur - dlast
    lhs = (diff * diff).astype(np.float64)
    rhs = (0.13 * dcur.astype(np.float64)) * (0.13 * dcur.astype(np.float64))                  # :599
    occ = use & rng_ok & ((lhs > rhs) | dyn_last)
    occlusion = np.bincount(lab[occ], minlength=12)[:12].astype(np.int64)
    # the point of every sampled pixel (:611-633): double arithmetic here, unlike the single-frame overload
    bad = (mask[m, n].astype(np.int32) >= 240) | (dcur.astype(np.float64) < 0.01) | (dcur.astype(np.float64) > 10)
    z = dcur
    x = ((n.astype(np.float64) - cx) * z.astype(np.float64) / fx).astype(np.float32)
    y = ((m.astype(np.float64) - cy) * z.astype(np.float64) / fy).astype(np.float32)
    xyz = np.stack([x, y, z], -1).astype(np.float32)
    xyz[bad] = NAN
    col = bgr[m, n]
    label_count = np.bincount(label.reshape(-1), minlength=256)[:12].astype(np.int64)
    kept = np.zeros(12, bool)
    mask_new = mask.copy()
    parts_xyz, parts_col = [], []
    for i in range(12):
        if i == 0 or occlusion[i] * 9 <= 0.4 * label_count[i]:                                # :643-657
            kept[i] = True
            s = use & (lab == i)
            parts_xyz.append(xyz[s])
            parts_col.append(col[s])
        else:
            mask_new[label == i] = 255
    xyz_all = np.concatenate(parts_xyz, 0) if parts_xyz else np.zeros((0, 3), np.float32)
    col_all = np.concatenate(parts_col, 0) if parts_col else np.zeros((0, 3), np.uint8)
    return dict(xyz=_transform(xyz_all, np.asarray(Twc, np.float64)), bgr=col_all, mask_new=mask_new, occlusion=occlusion,
                label_count=label_count, kept=kept, depth_new=depth_new)
