"""oracle/frame_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

CPU re-statement of what ORB_SLAM2::Frame does with the keypoints right after ORB extraction
(reference: ORB_SLAM2/src/Frame.cc:143-170): UndistortKeyPoints (:714-753, real cv2.undistortPoints),
ComputeImageBounds (:507-535), ComputeStereoFromRGBD (:284-298 of this fork: depth lookup + virtual right coordinate),
PosInGrid / AssignFeaturesToGrid (:453-463, 58-72; 64 x 48 cells).  SURVEY.md 8(f) row f2.  Parity unpinned by the
reference (no tests); pinned by the real OpenCV undistortPoints."""
from __future__ import annotations

import cv2
import numpy as np

GRID_COLS, GRID_ROWS = 64, 48       # Frame.h:37-38
f32 = np.float32


def c_round(x):
    """C round(): half away from zero."""
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)


def frame_features(kps_xy, depth_raw, width, height, fx, fy, cx, cy, dist, bf, depth_map_factor):
    """kps_xy: n x 2 float32 (distorted keypoints as ORBextractor returns them); depth_raw: H x W u16;
    depth_map_factor = 1 / DepthMapFactor (Tracking.cc:142-146).  Returns dict(keys_un, depth, u_right, bounds, grid)."""
    kps_xy = np.ascontiguousarray(kps_xy, np.float32).reshape(-1, 2)
    n = len(kps_xy)
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    dist = np.asarray(dist, np.float32)
    if dist[0] == 0.0:
        un = kps_xy.copy()
        bounds = (f32(0), f32(width), f32(0), f32(height))
    else:
        un = cv2.undistortPoints(kps_xy.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2) if n else kps_xy.copy()
        c = np.array([[0, 0], [width, 0], [0, height], [width, height]], np.float32)
        cu = cv2.undistortPoints(c.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
        bounds = (min(cu[0, 0], cu[2, 0]), max(cu[1, 0], cu[3, 0]), min(cu[0, 1], cu[1, 1]), max(cu[2, 1], cu[3, 1]))
    # imDepth.convertTo(CV_32F, mDepthMapFactor) then imDepth.at<float>(v, u) with truncated float indices
    v, u = kps_xy[:, 1].astype(np.int64), kps_xy[:, 0].astype(np.int64)
    d = depth_raw[v, u].astype(np.float32) * f32(depth_map_factor)
    depth = np.where(d > 0, d, f32(-1)).astype(np.float32)
    with np.errstate(divide="ignore"):
        u_right = np.where(d > 0, un[:, 0] - f32(bf) / d, f32(-1)).astype(np.float32)
    inv_w = f32(GRID_COLS) / f32(bounds[1] - bounds[0])
    inv_h = f32(GRID_ROWS) / f32(bounds[3] - bounds[2])
    px = c_round(((un[:, 0] - f32(bounds[0])) * inv_w).astype(np.float32).astype(np.float64))
    py = c_round(((un[:, 1] - f32(bounds[2])) * inv_h).astype(np.float32).astype(np.float64))
    ok = (px >= 0) & (px < GRID_COLS) & (py >= 0) & (py < GRID_ROWS)
    grid = [[[] for _ in range(GRID_ROWS)] for _ in range(GRID_COLS)]
    for i in np.nonzero(ok)[0]:
        grid[int(px[i])][int(py[i])].append(int(i))
    return dict(keys_un=un.astype(np.float32), depth=depth, u_right=u_right, bounds=np.array(bounds, np.float32), grid=grid)
