"""oracle/matcher_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

CPU re-statement of the frame-to-frame descriptor matching of Tracking::TrackWithMotionModel:
  ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)   ORBmatcher.cc:1328-1470
  ORBmatcher::DescriptorDistance                                                           ORBmatcher.cc:1647-1665
  ORBmatcher::ComputeThreeMaxima                                                           ORBmatcher.cc:1601-1643
  Frame::GetFeaturesInArea                                                                 Frame.cc:398-452
Parity unpinned by the reference (no tests, cannot be built here).  cv::Mat products of CV_32F matrices (un-vendored
OpenCV gemm) are restated as double accumulation rounded to float once (GEMMSingleMul<float, double>).

MapPoint objects are replaced by plain arrays: for every key point i of the last frame `valid[i]` = (mvpMapPoints[i] != NULL
&& !mvbOutlier[i]), `xyz_w[i]` = GetWorldPos(), `desc[i]` = GetDescriptor(), `observed[i]` = (Observations() > 0); for the
current frame `blocked[i2]` = (mvpMapPoints[i2] != NULL && Observations() > 0) on entry (all false in TrackWithMotionModel,
which clears the vector first)."""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
TH_HIGH, HISTO_LENGTH = 100, 30
GRID_COLS, GRID_ROWS = 64, 48

_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def descriptor_distance(a, b):
    return int(_POP[np.bitwise_xor(a, b)].sum())


def _gemv(R, x, t):
    """cv::Mat (3x3 float) * (3x1 float) + (3x1 float): double accumulation, one rounding."""
    R = R.astype(np.float64); x = x.astype(np.float64); t = t.astype(np.float64)
    return np.array([f32((R[i, 0] * x[0] + R[i, 1] * x[1] + R[i, 2] * x[2]) * 1.0 + t[i] * 1.0) for i in range(3)], np.float32)


def features_in_area(cur, x, y, r, min_level, max_level):
    minx, maxx, miny, maxy = [f32(v) for v in cur["bounds"]]
    inv_w = f32(GRID_COLS) / f32(maxx - minx)
    inv_h = f32(GRID_ROWS) / f32(maxy - miny)
    out = []
    c0 = max(0, int(math.floor(f32(f32(f32(x - minx) - r) * inv_w))))
    if c0 >= GRID_COLS:
        return out
    c1 = min(GRID_COLS - 1, int(math.ceil(f32(f32(f32(x - minx) + r) * inv_w))))
    if c1 < 0:
        return out
    r0 = max(0, int(math.floor(f32(f32(f32(y - miny) - r) * inv_h))))
    if r0 >= GRID_ROWS:
        return out
    r1 = min(GRID_ROWS - 1, int(math.ceil(f32(f32(f32(y - miny) + r) * inv_h))))
    if r1 < 0:
        return out
    check = (min_level > 0) or (max_level >= 0)
    off, idx = cur["grid_offsets"], cur["grid_indices"]
    for ix in range(c0, c1 + 1):
        for iy in range(r0, r1 + 1):
            c = ix * GRID_ROWS + iy
            for j in idx[off[c]:off[c + 1]]:
                if check:
                    o = cur["octave"][j]
                    if o < min_level:
                        continue
                    if max_level >= 0 and o > max_level:
                        continue
                dx = f32(cur["keys_un"][j, 0] - x)
                dy = f32(cur["keys_un"][j, 1] - y)
                if abs(dx) < r and abs(dy) < r:
                    out.append(int(j))
    return out


def three_maxima(sizes):
    max1 = max2 = max3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > max1:
            max3, max2, max1 = max2, max1, s
            i3, i2, i1 = i2, i1, i
        elif s > max2:
            max3, max2 = max2, s
            i3, i2 = i2, i
        elif s > max3:
            max3, i3 = s, i
    if max2 < f32(0.1) * f32(max1):
        i2 = i3 = -1
    elif max3 < f32(0.1) * f32(max1):
        i3 = -1
    return i1, i2, i3


def search_by_projection(cur, last, Tcw_cur, Tcw_last, fx, fy, cx, cy, bf, b, scale_factors, th, mono=False, check_orientation=True,
                         blocked=None):
    """cur: dict(keys_un n2x2, octave, angle, u_right, desc n2x32, bounds[4], grid_offsets, grid_indices);
    last: dict(xyz_w n1x3, valid, desc n1x32, octave, angle (undistorted key angle), observed).
    Returns (match[n2] = index of the last-frame point assigned to current key point i2 or -1, nmatches)."""
    Tc = np.asarray(Tcw_cur, np.float32); Tl = np.asarray(Tcw_last, np.float32)
    Rcw, tcw = Tc[:3, :3], Tc[:3, 3]
    Rlw, tlw = Tl[:3, :3], Tl[:3, 3]
    twc = _gemv(-Rcw.T, tcw, np.zeros(3, np.float32))
    tlc = _gemv(Rlw, twc, tlw)
    forward = bool(tlc[2] > f32(b)) and not mono
    backward = bool(-tlc[2] > f32(b)) and not mono
    fx, fy, cx, cy, bf = f32(fx), f32(fy), f32(cx), f32(cy), f32(bf)
    minx, maxx, miny, maxy = [f32(v) for v in cur["bounds"]]
    n2 = len(cur["keys_un"])
    match = np.full(n2, -1, np.int32)
    blk = np.zeros(n2, bool) if blocked is None else np.asarray(blocked, bool).copy()
    factor = f32(1.0) / f32(HISTO_LENGTH)
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    nmatches = 0
    for i in range(len(last["xyz_w"])):
        if not last["valid"][i]:
            continue
        xc, yc, zc = _gemv(Rcw, last["xyz_w"][i].astype(np.float32), tcw)
        with np.errstate(divide="ignore"):
            invzc = f32(np.float64(1.0) / np.float64(zc))
        if invzc < 0:
            continue
        u = f32(f32(f32(fx * xc) * invzc) + cx)
        v = f32(f32(f32(fy * yc) * invzc) + cy)
        if math.isnan(u) or math.isnan(v):
            continue   # zc == 0 with xc == 0: no comparison of the reference rejects NaN and its (int)floor(NaN) is undefined; skipped here
        if u < minx or u > maxx or v < miny or v > maxy:
            continue
        oct_l = int(last["octave"][i])
        radius = f32(f32(th) * f32(scale_factors[oct_l]))
        if forward:
            cand = features_in_area(cur, u, v, radius, oct_l, -1)
        elif backward:
            cand = features_in_area(cur, u, v, radius, 0, oct_l)
        else:
            cand = features_in_area(cur, u, v, radius, oct_l - 1, oct_l + 1)
        if not cand:
            continue
        best, best_i2 = 256, -1
        for i2 in cand:
            if blk[i2]:
                continue
            if cur["u_right"][i2] > 0:
                ur = f32(u - f32(bf * invzc))
                er = abs(f32(ur - cur["u_right"][i2]))
                if er > radius:
                    continue
            d = descriptor_distance(last["desc"][i], cur["desc"][i2])
            if d < best:
                best, best_i2 = d, i2
        if best <= TH_HIGH:
            match[best_i2] = i
            blk[best_i2] = bool(last["observed"][i])
            nmatches += 1
            if check_orientation:
                rot = f32(f32(last["angle"][i]) - f32(cur["angle"][best_i2]))
                if rot < 0.0:
                    rot = f32(rot + f32(360.0))
                x = float(f32(rot * factor))
                bin_ = int(math.floor(x + 0.5)) if x >= 0 else int(math.ceil(x - 0.5))
                if bin_ == HISTO_LENGTH:
                    bin_ = 0
                rot_hist[bin_].append(best_i2)
    if check_orientation:
        i1, i2_, i3 = three_maxima([len(h) for h in rot_hist])
        for i in range(HISTO_LENGTH):
            if i != i1 and i != i2_ and i != i3:
                for j in rot_hist[i]:
                    match[j] = -1
                    nmatches -= 1
    return match, nmatches
