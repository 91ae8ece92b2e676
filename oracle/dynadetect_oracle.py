"""oracle/dynadetect_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path is libsindyn_cuda.so.

CPU re-statement of ORB_SLAM2::DynaDetect (reference: ORB_SLAM2/src/DynaDetect.cc,
ORB_SLAM2/include/DynaDetect.h) in Python + numpy, calling cv2 (opencv-python-headless 4.13)
for every OpenCV primitive the reference calls.  Each function cites the reference lines it follows.

Pinning status: the reference ships NO tests or golden vectors for this
path and cannot be compiled here (needs OpenCV 4.2.0 C++ + contrib, PCL, Eigen), so the glue
logic restated here is PARITY UNPINNED by the reference itself.  What IS pinned: every OpenCV
primitive is executed by the real OpenCV (cv2 4.13; the reference pins 4.2.0 -- documented
drift), and the committed fixtures under tests/golden/ freeze this oracle's outputs.
The dense-flow engines (cv::cuda::BroxOpticalFlow / optflow DeepFlow) are absent from cv2-headless:
flow comes from oracle/brox_cpu.c (restated published algorithm) and everything downstream is
checked with identical injected flow, like the authors' own .flo hook (DynaDetect.cc:1149-1158).
"""
from __future__ import annotations

import ctypes
import os

import cv2
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

NUM_CLUSTER = 12          # DynaDetect.cc:46-47
DEPTH_WEIGHT = 1.5        # DynaDetect.cc:48
SCALE_ELEMENT = 0.6       # DynaDetect.cc:1033


def ellipse(k):
    """DynaDetect.cc:51-59 elementN = getStructuringElement(MORPH_ELLIPSE, Size(N, N))."""
    return cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))


# ----------------------------------------------------------------------------- CPU Brox
_brox = None


def _brox_lib():
    global _brox
    if _brox is None:
        path = os.path.join(_HERE, "_build", "libbrox_cpu.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE])
        lib = ctypes.CDLL(path)
        lib.brox_flow_cpu.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_float] * 3 \
            + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_void_p]
        lib.brox_flow_cpu.restype = ctypes.c_int
        lib.brox_cpu_threads.restype = ctypes.c_int
        _brox = lib
    return _brox


def brox_flow(I0, I1, alpha=0.197, gamma=50.0, scale=0.8, inner=10, outer=77, solver=10, omega=1.99):
    """cv::cuda::BroxOpticalFlow::create(0.197, 50, 0.8, 10, 77, 10)->calc(I0, I1) (DynaDetect.cc:1029,1072)."""
    I0 = np.ascontiguousarray(I0, np.float32)
    I1 = np.ascontiguousarray(I1, np.float32)
    h, w = I0.shape
    out = np.zeros((h, w, 2), np.float32)
    _brox_lib().brox_flow_cpu(I0.ctypes.data, I1.ctypes.data, w, h, alpha, gamma, scale, inner, outer, solver, omega, out.ctypes.data)
    return out


def brox_threads():
    return int(_brox_lib().brox_cpu_threads())


# ----------------------------------------------------------------------------- flow-branch prologue
def bgr2gray(bgr):
    """DynaDetect.cc:1390-1392."""
    return cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)


def flow_size(W, H):
    """cv::Size(scale_element * width, scale_element * height) truncates (DynaDetect.cc:1037)."""
    s = np.float32(SCALE_ELEMENT)
    return int(s * np.float32(W)), int(s * np.float32(H))


def gray_small(gray):
    """DynaDetect.cc:1037-1039 (cv::resize default INTER_LINEAR)."""
    H, W = gray.shape
    return cv2.resize(gray, flow_size(W, H))


def flow_magnitude_hist_large_motion(flow_neg, W, H):
    """Large-motion test on the NEGATED small flow (DynaDetect.cc:1080-1114). Returns (largeMotion, endFlow, endFlow2)."""
    mag, _ = cv2.cartToPolar(flow_neg[..., 0], flow_neg[..., 1], angleInDegrees=True)
    max_flow = float(mag.max())
    m8 = scale_to_u8(mag, max_flow)
    hist = np.bincount(m8.ravel(), minlength=256).astype(np.float32)
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        end_flow_f = np.float32(10.0) * np.float32(SCALE_ELEMENT) * np.float32(255.0) / np.float64(max_flow)
    end_flow = int(end_flow_f) if np.isfinite(end_flow_f) else 2 ** 31 - 1
    total = np.float32(np.float32(np.float32(W * H) * np.float32(SCALE_ELEMENT)) * np.float32(SCALE_ELEMENT))
    ratio = np.float32(0)
    end_flow2 = 0
    for i in range(255):
        ratio = np.float32(ratio + hist[i])
        if ratio > np.float32(0.3) * total:
            end_flow2 = i
            break
    return end_flow2 > end_flow, end_flow, end_flow2


def scale_to_u8(mag, max_val):
    """`img * (255.0/max)` then convertTo(CV_8UC1) (DynaDetect.cc:1090-1091, 1281-1282): float multiply by
    (float)(255.0/max), then round-half-even with saturation."""
    scale = np.float32(255.0 / float(max_val))
    v = mag.astype(np.float32) * scale
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def upsample_flow(flow_small, W, H):
    """DynaDetect.cc:1144-1147."""
    f = cv2.resize(flow_small, (W, H))
    return f * np.float32(1.0 / np.float32(SCALE_ELEMENT))


# ----------------------------------------------------------------------------- residual + thresholds
def homography_residual(flow, Hm):
    """DynaDetect.cc:1252-1271: flow - (x - Hx) in double, stored float; magnitude via cartToPolar."""
    H, W = flow.shape[:2]
    col, row = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    h = np.asarray(Hm, np.float64).ravel()
    den = h[6] * col + h[7] * row + h[8]
    fx2 = col - (h[0] * col + h[1] * row + h[2]) / den
    fy2 = row - (h[3] * col + h[4] * row + h[5]) / den
    rx = flow[..., 0] - fx2.astype(np.float32)
    ry = flow[..., 1] - fy2.astype(np.float32)
    mag, _ = cv2.cartToPolar(np.ascontiguousarray(rx), np.ascontiguousarray(ry), angleInDegrees=True)
    return mag


def pose_residual(flow, depth, T_old_cur, fx, fy, cx, cy, depth_scale):
    """north_star variant: predicted flow from depth back-projection + SE(3) pose (interface extension;
    the reference uses the homography)."""
    H, W = flow.shape[:2]
    T = np.asarray(T_old_cur, np.float64)
    col, row = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    z = (depth.astype(np.float32) * np.float32(1.0 / np.float32(depth_scale))).astype(np.float64)
    X = (col - np.float64(np.float32(cx))) * z / np.float64(np.float32(fx))
    Y = (row - np.float64(np.float32(cy))) * z / np.float64(np.float32(fy))
    xo = T[0, 0] * X + T[0, 1] * Y + T[0, 2] * z + T[0, 3]
    yo = T[1, 0] * X + T[1, 1] * Y + T[1, 2] * z + T[1, 3]
    zo = T[2, 0] * X + T[2, 1] * Y + T[2, 2] * z + T[2, 3]
    ok = (depth != 0) & (zo > 1e-6)
    with np.errstate(divide="ignore", invalid="ignore"):
        fx2 = col - (np.float64(np.float32(fx)) * xo / zo + np.float64(np.float32(cx)))
        fy2 = row - (np.float64(np.float32(fy)) * yo / zo + np.float64(np.float32(cy)))
    rx = np.where(ok, flow[..., 0] - fx2.astype(np.float32), np.float32(0)).astype(np.float32)
    ry = np.where(ok, flow[..., 1] - fy2.astype(np.float32), np.float32(0)).astype(np.float32)
    mag, _ = cv2.cartToPolar(np.ascontiguousarray(rx), np.ascontiguousarray(ry), angleInDegrees=True)
    return mag


def threshold_masks(mag):
    """DynaDetect.cc:1276-1367. Returns (low {0,128}, high {0,255}, [otsu, triangle, t_low, t_high], m8)."""
    H, W = mag.shape
    max_err = float(mag.max())
    max_f = np.float32(max_err)
    m8 = scale_to_u8(mag, max_err)
    t1, _ = cv2.threshold(m8, 80, 255, cv2.THRESH_OTSU)
    t2, _ = cv2.threshold(m8, 80, 255, cv2.THRESH_TRIANGLE)
    otsu, tri = np.float32(t1), np.float32(t2)
    thred1, thred2 = otsu, tri
    f = np.float32
    u17 = f(f(1.7) * f(255.0)) / max_f
    u30 = f(f(3.0) * f(255.0)) / max_f
    u02 = f(f(0.2) * f(255.0)) / max_f
    u100 = f(f(10.0) * f(255.0)) / max_f
    if thred1 < thred2:
        if thred1 < u17:
            thred1 = u17
        elif thred1 > u30:
            thred1 = u30
        if np.count_nonzero(m8 > thred1) > 0.5 * W * H:
            thred1 = f(thred1 + u02)
        m = max(u30, f(thred1 * f(1.2)))
        if thred2 < m:
            thred2 = m
        elif thred2 > u100:
            thred2 = u100
        tl, th = thred1, thred2
    else:
        if thred2 < u17:
            thred2 = u17
        elif thred2 > u30:
            thred2 = u30
        # DynaDetect.cc:1348: countNonZero(thred2) on the scalar -> the bump is dead code
        m = max(u30, f(thred2 * f(1.2)))
        if thred1 < m:
            thred1 = m
        elif thred1 > u100:
            thred1 = u100
        tl, th = thred2, thred1
    low = np.where(m8 > tl, 128, 0).astype(np.uint8)     # imgThhd * 0.5 -> 127.5 -> 128
    high = np.where(m8 > th, 255, 0).astype(np.uint8)
    return low, high, np.array([otsu, tri, tl, th], np.float32), m8


# ----------------------------------------------------------------------------- SegByKmeans
KM_FIX = 36  # fixed-point fraction bits of the order-independent centre accumulation ("oracle-fx")


def depth_pyramid(depth, levels=4):
    """DynaDetect.cc:324-337: successive cv::resize (default INTER_LINEAR) of the u16 image by 0.5."""
    pyr = [depth]
    for _ in range(1, levels):
        p = pyr[-1]
        pyr.append(cv2.resize(p, (int(p.shape[1] * np.float32(0.5)), int(p.shape[0] * np.float32(0.5)))))
    return pyr


def backproject_level(depth_lvl, s, fx, fy, cx, cy, depth_scale):
    """DynaDetect.cc:347-369 (float32 arithmetic, ushort truncation of depth*scale)."""
    f = np.float32
    s = f(s)
    h, w = depth_lvl.shape
    d = (depth_lvl.astype(np.float32) * s).astype(np.uint16)              # ushort depth = pyr * scales[level]
    df = d.astype(np.float32)
    invalid = ((df / f(depth_scale)) >= f(6.0)) | (d == 0)
    depth2 = df * (f(1.0) / f(depth_scale))
    col, row = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    X = ((col - f(cx) * s) * depth2) * (f(1.0) / (f(fx) * s))
    Y = ((row - f(cy) * s) * depth2) * (f(1.0) / (f(fy) * s))
    Z = depth2 * f(DEPTH_WEIGHT)
    pts = np.stack([X, Y, Z], -1).astype(np.float32)
    pts[invalid] = 0
    return pts.reshape(-1, 3)


def kmeans_fx(points, K, labels, max_iter=4, eps=0.07):
    """cv::kmeans(points, K, labels, (EPS+COUNT, 4, 0.07), 1, KMEANS_USE_INITIAL_LABELS, centers)
    (DynaDetect.cc:397,408; semantics SURVEY Appendix C.1) with ONE documented deviation: the centre sums are
    accumulated in 2^-36 fixed point (order independent) instead of OpenCV's sequential float32 sum, which no
    parallel reduction can reproduce.  Returns (labels, centers, counts)."""
    f = np.float32
    labels = labels.astype(np.int32).copy()
    N = points.shape[0]
    fixed = np.rint(points.astype(np.float64) * float(1 << KM_FIX)).astype(np.int64)
    centers = np.zeros((K, 3), np.float32)
    old = np.zeros((K, 3), np.float32)
    eps2 = eps * eps
    it = 0
    counts = None
    while True:
        centers, old = old, centers
        sums = np.zeros((K, 3), np.int64)
        for j in range(3):
            sums[:, j] = np.bincount(labels, weights=None, minlength=K) * 0  # placeholder shape
        counts = np.bincount(labels, minlength=K).astype(np.int64)
        for j in range(3):
            # exact int64 sums per label
            order = np.argsort(labels, kind="stable")
            cs = np.concatenate([[0], np.cumsum(fixed[order, j])])
            ends = np.cumsum(counts)
            starts = ends - counts
            sums[:, j] = cs[ends] - cs[starts]
        for k in range(K):
            if counts[k] != 0:
                continue
            max_k = int(np.argmax(counts))  # first maximum, like the reference loop
            scale = f(1.0) / f(counts[max_k])
            base = (sums[max_k].astype(np.float64) * 2.0 ** -KM_FIX).astype(np.float32) * scale
            old[max_k] = base  # the reference overwrites old_centers[max_k] with the normalised centre
            idx = np.nonzero(labels == max_k)[0]
            d = points[idx] - base
            dist = np.zeros(len(idx), np.float32)
            for j in range(3):
                dist = (dist + d[:, j] * d[:, j]).astype(np.float32)
            far = idx[len(idx) - 1 - int(np.argmax(dist[::-1]))]  # max_dist <= dist: the last maximum wins
            counts[max_k] -= 1
            counts[k] += 1
            labels[far] = k
            sums[max_k] -= fixed[far]
            sums[k] += fixed[far]
        shift = 0.0
        for k in range(K):
            c = (sums[k].astype(np.float64) * 2.0 ** -KM_FIX).astype(np.float32) * (f(1.0) / f(counts[k]))
            centers[k] = c
            if it > 0:
                t = centers[k].astype(np.float64) - old[k].astype(np.float64)
                shift = max(shift, float((t * t).sum()))
        if it == 0:
            shift = np.inf
        it += 1
        if it == max(max_iter, 2) or shift <= eps2:
            break
        # assign: argmin_k sum_d (x_d - c_kd)^2 in float32, first minimum wins
        best = np.full(N, np.inf, np.float32)
        lab = np.zeros(N, np.int32)
        for k in range(K):
            d = points - centers[k]
            dist = np.zeros(N, np.float32)
            for j in range(3):
                dist = (dist + d[:, j] * d[:, j]).astype(np.float32)
            m = dist < best
            best[m] = dist[m]
            lab[m] = k
        labels = lab
    return labels, centers.copy(), counts


def seg_by_kmeans(depth, label_last, fx, fy, cx, cy, depth_scale, kmeans_impl="fx"):
    """DynaDetect::SegByKmeans (DynaDetect.cc:315-420). kmeans_impl: 'fx' (order-independent sums, what the
    CUDA kernel is bit-exact against) or 'cv2' (cv2.kmeans itself, sequential float32 sums)."""
    H, W = depth.shape
    scales = [1.0, 0.5, 0.25, 0.125]
    pyr = depth_pyramid(depth, 4)
    lab_lvl = [None] * 4
    points = centers = None
    for level in (3, 2, 1, 0):
        s = np.float32(scales[level])
        hp, wp = int(np.float32(H) * s), int(np.float32(W) * s)
        pts = backproject_level(pyr[level], scales[level], fx, fy, cx, cy, depth_scale)
        if level == 3:
            if np.count_nonzero(label_last) == 0:
                br, bc = np.float32(hp) / np.float32(3), np.float32(wp) / np.float32(4)
                ii, jj = np.meshgrid(np.arange(hp, dtype=np.float32), np.arange(wp, dtype=np.float32), indexing="ij")
                lab = (np.floor(ii / br) * 4 + np.floor(jj / bc)).astype(np.int32)
            else:
                lab = np.rint(cv2.resize(label_last.astype(np.float32), (wp, hp))).astype(np.int32)
        else:
            lab = np.rint(cv2.resize(lab_lvl[level + 1].astype(np.float32), (wp, hp))).astype(np.int32)
        lab = lab.reshape(-1)
        if kmeans_impl == "cv2":
            crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 4, 0.07)
            _, l2, c2 = cv2.kmeans(pts, NUM_CLUSTER, lab.reshape(-1, 1).copy(), crit, 1, cv2.KMEANS_USE_INITIAL_LABELS)
            lab, ctr = l2.reshape(-1), c2
        else:
            lab, ctr, _ = kmeans_fx(pts, NUM_CLUSTER, lab)
        lab_lvl[level] = lab.reshape(hp, wp)
        if level == 0:
            points, centers = pts, ctr
    return lab_lvl[0].astype(np.uint8), points, centers


def cluster_order(labels, centers):
    """DynaDetect.cc:1425-1491. Returns (order of kept cluster ids = allLabels, imgLabelForSegEdge dilated 7x7, count0)."""
    H, W = labels.shape
    z = centers[:, 2].astype(np.float32).copy()
    z[z < 0.2] += np.float32(20.0)
    sort_idx = np.argsort(z, kind="stable")
    seg = np.zeros((H, W), np.uint8)
    kept = []
    ratio_area = np.float32(0)
    total = np.float32(H * W)
    count0 = 0
    for i in range(NUM_CLUSTER):
        idx = int(sort_idx[i])
        each = labels == idx
        cnt = int(np.count_nonzero(each))
        if cnt < 60:
            continue
        kept.append(idx)
        ratio = np.float32(cnt) * (np.float32(1.0) / total)
        ratio_area = np.float32(ratio_area + ratio)
        if count0 <= 5 and ratio_area < np.float32(0.6):
            seg[each] = 255
            count0 += 1
    seg = cv2.morphologyEx(seg, cv2.MORPH_DILATE, ellipse(7))
    return kept, seg, count0


# ----------------------------------------------------------------------------- CalOccluded (gradient part)
AROUND = [(0, -2), (1, -2), (2, -1), (2, 0), (2, 1), (1, 2), (0, 2), (-1, 2), (-2, 1), (-2, 0), (-2, -1), (-1, -2)]  # DynaDetect.h:113-125


def depth_edges(depth, depth_scale):
    """DynaDetect.cc:434-536. Returns (imgTotalArea, imgOccluded after OPEN4 (= imgOccludedForPlane), endpoints (x,y) after NMS)."""
    f = np.float32
    H, W = depth.shape
    d1 = depth.astype(np.float32)
    filt = cv2.medianBlur(d1, 5)
    depth_max = f(filt.max())
    total = np.zeros((H, W), np.uint8)
    occl = np.zeros((H, W), np.uint8)
    r = 3
    c = filt[r:H - r, r:W - r]
    total[r:H - r, r:W - r][(c > 0) & ((c / f(depth_scale)) < f(6.0))] = 255
    val_max = np.zeros_like(c)
    for i in range(5):
        for j in range(5):
            nb = filt[r + i - 2:H - r + i - 2, r + j - 2:W - r + j - 2]
            diff = c - nb
            skip = diff > depth_max * f(0.5)
            val_max = np.where(skip, val_max, np.maximum(np.abs(val_max), np.abs(diff)))
    occl[r:H - r, r:W - r][(val_max > c * f(0.03)) & (val_max > f(400.0))] = 255
    occl = cv2.morphologyEx(occl, cv2.MORPH_OPEN, ellipse(4))
    on = occl == 255
    cnt = np.zeros((H, W), np.int32)
    for dx, dy in AROUND:
        sh = np.zeros((H, W), bool)
        ys0, ys1 = max(0, -dy), min(H, H - dy)
        xs0, xs1 = max(0, -dx), min(W, W - dx)
        sh[ys0:ys1, xs0:xs1] = on[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
        cnt += sh
    cand = on & (cnt <= 4)
    cand[:3] = cand[-3:] = False
    cand[:, :3] = cand[:, -3:] = False
    ys, xs = np.nonzero(cand)  # raster order
    kept = []
    # applyNMS (DynaDetect.cc:110-143): sort by the never-assigned curvature (no-op), greedy 6-px suppression
    for x, y in zip(xs.tolist(), ys.tolist()):
        ok = True
        for (qx, qy) in reversed(kept):
            if qy < y - 6:
                break
            if (x - qx) * (x - qx) + (y - qy) * (y - qy) < 36:
                ok = False
                break
        if ok:
            kept.append((x, y))
    return total, occl, np.array(kept, np.int32).reshape(-1, 2)
