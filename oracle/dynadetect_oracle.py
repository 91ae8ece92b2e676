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


# ----------------------------------------------------------------------------- plane-edge filtering
def filter_plane_edges(plane_edges, grad_edges, endpoints):
    """DynaDetect.cc:598-641. plane_edges = imgEdgeByPlane from PEAC, grad_edges = imgOccludedForPlane.
    Returns (imgOccluded1, imgOccluded2)."""
    H, W = grad_edges.shape
    e = cv2.subtract(plane_edges, grad_edges)
    contours, _ = cv2.findContours(e, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    out = np.zeros((H, W), np.uint8)
    for i, c in enumerate(contours):
        if len(c) < 25:
            continue
        one = np.zeros((H, W), np.uint8)
        cv2.drawContours(one, contours, i, 255, 2)
        one = cv2.morphologyEx(one, cv2.MORPH_DILATE, ellipse(10))
        if any(one[y, x] == 255 for x, y in np.asarray(endpoints).reshape(-1, 2)):
            one = cv2.morphologyEx(one, cv2.MORPH_ERODE, ellipse(7))
            out = cv2.add(out, one)
    occl2 = out.copy()
    occl1 = cv2.morphologyEx(cv2.bitwise_or(grad_edges, out), cv2.MORPH_CLOSE, ellipse(3))
    return occl1, occl2


# ----------------------------------------------------------------------------- SegAndMergeV2
def _hist_depth(depth_norm, mask):
    """calcHist(&imgDepth, 1, 0, mask, hist, 1, {256}, {0,255}) (DynaDetect.cc:1691-1696): value 255 is dropped."""
    return cv2.calcHist([depth_norm], [0], mask, [256], [0, 255])


def cal_hist(img1, img2, depth_norm):
    """DynaDetect.cc:1685-1739. Returns [CORREL, 1 - BHATTACHARYYA, INTERSECT]."""
    h1 = _hist_depth(depth_norm, img1)
    h2 = _hist_depth(depth_norm, img2)
    m1, m2 = float(h1.max()), float(h2.max())
    if m1 > m2:
        h1 = cv2.normalize(h1, None, 0, 400, cv2.NORM_MINMAX, -1)
        h2 = h2 * np.float32(1.0 / (m1 / 400))
    else:
        h2 = cv2.normalize(h2, None, 0, 400, cv2.NORM_MINMAX, -1)
        h1 = h1 * np.float32(1.0 / (m2 / 400))
    return [cv2.compareHist(h1, h2, cv2.HISTCMP_CORREL), 1 - cv2.compareHist(h1, h2, cv2.HISTCMP_BHATTACHARYYA),
            cv2.compareHist(h1, h2, cv2.HISTCMP_INTERSECT)]


def center_z_fx(points, mask):
    """myCluster::calCenterPoint (DynaDetect.cc:256-293), z only (the only component consumed, :741).  The reference's
    OpenMP float reduction is order dependent; restated with the same 2^-36 fixed-point sum as the k-means centres."""
    idx = np.nonzero(mask.reshape(-1))[0]
    s = int(np.rint(points[idx, 2].astype(np.float64) * float(1 << KM_FIX)).astype(np.int64).sum())
    return np.float32(np.float32(s * 2.0 ** -KM_FIX) / np.float32(len(idx)))


def seg_and_merge_v2(kept, labels_km, occluded1, occluded2, seg_edge, points, depth, debug=None):
    """DynaDetect::SegAndMergeV2 (DynaDetect.cc:653-1018). kept = allLabels order (k-means ids by depth).
    Returns imgLabelNew (u8; 0 = invalid, 1..n)."""
    f = np.float32
    H, W = labels_km.shape
    clusters = []
    occl_dil = cv2.morphologyEx(occluded1, cv2.MORPH_DILATE, ellipse(10))
    for ki in kept[:-1]:                       # the last (farthest / invalid) cluster is not processed (:664)
        orig = np.where(labels_km == ki, 255, 0).astype(np.uint8)
        each = cv2.subtract(orig, occluded1)
        each = cv2.morphologyEx(each, cv2.MORPH_OPEN, ellipse(4))
        contours, _ = cv2.findContours(each, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        for c, cnt in enumerate(contours):
            if len(cnt) > 50 and cv2.contourArea(cnt) > 80:
                temp = np.zeros((H, W), np.uint8)
                cv2.drawContours(temp, contours, c, 255, cv2.FILLED)
                temp = cv2.morphologyEx(temp, cv2.MORPH_DILATE, ellipse(9))
                temp = cv2.bitwise_and(temp, orig)
                cl = dict(img=temp, area=f(np.count_nonzero(temp)), lianjie=None, km=ki)
                cl["dil"] = cv2.morphologyEx(temp, cv2.MORPH_DILATE, ellipse(7))
                draw = np.zeros((H, W), np.uint8)
                cv2.drawContours(draw, contours, c, 255, 2)
                temp1 = cv2.subtract(draw, occl_dil)
                temp1 = cv2.bitwise_and(temp1, seg_edge)
                if np.count_nonzero(temp1) > 20:
                    c2, _ = cv2.findContours(temp1, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
                    c2 = [q for q in c2 if len(q) >= 30]
                    if len(c2) > 0:
                        t = np.zeros((H, W), np.uint8)
                        cv2.drawContours(t, c2, -1, 255, cv2.FILLED)
                        cl["lianjie"] = t
                cl["z"] = center_z_fx(points, temp)
                clusters.append(cl)
    for cl in clusters:
        cl["score"] = f(cl["area"] * f(0.0003) - cl["z"])
    clusters.sort(key=lambda c: -c["score"])    # std::sort descending (ties: unspecified in the reference)
    n = len(clusters)
    total_img = np.full((H, W), (n + 1) & 255, np.uint8)
    for i, cl in enumerate(clusters):
        total_img[cl["img"] > 0] = i
    depth_max = float(depth.max())
    depth_norm = cv2.convertScaleAbs(depth, alpha=(1.0 / depth_max) * 255.0) if depth_max > 0 else np.zeros((H, W), np.uint8)
    M1 = np.zeros((n + 1, n + 1), np.float32)
    M2 = np.zeros((n + 1, n + 1), np.float32)
    M3 = np.zeros((n + 1, n + 1), np.float32)
    Wt = np.ones((n + 1, n + 1), np.float32)
    Rj = np.ones((n + 1, n + 1), np.float32)
    small_label = int(min(f(0.7) * f(n), f(15.0)))
    for i in range(n):
        c1 = clusters[i]
        for j in range(i + 1, n):
            c2 = clusters[j]
            v1 = v2 = v3 = f(0)
            if c1["area"] < c2["area"]:
                less_area, less_label = c1["area"], i
            else:
                less_area, less_label = c2["area"], j
            if less_label < 10:
                Wt[i, j] = Wt[j, i] = 0.7
            elif less_label > small_label:
                Wt[i, j] = Wt[j, i] = 2.0
            overlap = cv2.bitwise_and(c1["dil"], c2["dil"])
            is_must = False
            if np.count_nonzero(overlap) > min(f(200.0), f(less_area * f(0.4))):
                ov_edge = cv2.bitwise_and(overlap, occluded2)
                v1 = f(1.0)
                r = cal_hist(c1["img"], c2["img"], depth_norm)
                v3 = f(r[0] + r[1] + r[2] * 0.0005)
                if np.count_nonzero(ov_edge) > 100 and less_label < small_label:
                    Rj[i, j] = Rj[j, i] = 0.0
                    continue
                elif v3 < f(0.19) and less_label < small_label and not is_must:
                    Rj[i, j] = Rj[j, i] = 0.0
                    continue
                l1, l2 = c1["lianjie"], c2["lianjie"]
                if l1 is not None and l2 is not None:
                    ol = int(np.count_nonzero(cv2.bitwise_and(l1, l2)))
                    if ol > 0:
                        a1, a2 = int(np.count_nonzero(l1)), int(np.count_nonzero(l2))
                        if ol > min(50, int(0.5 * min(a1, a2))):
                            v2 = f(ol)
                            if ol > 0.62 * a1 or ol > 0.62 * a2:
                                v2 = f(max(250, ol))
                                is_must = True
                M1[i, j] = M1[j, i] = v1
                M2[i, j] = M2[j, i] = v2
                M3[i, j] = M3[j, i] = v3
    T = ((M2 * f(0.01) + M3) * Rj * Wt).astype(np.float32)
    if debug is not None:
        debug.update(clusters=clusters, total_img=total_img.copy(), T=T.copy(), M2=M2, M3=M3, Rj=Rj, Wt=Wt)
    labels = merge_and_relabel(T, total_img, n)
    return labels


def merge_and_relabel(T, total_img, n):
    """Greedy merge + relabel (DynaDetect.cc:894-1016) on the RAG matrix T ((n+1) x (n+1) float32)."""
    f = np.float32
    T = T.copy()
    H, W = total_img.shape
    count_merged = 0
    merge = [[] for _ in range(n + 1)]
    merged = [0] * (n + 1)
    i = 0
    while i < min(NUM_CLUSTER - 1 + count_merged, n):
        j = i + 1
        while j < min(NUM_CLUSTER - 1 + count_merged, n):
            score = T[j, i]
            if score > f(0.9):
                col = T[0:j, j].copy()
                to_merge = i
                val = T[j, i]
                for k in range(j):
                    if col[k] > val:
                        to_merge = k
                merged[j] = 1
                merge[to_merge].append(j)
                one_col = T[:, j].copy()
                T[:, to_merge] += one_col
                T[to_merge, :] += one_col
                T[:, j] = 0
                T[j, :] = 0
                count_merged += 1
            j += 1
        i += 1
    for i in range(min(NUM_CLUSTER - 1 + count_merged, n), n):
        mc, best = n, f(0.2)
        for j in range(i):
            s = T[j, i]
            if s > best:
                best, mc = s, j
        merged[i] = 1
        merge[mc].append(i)
        one_col = T[:, i].copy()
        T[:, mc] += one_col
        T[mc, :] += one_col
        T[:, i] = 0
        T[i, :] = 0
    out = np.zeros((H, W), np.uint8)
    idx = 1
    for i in range(n):
        if not merged[i]:
            m = total_img == i
            for a in merge[i]:
                m |= total_img == a
                for b in merge[a]:
                    m |= total_img == b
            out[m] = idx
            idx += 1
    return out


# ----------------------------------------------------------------------------- decision + final mask
def dynamic_decide(low_in, high, high_last, total_area, labels):
    """DynaDetect.cc:1553-1636. low_in = imgMaskLowError (0/128), high (0/255). Returns imgDyna {0,125,255}."""
    H, W = labels.shape
    dyna = np.zeros((H, W), np.uint8)
    low = cv2.bitwise_or(high_last, low_in)
    low[low > 0] = 128
    low = cv2.bitwise_and(low, total_area)
    low = cv2.morphologyEx(low, cv2.MORPH_DILATE, ellipse(5))
    nmax = int(labels.max())
    for n in range(1, nmax + 1):
        one = np.where(labels == n, 255, 0).astype(np.uint8)
        border = cv2.copyMakeBorder(one, 1, 1, 1, 1, cv2.BORDER_CONSTANT, value=0)
        border = cv2.bitwise_not(border)
        with_high = cv2.bitwise_and(one, high)
        if np.count_nonzero(with_high) > 100:
            cs, _ = cv2.findContours(with_high, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)
            for c in cs:
                area = cv2.contourArea(c)
                ln = cv2.arcLength(c, True)
                with np.errstate(divide="ignore", invalid="ignore"):
                    roundness = np.float64(4 * np.pi * area) / np.float64(ln * ln)
                seed = (0, 0)
                for p in c[:, 0, :]:
                    if low[p[1], p[0]] == 128:
                        seed = (int(p[0]), int(p[1]))
                        break
                if (area > 100.0 and roundness > 0.2) or area > 2000.0:
                    cv2.floodFill(low, border, seed, 50, 5, 5, 8 | cv2.FLOODFILL_MASK_ONLY | (50 << 8))
        filled = np.where(border[1:-1, 1:-1] == 50, 255, 0).astype(np.uint8)
        if np.count_nonzero(filled) > 0.5 * np.count_nonzero(one):
            dyna = cv2.bitwise_or(dyna, one)
        else:
            dyna = cv2.bitwise_or(dyna, filled)
    dyna = cv2.morphologyEx(dyna, cv2.MORPH_DILATE, ellipse(9))
    static2 = (cv2.subtract(total_area, dyna).astype(np.float32) * np.float32(125.0 / 255.0))
    static2 = np.rint(static2).astype(np.uint8)
    return cv2.add(dyna, static2)


# ----------------------------------------------------------------------------- sampling + homography
def sample_pairs(flow, dyna_last, label_last):
    """DynaDetect.cc:1163-1231: weighted 10-px grid samples, sorted by weight (descending), filtered by inBorder.
    Returns (inputPoints, inputPointsLast) as float32 n x 2 (x, y)."""
    f = np.float32
    H, W = dyna_last.shape
    cw = np.zeros(NUM_CLUSTER, np.float32)
    dyn0 = dyna_last == 255
    for i in range(1, NUM_CLUSTER):
        one = label_last == i
        cw[i] = f(np.count_nonzero(one & dyn0)) / (f(np.count_nonzero(one)) + f(1.0))
    rows = list(range(10, H, 10))
    cols = list(range(10, W, 10))
    cv2.setRNGSeed(12345)                       # cv::RNG rng(12345) (DynaDetect.cc:1163)
    g = np.zeros((len(rows) * len(cols), 1), np.float32)
    cv2.randn(g, 0, 0.5)                        # rng.gaussian(0.5) per sample, raster order (:1187)
    g = g.ravel()
    items = []
    k = 0
    for r in rows:
        for c in cols:
            rd = g[k]
            k += 1
            dl = int(dyna_last[r, c])
            if dl < 20:
                w = f(rd + f(1.0))
            elif 20 <= dl <= 230:
                w = f(rd + f(1.2) * (f(1.0) - cw[int(label_last[r, c])]))
            else:
                w = f(rd + f(0.4))
            items.append((c, r, w))
    items.sort(key=lambda t: -t[2])              # std::sort descending (stable here; ties are measure-zero)
    pts, last = [], []
    for c, r, _ in items:
        fx_, fy_ = flow[r, c]
        lx, ly = f(f(c) - fx_), f(f(r) - fy_)
        irow, icol = int(ly), int(lx)            # float -> int truncation toward zero at the inBorder call
        if 0 <= irow <= H and 0 <= icol <= W:    # (uint)(v) <= dim : inclusive upper bound (DynaDetect.cc:103-106)
            pts.append((f(c), f(r)))
            last.append((lx, ly))
    return np.array(pts, np.float32).reshape(-1, 2), np.array(last, np.float32).reshape(-1, 2)


def estimate_homography(pts, pts_last):
    """cv::findHomography(inputPoints, inputPointsLast, cv::noArray(), cv::RHO) (DynaDetect.cc:1235)."""
    Hm, _ = cv2.findHomography(pts, pts_last, cv2.RHO)
    return Hm


# ----------------------------------------------------------------------------- CPU flow engines of the reference
def variational_refine(I0_u8, I1_u8, flow):
    """cv::VariationalRefinement::create()->calc(I0, I1, flow) with default parameters, in place on a copy
    (DynaDetect.cc:1133-1143).  Executed by the real OpenCV (cv2.VariationalRefinement)."""
    f = np.ascontiguousarray(flow, np.float32).copy()
    cv2.VariationalRefinement_create().calc(I0_u8, I1_u8, f)
    return f


def deepflow_restated(I0_u8, I1_u8):
    """cv::optflow::createOptFlow_DeepFlow()->calc(I0, I1, flow) (DynaDetect.cc:1031,1075) -- the reference's default
    CPU flow.  opencv_contrib's optflow module is NOT in cv2-headless, so this restates its published structure
    (optflow/src/deepflow.cpp, 4.x: sigma 0.6 pre-blur, pyramid factor 0.95 down to >= 25 px, per level
    VariationalRefinement(alpha=4*1.0, delta=0.5/3, gamma=5.0/3, fixedPoint=5, sor=25, omega=1.6), bilinear
    up-sampling / 0.95) around the real cv2.VariationalRefinement solver.  Used only as the timed CPU baseline
    ("DeepFlow-restated") -- PARITY UNPINNED, never a parity oracle."""
    sigma, min_size, factor = 0.6, 25, 0.95
    a = cv2.GaussianBlur(I0_u8.astype(np.float32), (5, 5), sigma)
    b = cv2.GaussianBlur(I1_u8.astype(np.float32), (5, 5), sigma)
    pyr = [(a, b)]
    while True:
        h, w = pyr[-1][0].shape
        nw, nh = int(w * factor), int(h * factor)
        if min(nw, nh) < min_size:
            break
        pyr.append((cv2.resize(pyr[-1][0], (nw, nh), interpolation=cv2.INTER_LINEAR),
                    cv2.resize(pyr[-1][1], (nw, nh), interpolation=cv2.INTER_LINEAR)))
    vr = cv2.VariationalRefinement_create()
    vr.setAlpha(4 * 1.0)
    vr.setDelta(0.5 / 3)
    vr.setGamma(5.0 / 3)
    vr.setFixedPointIterations(5)
    vr.setSorIterations(25)
    vr.setOmega(1.6)
    h, w = pyr[-1][0].shape
    flow = np.zeros((h, w, 2), np.float32)
    for lvl in range(len(pyr) - 1, -1, -1):
        A, B = pyr[lvl]
        h, w = A.shape
        if flow.shape[:2] != (h, w):
            flow = cv2.resize(flow, (w, h), interpolation=cv2.INTER_LINEAR) * np.float32(1.0 / factor)
        flow = np.ascontiguousarray(flow)
        vr.calc(A, B, flow)
    return flow, len(pyr)


def flow_residual_cpu(bgr_cur, bgr_last, bgr_lastlast, dyna_last, label_last, engine="brox", refine=True, inject_flow=None):
    """DynaDetect::DetectDynaByDenseOpticalFLow (DynaDetect.cc:1023-1374) end to end on the CPU.
    engine: 'brox' (USECUDA build's solver, oracle/brox_cpu.c) or 'deepflow' (default CPU build, restated).
    inject_flow: (full-resolution flow, large_motion) -- the authors' own identical-flow hook (the .flo injection at
    DynaDetect.cc:1149-1158): everything AFTER the dense flow runs here (sample weighting, the real cv2.findHomography(RHO),
    residual, Otsu / Triangle thresholds, masks).
    Returns dict(low, high, thr, flow, H, large_motion)."""
    H, W = bgr_cur.shape[:2]
    if inject_flow is not None:
        full, lm = inject_flow
        full = np.ascontiguousarray(full, np.float32)
        p, q = sample_pairs(full, dyna_last, label_last)
        Hm = estimate_homography(p, q)
        mag = homography_residual(full, Hm)
        low, high, thr, _ = threshold_masks(mag)
        return dict(low=low, high=high, thr=thr, flow=full, H=Hm, large_motion=bool(lm))
    g = [gray_small(bgr2gray(x)) for x in (bgr_cur, bgr_last, bgr_lastlast)]

    def calc(i_ref):
        if engine == "brox":
            return brox_flow(g[0].astype(np.float32) * np.float32(1 / 255.0), g[i_ref].astype(np.float32) * np.float32(1 / 255.0))
        return deepflow_restated(g[0], g[i_ref])[0]

    flow = -calc(2)
    lm, _, _ = flow_magnitude_hist_large_motion(flow, W, H)
    i_ref = 2
    if lm:
        flow = -calc(1)
        i_ref = 1
    if refine:
        flow = variational_refine(g[0], g[i_ref], flow)
    full = upsample_flow(flow, W, H)
    p, q = sample_pairs(full, dyna_last, label_last)
    Hm = estimate_homography(p, q)
    mag = homography_residual(full, Hm)
    low, high, thr, _ = threshold_masks(mag)
    return dict(low=low, high=high, thr=thr, flow=full, H=Hm, large_motion=bool(lm))


# ----------------------------------------------------------------------------- DetectDynaArea end to end
class DynaDetectOracle:
    """ORB_SLAM2::DynaDetect (DynaDetect.h:95-189): constructor state + DetectDynaArea (DynaDetect.cc:1377-1666)."""

    def __init__(self, bgr_last, bgr_lastlast, fx, fy, cx, cy, depth_scale, plane_edges=False, engine="brox", refine=True, kmeans_impl="fx"):
        H, W = bgr_last.shape[:2]
        self.W, self.H = W, H
        self.fx, self.fy, self.cx, self.cy, self.depth_scale = fx, fy, cx, cy, depth_scale
        self.plane_edges, self.engine, self.refine, self.kmeans_impl = plane_edges, engine, refine, kmeans_impl
        self.rgb_last, self.rgb_lastlast = bgr_last.copy(), bgr_lastlast.copy()
        z = np.zeros((H, W), np.uint8)
        self.dyna_last, self.high_last, self.label_last = z.copy(), z.copy(), z.copy()

    def detect(self, bgr, depth, inject_masks=None, inject_flow=None):
        """Returns dict(mask, label, + every intermediate).  inject_masks = (low, high): skip the flow branch and use
        these masks instead; inject_flow = (flow, large_motion): skip only the dense-flow engine (the authors' own
        identical-flow hook, DynaDetect.cc:1149-1158) and run the sampling, the real RHO, the residual and the thresholds here."""
        out = {}
        if inject_masks is None:
            fr = flow_residual_cpu(bgr, self.rgb_last, self.rgb_lastlast, self.dyna_last, self.label_last, self.engine, self.refine,
                                   inject_flow=inject_flow)
            low, high = fr["low"], fr["high"]
            out["flow"] = fr
        else:
            low, high = inject_masks
        labels_km, points, centers = seg_by_kmeans(depth, self.label_last, self.fx, self.fy, self.cx, self.cy, self.depth_scale, self.kmeans_impl)
        kept, seg, _ = cluster_order(labels_km, centers)
        total_area, grad, ep = depth_edges(depth, self.depth_scale)
        if self.plane_edges:
            from oracle import peac_oracle
            plane = peac_oracle.plane_edges(depth, self.fx, self.fy, self.cx, self.cy, self.depth_scale)
        else:
            plane = np.zeros_like(grad)
        occl1, occl2 = filter_plane_edges(plane, grad, ep)
        dbg = {}
        labels = seg_and_merge_v2(kept, labels_km, occl1, occl2, seg, points, depth, debug=dbg)
        dyna = dynamic_decide(low, high, self.high_last, total_area, labels)
        out.update(mask=dyna, label=labels, low=low, high=high, labels_km=labels_km, kept=kept, seg=seg, total_area=total_area,
                   grad=grad, endpoints=ep, plane=plane, occl1=occl1, occl2=occl2, rag=dbg)
        # state roll (DynaDetect.cc:1660-1664)
        self.dyna_last = dyna.copy()
        self.rgb_lastlast, self.rgb_last = self.rgb_last, bgr.copy()
        self.high_last = high.copy()
        self.label_last = labels.copy()
        return out
