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
