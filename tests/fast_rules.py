"""numpy model of the FAST-9/16 score map + windowed non-maximum suppression that csrc/orb.cu implements
(SURVEY.md Appendix C.6); checked against cv2.FastFeatureDetector in test_orb_cpu.py."""
import numpy as np

RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def score_map(img):
    """s(p) = max over the 16 nine-long arcs of max(min_i d_i, min_i -d_i), d_i = I(p) - I(ring_i); 0 inside the 3-px margin."""
    I = img.astype(np.int32)
    H, W = I.shape
    c = I[3:H - 3, 3:W - 3]
    d = np.stack([c - I[3 + dy:H - 3 + dy, 3 + dx:W - 3 + dx] for dx, dy in RING])   # 16 x h x w
    d2 = np.concatenate([d, d[:8]], 0)
    best = np.zeros_like(c)
    for k in range(16):
        arc = d2[k:k + 9]
        best = np.maximum(best, np.maximum(arc.min(0), (-arc).min(0)))
    s = np.zeros((H, W), np.int32)
    s[3:H - 3, 3:W - 3] = best
    return s


def fast_window(img, threshold):
    """cv::FAST(window, kps, threshold, nonmaxSuppression=true): list of (x, y, response) in raster order."""
    s = score_map(img)
    H, W = s.shape
    cand = s > threshold
    sc = np.where(cand, s, 0)
    P = np.pad(sc, 1)
    keep = cand.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= sc > P[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
    ys, xs = np.nonzero(keep)
    return [(int(x), int(y), int(s[y, x]) - 1) for y, x in zip(ys, xs)]
