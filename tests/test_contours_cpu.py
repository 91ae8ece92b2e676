"""The closed-form contour rules (tests/contour_rules.py = what csrc/ccl.cu implements) against cv2's border
following on random blob images: contours.size(), contourArea, arcLength, drawContours FILLED / thickness 2,
RETR_EXTERNAL and RETR_CCOMP (DynaDetect.cc:605-617,675-713,1579-1603)."""
import cv2
import numpy as np
import pytest

import contour_rules as cr


def blobs(seed, shape=(96, 128), sigma=2.5, thr=0.5):
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.random(shape).astype(np.float32), (0, 0), sigma)
    img = (img - img.min()) / (img.max() - img.min())
    fg = img > thr
    if seed % 3 == 0:  # thin structures and specks
        fg ^= rng.random(shape) > 0.97
    if seed % 4 == 1:  # touch the image border
        fg[0, :] = True
        fg[:, -1] = True
    return fg


@pytest.mark.parametrize("seed", range(12))
def test_external_contours(seed):
    fg = blobs(seed)
    img = fg.astype(np.uint8) * 255
    contours, _ = cv2.findContours(img, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    mine = cr.external_components(fg)
    assert len(mine) == len(contours)
    # match by first pixel: cv2 contours start at the raster-first pixel of the component
    key = {}
    for F in mine:
        ys, xs = np.nonzero(F)
        key[(int(xs[ys == ys[0]].min()), int(ys[0]))] = F
    for i, c in enumerate(contours):
        F = key[tuple(int(v) for v in c[0, 0])]
        ref = np.zeros_like(img)
        cv2.drawContours(ref, contours, i, 255, cv2.FILLED)
        assert np.array_equal(ref > 0, F)
        axis, diag, g = cr.outer_stats(F)
        n = axis + diag
        assert (n if n else 1) == len(c)
        assert abs(g) / 2.0 == cv2.contourArea(c)
        assert abs((axis + diag * float(np.float32(np.sqrt(np.float32(2))))) - cv2.arcLength(c, True)) < 1e-3
        ref2 = np.zeros_like(img)
        cv2.drawContours(ref2, contours, i, 255, 2)
        got = cr.draw_thick2(F)
        assert np.array_equal(ref2 > 0, got), int(((ref2 > 0) != got).sum())


@pytest.mark.parametrize("seed", range(12))
def test_ccomp_contours(seed):
    fg = blobs(seed + 100, thr=0.47)
    img = fg.astype(np.uint8) * 255
    contours, hier = cv2.findContours(img, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)
    fl, nf, bl = cr.regions(fg)
    n_holes = int(bl.max())
    assert len(contours) == nf + len([h for h in range(1, n_holes + 1) if (bl == h).any()])
    for c, h in zip(contours, hier[0]):
        x, y = int(c[0, 0, 0]), int(c[0, 0, 1])
        if h[3] < 0:   # outer border of a foreground component: starts at its raster-first pixel
            comp = fl == fl[y, x]
            axis, diag, g = cr.outer_stats(cr.filled(comp))
        else:          # hole border: starts at the foreground pixel left of the hole's raster-first pixel
            hole = bl == bl[y, x + 1]
            assert bl[y, x + 1] > 0
            axis, diag, g = cr.hole_stats(cr.filled(hole, hole=True))
        n = axis + diag
        assert (n if n else 1) == len(c), (h[3], n, len(c))
        assert abs(g) / 2.0 == cv2.contourArea(c)
        assert abs((axis + diag * float(np.float32(np.sqrt(np.float32(2))))) - cv2.arcLength(c, True)) < 1e-3
