"""numpy model of the closed-form contour rules that sindslam_b200/csrc/ccl.cu implements on the device
(see ccl.cuh).  Test infrastructure: test_contours_cpu.py checks this model against cv2.findContours /
contourArea / arcLength / drawContours on random blob images, the GPU tests check the kernels against cv2
through the stages that use them."""
import numpy as np
from scipy import ndimage

S8 = np.ones((3, 3), bool)
S4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], bool)

# quad pattern m = a | b<<1 | c<<2 | d<<3 with a=(x,y) b=(x+1,y) c=(x,y+1) d=(x+1,y+1)
AXIS = {0x3: ("a", "b"), 0xC: ("d", "c"), 0x5: ("c", "a"), 0xA: ("b", "d")}
DIAG = {0xD: ("d", "a"), 0xB: ("a", "d"), 0xE: ("b", "c"), 0x7: ("c", "b")}


def quad_patterns(F):
    """Pattern image over all quads with top-left pixel (x, y), x in [-1, W-1], y in [-1, H-1]."""
    P = np.pad(F.astype(np.int32), 1)
    return P[:-1, :-1] | (P[:-1, 1:] << 1) | (P[1:, :-1] << 2) | (P[1:, 1:] << 3)


def outer_stats(F):
    """(#axis steps, #diagonal steps, signed Green sum = 2*area) of the outer border of the filled region F."""
    m = quad_patterns(F)
    ys, xs = np.mgrid[-1:F.shape[0], -1:F.shape[1]]
    corner = {"a": (xs, ys), "b": (xs + 1, ys), "c": (xs, ys + 1), "d": (xs + 1, ys + 1)}
    axis = diag = 0
    g = 0
    for pat, (p, q) in list(AXIS.items()) + list(DIAG.items()):
        sel = m == pat
        n = int(sel.sum())
        if pat in AXIS:
            axis += n
        else:
            diag += n
        px, py = corner[p]
        qx, qy = corner[q]
        g += int((px[sel].astype(np.int64) * qy[sel] - qx[sel].astype(np.int64) * py[sel]).sum())
    diag += 2 * int(((m == 0x9) | (m == 0x6)).sum())
    return axis, diag, g


def hole_stats(Hf):
    """Border of a hole, Hf = filled hole (the background region + everything nested in it): the polygon runs
    through the NON-member pixels around it."""
    m = quad_patterns(Hf)
    cm = (~m) & 0xF
    ys, xs = np.mgrid[-1:Hf.shape[0], -1:Hf.shape[1]]
    corner = {"a": (xs, ys), "b": (xs + 1, ys), "c": (xs, ys + 1), "d": (xs + 1, ys + 1)}
    axis = diag = 0
    g = 0
    diag += 2 * int(((cm == 0x9) | (cm == 0x6)).sum())
    for pat, (p, q) in list(AXIS.items()) + list(DIAG.items()):
        # complement pattern has 2 adjacent (axis) or 3 (diag) bits <-> m has 2 adjacent / 1 member
        sel = (cm == pat) & (m != 0x9) & (m != 0x6)
        n = int(sel.sum())
        if pat in AXIS:
            axis += n
        else:
            diag += n
        px, py = corner[p]
        qx, qy = corner[q]
        g -= int((px[sel].astype(np.int64) * qy[sel] - qx[sel].astype(np.int64) * py[sel]).sum())
    return axis, diag, g


def regions(fg):
    """8-connected foreground labels, 4-connected background labels (0 = exterior)."""
    fl, nf = ndimage.label(fg, S8)
    bp = np.pad(~fg, 1, constant_values=True)
    bl, nb = ndimage.label(bp, S4)
    ext = bl[0, 0]
    bl = bl[1:-1, 1:-1]
    bl = np.where(bl == ext, 0, bl)
    bl[fg] = -1
    return fl, nf, bl


def filled(comp, hole=False):
    """comp + everything nested in it.  A foreground component (8-connected) is surrounded by 4-connected
    background; a hole (4-connected background region) by 8-connected foreground."""
    return ndimage.binary_fill_holes(comp, S8 if hole else S4)


def external_components(fg):
    """Filled top-level components F(A) in raster order of their first pixel."""
    fl, nf, _ = regions(fg)
    out = []
    covered = np.zeros_like(fg)
    firsts = ndimage.find_objects(fl)
    order = []
    for i in range(1, nf + 1):
        ys, xs = np.nonzero(fl == i)
        order.append((ys[0] * fg.shape[1] + xs[np.nonzero(ys == ys[0])[0][0]], i))
    for _, i in sorted(order):
        comp = fl == i
        if covered[comp].any():
            continue  # nested inside an earlier top-level component (an enclosing component starts earlier in raster order)
        F = filled(comp)
        covered |= F
        out.append(F)
    return out


def draw_thick2(F):
    """drawContours(img, contour of F, thickness=2): cross-dilated boundary pixels + a 4x4-minus-corners block
    around every quad with a diagonal step."""
    H, W = F.shape
    P = np.pad(F, 1)
    bnd = F & ~(P[:-2, 1:-1] & P[2:, 1:-1] & P[1:-1, :-2] & P[1:-1, 2:])
    out = ndimage.binary_dilation(bnd, S4)
    m = quad_patterns(F)
    dq = np.isin(m, [0xD, 0xB, 0xE, 0x7, 0x9, 0x6])
    big = np.zeros((H + 4, W + 4), bool)   # offset: quad (x,y) top-left pixel -> big[y+2.., x+2..]
    ys, xs = np.nonzero(dq)                # indices are (y+1, x+1)
    blk = np.array([[0, 1, 1, 0], [1, 1, 1, 1], [1, 1, 1, 1], [0, 1, 1, 0]], bool)
    for y, x in zip(ys, xs):
        # quad top-left = (x-1, y-1); block covers pixels (x-2 .. x+1, y-2 .. y+1) -> big index +2
        big[y:y + 4, x:x + 4] |= blk
    return out | big[2:-2, 2:-2]
