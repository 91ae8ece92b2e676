"""oracle/peac_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

CPU re-statement of the plane-contour edge extraction that DynaDetect::CalOccluded runs through PEAC
(reference: ORB_SLAM2/src/DynaDetect.cc:558-593; ORB_SLAM2/include/PEAC/AHCPlaneFitter.hpp, AHCPlaneSeg.hpp,
AHCParamSet.hpp, DisjointSet.hpp, eig33sym.hpp, plane_fitter_pcl.hpp:160-317): organised cloud in METRES fed to the
mm-tuned agglomerative hierarchical clustering plane fitter (16x16 blocks, min-MSE heap, union-find), block erosion,
pixel-level region growing, final re-merge, and per-plane CLOSE 3x3 + external contours drawn with thickness 2.

Parity unpinned by the reference (no tests / golden vectors; PCL + Eigen are un-vendored).  Restatement choices where the
reference is implementation-defined (all measure-zero on real data): std::set<PlaneSeg*> neighbour order (pointer values,
AHCPlaneSeg.hpp:166) -> node creation order; std::priority_queue / std::sort tie order -> insertion order (stable);
Eigen::SelfAdjointEigenSolver (eig33sym.hpp:45-51) -> numpy.linalg.eigh.
"""
from __future__ import annotations

import heapq
import math

import cv2
import numpy as np

WIN = 16                       # windowWidth = windowHeight (AHCParamSet.hpp:140-147)
MIN_SUPPORT = 2000
DEPTH_SIGMA, STD_TOL_INIT, STD_TOL_MERGE = 3e-6, 10.0, 17.0      # AHCParamSet.hpp:49-58
Z_NEAR, Z_FAR = 500.0, 6000.0
ANGLE_NEAR, ANGLE_FAR = math.radians(10.0), math.radians(20.0)
SIM_MERGE, SIM_REFINE = math.cos(math.radians(15.0)), math.cos(math.radians(20.0))
DEPTH_ALPHA, DEPTH_CHANGE_TOL = 0.04, 20.0


def t_mse(init, z):
    return (DEPTH_SIGMA * z * z + (STD_TOL_INIT if init else STD_TOL_MERGE)) ** 2


def t_ang_init(z):
    cz = min(max(z, Z_NEAR), Z_FAR)
    factor = (ANGLE_FAR - ANGLE_NEAR) / (Z_FAR - Z_NEAR)
    return math.cos(factor * cz + ANGLE_NEAR - factor * Z_NEAR)


def organized_cloud(depth, fx, fy, cx, cy, depth_scale):
    """DynaDetect.cc:562-587: float32 arithmetic, NaN where d < 1e-3."""
    f = np.float32
    H, W = depth.shape
    d = depth.astype(np.float32)
    z = d * (f(1.0) / f(depth_scale))
    u, v = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    x = (u - f(cx)) * z / f(fx)
    y = (v - f(cy)) * z / f(fy)
    pts = np.stack([x, y, z], -1).astype(np.float32)
    pts[d < f(1e-3)] = np.nan
    return pts


class Seg:
    """ahc::PlaneSeg (AHCPlaneSeg.hpp:29-388)."""
    __slots__ = ("stats", "N", "rid", "mse", "center", "normal", "nouse", "nbs", "seq")
    _seq = 0

    def __init__(self, stats, N, rid):
        self.stats, self.N, self.rid = stats, N, rid
        self.nouse = False
        self.nbs = {}
        self.seq = Seg._seq
        Seg._seq += 1
        self.mse = float("nan")
        self.center = self.normal = None
        if N >= 4:
            self.compute()

    def compute(self):
        """Stats::compute (AHCPlaneSeg.hpp:84-116)."""
        sx, sy, sz, sxx, syy, szz, sxy, syz, sxz = self.stats
        sc = 1.0 / self.N
        c = np.array([sx * sc, sy * sc, sz * sc])
        K = np.array([[sxx - sx * sx * sc, sxy - sx * sy * sc, sxz - sx * sz * sc],
                      [0, syy - sy * sy * sc, syz - sy * sz * sc],
                      [0, 0, szz - sz * sz * sc]])
        K[1, 0], K[2, 0], K[2, 1] = K[0, 1], K[0, 2], K[1, 2]
        sv, V = np.linalg.eigh(K)
        n = V[:, 0]
        if n @ c > 0:
            n = -n
        self.center, self.normal, self.mse = c, n, float(sv[0] * sc)

    def sim(self, o):
        return abs(float(self.normal @ o.normal))

    def connect(self, o):
        self.nbs[o.seq] = o
        o.nbs[self.seq] = self

    def disconnect_all(self):
        for nb in self.nbs.values():
            nb.nbs.pop(self.seq, None)
        self.nbs = {}


class DisjointSet:
    def __init__(self, n):
        self.parent = list(range(n))
        self.size = [1] * n

    def find(self, x):
        while self.parent[x] != x:
            self.parent[x] = self.parent[self.parent[x]]
            x = self.parent[x]
        return x

    def union(self, x, y):
        xr, yr = self.find(x), self.find(y)
        if xr == yr:
            return xr
        if self.size[xr] < self.size[yr]:
            self.parent[xr] = yr
            self.size[yr] += self.size[xr]
            return yr
        self.parent[yr] = xr
        self.size[xr] += self.size[yr]
        return xr

    def set_size(self, x):
        return self.size[self.find(x)]


def ah_cluster(queue, ds, extracted):
    """PlaneFitter::ahCluster (AHCPlaneFitter.hpp:1050-1256)."""
    heap = [(s.mse, s.seq, s) for s in queue]
    heapq.heapify(heap)
    while heap:
        _, _, p = heapq.heappop(heap)
        if p.nouse:
            continue
        cand, cand_nb = None, None
        for key in sorted(p.nbs):
            nb = p.nbs[key]
            if p.sim(nb) < SIM_MERGE:
                continue
            m = Seg(tuple(a + b for a, b in zip(p.stats, nb.stats)), p.N + nb.N, p.rid if p.N >= nb.N else nb.rid)
            if cand is None or cand.mse > m.mse:
                cand, cand_nb = m, nb
        if cand is not None and cand.mse < t_mse(False, cand.center[2]):
            heapq.heappush(heap, (cand.mse, cand.seq, cand))
            ds.union(p.rid, cand_nb.rid)                      # mergeNbsFrom (AHCPlaneSeg.hpp:357-386)
            nbs = dict(p.nbs)
            nbs.update(cand_nb.nbs)
            nbs.pop(p.seq, None)
            nbs.pop(cand_nb.seq, None)
            p.disconnect_all()
            cand_nb.disconnect_all()
            cand.nbs = nbs
            for nb in nbs.values():
                nb.nbs[cand.seq] = cand
            p.nouse = cand_nb.nouse = True
        else:
            if p.N >= MIN_SUPPORT:
                extracted.append(p)
            p.disconnect_all()
    extracted.sort(key=lambda s: -s.N)                        # PlaneSegSizeCmp, stable


def plane_fit(pts):
    """PlaneFitter::run with doRefine (AHCPlaneFitter.hpp:186-236). Returns (membership image with final plane ids or -1,
    number of final planes, debug dict)."""
    H, W = pts.shape[:2]
    Nh, Nw = H // WIN, W // WIN
    ds = DisjointSet(Nh * Nw)
    P = pts.astype(np.float64)
    valid = ~np.isnan(P[..., 2])
    # ---- initGraph (AHCPlaneFitter.hpp:881-1039)
    G = [None] * (Nh * Nw)
    queue = []
    for i in range(Nh):
        for j in range(Nw):
            blk = P[i * WIN:(i + 1) * WIN, j * WIN:(j + 1) * WIN].reshape(-1, 3)
            if not valid[i * WIN:(i + 1) * WIN, j * WIN:(j + 1) * WIN].all():     # INIT_STRICT; T_dz never fires in metres
                continue
            x, y, z = blk[:, 0], blk[:, 1], blk[:, 2]
            # Stats::push: strictly sequential double sums in raster order (np.cumsum accumulates left to right)
            st = [float(np.cumsum(t)[-1]) for t in (x, y, z, x * x, y * y, z * z, x * y, y * z, x * z)]
            s = Seg(tuple(st), WIN * WIN, i * Nw + j)
            if s.mse < t_mse(True, s.center[2]):
                G[i * Nw + j] = s
                queue.append(s)
    for i in range(Nh):
        j = 1
        while j < Nw:
            c = i * Nw + j
            if G[c - 1] is None:
                j += 1
                continue
            if G[c] is None:
                j += 2
                continue
            if j < Nw - 1 and G[c + 1] is None:
                j += 3
                continue
            th = t_ang_init(G[c].center[2])
            if (j < Nw - 1 and G[c - 1].sim(G[c + 1]) >= th) or (j == Nw - 1 and G[c].sim(G[c - 1]) >= th):
                G[c].connect(G[c - 1])
                if j < Nw - 1:
                    G[c].connect(G[c + 1])
                j += 2
            else:
                j += 1
    for j in range(Nw):
        i = 1
        while i < Nh:
            c = i * Nw + j
            if G[c - Nw] is None:
                i += 1
                continue
            if G[c] is None:
                i += 2
                continue
            if i < Nh - 1 and G[c + Nw] is None:
                i += 3
                continue
            th = t_ang_init(G[c].center[2])
            if (i < Nh - 1 and G[c - Nw].sim(G[c + Nw]) >= th) or (i == Nh - 1 and G[c].sim(G[c - Nw]) >= th):
                G[c].connect(G[c - Nw])
                if i < Nh - 1:
                    G[c].connect(G[c + Nw])
                i += 2
            else:
                i += 1
    extracted = []
    ah_cluster(queue, ds, extracted)
    coarse = [(s.rid, s.N) for s in extracted]
    # ---- refineDetails: findBlockMembership (AHCPlaneFitter.hpp:603-705)
    rid2plid = {s.rid: k for k, s in enumerate(extracted)}
    member = np.full((H, W), -1, np.int32)
    blk_map = [-1] * (Nh * Nw)
    is_valid = [False] * len(extracted)
    rf = []
    npb = WIN * WIN
    for i in range(Nh):
        for j in range(Nw):
            b = i * Nw + j
            setid = ds.find(b)
            if ds.set_size(setid) * npb >= MIN_SUPPORT:
                nbs = []
                if j > 0: nbs.append(b - 1)
                if j < Nw - 1: nbs.append(b + 1)
                if i > 0: nbs.append(b - Nw)
                if i < Nh - 1: nbs.append(b + Nw)
                same = all(ds.find(n) == setid for n in nbs)               # ERODE_ALL_BORDER
                plid = rid2plid[setid]
                if same:
                    blk_map[b] = plid
                    member[i * WIN:(i + 1) * WIN, j * WIN:(j + 1) * WIN] = plid
                    is_valid[plid] = True
            if blk_map[b] < 0:
                if i > 0 and blk_map[b - Nw] >= 0:
                    sp = (i * WIN - 1) * W + j * WIN
                    rf.extend((sp + k, blk_map[b - Nw]) for k in range(1, WIN))
                if j > 0 and blk_map[b - 1] >= 0:
                    sp = (i * WIN) * W + j * WIN - 1
                    rf.extend((sp + k * W, blk_map[b - 1]) for k in range(0, WIN - 1))
            else:
                plid = blk_map[b]
                if i > 0 and blk_map[b - Nw] != plid:
                    sp = (i * WIN) * W + j * WIN
                    rf.extend((sp + k, plid) for k in range(0, WIN - 1))
                if j > 0 and blk_map[b - 1] != plid:
                    sp = (i * WIN) * W + j * WIN
                    rf.extend((sp + k * W, plid) for k in range(1, WIN))
    # ---- floodFill (AHCPlaneFitter.hpp:546-594)
    mem = member.reshape(-1)
    dist_map = np.full(H * W, np.finfo(np.float32).max, np.float32)
    Pf = P.reshape(-1, 3)
    vf = valid.reshape(-1)
    pl_c = [s.center for s in extracted]
    pl_n = [s.normal for s in extracted]
    pl_thr = [9 * s.mse + 1e-5 for s in extracted]
    k = 0
    while k < len(rf):
        s_idx, plid = rf[k]
        k += 1
        sy, sx = divmod(s_idx, W)
        nb = []
        if sx > 0: nb.append(s_idx - 1)
        if sx < W - 1: nb.append(s_idx + 1)
        if sy > 0: nb.append(s_idx - W)
        if sy < H - 1: nb.append(s_idx + W)
        for c in nb:
            trail = int(mem[c])
            if trail <= -6:
                continue
            if trail >= 0 and trail == plid:
                continue
            cy, cx = divmod(c, W)
            bx, by = cx // WIN, cy // WIN
            if by < Nh and bx < Nw and blk_map[by * Nw + bx] >= 0:
                continue
            ok = False
            if vf[c]:
                cdist = np.float32(abs(float(pl_n[plid] @ (Pf[c] - pl_c[plid]))))
                ok = float(cdist) ** 2 < pl_thr[plid]
            if ok:
                if trail >= 0:
                    a, b_ = extracted[trail], extracted[plid]
                    if b_.sim(a) >= SIM_REFINE:
                        a.connect(b_)
                if cdist < dist_map[c]:
                    mem[c] = plid
                    dist_map[c] = cdist
                    rf.append((c, plid))
                elif trail < 0:
                    mem[c] = trail - 1
            elif trail < 0:
                mem[c] = trail - 1
    # ---- final merge of the extracted planes that met during region growing (AHCPlaneFitter.hpp:291-323)
    old = extracted
    final = []
    ah_cluster([s for s, v in zip(old, is_valid) if v], ds, final)
    plidmap = [-1] * len(old)
    n_final = 0
    for i, op in enumerate(old):
        if not is_valid[i]:
            continue
        root = ds.find(op.rid)
        if root == op.rid:
            if plidmap[i] < 0:
                plidmap[i] = n_final
                n_final += 1
        else:
            npid = rid2plid[root]
            if plidmap[npid] < 0:
                plidmap[i] = plidmap[npid] = n_final
                n_final += 1
            else:
                plidmap[i] = plidmap[npid]
    lut = np.array(plidmap + [-1], np.int32)
    out = np.where(member >= 0, lut[np.clip(member, -1, len(old) - 1)], -1).astype(np.int32)
    return out, n_final, dict(coarse=coarse, blk_map=np.array(blk_map, np.int32).reshape(Nh, Nw), grown=member.copy(), n_final_sorted=len(final))


_peac_c = None


def _peac_lib():
    global _peac_c
    if _peac_c is None:
        import ctypes
        import os
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_build", "libpeac_cpu.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", here])
        lib = ctypes.CDLL(path)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        lib.peac_plane_fit.argtypes = [vp, ci, ci, vp, vp, vp, vp, ci, vp, vp]
        lib.peac_plane_fit.restype = ci
        _peac_c = lib
    return _peac_c


def plane_fit_c(pts):
    """Same as plane_fit, executed by oracle/peac_cpu.c (the C restatement of the same reference code; ~1000x faster than the
    Python loops above, which stay as the readable statement and are compared with it in tests/test_peac_cpu.py)."""
    import ctypes
    H, W = pts.shape[:2]
    Nh, Nw = H // WIN, W // WIN
    pts = np.ascontiguousarray(pts, np.float32)
    member = np.empty((H, W), np.int32)
    grown = np.empty((H, W), np.int32)
    blk = np.empty((Nh, Nw), np.int32)
    coarse = np.zeros((256, 2), np.int32)
    nc = ctypes.c_int(0)
    stats = np.zeros(8, np.int64)
    n = _peac_lib().peac_plane_fit(pts.ctypes.data, W, H, member.ctypes.data, grown.ctypes.data, blk.ctypes.data, coarse.ctypes.data, 256,
                                   ctypes.byref(nc), stats.ctypes.data)
    if n < 0:
        raise MemoryError("peac_plane_fit")
    names = ("pops", "seeds", "queue_entries", "levels", "max_level", "valid_blocks", "coarse_planes", "final_planes")
    return member, n, dict(coarse=[(int(r), int(c)) for r, c in coarse[:nc.value]], blk_map=blk, grown=grown,
                           stats={k: int(v) for k, v in zip(names, stats)})


def plane_edges(depth, fx, fy, cx, cy, depth_scale, debug=None, impl="c"):
    """imgEdgeByPlane of DynaDetect.cc:558-593: every final plane -> CLOSE 3x3 -> external contours, thickness 2
    (AHCPlaneFitter.hpp:366-399).  impl: 'c' (oracle/peac_cpu.c) or 'py' (the loops in this file) -- same restatement."""
    pts = organized_cloud(depth, fx, fy, cx, cy, depth_scale)
    member, n, dbg = plane_fit_c(pts) if impl == "c" else plane_fit(pts)
    H, W = depth.shape
    out = np.zeros((H, W), np.uint8)
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    for p in range(n):
        one = np.where(member == p, 255, 0).astype(np.uint8)
        one = cv2.morphologyEx(one, cv2.MORPH_CLOSE, se)
        contours, _ = cv2.findContours(one, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        cv2.drawContours(out, contours, -1, 255, 2)
    if debug is not None:
        debug.update(dbg, member=member, n=n)
    return out
