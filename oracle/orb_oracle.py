"""oracle/orb_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

CPU re-statement of ORB_SLAM2::ORBextractor (reference: ORB_SLAM2/src/ORBextractor.cc, include/ORBextractor.h)
in Python + numpy, calling cv2 (4.13) for the OpenCV primitives the reference calls (resize, copyMakeBorder, FAST,
GaussianBlur, fastAtan2).  Parity unpinned by the reference (it ships no tests / golden vectors); pinned here by the
real OpenCV primitives and by the known-answer constants checked in tests/test_orb_cpu.py (level sizes, umax,
feature quotas, pattern checksum).

Two documented restatement choices:
  * DistributeOctTree sorts (size, node pointer) pairs (ORBextractor.cc:677): ties between equal-sized nodes are
    broken by heap addresses in the reference.  Here the tie-break is the node creation sequence number.
  * cos/sin of the keypoint angle (ORBextractor.cc:113) are evaluated in double and rounded to float.
"""
from __future__ import annotations

import math
import os

import cv2
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATCH_SIZE, HALF_PATCH_SIZE, EDGE_THRESHOLD = 31, 15, 19     # ORBextractor.cc:72-74
f32 = np.float32


def cv_round(x):
    """cvRound: round half to even."""
    return int(np.rint(x))


def load_pattern():
    """bit_pattern_31_ (ORBextractor.cc:150-408), 256 x (x0, y0, x1, y1)."""
    return np.load(os.path.join(_HERE, "..", "tests", "golden", "orb_bit_pattern_31.npy")).astype(np.int32)


class OrbOracle:
    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        """ORBextractor::ORBextractor (ORBextractor.cc:410-470)."""
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        self.scale_factor = f32(scale_factor)
        sf = [f32(1.0)]
        for _ in range(1, nlevels):
            sf.append(f32(sf[-1] * self.scale_factor))
        self.scale = sf
        self.inv_scale = [f32(f32(1.0) / s) for s in sf]
        self.sigma2 = [f32(s * s) for s in sf]
        self.inv_sigma2 = [f32(f32(1.0) / s) for s in self.sigma2]
        factor = f32(1.0 / float(self.scale_factor))   # 1.0f / (double)scaleFactor, stored as float (ORBextractor.cc:433)
        nd = f32(f32(nfeatures) * f32(f32(1) - factor) / f32(f32(1) - f32(math.pow(float(factor), float(nlevels)))))
        self.per_level = []
        tot = 0
        for _ in range(nlevels - 1):
            self.per_level.append(cv_round(nd))
            tot += self.per_level[-1]
            nd = f32(nd * factor)
        self.per_level.append(max(nfeatures - tot, 0))
        # umax (ORBextractor.cc:450-467)
        umax = [0] * (HALF_PATCH_SIZE + 1)
        vmax = int(math.floor(float(f32(HALF_PATCH_SIZE * f32(math.sqrt(2.0)) / 2 + 1))))
        vmin = int(math.ceil(float(f32(HALF_PATCH_SIZE * f32(math.sqrt(2.0)) / 2))))
        hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(hp2 - v * v))
        v0 = 0
        for v in range(HALF_PATCH_SIZE, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax
        self.pattern = load_pattern()
        self.pyramid = []        # padded level images
        self.level_size = []

    # ------------------------------------------------------------------ ComputePyramid (ORBextractor.cc:1166-1191)
    def compute_pyramid(self, image):
        self.pyramid, self.level_size = [], []
        E = EDGE_THRESHOLD
        prev = None
        for level in range(self.nlevels):
            s = self.inv_scale[level]
            w, h = cv_round(f32(f32(image.shape[1]) * s)), cv_round(f32(f32(image.shape[0]) * s))
            if level == 0:
                cur = image
            else:
                cur = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
            self.pyramid.append(cv2.copyMakeBorder(cur, E, E, E, E, cv2.BORDER_REFLECT_101))
            self.level_size.append((w, h))
            prev = cur
        return self.pyramid

    def level_image(self, level):
        E = EDGE_THRESHOLD
        w, h = self.level_size[level]
        return self.pyramid[level][E:E + h, E:E + w]

    # ------------------------------------------------------------------ per-cell FAST (ORBextractor.cc:765-829)
    def cell_candidates(self, level):
        """vToDistributeKeys of one level: list of (x, y, response) in push order (coords relative to minBorder)."""
        E = EDGE_THRESHOLD
        w, h = self.level_size[level]
        pad = self.pyramid[level]
        min_bx = min_by = E - 3
        max_bx, max_by = w - E + 3, h - E + 3
        width, height = f32(max_bx - min_bx), f32(max_by - min_by)
        Wc = f32(30)
        n_cols, n_rows = int(width / Wc), int(height / Wc)
        w_cell, h_cell = int(math.ceil(float(f32(width / f32(n_cols))))), int(math.ceil(float(f32(height / f32(n_rows)))))
        fast_ini = cv2.FastFeatureDetector_create(self.ini_th, True)
        fast_min = cv2.FastFeatureDetector_create(self.min_th, True)
        out = []
        for i in range(n_rows):
            ini_y = f32(min_by + i * h_cell)
            max_y = f32(ini_y + h_cell + 6)
            if ini_y >= max_by - 3:
                continue
            if max_y > max_by:
                max_y = f32(max_by)
            for j in range(n_cols):
                ini_x = f32(min_bx + j * w_cell)
                max_x = f32(ini_x + w_cell + 6)
                if ini_x >= max_bx - 6:
                    continue
                if max_x > max_bx:
                    max_x = f32(max_bx)
                win = pad[E + int(ini_y):E + int(max_y), E + int(ini_x):E + int(max_x)]
                kps = fast_ini.detect(np.ascontiguousarray(win))
                if len(kps) == 0:
                    kps = fast_min.detect(np.ascontiguousarray(win))
                for kp in kps:
                    out.append((f32(kp.pt[0] + j * w_cell), f32(kp.pt[1] + i * h_cell), f32(kp.response)))
        return out, (min_bx, max_bx, min_by, max_by)

    # ------------------------------------------------------------------ DistributeOctTree (ORBextractor.cc:481-763)
    @staticmethod
    def distribute_octtree(keys, min_x, max_x, min_y, max_y, N):
        class Node:
            __slots__ = ("UL", "UR", "BL", "BR", "keys", "no_more", "seq")
        seq_counter = [0]

        def new_node():
            n = Node()
            n.keys, n.no_more = [], False
            n.seq = seq_counter[0]
            seq_counter[0] += 1
            return n

        def divide(p):
            half_x = int(math.ceil(float(f32(f32(p.UR[0] - p.UL[0]) / f32(2)))))
            half_y = int(math.ceil(float(f32(f32(p.BR[1] - p.UL[1]) / f32(2)))))
            n1, n2, n3, n4 = new_node(), new_node(), new_node(), new_node()
            n1.UL = p.UL; n1.UR = (p.UL[0] + half_x, p.UL[1]); n1.BL = (p.UL[0], p.UL[1] + half_y); n1.BR = (p.UL[0] + half_x, p.UL[1] + half_y)
            n2.UL = n1.UR; n2.UR = p.UR; n2.BL = n1.BR; n2.BR = (p.UR[0], p.UL[1] + half_y)
            n3.UL = n1.BL; n3.UR = n1.BR; n3.BL = p.BL; n3.BR = (n1.BR[0], p.BL[1])
            n4.UL = n3.UR; n4.UR = n2.BR; n4.BL = n3.BR; n4.BR = p.BR
            for kp in p.keys:
                if kp[0] < n1.UR[0]:
                    (n1 if kp[1] < n1.BR[1] else n3).keys.append(kp)
                elif kp[1] < n1.BR[1]:
                    n2.keys.append(kp)
                else:
                    n4.keys.append(kp)
            for n in (n1, n2, n3, n4):
                if len(n.keys) == 1:
                    n.no_more = True
            return n1, n2, n3, n4

        n_ini = int(math.floor(float(f32(f32(max_x - min_x) / f32(max_y - min_y))) + 0.5))   # C round()
        hx = f32(f32(max_x - min_x) / f32(n_ini))
        nodes = []                       # std::list order: index 0 = front
        ini = []
        for i in range(n_ini):
            n = new_node()
            n.UL = (int(hx * f32(i)), 0); n.UR = (int(hx * f32(i + 1)), 0)
            n.BL = (n.UL[0], max_y - min_y); n.BR = (n.UR[0], max_y - min_y)
            nodes.append(n)
            ini.append(n)
        for kp in keys:
            ini[int(f32(kp[0]) / hx)].keys.append(kp)
        kept = []
        for n in nodes:
            if len(n.keys) == 1:
                n.no_more = True
                kept.append(n)
            elif len(n.keys) > 0:
                kept.append(n)
        nodes = kept
        finish = False
        while not finish:
            prev_size = len(nodes)
            n_to_expand = 0
            size_and_node = []
            k = 0
            while k < len(nodes):
                cur = nodes[k]
                if cur.no_more:
                    k += 1
                    continue
                for c in divide(cur):
                    if len(c.keys) > 0:
                        nodes.insert(0, c)      # push_front
                        k += 1
                        if len(c.keys) > 1:
                            n_to_expand += 1
                            size_and_node.append((len(c.keys), c))
                nodes.pop(k)                    # lit = lNodes.erase(lit)
            if len(nodes) >= N or len(nodes) == prev_size:
                finish = True
            elif len(nodes) + n_to_expand * 3 > N:
                while not finish:
                    prev_size = len(nodes)
                    prev_list = sorted(size_and_node, key=lambda t: (t[0], t[1].seq))
                    size_and_node = []
                    for sz, node in reversed(prev_list):
                        for c in divide(node):
                            if len(c.keys) > 0:
                                nodes.insert(0, c)
                                if len(c.keys) > 1:
                                    size_and_node.append((len(c.keys), c))
                        nodes.remove(node)
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prev_size:
                        finish = True
        result = []
        for n in nodes:
            best = n.keys[0]
            for kp in n.keys[1:]:
                if kp[2] > best[2]:
                    best = kp
            result.append(best)
        return result

    # ------------------------------------------------------------------ IC_Angle (ORBextractor.cc:77-104)
    def ic_angle(self, level, x, y):
        E = EDGE_THRESHOLD
        pad = self.pyramid[level].astype(np.int32)
        cx, cy = E + cv_round(x), E + cv_round(y)
        m01 = m10 = 0
        for u in range(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1):
            m10 += u * int(pad[cy, cx + u])
        for v in range(1, HALF_PATCH_SIZE + 1):
            d = self.umax[v]
            plus = pad[cy + v, cx - d:cx + d + 1]
            minus = pad[cy - v, cx - d:cx + d + 1]
            us = np.arange(-d, d + 1)
            m01 += v * int((plus - minus).sum())
            m10 += int((us * (plus + minus)).sum())
        return f32(cv2.fastAtan2(float(m01), float(m10)))

    # ------------------------------------------------------------------ operator() (ORBextractor.cc:1043-1164)
    def extract(self, image, mask=None, debug=None):
        """Returns (keypoints: n x 6 float64 array [x, y, size, angle, response, octave], descriptors n x 32 u8)."""
        if image is None or image.size == 0:
            return np.zeros((0, 6)), np.zeros((0, 32), np.uint8)
        self.compute_pyramid(image)
        all_kps = []
        for level in range(self.nlevels):
            cand, (min_bx, max_bx, min_by, max_by) = self.cell_candidates(level)
            kps = self.distribute_octtree(cand, min_bx, max_bx, min_by, max_by, self.per_level[level])
            size = f32(int(f32(PATCH_SIZE) * self.scale[level]))
            lvl = []
            for (x, y, r) in kps:
                px, py = f32(x + f32(min_bx)), f32(y + f32(min_by))
                lvl.append([px, py, size, None, r, level])
            for kp in lvl:
                kp[3] = self.ic_angle(level, kp[0], kp[1])
            all_kps.append(lvl)
            if debug is not None:
                debug.setdefault("candidates", []).append(cand)
        copy = [list(l) for l in all_kps]
        if mask is not None and mask.size:
            for level in range(self.nlevels):
                sc = f32(math.pow(float(self.scale_factor), level))
                keep = []
                for kp in all_kps[level]:
                    if int(mask[int(f32(kp[1] * sc)), int(f32(kp[0] * sc))]) != 255:
                        keep.append(kp)
                all_kps[level] = keep
        if sum(len(l) for l in all_kps) < 250:          # "maybe lost" fallback (ORBextractor.cc:1105-1115)
            all_kps = copy
        out_k, out_d = [], []
        pat = self.pattern
        for level in range(self.nlevels):
            lvl = all_kps[level]
            if not lvl:
                continue
            work = cv2.GaussianBlur(np.ascontiguousarray(self.level_image(level)), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            for kp in lvl:
                out_d.append(self.descriptor(work, kp, pat))
                sc = self.scale[level]
                x, y = (kp[0], kp[1]) if level == 0 else (f32(kp[0] * sc), f32(kp[1] * sc))
                out_k.append([float(x), float(y), float(kp[2]), float(kp[3]), float(kp[4]), float(level)])
        if not out_k:
            return np.zeros((0, 6)), np.zeros((0, 32), np.uint8)
        return np.array(out_k, np.float64), np.array(out_d, np.uint8)

    @staticmethod
    def descriptor(img, kp, pat):
        """computeOrbDescriptor (ORBextractor.cc:108-147)."""
        angle = f32(f32(kp[3]) * f32(np.pi / f32(180.0)))
        a, b = f32(math.cos(float(angle))), f32(math.sin(float(angle)))
        cy, cx = cv_round(kp[1]), cv_round(kp[0])
        px = pat[:, [0, 2]].astype(np.float32)
        py = pat[:, [1, 3]].astype(np.float32)
        ry = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int32)
        rx = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int32)
        vals = img[cy + ry, cx + rx].astype(np.int32)          # 256 x 2
        bits = (vals[:, 0] < vals[:, 1]).astype(np.uint8).reshape(32, 8)
        return (bits << np.arange(8, dtype=np.uint8)).sum(1).astype(np.uint8)
