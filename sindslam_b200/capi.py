"""ctypes binding of libsindyn_cuda.so (include/sindyn.h).

This is the only way Python code in this repo reaches the kernels; it fails loudly when the
library or a CUDA device is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsindyn_cuda.so")

STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "NO_DEVICE", 4: "STATE", 5: "CAPACITY"}


class MatchParams(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("b", C.c_float),
                ("Tcw_cur", C.c_float * 16), ("Tcw_last", C.c_float * 16), ("th", C.c_float), ("mono", C.c_int), ("check_orientation", C.c_int)]


class SindynError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("depth_scale", C.c_float), ("flow_scale", C.c_float),
        ("brox_alpha", C.c_float), ("brox_gamma", C.c_float), ("brox_pyr_scale", C.c_float),
        ("brox_inner", C.c_int), ("brox_outer", C.c_int), ("brox_solver", C.c_int),
        ("brox_omega", C.c_float), ("refine", C.c_int),
        ("n_row_cluster", C.c_int), ("n_col_cluster", C.c_int), ("depth_weight", C.c_float),
        ("device", C.c_int), ("use_graphs", C.c_int), ("plane_edges", C.c_int),
        ("stage_timing", C.c_int),
    ]


class FrameParams(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3", "bf", "depth_map_factor")]


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int)]


_vp, _i, _sz, _f = C.c_void_p, C.c_int, C.c_size_t, C.c_float
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/sindyn.h declares
SIGNATURES = {
    "sindyn_version": (C.c_char_p, []),
    "sindyn_default_config": (None, [C.POINTER(Config), _i, _i]),
    "sindyn_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "sindyn_destroy": (_i, [_vp]),
    "sindyn_last_error": (C.c_char_p, [_vp]),
    "sindyn_set_stream": (_i, [_vp, _vp]),
    "sindyn_synchronize": (_i, [_vp]),
    "sindyn_launch_count": (C.c_ulonglong, [_vp]),
    "sindyn_set_prev_frames": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "sindyn_detect": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _i]),
    "sindyn_upload_frame": (_i, [_vp, _i, _vp, _sz, _vp, _sz]),
    "sindyn_detect_resident": (_i, [_vp, _i, _i]),
    "sindyn_get_detect_results": (_i, [_vp, _vp, _vp]),
    "sindyn_get_recluster_debug": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "sindyn_flow_residual": (_i, [_vp, _vp, _sz, _vp, _vp, _i]),
    "sindyn_flow_residual_resident": (_i, [_vp, _i, _i]),
    "sindyn_get_flow_results": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ip]),
    "sindyn_get_path_info": (_i, [_vp, _vp]),
    "sindyn_brox_profile": (_i, [_vp, _vp]),
    "sindyn_morph_ellipse": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _i, _i]),
    "sindyn_flow_brox": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "sindyn_gray_resize": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "sindyn_flow_branch": (_i, [_vp, _vp, _sz, _vp, _ip]),
    "sindyn_flow_refine": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "sindyn_find_homography_rho": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "sindyn_estimate_homography": (_i, [_vp, _vp, _vp, _ip]),
    "sindyn_sample_pairs": (_i, [_vp, _vp, _vp, _vp, _i, _ip]),
    "sindyn_residual_homography": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sindyn_residual_pose": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "sindyn_kmeans": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "sindyn_cluster_order": (_i, [_vp, _vp, _vp, _ip]),
    "sindyn_depth_edges": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _i, _ip]),
    "sindyn_plane_edges": (_i, [_vp, _vp, _sz, _vp]),
    "sindyn_get_peac_debug": (_i, [_vp, _vp, _vp, _ip, _ip]),
    "sindyn_filter_plane_edges": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "sindyn_recluster": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _ip]),
    "sindyn_dynamic_decide": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "sindyn_get_state": (_i, [_vp, _i, _vp]),
    "sindyn_set_state": (_i, [_vp, _i, _vp]),
    "sindyn_get_stage_ms": (_i, [_vp, _vp, _i]),
    "sindyn_orb_create": (_i, [_i, _f, _i, _i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "sindyn_orb_destroy": (_i, [_vp]),
    "sindyn_orb_extract": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _i, _ip]),
    "sindyn_track_frame": (_i, [_vp, _vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz, _vp, _sz, _vp, _vp, _i, _ip, _i]),
    "sindyn_track_frame_resident": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "sindyn_track_join": (_i, [_vp, _vp]),
    "sindyn_track_submit": (_i, [_vp, _vp, _vp, _sz, _vp, _sz, _i, _i, _i]),
    "sindyn_track_collect": (_i, [_vp, _vp, _vp, _sz, _vp, _sz, _vp, _vp, _i, _ip]),
    "sindyn_track_get_results": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _ip]),
    "sindyn_orb_get_pyramid_level": (_i, [_vp, _i, _vp, _ip, _ip]),
    "sindyn_orb_get_candidates": (_i, [_vp, _i, _vp, _i, _ip]),
    "sindyn_orb_get_plane": (_i, [_vp, _i, _i, _vp, _ip, _ip]),
    "sindyn_orb_frame_features": (_i, [_vp, _vp, _sz, C.POINTER(FrameParams), _vp, _vp, _vp, _vp, _vp, _vp, _i, _ip]),
    "sindyn_cloud_single": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _ip]),
    "sindyn_cloud_consistent": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp, _ip, _vp, _sz, _vp, _vp]),
    "sindyn_orb_search_by_projection": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ip, _ip]),
    "sindyn_orb_set_stream": (_i, [_vp, _vp]),
    "sindyn_orb_launch_count": (C.c_ulonglong, [_vp]),
    "sindyn_orb_last_error": (C.c_char_p, [_vp]),
}

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the in-tree library and bind every declared symbol. Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise SindynError(f"{path} not found: build it with `python -m sindslam_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8
    return a


class SinDyn:
    """One DynaDetect handle (one sequence, one CUDA device)."""

    def __init__(self, width=640, height=480, fx=535.4, fy=539.2, cx=320.1, cy=247.6, depth_scale=5000.0, device=0, **over):
        self.lib = load_library()
        cfg = Config()
        self.lib.sindyn_default_config(C.byref(cfg), width, height)
        cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.depth_scale, cfg.device = fx, fy, cx, cy, depth_scale, device
        for k, v in over.items():
            if not hasattr(cfg, k):
                raise KeyError(k)
            setattr(cfg, k, v)
        self.cfg = cfg
        self.W, self.H = width, height
        self.fw, self.fh = int(np.float32(cfg.flow_scale) * np.float32(width)), int(np.float32(cfg.flow_scale) * np.float32(height))
        self.h = _vp()
        st = self.lib.sindyn_create(C.byref(cfg), C.byref(self.h))
        if st != 0:
            msg = self.lib.sindyn_last_error(self.h).decode() if self.h else ""
            raise SindynError(f"sindyn_create failed: {STATUS.get(st, st)} {msg}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sindyn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st, what):
        if st != 0:
            raise SindynError(f"{what}: {STATUS.get(st, st)}: {self.lib.sindyn_last_error(self.h).decode()}")

    # -- plumbing
    def set_stream(self, stream_ptr):
        self._ck(self.lib.sindyn_set_stream(self.h, _vp(stream_ptr)), "set_stream")

    def synchronize(self):
        self._ck(self.lib.sindyn_synchronize(self.h), "synchronize")

    @property
    def launches(self):
        return int(self.lib.sindyn_launch_count(self.h))

    def stage_ms(self):
        ms = np.zeros(16, np.float32)
        self._ck(self.lib.sindyn_get_stage_ms(self.h, _p(ms), 16), "get_stage_ms")
        return ms

    # -- reference-shaped API
    def set_prev_frames(self, bgr_last, bgr_lastlast):
        a, b = _u8(bgr_last), _u8(bgr_lastlast)
        self._ck(self.lib.sindyn_set_prev_frames(self.h, _p(a), a.strides[0], _p(b), b.strides[0]), "set_prev_frames")

    def detect(self, bgr, depth, frame_idx):
        bgr = _u8(bgr)
        depth = np.ascontiguousarray(depth, np.uint16)
        mask = np.empty((self.H, self.W), np.uint8)
        label = np.empty((self.H, self.W), np.uint8)
        self._ck(self.lib.sindyn_detect(self.h, _p(bgr), bgr.strides[0], _p(depth), depth.strides[0], _p(mask), self.W,
                                        _p(label), self.W, frame_idx), "detect")
        return mask, label

    def upload_frame(self, slot, bgr, depth):
        bgr = _u8(bgr)
        depth = np.ascontiguousarray(depth, np.uint16)
        self._ck(self.lib.sindyn_upload_frame(self.h, slot, _p(bgr), bgr.strides[0], _p(depth), depth.strides[0]), "upload_frame")

    def detect_resident(self, slot, frame_idx):
        self._ck(self.lib.sindyn_detect_resident(self.h, slot, frame_idx), "detect_resident")

    def detect_results(self):
        mask = np.empty((self.H, self.W), np.uint8)
        label = np.empty((self.H, self.W), np.uint8)
        self._ck(self.lib.sindyn_get_detect_results(self.h, _p(mask), _p(label)), "get_detect_results")
        return mask, label

    def recluster_debug(self):
        cap = 129 * 129
        T = np.zeros(cap, np.float32)
        area = np.zeros(128, np.int32)
        score = np.zeros(128, np.float32)
        order = np.zeros(128, np.int32)
        n = self.lib.sindyn_get_recluster_debug(self.h, _p(T), cap, _p(area), _p(score), _p(order))
        return dict(n=n, T=T[:(n + 1) * (n + 1)].reshape(n + 1, n + 1).copy(), area=area[:n].copy(), score=score[:n].copy(), order=order[:n].copy())

    def flow_residual(self, bgr, roll=True, low=None, high=None):
        """DetectDynaByDenseOpticalFLow (DynaDetect.cc:1023-1374) -> (mask_low, mask_high)."""
        bgr = _u8(bgr)
        low = np.empty((self.H, self.W), np.uint8) if low is None else low
        high = np.empty((self.H, self.W), np.uint8) if high is None else high
        self._ck(self.lib.sindyn_flow_residual(self.h, _p(bgr), bgr.strides[0], _p(low), _p(high), int(roll)), "flow_residual")
        return low, high

    def flow_residual_resident(self, slot, roll=True):
        self._ck(self.lib.sindyn_flow_residual_resident(self.h, slot, int(roll)), "flow_residual_resident")

    def flow_results(self):
        flow = np.empty((self.H, self.W, 2), np.float32)
        Hm = np.empty((3, 3), np.float64)
        thr = np.empty(4, np.float32)
        lo = np.empty((self.H, self.W), np.uint8)
        hi = np.empty((self.H, self.W), np.uint8)
        lm = C.c_int(0)
        self._ck(self.lib.sindyn_get_flow_results(self.h, _p(flow), _p(Hm), _p(thr), _p(lo), _p(hi), C.byref(lm)), "get_flow_results")
        return dict(flow=flow, H=Hm, thr=thr, low=lo, high=hi, large_motion=bool(lm.value))

    def path_info(self):
        info = np.zeros(4, np.int32)
        self._ck(self.lib.sindyn_get_path_info(self.h, _p(info)), "get_path_info")
        return dict(flow_one_graph=bool(info[0]), flow_graph_broken=bool(info[1]), cluster_graph=bool(info[2]))

    def brox_profile(self):
        out = np.zeros(4, np.float64)
        self._ck(self.lib.sindyn_brox_profile(self.h, _p(out)), "brox_profile")
        return dict(sor_ms=float(out[0]), sor_launches=int(out[1]), solve_ms=float(out[2]), pixel_sweeps=int(out[3]))

    def morph_ellipse(self, img, k, op):
        img = _u8(img)
        out = np.empty_like(img)
        self._ck(self.lib.sindyn_morph_ellipse(self.h, _p(img), img.strides[0], _p(out), out.strides[0], img.shape[1], img.shape[0], k, op), "morph")
        return out

    # -- dense-map consumer (row f3): pubPointCloud.cc generatePointCloud
    POINT_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("z", np.float32), ("b", np.uint8), ("g", np.uint8), ("r", np.uint8), ("a", np.uint8)])

    @staticmethod
    def _intr(intr):
        return None if intr is None else np.ascontiguousarray(intr, np.float64)

    def cloud_single(self, bgr, depth, mask, Twc, intr=None):
        bgr, depth, mask = _u8(bgr), np.ascontiguousarray(depth, np.uint16), _u8(mask)
        Twc = np.ascontiguousarray(Twc, np.float64)
        out = np.empty(((self.H + 2) // 3) * ((self.W + 2) // 3), self.POINT_DTYPE)
        n = C.c_int(0)
        k = self._intr(intr)
        self._ck(self.lib.sindyn_cloud_single(self.h, _p(bgr), bgr.strides[0], _p(depth), depth.strides[0], _p(mask), mask.strides[0], _p(Twc),
                                              _p(k) if k is not None else None, _p(out), C.byref(n)), "cloud_single")
        return out[: n.value]

    def cloud_consistent(self, bgr, depth, depth_last, mask, mask_last, label, T_rel, Twc, intr=None):
        bgr, mask, mask_last, label = _u8(bgr), _u8(mask), _u8(mask_last), _u8(label)
        depth, depth_last = np.ascontiguousarray(depth, np.uint16), np.ascontiguousarray(depth_last, np.uint16)
        T_rel, Twc = np.ascontiguousarray(T_rel, np.float64), np.ascontiguousarray(Twc, np.float64)
        out = np.empty(((self.H + 1) // 2) * ((self.W + 1) // 2), self.POINT_DTYPE)
        mask_new = np.empty((self.H, self.W), np.uint8)
        depth_new = np.empty((self.H, self.W), np.uint16)
        stats = np.zeros(36, np.int32)
        n = C.c_int(0)
        k = self._intr(intr)
        self._ck(self.lib.sindyn_cloud_consistent(self.h, _p(bgr), bgr.strides[0], _p(depth), depth.strides[0], _p(depth_last), depth_last.strides[0],
                                                  _p(mask), mask.strides[0], _p(mask_last), mask_last.strides[0], _p(label), label.strides[0],
                                                  _p(T_rel), _p(Twc), _p(k) if k is not None else None, _p(out), C.byref(n), _p(mask_new),
                                                  mask_new.strides[0], _p(stats), _p(depth_new)), "cloud_consistent")
        return dict(points=out[: n.value], mask_new=mask_new, occlusion=stats[:12].copy(), label_count=stats[12:24].copy(),
                    kept=stats[24:36].astype(bool), depth_new=depth_new)

    # -- stage level
    def flow_brox(self, I0, I1):
        I0 = np.ascontiguousarray(I0, np.float32)
        I1 = np.ascontiguousarray(I1, np.float32)
        hh, ww = I0.shape
        out = np.empty((hh, ww, 2), np.float32)
        self._ck(self.lib.sindyn_flow_brox(self.h, _p(I0), _p(I1), ww, hh, _p(out)), "flow_brox")
        return out

    def gray_resize(self, bgr):
        bgr = _u8(bgr)
        g = np.empty((self.H, self.W), np.uint8)
        s = np.empty((self.fh, self.fw), np.uint8)
        self._ck(self.lib.sindyn_gray_resize(self.h, _p(bgr), bgr.strides[0], _p(g), _p(s)), "gray_resize")
        return g, s

    def flow_branch(self, bgr):
        bgr = _u8(bgr)
        out = np.empty((self.H, self.W, 2), np.float32)
        lm = C.c_int(0)
        self._ck(self.lib.sindyn_flow_branch(self.h, _p(bgr), bgr.strides[0], _p(out), C.byref(lm)), "flow_branch")
        return out, bool(lm.value)

    def flow_refine(self, I0, I1, flow):
        I0, I1 = _u8(I0), _u8(I1)
        f = np.ascontiguousarray(flow, np.float32).copy()
        self._ck(self.lib.sindyn_flow_refine(self.h, _p(I0), _p(I1), I0.shape[1], I0.shape[0], _p(f)), "flow_refine")
        return f

    def estimate_homography(self, flow):
        flow = np.ascontiguousarray(flow, np.float32)
        Hm = np.zeros((3, 3), np.float64)
        n = C.c_int(0)
        self._ck(self.lib.sindyn_estimate_homography(self.h, _p(flow), _p(Hm), C.byref(n)), "estimate_homography")
        return Hm, n.value

    def find_homography_rho(self, src, dst):
        """cv::findHomography(src, dst, noArray(), RHO) on an ordered correspondence list -> (H 3x3 or None, mask n x 1, info)."""
        src = np.ascontiguousarray(src, np.float32).reshape(-1, 2)
        dst = np.ascontiguousarray(dst, np.float32).reshape(-1, 2)
        Hm = np.zeros((3, 3), np.float64)
        mask = np.zeros(len(src), np.uint8)
        info = np.zeros(12, np.int32)
        self._ck(self.lib.sindyn_find_homography_rho(self.h, _p(src), _p(dst), len(src), _p(Hm), _p(mask), _p(info)), "find_homography_rho")
        return (Hm if info[1] >= 4 else None), mask, info

    def sample_pairs(self, flow, capacity=4096):
        flow = np.ascontiguousarray(flow, np.float32)
        a = np.zeros((capacity, 2), np.float32)
        b = np.zeros((capacity, 2), np.float32)
        n = C.c_int(0)
        self._ck(self.lib.sindyn_sample_pairs(self.h, _p(flow), _p(a), _p(b), capacity, C.byref(n)), "sample_pairs")
        return a[:n.value], b[:n.value]

    def residual_homography(self, flow, Hm):
        flow = np.ascontiguousarray(flow, np.float32)
        Hm = np.ascontiguousarray(Hm, np.float64)
        mag = np.empty((self.H, self.W), np.float32)
        lo = np.empty((self.H, self.W), np.uint8)
        hi = np.empty((self.H, self.W), np.uint8)
        thr = np.zeros(4, np.float32)
        self._ck(self.lib.sindyn_residual_homography(self.h, _p(flow), _p(Hm), _p(mag), _p(lo), _p(hi), _p(thr)), "residual_homography")
        return mag, lo, hi, thr

    def residual_pose(self, flow, depth, T_old_cur):
        flow = np.ascontiguousarray(flow, np.float32)
        depth = np.ascontiguousarray(depth, np.uint16)
        T = np.ascontiguousarray(np.asarray(T_old_cur, np.float64)[:3, :4])
        mag = np.empty((self.H, self.W), np.float32)
        lo = np.empty((self.H, self.W), np.uint8)
        hi = np.empty((self.H, self.W), np.uint8)
        thr = np.zeros(4, np.float32)
        self._ck(self.lib.sindyn_residual_pose(self.h, _p(flow), _p(depth), depth.strides[0], _p(T), _p(mag), _p(lo), _p(hi), _p(thr)), "residual_pose")
        return mag, lo, hi, thr

    def kmeans(self, depth):
        depth = np.ascontiguousarray(depth, np.uint16)
        labels = np.empty((self.H, self.W), np.uint8)
        pts = np.empty((self.H * self.W, 3), np.float32)
        ctr = np.empty((self.cfg.n_row_cluster * self.cfg.n_col_cluster, 3), np.float32)
        self._ck(self.lib.sindyn_kmeans(self.h, _p(depth), depth.strides[0], _p(labels), _p(pts), _p(ctr)), "kmeans")
        return labels, pts, ctr

    def cluster_order(self):
        img = np.empty((self.H, self.W), np.uint8)
        order = np.full(16, -1, np.int32)
        n = C.c_int(0)
        self._ck(self.lib.sindyn_cluster_order(self.h, _p(img), _p(order), C.byref(n)), "cluster_order")
        return img, order[:n.value].copy()

    def depth_edges(self, depth, capacity=8192):
        depth = np.ascontiguousarray(depth, np.uint16)
        ta = np.empty((self.H, self.W), np.uint8)
        ge = np.empty((self.H, self.W), np.uint8)
        ep = np.zeros((capacity, 2), np.int32)
        n = C.c_int(0)
        self._ck(self.lib.sindyn_depth_edges(self.h, _p(depth), depth.strides[0], _p(ta), _p(ge), _p(ep), capacity, C.byref(n)), "depth_edges")
        return ta, ge, ep[:n.value].copy()

    def plane_edges(self, depth):
        depth = np.ascontiguousarray(depth, np.uint16)
        out = np.empty((self.H, self.W), np.uint8)
        self._ck(self.lib.sindyn_plane_edges(self.h, _p(depth), depth.strides[0], _p(out)), "plane_edges")
        return out

    def peac_debug(self):
        member = np.zeros((self.H, self.W), np.int32)
        planes = np.zeros((64, 3), np.int32)
        n, nf = C.c_int(0), C.c_int(0)
        self._ck(self.lib.sindyn_get_peac_debug(self.h, _p(member), _p(planes), C.byref(n), C.byref(nf)), "get_peac_debug")
        return dict(member=member, planes=planes[:n.value].copy(), n_final=nf.value)

    def filter_plane_edges(self, plane_edges, grad_edges, endpoints):
        pe, ge = _u8(plane_edges), _u8(grad_edges)
        ep = np.ascontiguousarray(endpoints, np.int32).reshape(-1, 2)
        o1 = np.empty((self.H, self.W), np.uint8)
        o2 = np.empty((self.H, self.W), np.uint8)
        self._ck(self.lib.sindyn_filter_plane_edges(self.h, _p(pe), _p(ge), _p(ep), ep.shape[0], _p(o1), _p(o2)), "filter_plane_edges")
        return o1, o2

    def recluster(self, occluded1, occluded2, depth):
        o1, o2 = _u8(occluded1), _u8(occluded2)
        depth = np.ascontiguousarray(depth, np.uint16)
        out = np.empty((self.H, self.W), np.uint8)
        n = C.c_int(0)
        self._ck(self.lib.sindyn_recluster(self.h, _p(o1), _p(o2), _p(depth), depth.strides[0], _p(out), C.byref(n)), "recluster")
        return out, n.value

    def dynamic_decide(self, low, high, total_area, labels):
        out = np.empty((self.H, self.W), np.uint8)
        self._ck(self.lib.sindyn_dynamic_decide(self.h, _p(_u8(low)), _p(_u8(high)), _p(_u8(total_area)), _p(_u8(labels)), _p(out)), "dynamic_decide")
        return out

    def get_state(self, which):
        n = self.H * self.W * (3 if which >= 3 else 1)
        out = np.empty(n, np.uint8)
        self._ck(self.lib.sindyn_get_state(self.h, which, _p(out)), "get_state")
        return out.reshape((self.H, self.W, 3) if which >= 3 else (self.H, self.W))

    def set_state(self, which, img):
        img = _u8(img)
        self._ck(self.lib.sindyn_set_state(self.h, which, _p(img)), "set_state")


class Orb:
    """ORBextractor handle (include/ORBextractor.h:54-88)."""

    def __init__(self, nfeatures=1500, scale_factor=1.2, nlevels=8, ini_th=15, min_th=5, width=640, height=480, device=0):
        self.lib = load_library()
        self.h = _vp()
        self.W, self.H, self.nfeatures, self.nlevels = width, height, nfeatures, nlevels
        st = self.lib.sindyn_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th, width, height, device, C.byref(self.h))
        if st != 0:
            raise SindynError(f"sindyn_orb_create failed: {STATUS.get(st, st)}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sindyn_orb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        st = self.lib.sindyn_orb_set_stream(self.h, _vp(stream_ptr))
        if st != 0:
            raise SindynError("orb set_stream")

    @property
    def launches(self):
        return int(self.lib.sindyn_orb_launch_count(self.h))

    def extract(self, gray, mask=None):
        gray = _u8(gray)
        cap = self.nfeatures * 2 + 64
        kps = (Keypoint * cap)()
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        m = None if mask is None else _u8(mask)
        st = self.lib.sindyn_orb_extract(self.h, _p(gray), gray.strides[0], _p(m), 0 if m is None else m.strides[0],
                                         C.cast(kps, C.c_void_p), _p(desc), cap, C.byref(n))
        if st != 0:
            raise SindynError(f"orb_extract: {STATUS.get(st, st)}: {self.lib.sindyn_orb_last_error(self.h).decode()}")
        arr = np.frombuffer(kps, dtype=np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"), ("octave", "i4")]))[:n.value].copy()
        return arr, desc[:n.value].copy()

    KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"), ("octave", "i4")])

    def track_frame(self, sd, bgr, depth, frame_idx, rgb_order=1, dilate_k=15, mask_out=None, label_out=None, kps_out=None, desc_out=None):
        """rgbd_tum_noros.cc:132-139 + Tracking::GrabImageRGBD + ORBextractor::operator() in one call (sindyn_track_frame).
        Returns (dilated mask, labels, key points, descriptors)."""
        bgr = _u8(bgr)
        depth = np.ascontiguousarray(depth, np.uint16)
        cap = self.nfeatures * 2 + 64
        mask = np.empty((self.H, self.W), np.uint8) if mask_out is None else mask_out
        label = np.empty((self.H, self.W), np.uint8) if label_out is None else label_out
        kps = np.zeros(cap, self.KP_DTYPE) if kps_out is None else kps_out
        desc = np.zeros((cap, 32), np.uint8) if desc_out is None else desc_out
        n = C.c_int(0)
        st = self.lib.sindyn_track_frame(sd.h, self.h, _p(bgr), bgr.strides[0], _p(depth), depth.strides[0], int(rgb_order), int(dilate_k),
                                         _p(mask), mask.strides[0], _p(label), label.strides[0], _p(kps), _p(desc), cap, C.byref(n), frame_idx)
        if st != 0:
            raise SindynError(f"track_frame: {STATUS.get(st, st)}: {sd.lib.sindyn_last_error(sd.h).decode()}")
        return mask, label, kps[: n.value], desc[: n.value]

    def track_frame_resident(self, sd, slot, frame_idx, rgb_order=1, dilate_k=15):
        st = self.lib.sindyn_track_frame_resident(sd.h, self.h, slot, int(rgb_order), int(dilate_k), frame_idx)
        if st != 0:
            raise SindynError(f"track_frame_resident: {STATUS.get(st, st)}: {sd.lib.sindyn_last_error(sd.h).decode()}")

    def track_submit(self, sd, bgr, depth, frame_idx, rgb_order=1, dilate_k=15):
        """Asynchronous sindyn_track_frame: upload + enqueue frame `frame_idx` without waiting (at most three frames in flight).
        bgr / depth must stay alive (and unchanged, if pinned) until the frame is collected."""
        bgr, depth = _u8(bgr), np.ascontiguousarray(depth, np.uint16)
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append((bgr, depth))
        st = self.lib.sindyn_track_submit(sd.h, self.h, _p(bgr), bgr.strides[0], _p(depth), depth.strides[0], int(rgb_order), int(dilate_k), frame_idx)
        if st != 0:
            self._inflight.pop()
            raise SindynError(f"track_submit: {STATUS.get(st, st)}: {sd.lib.sindyn_last_error(sd.h).decode()}")

    def track_collect(self, sd, mask_out=None, label_out=None, kps_out=None, desc_out=None):
        """Results of the oldest submitted frame: (mask, labels, keypoints, descriptors)."""
        cap = self.nfeatures * 2 + 64
        mask = mask_out if mask_out is not None else np.empty((self.H, self.W), np.uint8)
        label = label_out if label_out is not None else np.empty((self.H, self.W), np.uint8)
        kps = kps_out if kps_out is not None else np.zeros(cap, self.KP_DTYPE)
        desc = desc_out if desc_out is not None else np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        st = self.lib.sindyn_track_collect(sd.h, self.h, _p(mask), 0, _p(label), 0, _p(kps), _p(desc), min(cap, len(kps)), C.byref(n))
        if getattr(self, "_inflight", None):
            self._inflight.pop(0)
        if st != 0:
            raise SindynError(f"track_collect: {STATUS.get(st, st)}: {sd.lib.sindyn_last_error(sd.h).decode()}")
        return mask, label, kps[: n.value], desc[: n.value]

    def track_join(self, sd):
        """Everything the resident frames have enqueued (several streams) precedes the next operation on sd's stream."""
        st = self.lib.sindyn_track_join(sd.h, self.h)
        if st != 0:
            raise SindynError(f"track_join: {STATUS.get(st, st)}: {sd.lib.sindyn_last_error(sd.h).decode()}")

    def track_results(self, sd):
        cap = self.nfeatures * 2 + 64
        mask = np.empty((self.H, self.W), np.uint8)
        label = np.empty((self.H, self.W), np.uint8)
        kps = np.zeros(cap, self.KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        st = self.lib.sindyn_track_get_results(sd.h, self.h, _p(mask), _p(label), _p(kps), _p(desc), cap, C.byref(n))
        if st != 0:
            raise SindynError(f"track_get_results: {STATUS.get(st, st)}: {sd.lib.sindyn_last_error(sd.h).decode()}")
        return mask, label, kps[: n.value].copy(), desc[: n.value].copy()

    def search_by_projection(self, last, Tcw_cur, Tcw_last, fx, fy, cx, cy, bf, b, th, mono=False, check_orientation=True, blocked=None):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) against the frame resident in this handle.
        last: dict(xyz_w, valid, desc, octave, angle, observed).  Returns (match[n_cur], nmatches)."""
        p = MatchParams()
        p.fx, p.fy, p.cx, p.cy, p.bf, p.b, p.th = fx, fy, cx, cy, bf, b, th
        p.mono, p.check_orientation = int(mono), int(check_orientation)
        p.Tcw_cur[:] = [float(v) for v in np.asarray(Tcw_cur, np.float32).reshape(-1)]
        p.Tcw_last[:] = [float(v) for v in np.asarray(Tcw_last, np.float32).reshape(-1)]
        xyz = np.ascontiguousarray(last["xyz_w"], np.float32).reshape(-1, 3)
        n_last = len(xyz)
        valid = np.ascontiguousarray(np.asarray(last["valid"]).astype(np.uint8))
        observed = np.ascontiguousarray(np.asarray(last["observed"]).astype(np.uint8))
        desc = _u8(last["desc"]).reshape(-1, 32)
        octave, angle = np.ascontiguousarray(last["octave"], np.int32), np.ascontiguousarray(last["angle"], np.float32)
        cap = self.nfeatures * 2 + 64
        match = np.full(cap, -1, np.int32)
        n_cur, nm = C.c_int(0), C.c_int(0)
        blk = np.ascontiguousarray(np.asarray(blocked).astype(np.uint8)) if blocked is not None else None
        st = self.lib.sindyn_orb_search_by_projection(self.h, C.byref(p), n_last, _p(xyz), _p(valid), _p(desc), _p(octave), _p(angle), _p(observed),
                                                      _p(blk) if blk is not None else None, _p(match), cap, C.byref(n_cur), C.byref(nm))
        if st != 0:
            raise SindynError(f"search_by_projection: {STATUS.get(st, st)}: {self.lib.sindyn_orb_last_error(self.h).decode()}")
        return match[: n_cur.value].copy(), nm.value

    def frame_features(self, depth_raw, fx, fy, cx, cy, dist, bf, depth_map_factor):
        """Frame.cc:143-170 on the keypoints of the last extract(): (keys_un, depth, u_right, bounds, offsets, indices)."""
        depth_raw = np.ascontiguousarray(depth_raw, np.uint16)
        p = FrameParams(fx, fy, cx, cy, *[float(v) for v in dist], bf, depth_map_factor)
        cap = self.nfeatures * 2 + 64
        un = np.zeros((cap, 2), np.float32)
        dep = np.zeros(cap, np.float32)
        ur = np.zeros(cap, np.float32)
        b = np.zeros(4, np.float32)
        off = np.zeros(64 * 48 + 1, np.int32)
        idx = np.zeros(cap, np.int32)
        n = C.c_int(0)
        st = self.lib.sindyn_orb_frame_features(self.h, _p(depth_raw), depth_raw.strides[0], C.byref(p), _p(un), _p(dep), _p(ur), _p(b),
                                                _p(off), _p(idx), cap, C.byref(n))
        if st != 0:
            raise SindynError(f"orb_frame_features: {STATUS.get(st, st)}: {self.lib.sindyn_orb_last_error(self.h).decode()}")
        k = n.value
        return un[:k].copy(), dep[:k].copy(), ur[:k].copy(), b, off, idx[:int(off[-1])].copy()

    def candidates(self, level, capacity=16384):
        buf = np.zeros((capacity, 3), np.int32)
        n = C.c_int(0)
        st = self.lib.sindyn_orb_get_candidates(self.h, level, _p(buf), capacity, C.byref(n))
        if st != 0:
            raise SindynError("orb candidates")
        return buf[:n.value].copy()

    def plane(self, level, which):
        out = np.zeros((self.W + 38) * (self.H + 38), np.uint8)
        w, h = C.c_int(0), C.c_int(0)
        st = self.lib.sindyn_orb_get_plane(self.h, level, which, _p(out), C.byref(w), C.byref(h))
        if st != 0:
            raise SindynError("orb plane")
        return out[: w.value * h.value].reshape(h.value, w.value).copy()

    def pyramid_level(self, level):
        out = np.zeros(self.W * self.H, np.uint8)
        w, h = C.c_int(0), C.c_int(0)
        st = self.lib.sindyn_orb_get_pyramid_level(self.h, level, _p(out), C.byref(w), C.byref(h))
        if st != 0:
            raise SindynError("orb pyramid_level")
        return out[: w.value * h.value].reshape(h.value, w.value).copy()
