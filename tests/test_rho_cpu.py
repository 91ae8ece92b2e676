"""Pins oracle/rho_cpu.c -- the restatement of OpenCV's RHO homography estimator that the CUDA kernel k_rho follows step by
step (DynaDetect.cc:1235: cv::findHomography(pts, ptsLast, noArray(), RHO); calib3d rho.cpp is un-vendored) -- against the
REAL library: cv2.findHomography(..., cv2.RHO) must return the same inlier mask and a bit-identical H."""
import ctypes
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from sindslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rho():
    path = os.path.join(ROOT, "oracle", "_build", "librho_cpu.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    lib.rho_find_homography.argtypes = [vp, vp, ctypes.c_uint, ctypes.c_float, ctypes.c_uint, ctypes.c_double, ctypes.c_double, vp, vp, vp]
    lib.rho_find_homography.restype = ctypes.c_uint

    def run(src, dst):
        src, dst = np.ascontiguousarray(src, np.float32), np.ascontiguousarray(dst, np.float32)
        H = np.zeros(9)
        m = np.zeros(len(src), np.uint8)
        d = np.zeros(8, np.uint32)
        n = lib.rho_find_homography(src.ctypes.data, dst.ctypes.data, len(src), 3.0, 2000, 0.995, 0.35, H.ctypes.data, m.ctypes.data, d.ctypes.data)
        return (H.reshape(3, 3) if n else None), m, d
    return run


def _case(rng, N, outl, noise, grid):
    if grid:
        g = np.stack(np.meshgrid(np.arange(10, 640, 10), np.arange(10, 480, 10)), -1).reshape(-1, 2).astype(np.float32)
        g = g[rng.permutation(len(g))[:N]]
    else:
        g = (rng.random((N, 2)) * [640, 480]).astype(np.float32)
    Ht = np.eye(3) + rng.normal(0, 1, (3, 3)) * [[0.01, 0.01, 3], [0.01, 0.01, 3], [1e-5, 1e-5, 0]]
    ph = np.concatenate([g, np.ones((N, 1))], 1) @ Ht.T
    d = (ph[:, :2] / ph[:, 2:]).astype(np.float32) + rng.normal(0, noise, (N, 2)).astype(np.float32)
    k = int(outl * N)
    idx = rng.permutation(N)[:k]
    d[idx] += rng.normal(0, 30, (k, 2)).astype(np.float32)
    return np.ascontiguousarray(g, np.float32), np.ascontiguousarray(d, np.float32)


def test_rho_restatement_equals_cv2_random(rho):
    rng = np.random.default_rng(7)
    n = 0
    for t in range(200):
        N = int(rng.choice([5, 6, 12, 50, 300, 1500, 2961]))
        s, d = _case(rng, N, float(rng.choice([0, 0.1, 0.3, 0.5, 0.6])), float(rng.choice([0, 0.02, 0.1, 0.5])), t % 2 == 0)
        Hc, mc = cv2.findHomography(s, d, cv2.RHO)
        Hm, mm, diag = rho(s, d)
        if Hc is None:
            assert Hm is None
            continue
        assert Hm is not None and np.array_equal(mm, mc.ravel()), (t, N)
        assert np.array_equal(Hm, Hc), (t, N, float(np.abs(Hm - Hc).max()), diag.tolist())
        n += 1
    assert n >= 180


def test_rho_restatement_equals_cv2_on_frame_sample_lists(rho, seq_c1):
    """The lists the pipeline actually feeds: weighted, sorted 10-px grid samples of a dense flow (DynaDetect.cc:1163-1231)."""
    scene, frames = seq_c1
    cam = synth.TUM3
    z = np.zeros((cam.height, cam.width), np.uint8)
    dyn = np.full((cam.height, cam.width), 125, np.uint8)
    dyn[frames[2].dyn_mask] = 255
    lab = (np.arange(cam.width)[None, :] // 80 + 1).repeat(cam.height, 0).astype(np.uint8)
    rng = np.random.default_rng(2)
    for k, (dl, ll) in enumerate(((z, z), (dyn, lab))):
        flow = -synth.gt_flow(scene, cam, 10, 8, frames[2]) + rng.normal(0, 0.05 + 0.2 * k, (cam.height, cam.width, 2)).astype(np.float32)
        p, q = orc.sample_pairs(flow.astype(np.float32), dl, ll)
        Hc, mc = cv2.findHomography(p, q, cv2.RHO)
        Hm, mm, diag = rho(p, q)
        print("frame list %d: %d samples, %d inliers, %d models, %d LM iterations" % (k, len(p), int(mm.sum()), diag[1], diag[6]))
        assert np.array_equal(mm, mc.ravel()) and np.array_equal(Hm, Hc)
