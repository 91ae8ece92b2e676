"""Developer tool: where k_rho spends its cycles on the sample lists of real frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sindslam_b200 import synth
from sindslam_b200.capi import SinDyn
cam = synth.TUM3
_, frames = synth.make_sequence(6, cam, seq=3, kind="box", start=8, hole_rate=0.0005)
sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1, stage_timing=1)
sd.set_prev_frames(frames[0].bgr, frames[0].bgr)
for k in range(1, 6):
    sd.detect(frames[k].bgr, frames[k].depth, k)
    fr = sd.flow_results()
    st = sd.get_state  # noqa
    p, q = sd.sample_pairs(fr["flow"])
    H, m, info = sd.find_homography_rho(p, q)
    print("frame", k, "n", info[0], "inliers", info[1], "models", info[2], "iters", info[7], "lm", info[3], "kcycles loop/nstar/lm", info[4], info[5], info[6], "sample / eval+sprt / update / fallbacks", info[8:12].tolist(),
          "homography stage ms %.3f" % sd.stage_ms()[3])
