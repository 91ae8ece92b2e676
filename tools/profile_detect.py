"""Profiling target: N frames of the full per-frame path (sindyn_detect + 15x15 dilation + masked ORB) without CUDA
graphs, so that every kernel shows up in the ncu launch list.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file out.csv python tools/profile_detect.py 3
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2

from sindslam_b200 import synth
from sindslam_b200.capi import Orb, SinDyn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cam = synth.TUM3
_, frames = synth.make_sequence(n + 2, cam, seq=0, kind="box", start=8, hole_rate=0.0003)
sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, use_graphs=0, stage_timing=1)
orb = Orb(1500, 1.2, 8, 15, 5, cam.width, cam.height)
sd.set_prev_frames(frames[1].bgr, frames[0].bgr)
for k in range(2, n + 2):
    mask, label = sd.detect(frames[k].bgr, frames[k].depth, k)
    mask = sd.morph_ellipse(mask, 15, 0)
    kps, _ = orb.extract(cv2.cvtColor(frames[k].bgr, cv2.COLOR_RGB2GRAY), mask)
    print("frame", k, "dyn px", int((mask == 255).sum()), "labels", int(label.max()), "kps", len(kps), "stage ms", [round(float(v), 3) for v in sd.stage_ms()[:11]])
