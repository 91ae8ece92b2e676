"""One un-graphed Brox solve on resident synthetic frames (profiling target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sindslam_b200 import synth
from sindslam_b200.capi import SinDyn
cam = synth.TUM3
_, frames = synth.make_sequence(3, cam, seq=0, kind="box", start=8)
sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=0, use_graphs=0)
sd.set_prev_frames(frames[1].bgr, frames[0].bgr)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    p = sd.brox_profile()
print(p)
