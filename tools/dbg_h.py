import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sindslam_b200 import synth
from sindslam_b200.capi import SinDyn
cam = synth.TUM3
_, frames = synth.make_sequence(5, cam, seq=0, kind="box", start=8)
sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=1)
sd.set_prev_frames(frames[1].bgr, frames[0].bgr)
for k in (2, 3):
    sd.flow_residual(frames[k].bgr, roll=True)
    r = sd.flow_results()
    np.save("gpurun_out/dbg_flow%d.npy" % k, r["flow"]); np.save("gpurun_out/dbg_H%d.npy" % k, r["H"])
print("ok")
