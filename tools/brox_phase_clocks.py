"""Per-phase clock breakdown of k_brox_sor (32x24 tiles); needs a build with SINDYN_NVCC_EXTRA=-DSINDYN_BROX_PHASE_CLOCKS."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sindslam_b200 import synth, capi
from sindslam_b200.capi import SinDyn
cam = synth.TUM3
_, frames = synth.make_sequence(3, cam, seq=0, kind="box", start=8)
sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=0, use_graphs=0)
sd.set_prev_frames(frames[1].bgr, frames[0].bgr)
lib = ctypes.CDLL(capi.LIB_PATH)
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3): sd.brox_profile()
lib.sindyn_dbg_brox_phase_clocks(buf, 1)
for _ in range(5): sd.brox_profile()
lib.sindyn_dbg_brox_phase_clocks(buf, 0)
n = buf[15]
names = ["stage (cp.async issue)", "table + systems + wait", "sweeps", "write"]
tot = sum(buf[i] for i in range(4))
for i, nm in enumerate(names): print("%-26s %8.0f cyc/launch  %5.1f %%" % (nm, buf[i] / n, 100.0 * buf[i] / tot))
print("launches", n, "total cyc/launch", tot / n)
