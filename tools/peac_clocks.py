"""Developer tool: phase clocks of k_peac_ahc (build with SINDYN_NVCC_EXTRA=-DPEAC_CLOCKS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sindslam_b200 import synth
from sindslam_b200.capi import SinDyn
cam = synth.TUM3
_, frames = synth.make_sequence(3, cam, seq=3, kind="box", start=8, hole_rate=0.0005)
sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1)
for f in frames:
    sd.plane_edges(f.depth)
    sd.peac_debug()
    print("----")
