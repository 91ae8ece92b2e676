import os, sys
sys.path.insert(0, "/root/repo")
import torch
from sindslam_b200 import synth
from sindslam_b200.capi import Orb, SinDyn
n = 40
cam = synth.TUM3
_, frames = synth.make_sequence_parallel(n, cam, seq=3, kind="box", start=0, hole_rate=0.0005)
order = (list(range(1, n)) + list(range(n - 2, -1, -1))) * 4
for inner in (10, 5, 2):
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=1, plane_edges=1, brox_inner=inner)
    orb = Orb(1500, 1.2, 8, 15, 5, cam.width, cam.height)
    stream = torch.cuda.Stream()
    sd.set_stream(stream.cuda_stream)
    for i, f in enumerate(frames):
        sd.upload_frame(i, f.bgr, f.depth)
    sd.set_prev_frames(frames[0].bgr, frames[0].bgr)
    for j in range(20):
        orb.track_frame_resident(sd, order[j], j)
    orb.track_join(sd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    m = 120
    for j in range(20, 20 + m):
        orb.track_frame_resident(sd, order[j], j)
    orb.track_join(sd)
    e1.record(stream)
    torch.cuda.synchronize()
    print("brox_inner=%d: %.3f ms per frame pair" % (inner, e0.elapsed_time(e1) / m))
    sd.close()
