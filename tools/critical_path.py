"""Developer tool: which branch of the per-frame graph limits the frame time?  Streams the resident full path
(sindyn_track_frame_resident: detect + 15x15 dilation + masked ORB, CUDA graphs on) over a short synthetic sequence with the
PEAC plane-edge branch on and off and prints the device time per frame pair of both.

    python tools/critical_path.py [frames]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sindslam_b200 import synth
from sindslam_b200.capi import Orb, SinDyn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cam = synth.TUM3
_, frames = synth.make_sequence_parallel(n, cam, seq=3, kind="box", start=0, hole_rate=0.0005)
order = (list(range(1, n)) + list(range(n - 2, -1, -1))) * 4
for pe in (1, 0, 1, 0):
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=1, plane_edges=pe)
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
    orb.track_results(sd)
    print("plane_edges=%d: %.3f ms per frame pair (%.1f pairs/s)" % (pe, e0.elapsed_time(e1) / m, 1e3 * m / e0.elapsed_time(e1)))
    sd.close()
