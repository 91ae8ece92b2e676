"""Developer aid: the streamed full-detect parity check of tests/test_sequence_gpu.py over many random synthetic sequences
(seeds, object kinds, depth-hole rates, both camera models).  Any label / mask / key-point mismatch raises."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from sindslam_b200 import synth
from test_sequence_gpu import _stream

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
t0 = time.time()
n = 0
seq = 1000
while time.time() - t0 < budget:
    kind = ("box", "humanoid")[seq % 2]
    cam = (synth.TUM3, synth.D455_848)[(seq // 2) % 2]
    hole = (0.015, 0.0003, 0.05)[seq % 3]
    _, frames = synth.make_sequence(7, cam, seq=seq, kind=kind, start=3 + seq % 5, hole_rate=hole)
    order = list(range(7)) if seq % 4 else [0, 1, 2, 6, 3, 4, 5]     # every 4th sequence contains frame jumps
    n_large, ious = _stream(cam, frames, order, orb_every=2)
    print("seq %d %s %dx%d holes %.4f: ok, large-motion frames %d, IoU vs truth max %.3f" % (seq, kind, cam.width, cam.height, hole, n_large, max(ious)), flush=True)
    n += 1
    seq += 1
print("fuzz parity: %d sequences bit-exact in %.0f s" % (n, time.time() - t0))
