"""Developer aid: streams random synthetic sequences through the full sindyn_detect WITH the PEAC plane-contour edges enabled
and checks invariants (no error, mask values, label range, masks identical when the same stream is replayed on a second
handle = determinism of the three-stream graph)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from sindslam_b200 import synth
from sindslam_b200.capi import SinDyn

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
t0 = time.time()
seq, n = 2000, 0
while time.time() - t0 < budget:
    kind = ("box", "humanoid")[seq % 2]
    cam = (synth.TUM3, synth.D455_848)[(seq // 2) % 2]
    hole = (0.0003, 0.0, 0.002)[seq % 3]
    _, frames = synth.make_sequence(8, cam, seq=seq, kind=kind, start=2 + seq % 7, hole_rate=hole)
    outs = []
    for rep in range(2):
        s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1)
        s.set_prev_frames(frames[0].bgr, frames[0].bgr)
        res = []
        for k in range(1, 8):
            mask, label = s.detect(frames[k].bgr, frames[k].depth, k)
            assert set(np.unique(mask)) <= {0, 125, 255}, np.unique(mask)
            assert label.max() < 128
            res.append((mask.copy(), label.copy()))
        planes = s.peac_debug()["n_final"]
        s.close()
        outs.append(res)
    for (m0, l0), (m1, l1) in zip(*outs):
        assert np.array_equal(m0, m1) and np.array_equal(l0, l1), "non-deterministic result"
    print("seq %d %s %dx%d holes %.4f: ok (planes in the last frame: %s, dynamic px %d)" % (seq, kind, cam.width, cam.height, hole, planes, int((outs[0][-1][0] == 255).sum())), flush=True)
    seq += 1
    n += 1
print("PEAC fuzz: %d sequences, deterministic, in %.0f s" % (n, time.time() - t0))
