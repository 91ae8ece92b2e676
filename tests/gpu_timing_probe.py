"""Ad-hoc timing probe run on the GPU box (not a test)."""
import time
import numpy as np
from sindslam_b200 import synth
from sindslam_b200.capi import SinDyn
from oracle import dynadetect_oracle as orc

scene, fr = synth.make_sequence(3)
I0 = orc.gray_small(orc.bgr2gray(fr[2].bgr)).astype(np.float32) / 255
I1 = orc.gray_small(orc.bgr2gray(fr[0].bgr)).astype(np.float32) / 255
for graphs in (0, 1):
    s = SinDyn(640, 480, use_graphs=graphs)
    for _ in range(3):
        s.flow_brox(I0, I1)
    t = time.time()
    n = 20
    for _ in range(n):
        s.flow_brox(I0, I1)
    dt = (time.time() - t) / n
    print("brox graphs=%d: %.3f ms per call incl. copies (launches so far %d)" % (graphs, dt * 1e3, s.launches))
    t = time.time()
    for _ in range(n):
        s.kmeans(fr[2].depth)
    print("kmeans: %.3f ms per call incl. copies" % ((time.time() - t) / n * 1e3))
    t = time.time()
    for _ in range(n):
        s.depth_edges(fr[2].depth)
    print("depth_edges: %.3f ms per call incl. copies" % ((time.time() - t) / n * 1e3))
    s.close()
