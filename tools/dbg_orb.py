import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from sindslam_b200 import synth
from sindslam_b200.capi import Orb
_, frames = synth.make_sequence(1, synth.TUM3, seq=0, kind="box", start=8)
gray = cv2.cvtColor(frames[0].bgr, cv2.COLOR_BGR2GRAY)
orb = Orb(1500, 1.2, 8, 15, 5, 640, 480)
try:
    k, d = orb.extract(gray, None)
    np.save("gpurun_out/dbg_kps.npy", k); np.save("gpurun_out/dbg_desc.npy", d)
except Exception as e:
    print("extract failed:", e)
for l in range(8):
    np.save("gpurun_out/dbg_cand%d.npy" % l, orb.candidates(l))
for which in range(3):
    np.save("gpurun_out/dbg_plane%d.npy" % which, orb.plane(0, which))
print("done")
