import sys, time, numpy as np, cv2
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sindslam_b200 import synth
from sindslam_b200.capi import Orb
cam = synth.TUM3
_, frames = synth.make_sequence(6, cam, seq=1, kind="box", start=4)
orb = Orb(1500, 1.2, 8, 15, 5, 640, 480)
fl, fc = frames[2], frames[3]
kps, desc = orb.extract(cv2.cvtColor(fl.bgr, cv2.COLOR_BGR2GRAY), None)
un, dep, ur, b, off, idx = orb.frame_features(fl.depth, cam.fx, cam.fy, cam.cx, cam.cy, (0,0,0,0,0), 40.0, 1/5000.0)
n = len(kps); z = dep[:n]
pc = np.stack([(un[:n,0]-cam.cx)*z/cam.fx, (un[:n,1]-cam.cy)*z/cam.fy, z, np.ones(n)], 1)
last = dict(xyz_w=(fl.T_wc @ pc.T).T[:, :3].astype(np.float32), valid=z>0, desc=desc, octave=kps["octave"], angle=kps["angle"], observed=np.zeros(n, bool))
orb.extract(cv2.cvtColor(fc.bgr, cv2.COLOR_BGR2GRAY), None)
orb.frame_features(fc.depth, cam.fx, cam.fy, cam.cx, cam.cy, (0,0,0,0,0), 40.0, 1/5000.0)
Tc, Tl = np.linalg.inv(fc.T_wc), np.linalg.inv(fl.T_wc)
for th in (15.0, 30.0):
    for _ in range(3): orb.search_by_projection(last, Tc, Tl, cam.fx, cam.fy, cam.cx, cam.cy, 40.0, 0.0747, th)
    t0 = time.perf_counter()
    for _ in range(20): m, nm = orb.search_by_projection(last, Tc, Tl, cam.fx, cam.fy, cam.cx, cam.cy, 40.0, 0.0747, th)
    print("th", th, "n_last", n, "matches", nm, "ms per call (wall, incl. python + H2D/D2H)", 1e3*(time.perf_counter()-t0)/20)
