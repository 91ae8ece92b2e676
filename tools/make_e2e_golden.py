#!/usr/bin/env python
"""Generates tests/golden/e2e_<name>.npz: the CPU oracle (oracle/dynadetect_oracle.py) streamed FREE-RUNNING and UN-INJECTED
over a synthetic sequence -- its own CPU Brox flow (oracle/brox_cpu.c), the real cv2.VariationalRefinement, the real
cv2.findHomography(RHO), PEAC plane edges on, its own state recurrence (DynaDetect.cc:1377-1666, driver loop
rgbd_tum_noros.cc:113-139).  The oracle run does not depend on anything the GPU computes, so it is generated here (CPU
container, ~3 s per frame) and committed; tests/test_e2e_gpu.py regenerates the same frames (synth is deterministic, a frame
checksum is stored) and compares sindyn_detect's free-running outputs with these masks / labels.

    python tools/make_e2e_golden.py c3 | c4 | all
"""
from __future__ import annotations

import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from oracle import dynadetect_oracle as orc
from sindslam_b200 import synth

# name -> (camera, kind, seq, n_frames, hole_rate): BASELINE.json configs[2] and configs[3]
CONFIGS = {
    "c3": ("TUM3", "box", 3, 300, 0.0005),           # 300-frame walking_xyz-shaped 640x480 sequence
    "c4": ("D455_848", "humanoid", 4, 72, 0.0005),   # 848x480 D455-shaped, humanoid-sized dynamic region
}


def frames_iter(name, n=None):
    cam_name, kind, seq, n_frames, hole = CONFIGS[name]
    cam = getattr(synth, cam_name)
    scene = synth.Scene(synth.BASE_SEED + seq, kind)
    scene.hole_rate = hole
    for i in range(n_frames if n is None else n):
        yield cam, synth.render_frame(scene, cam, i)


def frame_crc(f):
    return zlib.crc32(f.depth.tobytes(), zlib.crc32(f.bgr.tobytes())) & 0xFFFFFFFF


def generate(name, n=None):
    masks, labels, crcs, thr, Hs, lm, t_all = [], [], [], [], [], [], time.time()
    o = None
    for k, (cam, f) in enumerate(frames_iter(name, n)):
        crcs.append(frame_crc(f))
        if k == 0:
            # rgbd_tum_noros.cc:103-107: the detector is primed with frame 0 twice; frame 0 itself is not detected
            o = orc.DynaDetectOracle(f.bgr, f.bgr, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=True, engine="brox", refine=True)
            continue
        t0 = time.time()
        r = o.detect(f.bgr, f.depth)
        masks.append(r["mask"]); labels.append(r["label"])
        thr.append(r["flow"]["thr"]); Hs.append(r["flow"]["H"]); lm.append(r["flow"]["large_motion"])
        print("%s frame %d: %.1f s, dyn px %d, labels %d, lm %d" % (name, k, time.time() - t0, int((r["mask"] == 255).sum()), int(r["label"].max()), lm[-1]), flush=True)
    out = os.path.join(ROOT, "tests", "golden", "e2e_%s.npz" % name)
    np.savez_compressed(out, mask=np.stack(masks), label=np.stack(labels), crc=np.array(crcs, np.uint32), thr=np.stack(thr).astype(np.float32),
                        H=np.stack(Hs), large_motion=np.array(lm, np.uint8), config=np.array(list(map(str, CONFIGS[name]))))
    print("wrote", out, os.path.getsize(out) >> 10, "KiB in %.0f s" % (time.time() - t_all))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else None
    for nm in (CONFIGS if which == "all" else [which]):
        generate(nm, n)
