"""GPU parity of the PEAC plane-contour edge stage (sindyn_plane_edges = DynaDetect.cc:558-593 + include/PEAC) against
oracle/peac_oracle.py / oracle/peac_cpu.c.  Everything is bit-exact: the block-level clustering (which blocks form which
plane, in which order), the pixel-level region growing (the device replays the reference's single FIFO queue level by
level, AHCPlaneFitter.hpp:546-594), the final re-merge, and the u8 plane-edge image."""
import cv2
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from oracle import peac_oracle as po
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu


def _iou(a, b):
    u = (a | b).sum()
    return 1.0 if u == 0 else float((a & b).sum()) / float(u)


@pytest.mark.parametrize("cam_name,kind,hole", [("TUM3", "box", 0.0003), ("D455_848", "humanoid", 0.0003), ("TUM3", "box", 0.002), ("D455_848", "humanoid", 0.0)])
def test_plane_edges_vs_oracle(cam_name, kind, hole):
    from sindslam_b200.capi import SinDyn
    cam = getattr(synth, cam_name)
    _, frames = synth.make_sequence(3, cam, seq=3, kind=kind, start=8, hole_rate=hole)
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1)
    for f in frames:
        got = sd.plane_edges(f.depth)
        dbg = sd.peac_debug()
        ref_dbg = {}
        ref = po.plane_edges(f.depth, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, ref_dbg)
        # block-level AHC: same extracted planes (root block id, size), same order
        assert [(int(r), int(n)) for r, n, _ in dbg["planes"]] == ref_dbg["coarse"]
        assert dbg["n_final"] == ref_dbg["n"]
        gm, rm = dbg["member"], ref_dbg["member"]
        print("%s hole %.4f: planes %d -> %d, membership differing px %d, edge differing px %d, FIFO stats %s" % (
            cam_name, hole, len(ref_dbg["coarse"]), ref_dbg["n"], int((gm != rm).sum()), int((got != ref).sum()), ref_dbg.get("stats")))
        assert np.array_equal(gm, rm)          # final plane id of every pixel
        assert np.array_equal(got, ref)        # imgEdgeByPlane
    sd.close()


def test_detect_with_plane_edges_mask_iou(seq_c1):
    """End to end with the PEAC stage on: dynamic mask of sindyn_detect vs the oracle fed with the same low/high masks."""
    from sindslam_b200.capi import SinDyn
    cam = synth.TUM3
    _, frames = synth.make_sequence(5, cam, seq=0, kind="box", start=8, hole_rate=0.0003)
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1)
    s.set_prev_frames(frames[0].bgr, frames[0].bgr)
    o = orc.DynaDetectOracle(frames[0].bgr, frames[0].bgr, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=True)
    ious = []
    for k in range(1, 5):
        mask, label = s.detect(frames[k].bgr, frames[k].depth, k)
        fr = s.flow_results()
        r = o.detect(frames[k].bgr, frames[k].depth, inject_masks=(fr["low"], fr["high"]))
        iou = _iou(mask == 255, r["mask"] == 255)
        ious.append(iou)
        print("frame %d: mask IoU(255) %.4f, label agreement %.4f, labels %d/%d" % (k, iou, float((label == r["label"]).mean()), int(label.max()), int(r["label"].max())))
        assert np.array_equal(mask, r["mask"]) and np.array_equal(label, r["label"]), k     # free-running state on both sides
    assert min(ious) >= 0.99
    s.close()


def _crafted_depth(kind, W=640, H=480, factor=5000.0, seed=5):
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    if kind == "wall":            # one slanted plane filling the image: a single graph component of 1200 blocks
        z = 2.0 + 0.4 * u / W + 0.1 * v / H
    elif kind == "corner":        # two walls meeting at a vertical edge + a floor
        z = np.where(u < W * 0.45, 1.5 + 1.2 * (W * 0.45 - u) / W, 1.5 + 0.9 * (u - W * 0.45) / W)
        z = np.where(v > H * 0.7, np.minimum(z, 1.0 + 2.5 * (H - v) / H), z)
    else:                         # "steps": many small fronto-parallel patches (lots of small components, nothing reaches minSupport in places)
        z = 1.0 + 0.25 * ((u // 48).astype(int) % 5) + 0.2 * ((v // 40).astype(int) % 3)
    z = z + rng.normal(0, 0.0015, z.shape)
    d = np.clip(np.rint(z * factor), 0, 65535).astype(np.uint16)
    d[rng.random(d.shape) < 0.0002] = 0
    return d


@pytest.mark.parametrize("kind", ["wall", "corner", "steps"])
def test_plane_edges_crafted_scenes(kind):
    """Extremes of the block graph: ONE component that covers the whole image (the longest serial chain, the widest
    candidate lists), two large planes meeting, and a scene of many small patches -- still bit-exact."""
    from sindslam_b200.capi import SinDyn
    cam = synth.TUM3
    depth = _crafted_depth(kind)
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1)
    got = sd.plane_edges(depth)
    dbg = sd.peac_debug()
    ref_dbg = {}
    ref = po.plane_edges(depth, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, ref_dbg)
    print(kind, "planes", ref_dbg["coarse"], "->", ref_dbg["n"], ref_dbg.get("stats"))
    assert [(int(r), int(n)) for r, n, _ in dbg["planes"]] == ref_dbg["coarse"]
    assert dbg["n_final"] == ref_dbg["n"]
    assert np.array_equal(dbg["member"], ref_dbg["member"])
    assert np.array_equal(got, ref)
    sd.close()
