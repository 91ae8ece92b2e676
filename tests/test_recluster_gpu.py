"""GPU parity of the clustering / decision half of DetectDynaArea against the CPU oracle (bit-exact given identical
upstream inputs): plane-edge filter (DynaDetect.cc:598-641), SegAndMergeV2 (:653-1018), per-cluster decision
(:1553-1636), and the full sindyn_detect (DynaDetect.cc:1377-1666) streamed over a synthetic sequence."""
import cv2
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu


def _iou(a, b):
    u = (a | b).sum()
    return 1.0 if u == 0 else float((a & b).sum()) / float(u)


@pytest.fixture(scope="module")
def sd():
    from sindslam_b200.capi import SinDyn
    cam = synth.TUM3
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=0, refine=0)
    yield s
    s.close()


def _fake_plane_edges(shape, seed):
    """Synthetic imgEdgeByPlane: a few thick polylines / rectangles (what PEAC plane contours look like)."""
    rng = np.random.default_rng(seed)
    img = np.zeros(shape, np.uint8)
    for _ in range(6):
        x0, y0 = int(rng.integers(20, shape[1] - 200)), int(rng.integers(20, shape[0] - 150))
        cv2.rectangle(img, (x0, y0), (x0 + int(rng.integers(40, 180)), y0 + int(rng.integers(30, 130))), 255, 2)
    for _ in range(4):
        p = rng.integers(0, min(shape), (2, 2))
        cv2.line(img, tuple(int(v) for v in p[0]), tuple(int(v) for v in p[1]), 255, 2)
    return img


def test_filter_plane_edges_bit_exact(sd, seq_c1):
    _, frames = seq_c1
    cam = synth.TUM3
    for k, f in enumerate(frames[:3]):
        _, grad, ep = orc.depth_edges(f.depth, cam.depth_factor)
        plane = _fake_plane_edges(grad.shape, k)
        o1, o2 = orc.filter_plane_edges(plane, grad, ep)
        g1, g2 = sd.filter_plane_edges(plane, grad, ep)
        print("occl2 px", int((o2 > 0).sum()), "mismatch", int((o1 != g1).sum()), int((o2 != g2).sum()))
        assert np.array_equal(g2, o2)
        assert np.array_equal(g1, o1)
    # no plane edges: occluded1 = CLOSE3(gradient edges), occluded2 empty
    z = np.zeros_like(grad)
    g1, g2 = sd.filter_plane_edges(z, grad, ep)
    o1, o2 = orc.filter_plane_edges(z, grad, ep)
    assert np.array_equal(g1, o1) and not g2.any()


def test_recluster_bit_exact(sd, seq_c1):
    _, frames = seq_c1
    cam = synth.TUM3
    z = np.zeros((cam.height, cam.width), np.uint8)
    sd.set_state(2, z)
    for k, f in enumerate(frames[:3]):
        lab, pts, ctr = sd.kmeans(f.depth)
        olab, opts, octr = orc.seg_by_kmeans(f.depth, z, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, "fx")
        assert np.array_equal(lab, olab)
        kept, seg, _ = orc.cluster_order(olab, octr)
        _, grad, ep = orc.depth_edges(f.depth, cam.depth_factor)
        plane = _fake_plane_edges(grad.shape, 10 + k) if k else np.zeros_like(grad)
        o1, o2 = orc.filter_plane_edges(plane, grad, ep)
        dbg = {}
        ref = orc.seg_and_merge_v2(kept, olab, o1, o2, seg, opts, f.depth, debug=dbg)
        got, n = sd.recluster(o1, o2, f.depth)
        d = sd.recluster_debug()
        print("frame", k, "components", n, len(dbg["clusters"]), "labels", int(ref.max()), int(got.max()), "mismatch px", int((got != ref).sum()))
        assert n == len(dbg["clusters"])
        assert np.array_equal(np.sort(d["area"]), np.sort(np.array([c["area"] for c in dbg["clusters"]], np.int32)))
        assert np.allclose(d["T"], dbg["T"], rtol=1e-5, atol=1e-6), float(np.abs(d["T"] - dbg["T"]).max())
        assert np.array_equal(got, ref)


def test_dynamic_decide_bit_exact(sd, seq_c1):
    _, frames = seq_c1
    cam = synth.TUM3
    rng = np.random.default_rng(5)
    z = np.zeros((cam.height, cam.width), np.uint8)
    o = orc.DynaDetectOracle(frames[1].bgr, frames[0].bgr, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=False)
    high_last = z
    for k in (2, 3, 4):
        f = frames[k]
        # plausible masks: the true dynamic mask plus blobs of noise, high is a subset of low
        noise = cv2.GaussianBlur(rng.random(z.shape).astype(np.float32), (0, 0), 6)
        dm = f.dyn_mask.astype(np.uint8)
        low = (((dm > 0) | (noise > np.quantile(noise, 0.93))).astype(np.uint8)) * 128
        high = (((cv2.erode(dm, np.ones((5, 5), np.uint8)) > 0) | (noise > np.quantile(noise, 0.985))).astype(np.uint8)) * 255
        r = o.detect(f.bgr, f.depth, inject_masks=(low, high))
        sd.set_state(1, high_last)
        got = sd.dynamic_decide(low, high, r["total_area"], r["label"])
        print("frame", k, "dyn px", int((r["mask"] == 255).sum()), "mismatch", int((got != r["mask"]).sum()))
        assert np.array_equal(got, r["mask"])
        assert np.array_equal(sd.get_state(0), r["mask"]) and np.array_equal(sd.get_state(1), high)
        high_last = high


def test_detect_stream_bit_exact_given_identical_flow(seq_c1):
    """Full sindyn_detect over a sequence.  The oracle is fed the GPU's own low/high masks (identical flow), so
    labels and the dynamic mask must be bit-exact; the state recurrence (labels -> next k-means init, dyna -> sample
    weights) is exercised because both sides roll their own state."""
    from sindslam_b200.capi import SinDyn
    _, frames = seq_c1
    cam = synth.TUM3
    s = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=0, refine=0, stage_timing=1)
    s.set_prev_frames(frames[1].bgr, frames[0].bgr)
    o = orc.DynaDetectOracle(frames[1].bgr, frames[0].bgr, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=False)
    for k in (2, 3, 4):
        mask, label = s.detect(frames[k].bgr, frames[k].depth, k)
        fr = s.flow_results()
        r = o.detect(frames[k].bgr, frames[k].depth, inject_masks=(fr["low"], fr["high"]))
        iou_gt = _iou(mask == 255, cv2.dilate(frames[k].dyn_mask.astype(np.uint8), orc.ellipse(9)) > 0)
        print("frame %d: labels %d, dyn px %d, mismatch mask %d label %d, IoU vs rendered truth %.3f, stage ms %s" % (
            k, int(label.max()), int((mask == 255).sum()), int((mask != r["mask"]).sum()), int((label != r["label"]).sum()), iou_gt,
            np.round(s.stage_ms()[:11], 3)))
        assert np.array_equal(label, r["label"])
        assert np.array_equal(mask, r["mask"])
        assert set(np.unique(mask)) <= {0, 125, 255}
    s.close()
