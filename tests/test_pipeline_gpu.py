"""GPU parity of the fused flow + residual branch (sindyn_flow_residual = DetectDynaByDenseOpticalFLow,
DynaDetect.cc:1023-1374) against the CPU oracle, streamed over a short synthetic sequence."""
import numpy as np
import pytest

from oracle import dynadetect_oracle as orc
from sindslam_b200 import synth

pytestmark = pytest.mark.gpu

FLOW_EPE_TOL = 0.08     # px at 640x480 (= 0.05 px on the 384x288 grid / 0.6), mean EPE GPU branch vs CPU branch
MASK_IOU_MIN = 0.97     # masks from independently solved flows (GPU Brox vs CPU Brox); bit-exactness is checked below with identical flow


def _iou(a, b):
    a, b = a > 0, b > 0
    u = (a | b).sum()
    return 1.0 if u == 0 else float((a & b).sum()) / float(u)


@pytest.mark.parametrize("refine", [0, 1])
def test_flow_residual_stream(seq_c1, refine):
    from sindslam_b200.capi import SinDyn
    scene, frames = seq_c1
    cam = synth.TUM3
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=refine, stage_timing=1)
    sd.set_prev_frames(frames[1].bgr, frames[0].bgr)
    z = np.zeros((cam.height, cam.width), np.uint8)
    n_checked = [0]
    for k in range(2, 5):
        lo, hi = sd.flow_residual(frames[k].bgr, roll=True)
        res = sd.flow_results()
        assert np.array_equal(lo, res["low"]) and np.array_equal(hi, res["high"])
        ref = orc.flow_residual_cpu(frames[k].bgr, frames[k - 1].bgr, frames[k - 2].bgr, z, z, "brox", refine=bool(refine))
        assert res["large_motion"] == ref["large_motion"]
        epe = float(np.sqrt(((res["flow"] - ref["flow"]) ** 2).sum(-1)).mean())
        print("frame %d refine %d: flow EPE %.4f, IoU low %.4f high %.4f, thr gpu %s cpu %s, stage ms %s" % (
            k, refine, epe, _iou(lo, ref["low"]), _iou(hi, ref["high"]), res["thr"], ref["thr"], np.round(sd.stage_ms()[:5], 3)))
        assert epe <= FLOW_EPE_TOL
        # k_rho restates cv::findHomography(RHO) (DynaDetect.cc:1235): on the device's own flow and state the device H must be
        # the library's H bit for bit (the oracle's sample list is bit-exact with the device's, tests/test_homography_gpu.py)
        p, q = orc.sample_pairs(res["flow"], z, z)      # the flow-only entry point leaves imgDynaLast / imgLabelLast at zero
        assert np.array_equal(res["H"], orc.estimate_homography(p, q)), k
        # masks from independently solved flows (GPU Brox vs CPU Brox, EPE ~0.003 px): a stated IoU floor on every frame
        assert _iou(lo, ref["low"]) >= MASK_IOU_MIN and _iou(hi, ref["high"]) >= MASK_IOU_MIN
        n_checked[0] += 1
        # identical flow + identical H -> residual / thresholds / masks bit-exact
        mag = orc.homography_residual(res["flow"], res["H"])
        olo, ohi, othr, _ = orc.threshold_masks(mag)
        assert np.array_equal(othr, res["thr"])
        assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)
    assert n_checked[0] == 3
    sd.close()


def test_resident_equals_host_path(seq_c1):
    from sindslam_b200.capi import SinDyn
    _, frames = seq_c1
    cam = synth.TUM3
    a = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=0)
    b = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=0)
    for s in (a, b):
        s.set_prev_frames(frames[1].bgr, frames[0].bgr)
    for i, f in enumerate(frames):
        b.upload_frame(i, f.bgr, f.depth)
    for k in range(2, 5):
        lo, hi = a.flow_residual(frames[k].bgr, roll=True)
        b.flow_residual_resident(k, roll=True)
        r = b.flow_results()
        assert np.array_equal(lo, r["low"]) and np.array_equal(hi, r["high"])
    p = b.brox_profile()
    # k_brox_sor runs the 9 finest of the 15 levels (> 2100 px), two launches of 5 sweeps per inner iteration; the 6 coarsest
    # levels run all inner iterations in one k_brox_level launch
    assert p["sor_launches"] == 180 and p["pixel_sweeps"] == 100 * 302822 and p["sor_ms"] > 0
    a.close()
    b.close()


@pytest.mark.parametrize("entry", ["detect", "flow_residual"])
def test_one_graph_flow_branch_equals_classic_path(entry):
    """The default path (use_graphs=1, stage_timing=0) runs the whole flow branch as ONE CUDA graph: the large-motion decision
    (DynaDetect.cc:1097-1114) is a conditional IF node whose body is the second Brox solve, and the refinement picks its
    reference image from the device flag.  It must be bit-identical, frame by frame, to the classic path (use_graphs=0:
    host decision after a D2H copy, like the reference, DynaDetect.cc:1073) on a sequence WITH a large-motion frame jump
    (7 -> 23 -> 9) and refine=1: flow, H, thresholds, both masks, the large-motion flag -- and for sindyn_detect the final
    mask and labels."""
    from sindslam_b200.capi import SinDyn
    cam = synth.TUM3
    _, frames = synth.make_sequence(24, cam, seq=1, kind="box", start=4)
    order = list(range(0, 8)) + [23] + list(range(9, 12))
    mk = lambda **kw: SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, refine=1, plane_edges=0, **kw)
    a, b = mk(), mk(use_graphs=0)
    for s in (a, b):
        s.set_prev_frames(frames[order[0]].bgr, frames[order[0]].bgr)
    n_large = n_graph = 0
    for n, k in enumerate(order[1:], 1):
        f = frames[k]
        if entry == "detect":
            ra, rb = a.detect(f.bgr, f.depth, n), b.detect(f.bgr, f.depth, n)
        else:
            ra, rb = a.flow_residual(f.bgr, roll=True), b.flow_residual(f.bgr, roll=True)
        pa, pb = a.path_info(), b.path_info()
        fa, fb = a.flow_results(), b.flow_results()
        assert not pb["flow_one_graph"] and not pa["flow_graph_broken"]
        n_graph += int(pa["flow_one_graph"])
        assert fa["large_motion"] == fb["large_motion"], n
        n_large += int(fa["large_motion"])
        for key in ("flow", "H", "thr", "low", "high"):
            assert np.array_equal(fa[key], fb[key]), (n, key)
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]), n
    assert n_large >= 1, "the frame jump must trigger the large-motion fallback (second Brox solve inside the IF node)"
    assert n_graph == len(order) - 1, "the default handle must take the one-graph path on every frame"
    a.close()
    b.close()
