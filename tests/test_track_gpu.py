"""sindyn_track_frame (one driver iteration: rgbd_tum_noros.cc:132-139 + Tracking::GrabImageRGBD's colour conversion +
ORBextractor::operator() on the dilated mask) must equal the three reference-shaped calls it fuses -- sindyn_detect,
sindyn_morph_ellipse(15, dilate), sindyn_orb_extract on cv2's RGB2GRAY / BGR2GRAY image -- bit for bit, for host buffers
(pageable and pinned) and for device-resident frames."""
import cv2
import numpy as np
import pytest

from sindslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rgb_order", [1, 0])
def test_track_frame_equals_separate_calls(rgb_order):
    from sindslam_b200.capi import Orb, SinDyn
    cam = synth.TUM3
    _, frames = synth.make_sequence(6, cam, seq=7, kind="box", start=5, hole_rate=0.0005)
    mk = lambda: (SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1), Orb(1500, 1.2, 8, 15, 5, cam.width, cam.height))
    (sa, oa), (sb, ob), (sc, oc) = mk(), mk(), mk()
    for s in (sa, sb, sc):
        s.set_prev_frames(frames[0].bgr, frames[0].bgr)
    for i, f in enumerate(frames):
        sc.upload_frame(i, f.bgr, f.depth)
    code = cv2.COLOR_RGB2GRAY if rgb_order else cv2.COLOR_BGR2GRAY
    for k in range(1, 6):
        f = frames[k]
        mask, label = sa.detect(f.bgr, f.depth, k)
        dil = sa.morph_ellipse(mask, 15, 0)
        kps, desc = oa.extract(cv2.cvtColor(f.bgr, code), dil)
        m2, l2, k2, d2 = ob.track_frame(sb, f.bgr, f.depth, k, rgb_order=rgb_order, dilate_k=15)
        assert np.array_equal(m2, dil) and np.array_equal(l2, label), k
        assert len(k2) == len(kps) and np.array_equal(d2, desc), k
        for name in ("x", "y", "angle", "response", "octave"):
            assert np.array_equal(k2[name], kps[name]), (k, name)
        oc.track_frame_resident(sc, k, k, rgb_order=rgb_order, dilate_k=15)
        m3, l3, k3, d3 = oc.track_results(sc)
        assert np.array_equal(m3, dil) and np.array_equal(l3, label) and np.array_equal(d3, desc), k
        assert np.array_equal(k3["x"], kps["x"]) and np.array_equal(k3["octave"], kps["octave"]), k
    for h in (oa, ob, oc, sa, sb, sc):
        h.close()


def test_pipelined_frames_equal_frame_at_a_time():
    """The frame pipeline (pipe.cu): frames enqueued back to back without a synchronisation (device-resident slots), and frames
    submitted two ahead through sindyn_track_submit / sindyn_track_collect (pinned and pageable host buffers), must return exactly
    what the frame-at-a-time entry returns -- 24 consecutive frames with free-running state, plane edges on."""
    import torch
    from sindslam_b200.capi import Orb, SinDyn
    cam = synth.TUM3
    n = 25
    _, frames = synth.make_sequence(n, cam, seq=11, kind="box", start=3, hole_rate=0.0005)
    mk = lambda: (SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, plane_edges=1), Orb(1500, 1.2, 8, 15, 5, cam.width, cam.height))
    (sa, oa), (sb, ob), (sc, oc) = mk(), mk(), mk()
    for s in (sa, sb, sc):
        s.set_prev_frames(frames[0].bgr, frames[0].bgr)
    ref = [oa.track_frame(sa, frames[k].bgr, frames[k].depth, k) for k in range(1, n)]
    ref = [(m.copy(), l.copy(), kp.copy(), d.copy()) for m, l, kp, d in ref]

    def same(got, want, k):
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), k
        assert len(got[2]) == len(want[2]) and np.array_equal(got[3], want[3]), k
        for name in ("x", "y", "angle", "response", "octave"):
            assert np.array_equal(got[2][name], want[2][name]), (k, name)

    # (a) submit one frame ahead: odd frames from pinned buffers, even frames from pageable ones
    pinned = [(torch.from_numpy(frames[k].bgr.copy()).pin_memory().numpy(), torch.from_numpy(frames[k].depth.view(np.int16).copy()).pin_memory().numpy().view(np.uint16))
              if k & 1 else (frames[k].bgr, frames[k].depth) for k in range(n)]
    ob.track_submit(sb, pinned[1][0], pinned[1][1], 1)
    ob.track_submit(sb, pinned[2][0], pinned[2][1], 2)
    for k in range(3, n):
        ob.track_submit(sb, pinned[k][0], pinned[k][1], k)      # three frames in flight
        if k == 3:
            with pytest.raises(Exception):
                ob.track_submit(sb, pinned[4][0], pinned[4][1], 4)   # a fourth is refused
        same(ob.track_collect(sb), ref[k - 3], k - 2)
    same(ob.track_collect(sb), ref[n - 3], n - 2)
    same(ob.track_collect(sb), ref[n - 2], n - 1)
    with pytest.raises(Exception):
        ob.track_collect(sb)               # nothing in flight
    # the frame-at-a-time entry continues the same state afterwards
    extra = synth.make_sequence(n + 1, cam, seq=11, kind="box", start=3, hole_rate=0.0005)[1][n]
    want = oa.track_frame(sa, extra.bgr, extra.depth, n)
    same(ob.track_frame(sb, extra.bgr, extra.depth, n), want, n)
    # (b) resident slots, all frames enqueued without waiting; only the last frame's results are visible afterwards
    for i, f in enumerate(frames):
        sc.upload_frame(i, f.bgr, f.depth)
    for k in range(1, n):
        oc.track_frame_resident(sc, k, k)
    oc.track_join(sc)
    same(oc.track_results(sc), ref[n - 2], n - 1)
    # (c) state injected from outside (sindyn_set_state) between pipelined frames is picked up: labels (k-means warm start,
    # sample weights) and dynamic mask (sample weights).  sa has seen frame n in (a): sc catches up first.
    more = synth.make_sequence(n + 4, cam, seq=11, kind="box", start=3, hole_rate=0.0005)[1]
    sc.upload_frame(0, more[n].bgr, more[n].depth)
    oc.track_frame_resident(sc, 0, n)
    same(oc.track_results(sc), want, n)
    rng = np.random.default_rng(4)
    lab = rng.integers(0, 6, (cam.height // 40, cam.width // 40)).repeat(40, 0).repeat(40, 1).astype(np.uint8)
    dyn = np.full((cam.height, cam.width), 125, np.uint8)
    dyn[100:220, 300:420] = 255
    for s in (sa, sc):
        s.set_state(2, lab)
        s.set_state(0, dyn)
    for k in range(n + 1, n + 4):
        sc.upload_frame(k % 8, more[k].bgr, more[k].depth)
        oc.track_frame_resident(sc, k % 8, k)
        same(oc.track_results(sc), oa.track_frame(sa, more[k].bgr, more[k].depth, k), k)
    for h in (oa, ob, oc, sa, sb, sc):
        h.close()
