"""Row f4: ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1328-1470) on the
device against oracle/matcher_oracle.py; the current frame is the one resident in the extractor handle."""
import cv2
import numpy as np
import pytest

from sindslam_b200 import synth
from oracle import matcher_oracle as mo

pytestmark = pytest.mark.gpu

BF, B = 40.0, 40.0 / 535.4
DMF = 1.0 / 5000.0


@pytest.fixture(scope="module")
def scene():
    """Two frames of a synthetic sequence: `last` as arrays (map points from its depth), `cur` resident in the handle."""
    from sindslam_b200.capi import Orb
    cam = synth.TUM3
    _, frames = synth.make_sequence(6, cam, seq=1, kind="box", start=4)
    orb = Orb(1000, 1.2, 8, 20, 7, 640, 480)
    fl, fc = frames[2], frames[3]
    kps, desc = orb.extract(cv2.cvtColor(fl.bgr, cv2.COLOR_BGR2GRAY), None)
    un, dep, ur, b, off, idx = orb.frame_features(fl.depth, cam.fx, cam.fy, cam.cx, cam.cy, (0, 0, 0, 0, 0), BF, DMF)
    n = len(kps)
    # Frame::UnprojectStereo (Frame.cc:686-700) for every key point with depth; the map point descriptor is the key point's
    z = dep[:n]
    x = (un[:n, 0] - np.float32(cam.cx)) * z / np.float32(cam.fx)
    y = (un[:n, 1] - np.float32(cam.cy)) * z / np.float32(cam.fy)
    pc = np.stack([x, y, z, np.ones(n, np.float32)], 1).astype(np.float64)
    xyz_w = (fl.T_wc @ pc.T).T[:, :3].astype(np.float32)
    rng = np.random.default_rng(9)
    last = dict(xyz_w=xyz_w, valid=(z > 0), desc=desc.copy(), octave=kps["octave"].astype(np.int32), angle=kps["angle"].astype(np.float32),
                observed=rng.random(n) < 0.5)
    kc, dc = orb.extract(cv2.cvtColor(fc.bgr, cv2.COLOR_BGR2GRAY), None)
    un2, dep2, ur2, b2, off2, idx2 = orb.frame_features(fc.depth, cam.fx, cam.fy, cam.cx, cam.cy, (0, 0, 0, 0, 0), BF, DMF)
    n2 = len(kc)
    cur = dict(keys_un=un2[:n2], octave=kc["octave"].astype(np.int32), angle=kc["angle"].astype(np.float32), u_right=ur2[:n2], desc=dc.copy(),
               bounds=b2, grid_offsets=off2, grid_indices=idx2)
    scale = [np.float32(1.0)]
    for _ in range(7):
        scale.append(np.float32(scale[-1] * np.float32(1.2)))      # ORBextractor.cc:419-424
    yield dict(orb=orb, cam=cam, last=last, cur=cur, Tcw_last=np.linalg.inv(fl.T_wc), Tcw_cur=np.linalg.inv(fc.T_wc), scale=scale)
    orb.close()


def _run(sc, Tcw_cur, th=15.0, mono=False, check=True, blocked=None, last=None):
    cam, last = sc["cam"], (last or sc["last"])
    got, ng = sc["orb"].search_by_projection(last, Tcw_cur, sc["Tcw_last"], cam.fx, cam.fy, cam.cx, cam.cy, BF, B, th, mono, check, blocked)
    ref, nr = mo.search_by_projection(sc["cur"], last, Tcw_cur, sc["Tcw_last"], cam.fx, cam.fy, cam.cx, cam.cy, BF, B, sc["scale"], th, mono, check,
                                      blocked)
    assert np.array_equal(got, ref)
    assert ng == nr
    return got, ng


def test_search_by_projection_true_pose(scene):
    got, n = _run(scene, scene["Tcw_cur"])
    print("matches with the true pose:", n, "of", int(scene["last"]["valid"].sum()), "map points")
    assert n > 200
    # matched pairs are geometrically consistent: the matched key point lies near the projection
    assert (got >= 0).sum() <= n   # overwritten assignments are counted by the reference, too


@pytest.mark.parametrize("case", ["perturbed", "forward", "backward", "mono", "no_orientation", "wide", "blocked", "all_observed", "rotated"])
def test_search_by_projection_variants(scene, case):
    T = scene["Tcw_cur"].copy()
    kw = {}
    if case == "perturbed":
        T[:3, 3] += (0.01, -0.015, 0.02)
    elif case == "forward":
        T[2, 3] -= 0.3          # camera moved forward along the optical axis by more than the baseline
    elif case == "backward":
        T[2, 3] += 0.3
    elif case == "mono":
        kw["mono"] = True
    elif case == "no_orientation":
        kw["check"] = False
    elif case == "wide":
        kw["th"] = 30.0         # the retry of TrackWithMotionModel (2 * th, Tracking.cc:891): long candidate lists
    elif case == "blocked":
        rng = np.random.default_rng(2)
        kw["blocked"] = rng.random(len(scene["cur"]["keys_un"])) < 0.3
    elif case == "all_observed":
        last = dict(scene["last"]); last["observed"] = np.ones(len(last["valid"]), bool)
        kw["last"] = last
    elif case == "rotated":
        a = np.deg2rad(25.0)
        Rz = np.array([[np.cos(a), -np.sin(a), 0, 0], [np.sin(a), np.cos(a), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        T = Rz @ T              # in-plane rotation: projections move, the rotation histogram filter bites
    got, n = _run(scene, T, **kw)
    print(case, "matches", n)
    if case in ("perturbed", "mono", "no_orientation", "all_observed"):
        assert n > 100


def test_search_by_projection_degenerate_inputs(scene):
    cam = scene["cam"]
    last = scene["last"]
    empty = dict(xyz_w=np.zeros((0, 3), np.float32), valid=np.zeros(0, bool), desc=np.zeros((0, 32), np.uint8), octave=np.zeros(0, np.int32),
                 angle=np.zeros(0, np.float32), observed=np.zeros(0, bool))
    got, n = _run(scene, scene["Tcw_cur"], last=empty)
    assert n == 0 and (got == -1).all()
    none_valid = dict(last); none_valid["valid"] = np.zeros(len(last["valid"]), bool)
    got, n = _run(scene, scene["Tcw_cur"], last=none_valid)
    assert n == 0
    behind = dict(last); behind["xyz_w"] = last["xyz_w"].copy(); behind["xyz_w"][::2] *= -1    # half the points behind the camera / off-image
    _run(scene, scene["Tcw_cur"], last=behind)
    from sindslam_b200.capi import SindynError
    with pytest.raises(SindynError, match="CAPACITY"):   # absurd radius: the fixed-size candidate lists overflow, loudly
        scene["orb"].search_by_projection(last, scene["Tcw_cur"], scene["Tcw_last"], cam.fx, cam.fy, cam.cx, cam.cy, BF, B, 400.0)
    bad = dict(last); bad["octave"] = last["octave"].copy(); bad["octave"][np.nonzero(last["valid"])[0][0]] = 9
    with pytest.raises(SindynError):
        scene["orb"].search_by_projection(bad, scene["Tcw_cur"], scene["Tcw_last"], cam.fx, cam.fy, cam.cx, cam.cy, BF, B, 15.0)
