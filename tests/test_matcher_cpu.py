"""Oracle of row f4 (oracle/matcher_oracle.py): known answers without a GPU."""
import numpy as np

from oracle import matcher_oracle as mo


def test_descriptor_distance_matches_bit_count():
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
    assert mo.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())
    assert mo.descriptor_distance(a, a) == 0 and mo.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_three_maxima_reference_cases():
    assert mo.three_maxima([0] * 30) == (-1, -1, -1)
    s = [0] * 30; s[4] = 50; s[7] = 20; s[9] = 10
    assert mo.three_maxima(s) == (4, 7, 9)
    s[9] = 4                                  # third maximum below 10 % of the first
    assert mo.three_maxima(s) == (4, 7, -1)
    s[7] = 4
    assert mo.three_maxima(s) == (4, -1, -1)
    assert mo.three_maxima([3, 3, 3, 3]) == (0, 1, 2)    # ties keep the earlier bins


def _one_cell_frame(pts, octave, angle, desc):
    n = len(pts)
    off = np.zeros(64 * 48 + 1, np.int32)
    cells = (np.round(pts[:, 0] * 64 / 640).astype(int)) * 48 + np.round(pts[:, 1] * 48 / 480).astype(int)
    order = np.argsort(cells, kind="stable")
    cnt = np.bincount(cells, minlength=64 * 48)
    off[1:] = np.cumsum(cnt)
    return dict(keys_un=pts.astype(np.float32), octave=np.asarray(octave, np.int32), angle=np.asarray(angle, np.float32),
                u_right=np.full(n, -1, np.float32), desc=desc, bounds=np.array([0, 640, 0, 480], np.float32), grid_offsets=off,
                grid_indices=order.astype(np.int32))


def test_search_order_dependence_and_rotation_filter():
    # two map points project onto the same pair of key points; both prefer key point 0
    d0 = np.zeros(32, np.uint8); d1 = np.zeros(32, np.uint8); d1[0] = 0b111
    cur = _one_cell_frame(np.array([[320.0, 240.0], [322.0, 240.0]]), [0, 0], [10.0, 10.0], np.stack([d0, d1]))
    K = dict(fx=500.0, fy=500.0, cx=320.0, cy=240.0, bf=40.0, b=0.08)
    P = np.array([[0.0, 0.0, 2.0], [0.004, 0.0, 2.0]], np.float32)      # project to (320, 240) and (321, 240)
    last = dict(xyz_w=P, valid=np.ones(2, bool), desc=np.stack([d0, d0]), octave=np.zeros(2, np.int32), angle=np.array([10.0, 10.0], np.float32),
                observed=np.array([False, False]))
    scale = [1.0, 1.2]
    m, n = mo.search_by_projection(cur, last, np.eye(4), np.eye(4), scale_factors=scale, th=15.0, mono=True, **K)
    assert list(m) == [1, -1] and n == 2          # the second point overwrites the first assignment; both are counted
    last["observed"] = np.array([True, False])
    m, n = mo.search_by_projection(cur, last, np.eye(4), np.eye(4), scale_factors=scale, th=15.0, mono=True, **K)
    assert list(m) == [0, 1] and n == 2           # key point 0 is blocked for the second point, which takes its next best
    # rotation consistency: a lone match with a different rotation is dropped when the dominant bin is >10x larger -- here 1 vs 1,
    # so both survive; with check off nothing is filtered either
    last["angle"] = np.array([10.0, 100.0], np.float32)
    m2, n2 = mo.search_by_projection(cur, last, np.eye(4), np.eye(4), scale_factors=scale, th=15.0, mono=True, **K)
    assert n2 == 2
    # descriptor too far (> TH_HIGH = 100 bits): no match
    far = np.full(32, 255, np.uint8)
    last["desc"] = np.stack([far, far])
    m3, n3 = mo.search_by_projection(cur, last, np.eye(4), np.eye(4), scale_factors=scale, th=15.0, mono=True, **K)
    assert n3 == 0 and (m3 == -1).all()
