"""Deterministic synthetic TUM-format RGB-D sequences.

TUM / Bonn data are not available offline, so parity and throughput are measured on
rendered sequences: a textured room (floor, ceiling, three walls, two static boxes),
one moving object, a walking_xyz-shaped camera trajectory, u16 depth with noise,
holes and shadow bands.  Every frame also carries the analytic ground-truth flow and
the ground-truth dynamic mask.

The on-disk layout written by :func:`write_tum_sequence` is the one
`rgbd_tum_noros.cc:217-242` (LoadImages) reads: ``associations.txt`` lines
``t rgb/t.png t depth/t.png`` plus ``groundtruth.txt`` ``t tx ty tz qx qy qz qw``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

BASE_SEED = 20241108
T0 = 1341846313.553992  # same epoch as EVO/CameraTrajectory.txt:1


@dataclass
class CameraConfig:
    width: int = 640
    height: int = 480
    fx: float = 535.4
    fy: float = 539.2
    cx: float = 320.1
    cy: float = 247.6
    depth_factor: float = 5000.0
    name: str = "TUM3"


TUM3 = CameraConfig()
D455_848 = CameraConfig(848, 480, 425.0, 425.0, 424.0, 240.0, 1000.0, "D455-848")


@dataclass
class Frame:
    bgr: np.ndarray          # H x W x 3 u8
    depth: np.ndarray        # H x W u16 (raw units, depth_factor per metre)
    timestamp: float
    T_wc: np.ndarray         # 4x4 camera-to-world
    dyn_mask: np.ndarray     # H x W bool, ground-truth moving-object pixels
    points_w: np.ndarray = field(repr=False, default=None)  # H x W x 3 world hit points (noise free)
    obj_id: np.ndarray = field(repr=False, default=None)


# ----------------------------------------------------------------------------- texture
def _hash3(ix, iy, iz, seed):
    h = (ix.astype(np.uint32) * np.uint32(374761393)
         + iy.astype(np.uint32) * np.uint32(668265263)
         + iz.astype(np.uint32) * np.uint32(2147483647)
         + np.uint32((int(seed) * 1274126177) & 0xFFFFFFFF))
    h = (h ^ (h >> np.uint32(13))) * np.uint32(1274126177)
    h = h ^ (h >> np.uint32(16))
    return (h & np.uint32(0xFFFF)).astype(np.float32) * np.float32(1.0 / 65535.0)


def _value_noise3(p, seed):
    """Trilinear value noise on the integer lattice of p (N x 3 float)."""
    f = np.floor(p)
    t = (p - f).astype(np.float32)
    t = t * t * (3.0 - 2.0 * t)
    i = f.astype(np.int64)
    out = np.zeros(p.shape[0], np.float32)
    for dx in (0, 1):
        wx = t[:, 0] if dx else 1.0 - t[:, 0]
        for dy in (0, 1):
            wy = t[:, 1] if dy else 1.0 - t[:, 1]
            for dz in (0, 1):
                wz = t[:, 2] if dz else 1.0 - t[:, 2]
                out += wx * wy * wz * _hash3(i[:, 0] + dx, i[:, 1] + dy, i[:, 2] + dz, seed)
    return out


def _albedo(p_obj, seed):
    """Multi-octave value noise in object coordinates -> u8-range albedo 30..225."""
    acc = np.zeros(p_obj.shape[0], np.float32)
    amp, freq, tot = 1.0, 4.0, 0.0
    for o in range(5):
        acc += amp * _value_noise3(p_obj * freq + 17.0 * o, seed + o)
        tot += amp
        amp *= 0.62
        freq *= 2.3
    acc /= tot
    acc = np.clip((acc - 0.5) * 2.6 + 0.5, 0.0, 1.0)
    return 30.0 + 195.0 * acc


# ----------------------------------------------------------------------------- scene
@dataclass
class Box:
    center: np.ndarray
    half: np.ndarray
    tex_seed: int
    dynamic: bool = False
    tint: tuple = (1.0, 1.0, 1.0)


class Scene:
    """Room [-2.5,2.5] x [-1.3,1.2] x [-1.0,4.0] (x right, y down, z forward) seen from inside."""

    def __init__(self, seed: int, kind: str = "box", fps: float = 30.0):
        self.seed = int(seed)
        self.kind = kind
        self.fps = fps
        self.room_min = np.array([-2.5, -1.3, -1.0])
        self.room_max = np.array([2.5, 1.2, 4.0])
        self.static_boxes = [
            Box(np.array([-1.4, 0.8, 2.8]), np.array([0.35, 0.4, 0.3]), seed + 101, tint=(1.0, 0.9, 0.8)),
            Box(np.array([1.5, 0.7, 3.1]), np.array([0.3, 0.5, 0.35]), seed + 202, tint=(0.8, 1.0, 0.9)),
        ]

    # -- dynamic object(s) at time t: list of Boxes (humanoid = several boxes)
    def dynamic_boxes(self, t: float):
        if self.kind == "box":
            x = -0.9 + 0.6 * t
            # bounce between -1.1 and 1.1
            span = 2.2
            xx = (x + 1.1) % (2 * span)
            x = -1.1 + (xx if xx < span else 2 * span - xx)
            return [Box(np.array([x, 0.35, 2.0]), np.array([0.25, 0.85, 0.15]), self.seed + 303, True, (1.0, 0.8, 0.7))]
        # humanoid: torso, head, two legs, two arms (boxes standing in for capsules)
        x = -0.5 + 0.35 * t
        span = 1.4
        xx = (x + 0.7) % (2 * span)
        x = -0.7 + (xx if xx < span else 2 * span - xx)
        z = 1.5
        sw = 0.18 * np.sin(2 * np.pi * 1.1 * t)
        s = self.seed
        return [
            Box(np.array([x, 0.05, z]), np.array([0.23, 0.35, 0.12]), s + 303, True, (1.0, 0.8, 0.7)),
            Box(np.array([x, -0.45, z]), np.array([0.12, 0.13, 0.12]), s + 304, True, (1.0, 0.85, 0.75)),
            Box(np.array([x - 0.11, 0.8, z + sw]), np.array([0.09, 0.4, 0.09]), s + 305, True, (0.7, 0.7, 1.0)),
            Box(np.array([x + 0.11, 0.8, z - sw]), np.array([0.09, 0.4, 0.09]), s + 306, True, (0.7, 0.7, 1.0)),
            Box(np.array([x - 0.31, 0.05, z - sw]), np.array([0.06, 0.33, 0.06]), s + 307, True, (1.0, 0.8, 0.7)),
            Box(np.array([x + 0.31, 0.05, z + sw]), np.array([0.06, 0.33, 0.06]), s + 308, True, (1.0, 0.8, 0.7)),
        ]

    def camera_pose(self, t: float) -> np.ndarray:
        """walking_xyz-shaped: small sinusoidal translation, <=2 deg yaw/pitch."""
        tr = np.array([0.15 * np.sin(2 * np.pi * 0.5 * t),
                       0.05 * np.sin(2 * np.pi * 0.8 * t),
                       0.10 * np.sin(2 * np.pi * 0.3 * t)])
        yaw = np.deg2rad(2.0) * np.sin(2 * np.pi * 0.4 * t)
        pitch = np.deg2rad(1.5) * np.sin(2 * np.pi * 0.25 * t + 0.7)
        cy_, sy_ = np.cos(yaw), np.sin(yaw)
        cp, sp = np.cos(pitch), np.sin(pitch)
        Ry = np.array([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]])
        Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        T = np.eye(4)
        T[:3, :3] = Ry @ Rx
        T[:3, 3] = tr
        return T


def _ray_box(o, d, bmin, bmax, inside=False):
    """Slab intersection. Returns hit distance (inf if none). inside=True -> exit distance."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t1 = (bmin - o) * inv
        t2 = (bmax - o) * inv
    tmin = np.minimum(t1, t2).max(axis=-1)
    tmax = np.maximum(t1, t2).min(axis=-1)
    if inside:
        return tmax
    hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 1e-6)
    return np.where(hit, tmin, np.inf)


def render_frame(scene: Scene, cam: CameraConfig, idx: int, rng_seed: int | None = None, noise: bool = True) -> Frame:
    t = idx / scene.fps
    H, W = cam.height, cam.width
    T_wc = scene.camera_pose(t)
    R, o = T_wc[:3, :3], T_wc[:3, 3]
    uu, vv = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    dirs_c = np.stack([(uu - cam.cx) / cam.fx, (vv - cam.cy) / cam.fy, np.ones_like(uu)], -1).reshape(-1, 3)
    d = dirs_c @ R.T
    N = d.shape[0]
    best = _ray_box(o, d, scene.room_min, scene.room_max, inside=True)
    obj = np.zeros(N, np.int32)  # 0 = room
    boxes = scene.static_boxes + scene.dynamic_boxes(t)
    for bi, b in enumerate(boxes):
        tb = _ray_box(o, d, b.center - b.half, b.center + b.half)
        closer = tb < best
        best = np.where(closer, tb, best)
        obj = np.where(closer, bi + 1, obj)
    P = o + d * best[:, None]
    # albedo in object coordinates
    gray = np.zeros(N, np.float32)
    tint = np.ones((N, 3), np.float32)
    m = obj == 0
    gray[m] = _albedo(P[m] * 0.55, scene.seed)
    for bi, b in enumerate(boxes):
        m = obj == bi + 1
        if m.any():
            gray[m] = _albedo((P[m] - b.center) * 1.4, b.tex_seed)
            tint[m] = np.array(b.tint, np.float32)
    # simple shading by dominant face normal so that planes differ in brightness
    bgr = gray[:, None] * tint
    zc = best * 1.0  # depth along camera z: dirs_c has z = 1, so z_c = t
    dyn = np.zeros(N, bool)
    nstat = len(scene.static_boxes)
    dyn[obj > nstat] = True
    rng = np.random.default_rng(BASE_SEED * 7919 + (rng_seed if rng_seed is not None else scene.seed) * 100003 + idx)
    z = zc.reshape(H, W).copy()
    if noise:
        bgr = bgr + rng.normal(0.0, 1.5, bgr.shape).astype(np.float32)
        z = z + rng.normal(0.0, 1.0, z.shape) * (0.0012 * z * z)
        holes = rng.random(z.shape) < getattr(scene, "hole_rate", 0.015)
        # 2-px zero band on the left (shadow) side of depth discontinuities
        jump = np.zeros_like(holes)
        dz = z[:, 1:] - z[:, :-1]
        edge = np.abs(dz) > 0.25
        jump[:, 1:] |= edge
        jump[:, :-1] |= edge & (dz < 0)
        band = jump.copy()
        band[:, :-1] |= jump[:, 1:]
        far_side = np.zeros_like(z, bool)
        far_side[:, :-1] = z[:, :-1] > z[:, 1:]
        far_side[:, 1:] |= z[:, 1:] > z[:, :-1]
        z[holes | (band & far_side)] = 0.0
    raw = np.clip(np.rint(z * cam.depth_factor), 0, 65535).astype(np.uint16)
    img = np.clip(np.rint(bgr), 0, 255).astype(np.uint8).reshape(H, W, 3)
    return Frame(img, raw, T0 + t, T_wc, dyn.reshape(H, W), P.reshape(H, W, 3), obj.reshape(H, W))


def gt_flow(scene: Scene, cam: CameraConfig, idx_cur: int, idx_old: int, cur: Frame) -> np.ndarray:
    """Analytic displacement w with I_cur(x) ~ I_old(x + w(x)) (the raw sign of calc(cur, old),
    DynaDetect.cc:1072): where each surface point of the current frame was in the older frame."""
    t_cur, t_old = idx_cur / scene.fps, idx_old / scene.fps
    H, W = cam.height, cam.width
    P = cur.points_w.reshape(-1, 3).copy()
    obj = cur.obj_id.reshape(-1)
    nstat = len(scene.static_boxes)
    b_cur, b_old = scene.dynamic_boxes(t_cur), scene.dynamic_boxes(t_old)
    for k in range(len(b_cur)):
        m = obj == nstat + 1 + k
        P[m] += b_old[k].center - b_cur[k].center
    T_old = scene.camera_pose(t_old)
    Pc = (P - T_old[:3, 3]) @ T_old[:3, :3]
    u = cam.fx * Pc[:, 0] / Pc[:, 2] + cam.cx
    v = cam.fy * Pc[:, 1] / Pc[:, 2] + cam.cy
    uu, vv = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    return np.stack([u.reshape(H, W) - uu, v.reshape(H, W) - vv], -1).astype(np.float32)


def make_sequence(n_frames: int, cam: CameraConfig = TUM3, seq: int = 0, kind: str = "box", start: int = 0, hole_rate: float = 0.015):
    """hole_rate: probability of an isolated zero-depth pixel (1.5 %).  The PEAC plane fitter rejects every
    16x16 block that contains a hole (INIT_STRICT), so its tests use a sensor-like low rate of isolated holes."""
    scene = Scene(BASE_SEED + seq, kind)
    scene.hole_rate = hole_rate
    return scene, [render_frame(scene, cam, start + i) for i in range(n_frames)]


def _quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = np.copysign(np.sqrt(max(0.0, 1 + R[0, 0] - R[1, 1] - R[2, 2])) / 2, R[2, 1] - R[1, 2])
    y = np.copysign(np.sqrt(max(0.0, 1 - R[0, 0] + R[1, 1] - R[2, 2])) / 2, R[0, 2] - R[2, 0])
    zq = np.copysign(np.sqrt(max(0.0, 1 - R[0, 0] - R[1, 1] + R[2, 2])) / 2, R[1, 0] - R[0, 1])
    return x, y, zq, w


def write_tum_sequence(root: str, frames, cam: CameraConfig, raw: bool = False):
    """Write rgb/, depth/, associations.txt, groundtruth.txt (TUM format).  raw=True additionally writes every image
    as binary PPM (B,G,R byte order, what cv::imread delivers) / 16-bit big-endian PGM for the OpenCV-free C++ driver."""
    import cv2
    os.makedirs(os.path.join(root, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(root, "depth"), exist_ok=True)
    with open(os.path.join(root, "associations.txt"), "w") as fa, open(os.path.join(root, "groundtruth.txt"), "w") as fg:
        fg.write("# ground truth trajectory\n# synthetic (sindslam_b200.synth)\n# timestamp tx ty tz qx qy qz qw\n")
        for f in frames:
            ts = "%.6f" % f.timestamp
            cv2.imwrite(os.path.join(root, "rgb", ts + ".png"), f.bgr)
            cv2.imwrite(os.path.join(root, "depth", ts + ".png"), f.depth)
            if raw:
                h, w = f.depth.shape
                with open(os.path.join(root, "rgb", ts + ".ppm"), "wb") as fp:
                    fp.write(b"P6\n%d %d\n255\n" % (w, h) + np.ascontiguousarray(f.bgr).tobytes())
                with open(os.path.join(root, "depth", ts + ".pgm"), "wb") as fp:
                    fp.write(b"P5\n%d %d\n65535\n" % (w, h) + f.depth.astype(">u2").tobytes())
            fa.write(f"{ts} rgb/{ts}.png {ts} depth/{ts}.png\n")
            q = _quat(f.T_wc[:3, :3])
            p = f.T_wc[:3, 3]
            fg.write("%s %.6f %.6f %.6f %.6f %.6f %.6f %.6f\n" % (ts, p[0], p[1], p[2], *q))


def load_tum_sequence(root: str):
    """Mirror of LoadImages (rgbd_tum_noros.cc:217-242): returns (rgb files, depth files, timestamps)."""
    rgb, dep, ts = [], [], []
    with open(os.path.join(root, "associations.txt")) as f:
        for line in f:
            s = line.split()
            if len(s) >= 4:
                ts.append(float(s[0])); rgb.append(os.path.join(root, s[1])); dep.append(os.path.join(root, s[3]))
    return rgb, dep, ts



def make_sequence_parallel(n_frames: int, cam: CameraConfig = TUM3, seq: int = 0, kind: str = "box", start: int = 0, hole_rate: float = 0.015,
                           workers: int | None = None):
    """Same frames as :func:`make_sequence` (every frame is rendered independently from (seed, index)), rendered by worker
    PROCESSES -- long sequences (BASELINE.json configs[2]: 300 frames) take ~0.9 s per frame on one core.  The workers are
    plain `python -m sindslam_b200.synth` subprocesses writing .npz chunks: no fork of a process that may hold a CUDA context.
    Frames come back without points_w / obj_id."""
    import subprocess
    import sys
    import tempfile
    workers = max(1, min(workers or (os.cpu_count() or 1), 32, n_frames // 2))
    scene = Scene(BASE_SEED + seq, kind)
    scene.hole_rate = hole_rate
    if workers <= 1:
        return scene, [render_frame(scene, cam, start + i) for i in range(n_frames)]
    bounds = np.linspace(0, n_frames, workers + 1).astype(int)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for w in range(workers):
            a, b = int(bounds[w]), int(bounds[w + 1])
            out = os.path.join(tmp, "chunk%03d.npz" % w)
            cmd = [sys.executable, "-m", "sindslam_b200.synth", "--render", str(seq), kind, repr(hole_rate), cam.name, str(start + a), str(b - a), out]
            procs.append((subprocess.Popen(cmd, cwd=root, env=dict(os.environ, OMP_NUM_THREADS="1")), out))
        frames = []
        for p, out in procs:
            if p.wait(timeout=1800) != 0:
                raise RuntimeError("synth worker failed")
            z = np.load(out)
            for i in range(len(z["timestamp"])):
                frames.append(Frame(z["bgr"][i], z["depth"][i], float(z["timestamp"][i]), z["T_wc"][i], z["dyn"][i]))
    return scene, frames


def _worker_main(argv):
    seq, kind, hole, cam_name, a, n, out = int(argv[0]), argv[1], float(argv[2]), argv[3], int(argv[4]), int(argv[5]), argv[6]
    cam = {c.name: c for c in (TUM3, D455_848)}[cam_name]
    scene = Scene(BASE_SEED + seq, kind)
    scene.hole_rate = hole
    fr = [render_frame(scene, cam, a + i) for i in range(n)]
    np.savez(out, bgr=np.stack([f.bgr for f in fr]), depth=np.stack([f.depth for f in fr]), dyn=np.stack([f.dyn_mask for f in fr]),
             T_wc=np.stack([f.T_wc for f in fr]), timestamp=np.array([f.timestamp for f in fr], np.float64))


if __name__ == "__main__":
    import sys
    if len(sys.argv) >= 9 and sys.argv[1] == "--render":
        _worker_main(sys.argv[2:])
