#!/usr/bin/env python
"""bench.py -- dyn-detect frame pairs/s at 640x480 (BASELINE.json metric) on the metric's workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sequences S]

Workload (BASELINE.json configs[2]): the FULL per-frame dynamic-region pipeline of the driver loop
(rgbd_tum_noros.cc:113-192) over a 300-frame synthetic walking_xyz-shaped 640x480 TUM-format sequence:
DynaDetect::DetectDynaArea (DynaDetect.cc:1377-1666: gray/resize, Brox dense flow with the reference's parameters,
large-motion test, variational refinement, up-sampling, sample weighting, robust homography, residual, Otsu/Triangle
masks || 4-level k-means, gradient depth edges, PEAC plane edges, plane-edge filter, split / RAG / merge re-clustering ||
per-cluster decision, final mask, state roll), the driver's 15x15 dilation (rgbd_tum_noros.cc:136-139) and the masked
ORBextractor::operator() (ORBextractor.cc:1043-1164; 1500 features, 8 levels, TUM3.yaml).  Depth holes are isolated
pixels at a 0.05 % rate, so the PEAC plane fitter has planes to fit (it rejects every 16x16 block containing a hole)
and is busy on every frame.

One STEP = FRAMES_PER_STEP (15) consecutive frame pairs of the sequence; the sequence is played forward and then
backward (a camera walking back the same path), so K = 20 steps cover the 300-frame sequence once and no artificial
frame jump is introduced.  value = pairs/s = frames / time.

value    : frames resident in HBM (device slots, 300 x 1.5 MB = 460 MB per sequence, larger than the 126 MB L2 and each
           read once per pass; a 256 MiB flush is written between steps as well), one call sindyn_track_frame_resident per
           frame (frame pipeline: the image-only stages of the next frames overlap the decision of the current one, so a
           step boundary is not a synchronisation point), one pair of CUDA events on the handle's stream around the K
           steps, the end event after a join of all the pipeline's streams; max over ranks.
e2e      : the same frames through the host C-ABI pair sindyn_track_submit / sindyn_track_collect with pinned HOST buffers
           (frame i + 2 is submitted before frame i is collected): H2D of the BGR + depth frame and D2H of the dilated
           mask, the label image, the key points and the descriptors inside the timed region.  e2e.frame_at_a_time: the
           same through sindyn_track_frame (what the C++ classes of include/sindyn_classes.hpp call), which returns a
           frame's results before it accepts the next frame.
roofline : the dominant kernel, k_brox_sor (temporally blocked red-black SOR), timed per launch with CUDA events on the
           handle's stream (sindyn_brox_profile); algorithmic bytes = 52 B per pixel and sweep.
roofline_per_stage: SURVEY.md 8(d) algorithmic bytes of every stage / its device ms / the measured HBM peak.
cpu_baseline / --impl reference: the reference's CPU path restated in oracle/ (checker code): the full oracle pipeline
           (flow + k-means + edges + PEAC + merge + decision + dilation + masked ORB) on the host cores of this box, on a
           bounded sample of the same sequence.
--sequences S: BASELINE configs[4] as written -- S independent sequences spread over the N GPUs (S/N per GPU, strong
           scaling of a fixed job); default S = N (one sequence per GPU, weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "dyn-detect frame pairs/sec @640x480"
UNIT = "pairs/s"
N_FRAMES = 300               # BASELINE.json configs[2]
FRAMES_PER_STEP = 15
HOLE_RATE = 0.0005
SEQ_BASE = 3                 # sequence s of rank r uses seed BASE_SEED + SEQ_BASE + ...
ORB_CFG = (1500, 1.2, 8, 15, 5)   # TUM3.yaml ORBextractor.*
ALGO_BYTES_PER_PX_SWEEP = 52.0    # 52 B per pixel and red-black sweep

# SURVEY.md 8(d) algorithmic bytes per frame pair at 640x480 (C1), per stage
STAGE_BYTES_C1 = {
    "gray_resize_upsample": 4.0e6,
    "brox": 1.925e9,
    "refine_largemotion": 0.197e9,
    "homography": 2961 * 2 * 2 * 4.0,
    "residual_masks": 7.7e6,
    "kmeans": 45.7e6,
    "depth_edges": 3.7e6,
    "plane_edges_peac": 0.9e6,
    "filter_recluster_decide": 24.6e6,
    "orb": 5.7e6 + 0.8e6,
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profile():
    p = os.path.join(ROOT, "profiles", "brox_sor_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []          # (host time of arrival, csv line)
        self.t_begin = None      # samples before this host time are dropped (nvidia-smi is started early: it needs ~0.5 s to come up)

    def begin_window(self):
        self.t_begin = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, ln in self.lines:
            if self.t_begin is not None and t < self.t_begin:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_frames(seq_index, n=N_FRAMES, workers=None):
    from sindslam_b200 import synth
    cam = synth.TUM3
    _, frames = synth.make_sequence_parallel(n, cam, seq=SEQ_BASE + seq_index, kind="box", start=0, hole_rate=HOLE_RATE, workers=workers)
    return cam, frames


def frame_order(n_frames, count):
    """Indices of `count` consecutive frames: 1, 2, ..., n-1, n-2, ..., 0, 1, ... (forward, then backward: no jumps)."""
    period = list(range(1, n_frames)) + list(range(n_frames - 2, -1, -1))
    return [period[i % len(period)] for i in range(count)]


def workload_config(cam, extra=None):
    c = {"workload": "configs[2]: full dynamic-region pipeline (DetectDynaArea: Brox flow alpha 0.197 gamma 50 scale 0.8 10 inner 77 outer 10 SOR "
                     "+ refinement + RHO homography ego-motion residual + Otsu/Triangle masks, 4-level k-means re-clustering, depth edges, "
                     "PEAC plane edges, split/RAG/merge, per-cluster decision, mask morphology; 15x15 dilation; masked ORB 1500 features "
                     f"x 8 levels with key-point erasure) over a {N_FRAMES}-frame synthetic walking_xyz-shaped 640x480 sequence per GPU; "
                     f"one step = {FRAMES_PER_STEP} consecutive frame pairs",
         "width": cam.width, "height": cam.height, "flow_grid": "384x288", "frames": N_FRAMES, "frames_per_step": FRAMES_PER_STEP,
         "depth_hole_rate": HOLE_RATE, "plane_edges": True,
         "l2": "inputs (460 MB of resident frames per sequence, each read once per pass) exceed the 126 MB L2; a 256 MiB device write "
               "also flushes it between timed steps",
         "parallelism": "replicas (independent sequences per GPU, no collectives)"}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------- CPU arms (oracle = checker code)
def cpu_full_pipeline(frames, cam, engine, kmeans_impl, budget_s, max_pairs, start=1):
    """The reference's whole per-frame path restated in oracle/ on a bounded sample: DetectDynaArea (oracle flow engine,
    k-means, edges, PEAC, merge, decision) + 15x15 dilation + masked ORB (oracle extractor).  Returns (pairs/s, n, seconds)."""
    import cv2
    from oracle import dynadetect_oracle as orc
    from oracle import orb_oracle as oo
    o = orc.DynaDetectOracle(frames[start - 1].bgr, frames[start - 1].bgr, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor,
                             plane_edges=True, engine=engine, refine=True, kmeans_impl=kmeans_impl)
    orb = oo.OrbOracle(*ORB_CFG)
    el = orc.ellipse(15)

    def one(f):
        r = o.detect(f.bgr, f.depth)
        dil = cv2.dilate(r["mask"], el)
        orb.extract(cv2.cvtColor(f.bgr, cv2.COLOR_RGB2GRAY), dil)

    one(frames[start])      # warm-up (thread pools, library loads); also moves the state off the all-zero first frame
    t0 = time.perf_counter()
    n = 0
    for k in range(start + 1, min(start + 1 + max_pairs, len(frames))):
        one(frames[k])
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)     # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses all host threads at every N
    cv2.setNumThreads(cores)
    total = args.warmup + args.steps
    per = 2                                         # bounded sample: 2 frame pairs of the step's 15 (+1 warm-up pair per step)
    need = min(N_FRAMES, 2 + total * per + 2)
    cam, frames = make_frames(0, n=need)
    per_step, n_pairs = [], 0
    for s in range(total):
        a = 1 + (s * per) % max(1, need - per - 2)
        v, n, dt = cpu_full_pipeline(frames, cam, "deepflow", "cv2", budget_s=20.0, max_pairs=per, start=a)
        if s >= args.warmup:
            per_step.append(dt / n)
            n_pairs += n
    sec_per_pair = float(np.mean(per_step))
    value = 1.0 / sec_per_pair
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_pair * 1e3 * FRAMES_PER_STEP, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(cam),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_pairs} frame pairs ({per} of every step's {FRAMES_PER_STEP}) of the same synthetic sequence through the full oracle "
                                   "pipeline: DeepFlow-restated flow (the reference's default CPU engine; cv2.VariationalRefinement pyramid, optflow is "
                                   "not in cv2-headless) + cv2 refinement / RHO / Otsu / Triangle + cv2.kmeans + depth edges + oracle/peac_cpu.c + "
                                   "split/RAG/merge + decision + 15x15 dilation + oracle ORB extractor; ms_per_step is scaled to the step's 15 pairs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
STAGE_NAMES = ["upload_gray_resize", "brox", "largemotion_refine_upsample", "homography", "residual_masks", "kmeans", "depth_edges",
               "plane_edge_filter", "recluster", "decide", "total_device", "plane_edges_peac"]


def stage_profile(cam, frames, device, n=24):
    """Per-stage device ms (CUDA events inside the library, stage_timing=1: classic launch path, no graphs) and the ORB time,
    averaged over n frames of the sequence."""
    import torch
    from sindslam_b200.capi import Orb, SinDyn
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=device, refine=1, plane_edges=1, stage_timing=1)
    orb = Orb(*ORB_CFG, cam.width, cam.height, device=device)
    st = torch.cuda.Stream(device=device)
    orb.set_stream(st.cuda_stream)
    sd.set_prev_frames(frames[0].bgr, frames[0].bgr)
    import cv2
    acc = np.zeros(16)
    orb_ms, cnt = 0.0, 0
    for k in range(1, n + 4):
        mask, label = sd.detect(frames[k].bgr, frames[k].depth, k)
        dil = sd.morph_ellipse(mask, 15, 0)
        gray = cv2.cvtColor(frames[k].bgr, cv2.COLOR_RGB2GRAY)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        orb.extract(gray, dil)
        e1.record(st)
        e1.synchronize()
        if k >= 4:
            acc += sd.stage_ms()
            orb_ms += e0.elapsed_time(e1)
            cnt += 1
    sd.close()
    orb.close()
    acc /= cnt
    ms = {nm: float(acc[i]) for i, nm in enumerate(STAGE_NAMES)}
    ms["orb_extract_incl_copies"] = orb_ms / cnt
    return ms


def roofline_per_stage(ms, peak):
    t = {
        "gray_resize_upsample": ms["upload_gray_resize"],
        "brox": ms["brox"],
        "refine_largemotion": ms["largemotion_refine_upsample"],
        "homography": ms["homography"],
        "residual_masks": ms["residual_masks"],
        "kmeans": ms["kmeans"],
        "depth_edges": ms["depth_edges"],
        "plane_edges_peac": ms["plane_edges_peac"],
        "filter_recluster_decide": ms["plane_edge_filter"] + ms["recluster"] + ms["decide"],
        "orb": ms["orb_extract_incl_copies"],
    }
    out = {}
    for k, b in STAGE_BYTES_C1.items():
        if t[k] > 0:
            gbs = b / (t[k] * 1e-3) / 1e9
            out[k] = {"ms": round(t[k], 4), "algorithmic_bytes": b, "achieved_gbs": round(gbs, 2), "frac": round(gbs / peak, 5)}
    return out


def multi_sequence_stats(cam, device, n_seq=4, n_frames=76, steps=4):
    """Batched throughput on ONE GPU ('report both single-frame latency and batched throughput'): n_seq independent
    sequences, one detector + extractor handle, stream and host thread each, frames resident, the same full per-frame path as
    the headline.  A single sequence leaves most of the chip idle (serial chains on one or a few SMs), so sequences overlap."""
    import torch
    from sindslam_b200 import synth
    from sindslam_b200.capi import Orb, SinDyn
    hs = []
    for s in range(n_seq):
        _, fr = synth.make_sequence_parallel(n_frames, cam, seq=200 + s, kind="box", start=0, hole_rate=HOLE_RATE)
        sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=device, refine=1, plane_edges=1)
        orb = Orb(*ORB_CFG, cam.width, cam.height, device=device)
        st = torch.cuda.Stream(device=device)
        sd.set_stream(st.cuda_stream)
        for i, f in enumerate(fr):
            sd.upload_frame(i, f.bgr, f.depth)
        sd.set_prev_frames(fr[0].bgr, fr[0].bgr)
        hs.append((sd, orb))
    order = frame_order(n_frames, (steps + 1) * FRAMES_PER_STEP)

    def work(h, a, b):
        for j in range(a, b):
            h[1].track_frame_resident(h[0], order[j], j)
        h[0].synchronize()

    def run(a, b):
        th = [threading.Thread(target=work, args=(h, a, b)) for h in hs]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()

    run(0, FRAMES_PER_STEP)                         # warm-up: graph capture, first frames
    t0 = time.perf_counter()
    run(FRAMES_PER_STEP, (steps + 1) * FRAMES_PER_STEP)
    dt = time.perf_counter() - t0
    for sd, orb in hs:
        orb.track_results(sd)
        orb.close()
        sd.close()
    return {"sequences_per_gpu": n_seq, "pairs_per_s": n_seq * steps * FRAMES_PER_STEP / dt, "frames_per_sequence": steps * FRAMES_PER_STEP,
            "note": "same full per-frame path as the headline, n_seq concurrent sequences on one GPU, host wall clock"}


def other_configs(device):
    """The other single-GPU BASELINE configurations, outside the headline: configs[1] -- the flow + ego-motion-residual branch alone
    on 640x480 pairs (sindyn_flow_residual_resident: Brox, refinement, RHO, residual, thresholds, masks; frame at a time) -- and
    configs[3] -- the full per-frame path on an 848x480 D455-shaped sequence with a humanoid-sized dynamic region (frame pipeline,
    like the headline)."""
    import torch
    from sindslam_b200 import synth
    from sindslam_b200.capi import Orb, SinDyn
    out = {}
    # ---- configs[1]
    cam = synth.TUM3
    n = 46
    _, fr = synth.make_sequence_parallel(n, cam, seq=300, kind="box", start=0, hole_rate=HOLE_RATE)
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=device, refine=1, plane_edges=1)
    st = torch.cuda.Stream(device=device)
    sd.set_stream(st.cuda_stream)
    for i, f in enumerate(fr):
        sd.upload_frame(i, f.bgr, f.depth)
    sd.set_prev_frames(fr[0].bgr, fr[0].bgr)
    for k in range(1, 16):
        sd.flow_residual_resident(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for k in range(16, n):
        sd.flow_residual_resident(k)
    e1.record(st)
    torch.cuda.synchronize()
    out["configs1_flow_residual_branch"] = {"pairs_per_s": 1e3 * (n - 16) / e0.elapsed_time(e1), "ms_per_pair": e0.elapsed_time(e1) / (n - 16),
                                            "note": "640x480, frames resident, one pair at a time (no look-ahead), device time"}
    sd.close()
    # ---- configs[3]
    cam = synth.D455_848
    n = 61
    _, fr = synth.make_sequence_parallel(n, cam, seq=301, kind="humanoid", start=0, hole_rate=HOLE_RATE)
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=device, refine=1, plane_edges=1)
    orb = Orb(*ORB_CFG, cam.width, cam.height, device=device)
    st = torch.cuda.Stream(device=device)
    sd.set_stream(st.cuda_stream)
    for i, f in enumerate(fr):
        sd.upload_frame(i, f.bgr, f.depth)
    sd.set_prev_frames(fr[0].bgr, fr[0].bgr)
    for k in range(1, 16):
        orb.track_frame_resident(sd, k, k)
    orb.track_join(sd)
    torch.cuda.synchronize()
    e0.record(st)
    for k in range(16, n):
        orb.track_frame_resident(sd, k, k)
    orb.track_join(sd)
    e1.record(st)
    torch.cuda.synchronize()
    orb.track_results(sd)        # surfaces capacity errors
    out["configs3_848x480_humanoid"] = {"pairs_per_s": 1e3 * (n - 16) / e0.elapsed_time(e1), "ms_per_pair": e0.elapsed_time(e1) / (n - 16),
                                        "note": "848x480 D455-shaped, humanoid-sized dynamic region, full per-frame path incl. 15x15 dilation and masked ORB, "
                                                "frames resident, frame pipeline, device time"}
    orb.close()
    sd.close()
    return out


def run_ours(args):
    import torch
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    from sindslam_b200 import replicas
    rank, world, local = replicas.rank_info()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # replicas only: the process group carries the barrier and the max-over-ranks of the timings, nothing on the data path
        dist.init_process_group("gloo")
    from sindslam_b200.capi import Orb, SinDyn

    n_seq_total = args.sequences if args.sequences else world
    my_seqs = [s for s in range(n_seq_total) if s % world == rank]
    cam = None
    seqs = []
    workers = max(1, (os.cpu_count() or 1) // world)
    for s in my_seqs:
        cam, frames = make_frames(s, workers=workers)
        seqs.append(frames)
    assert seqs, "more ranks than sequences"

    # one explicit (non-default) stream per sequence for the handle, the L2 flush and the timing events
    handles = []
    for frames in seqs:
        sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=local, refine=1, plane_edges=1)
        orb = Orb(*ORB_CFG, cam.width, cam.height, device=local)
        stream = torch.cuda.Stream(device=local)
        sd.set_stream(stream.cuda_stream)
        for i, f in enumerate(frames):
            sd.upload_frame(i, f.bgr, f.depth)
        sd.set_prev_frames(frames[0].bgr, frames[0].bgr)          # rgbd_tum_noros.cc:103-107
        handles.append((sd, orb, stream, frames))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    total_steps = args.warmup + args.steps
    order = frame_order(N_FRAMES, total_steps * FRAMES_PER_STEP)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.6)
        clocks.begin_window()

    # ---------------- device-resident throughput (value)
    def run_steps_resident(sd, orb, stream, s0, s1, ev):
        # ev: one (start, end) event pair around ALL the steps.  The frame pipeline keeps several frames in flight on its own
        # streams, so a step boundary is not a synchronisation point: the L2 flush of a step is enqueued on the handle's stream
        # next to the frames (it evicts the lines as intended and its 40 us are inside the timed region), the end event follows
        # a join of all the pipeline's streams.
        if ev is not None:
            ev[0].record(stream)
        for s in range(s0, s1):
            with torch.cuda.stream(stream):
                flush.fill_(s & 255)
            for j in range(FRAMES_PER_STEP):
                orb.track_frame_resident(sd, order[s * FRAMES_PER_STEP + j], s * FRAMES_PER_STEP + j)
        orb.track_join(sd)
        if ev is not None:
            ev[1].record(stream)

    def over_handles(fn):
        if len(handles) == 1:
            fn(handles[0])
            return
        th = [threading.Thread(target=fn, args=(h,)) for h in handles]
        for t in th:
            t.start()
        for t in th:
            t.join()

    over_handles(lambda h: run_steps_resident(h[0], h[1], h[2], 0, args.warmup, None))
    barrier()
    l0 = sum(h[0].launches + h[1].launches for h in handles)
    evs = {id(h[0]): (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for h in handles}
    t_host0 = time.perf_counter()
    over_handles(lambda h: run_steps_resident(h[0], h[1], h[2], args.warmup, total_steps, evs[id(h[0])]))
    barrier()
    t_host1 = time.perf_counter()
    gpu_launches = sum(h[0].launches + h[1].launches for h in handles) - l0
    # one sequence: device time of the K steps between two events on the handle's stream; several concurrent sequences on one
    # GPU: the wall time of the region (events of different streams overlap)
    if len(handles) == 1:
        dev_ms = evs[id(handles[0][0])][0].elapsed_time(evs[id(handles[0][0])][1])
    else:
        dev_ms = (t_host1 - t_host0) * 1e3
    for h in handles:   # surface capacity errors of the resident path
        h[1].track_results(h[0])

    # ---------------- end to end through the host C-ABI (pinned host buffers)
    sd, orb, stream, frames = handles[0]
    pin_bgr = [torch.from_numpy(np.ascontiguousarray(f.bgr)).pin_memory() for f in frames]
    pin_dep = [torch.from_numpy(np.ascontiguousarray(f.depth).view(np.int16)).pin_memory() for f in frames]
    bgr_np = [t.numpy() for t in pin_bgr]
    dep_np = [t.numpy().view(np.uint16) for t in pin_dep]
    cap = ORB_CFG[0] * 2 + 64
    pin_mask = torch.empty((cam.height, cam.width), dtype=torch.uint8).pin_memory().numpy()
    pin_label = torch.empty((cam.height, cam.width), dtype=torch.uint8).pin_memory().numpy()
    pin_kps = torch.empty(cap * 24, dtype=torch.uint8).pin_memory().numpy().view(Orb.KP_DTYPE)
    pin_desc = torch.empty((cap, 32), dtype=torch.uint8).pin_memory().numpy()
    # the e2e stream continues the sequence where the resident run stopped (same state recurrence)
    order2 = frame_order(N_FRAMES, 2 * total_steps * FRAMES_PER_STEP)[total_steps * FRAMES_PER_STEP:]
    n_kp = 0

    def e2e_steps(s0, s1):
        # the driver loop reads a recorded sequence (rgbd_tum_noros.cc:113-192), so frame i + 1 is at hand while frame i is being
        # processed: it is submitted first (upload + image-only stages overlap frame i), then frame i's results are collected on
        # the host -- every frame's inputs cross the bus inside the timed region and every frame's results come back before
        # frame i + 2 is accepted
        nonlocal n_kp
        ks = [order2[s * FRAMES_PER_STEP + j] for s in range(s0, s1) for j in range(FRAMES_PER_STEP)]
        if not ks:
            return
        ahead = 2                     # frames submitted before the oldest one is collected (at most three in flight)
        for i in range(min(ahead, len(ks))):
            orb.track_submit(sd, bgr_np[ks[i]], dep_np[ks[i]], ks[i])
        for i in range(len(ks)):
            if i + ahead < len(ks):
                orb.track_submit(sd, bgr_np[ks[i + ahead]], dep_np[ks[i + ahead]], ks[i + ahead])
            _, _, kps, _ = orb.track_collect(sd, mask_out=pin_mask, label_out=pin_label, kps_out=pin_kps, desc_out=pin_desc)
            n_kp += len(kps)

    e2e_steps(0, args.warmup)
    barrier()
    n_kp = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e2e_steps(args.warmup, total_steps)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    # the same frames one at a time (sindyn_track_frame: a frame's results are on the host before the next frame is accepted --
    # the latency-bound figure of a caller that cannot look one frame ahead), a few steps, reported next to the headline
    n_sync = min(args.steps, 4)
    order3 = frame_order(N_FRAMES, (2 * total_steps + n_sync + 1) * FRAMES_PER_STEP)[2 * total_steps * FRAMES_PER_STEP:]
    for j in range(FRAMES_PER_STEP):
        orb.track_frame(sd, bgr_np[order3[j]], dep_np[order3[j]], j, mask_out=pin_mask, label_out=pin_label, kps_out=pin_kps, desc_out=pin_desc)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for j in range(FRAMES_PER_STEP, (n_sync + 1) * FRAMES_PER_STEP):
        orb.track_frame(sd, bgr_np[order3[j]], dep_np[order3[j]], j, mask_out=pin_mask, label_out=pin_label, kps_out=pin_kps, desc_out=pin_desc)
    e3.record(stream)
    barrier()
    sync_pairs_s = n_sync * FRAMES_PER_STEP / (e2.elapsed_time(e3) * 1e-3)
    clk = clocks.stop() if rank == 0 else None
    kp_per_frame = n_kp / max(1, args.steps * FRAMES_PER_STEP)

    dev_ms, e2e_ms = replicas.max_over_ranks([dev_ms, e2e_ms], dist, "cpu" if dist is not None else "cuda")
    if rank == 0:
        frames_per_rank_value = args.steps * FRAMES_PER_STEP * len(handles)
        value = world * frames_per_rank_value / (dev_ms * 1e-3)
        e2e = world * args.steps * FRAMES_PER_STEP / (e2e_ms * 1e-3)
        prof = sd.brox_profile()
        prof = sd.brox_profile()  # second call: warm
        peak, which = peaks()
        algo_bytes = ALGO_BYTES_PER_PX_SWEEP * prof["pixel_sweeps"]
        ach = algo_bytes / (prof["sor_ms"] * 1e-3) / 1e9
        extras = {}
        if not args.headline_only:
            ms = stage_profile(cam, frames, local)
            extras["stage_ms_device"] = ms
            extras["roofline_per_stage"] = roofline_per_stage(ms, peak)
            extras["multi_sequence"] = multi_sequence_stats(cam, local) if world == 1 else None
            extras["other_configs"] = other_configs(local) if world == 1 else None
            cpu_v, cpu_n, cpu_dt = cpu_full_pipeline(frames, cam, "brox", "fx", budget_s=15.0, max_pairs=40)
            extras["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                      "sample": f"{cpu_n} frame pairs in {cpu_dt:.1f} s through the full oracle pipeline (oracle/brox_cpu.c OpenMP Brox + cv2 "
                                                "refinement/RHO/thresholds + fixed-point k-means + edges + oracle/peac_cpu.c + merge + decision + dilation + "
                                                "oracle ORB); the reference's default CPU engine (DeepFlow) is timed by --impl reference"}
        else:
            extras["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "skipped (--headline-only)"}
        d2h = 2 * cam.width * cam.height + (ORB_CFG[0] * 2 + 64) * (24 + 32) + 80   # mask + labels + the fixed-capacity key-point / descriptor block + counters
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if not args.sequences else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(cam),     # identical for both arms (--impl reference prints the same dict)
            "run": {"refine": True, "sequences_total": n_seq_total, "sequences_per_gpu": len(handles), "keypoints_per_frame": round(kp_per_frame, 1)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": FRAMES_PER_STEP * cam.width * cam.height * 5,
                    "d2h_bytes_per_step": FRAMES_PER_STEP * d2h, "ms_per_step": e2e_ms / args.steps,
                    "frame_at_a_time": sync_pairs_s,
                    "note": "one sequence per GPU through sindyn_track_submit / sindyn_track_collect (pinned host buffers; frame i + 2 is submitted "
                            "before frame i is collected, at most three frames in flight); frame_at_a_time = the same through sindyn_track_frame, "
                            "which returns a frame's results before it accepts the next one"},
            "ms_per_frame": dev_ms / (args.steps * FRAMES_PER_STEP),
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "kernel": "k_brox_sor (temporally blocked red-black SOR, 5 sweeps per launch; the 9 finest pyramid levels)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic_from_profile(),
                         "peak_source": which, "algorithmic_bytes_per_launch": algo_bytes / max(prof["sor_launches"], 1),
                         "launches_per_solve": prof["sor_launches"], "avg_launch_us": 1e3 * prof["sor_ms"] / max(prof["sor_launches"], 1),
                         "sor_share_of_solve": prof["sor_ms"] / prof["solve_ms"]},
            "clocks": clk,
        }
        line.update(extras)
        print(json.dumps(line))
    for h in handles:
        h[1].close()
        h[0].close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # torchrun exports OMP_NUM_THREADS=1; the CPU legs (cpu_baseline on rank 0, --impl reference) use all host threads at every N,
    # so that their figures are comparable across N.  Must happen before torch / cv2 / the OpenMP oracle load their runtimes.
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and int(os.environ.get("RANK", "0")) == 0:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sequences", type=int, default=0, help="total number of independent sequences spread over the GPUs (default: one per GPU)")
    ap.add_argument("--headline-only", action="store_true", help="skip the per-stage profile and the CPU baseline (ncu runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
