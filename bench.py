#!/usr/bin/env python
"""bench.py -- dyn-detect frame pairs/s at 640x480 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): one 640x480 frame pair per step through the flow + ego-motion
residual branch (DynaDetect::DetectDynaByDenseOpticalFLow, DynaDetect.cc:1023-1374): gray/resize, Brox dense
flow with the reference's parameters, large-motion test, variational refinement, up-sampling, sample weighting,
robust homography, residual, Otsu/Triangle thresholds, two masks.  A synthetic TUM-format RGB-D sequence is
streamed in order (one per rank, seed 20241108 + rank); the wrap-around of the frame cycle triggers the
reference's large-motion fallback (a second Brox solve) exactly as a real sequence would.

value    : pairs/s, frames resident in HBM (device slots), CUDA-event time summed over the K steps, L2 flushed
           between steps (256 MiB write), max over ranks.
e2e      : pairs/s through the host C-ABI call sindyn_flow_residual with pinned HOST buffers: H2D of the BGR frame
           and D2H of both masks inside the timed region.
roofline : the temporally blocked red-black SOR kernel (k_brox_sor: 5 sweeps of one lagged-nonlinearity iteration of
           one pyramid level per launch) timed per launch with CUDA events on the handle's stream
           (sindyn_brox_profile); algorithmic bytes = sweeps of the launch x 52 B per pixel x the
           pixels one launch covers.
cpu_baseline / --impl reference: the reference's CPU path restated in oracle/ (checker code), timed on the
           host cores of this box on a bounded sample of the same sequence.
"""
from __future__ import annotations

import argparse
import ctypes
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
N_FRAMES = 16
ALGO_BYTES_PER_PX_SWEEP = 52.0   # 52 B per pixel and red-black sweep (the 100 B coefficient preparation is k_brox_system's)


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


def make_frames(rank):
    from sindslam_b200 import synth
    cam = synth.TUM3
    scene, frames = synth.make_sequence(N_FRAMES, cam, seq=rank, kind="box", start=8)
    return cam, frames


def cpu_pairs_per_s(frames, engine, budget_s, max_pairs):
    """The reference's CPU flow+residual path (oracle restatement) on a bounded sample: frames 2.. of the sequence."""
    import cv2
    from oracle import dynadetect_oracle as orc
    cv2.setNumThreads(os.cpu_count() or 1)
    z = np.zeros(frames[0].bgr.shape[:2], np.uint8)
    orc.flow_residual_cpu(frames[2].bgr, frames[1].bgr, frames[0].bgr, z, z, engine)  # warm-up (thread pools, lib load)
    t0 = time.perf_counter()
    n = 0
    for i in range(2, 2 + max_pairs):
        k = 2 + (i - 2) % (len(frames) - 2)
        orc.flow_residual_cpu(frames[k].bgr, frames[k - 1].bgr, frames[k - 2].bgr, z, z, engine)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cam, frames = make_frames(0)
    per_step = []
    n_pairs = 0
    for s in range(args.warmup + args.steps):
        v, n, dt = cpu_pairs_per_s(frames, "deepflow", budget_s=min(3.0, 150.0 / (args.warmup + args.steps)), max_pairs=12)
        if s >= args.warmup:
            per_step.append(dt / n)
            n_pairs += n
    sec_per_pair = float(np.mean(per_step))
    value = 1.0 / sec_per_pair
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_pair * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(cam, extra={"engine": "DeepFlow-restated (cv2.VariationalRefinement pyramid; optflow not in cv2-headless) "
                                               "+ cv2 refinement/RHO homography/Otsu/Triangle, all host threads"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_pairs} frame pairs of the synthetic 640x480 sequence (each step = up to 12 pairs / 3 s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(cam, extra=None):
    c = {"workload": "configs[1]: single 640x480 frame pair, Brox dense flow (alpha 0.197, gamma 50, scale 0.8, 10 inner, 77 outer, "
                     "10 SOR) + refinement + homography ego-motion residual + thresholds -> two masks, streamed over a "
                     f"{N_FRAMES}-frame synthetic TUM-format sequence per GPU",
         "width": cam.width, "height": cam.height, "flow_grid": "384x288", "frames": N_FRAMES,
         "l2": "flushed between timed steps (256 MiB device write)", "parallelism": "replicas (one sequence per GPU, no collectives)"}
    if extra:
        c.update(extra)
    return c


def full_pipeline_stats(cam, frames, device, refine):
    """BASELINE configs[2] in miniature (reported next to the headline, not the headline): the whole per-frame path --
    sindyn_detect (flow + residual + k-means + depth edges + re-clustering + decision), 15x15 dilation and the masked ORB
    extraction -- through the host C ABI, with per-stage device milliseconds."""
    import cv2
    from sindslam_b200.capi import Orb, SinDyn
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=device, refine=refine, stage_timing=1)
    orb = Orb(1500, 1.2, 8, 15, 5, cam.width, cam.height, device=device)
    sd.set_prev_frames(frames[1].bgr, frames[0].bgr)
    grays = [cv2.cvtColor(f.bgr, cv2.COLOR_RGB2GRAY) for f in frames]
    acc = np.zeros(16)
    n = 0
    t_det = t_orb = t_ff = t_cloud = 0.0
    nkp = 0
    prev = last_pts = None
    t_match = 0.0
    for k in range(2, len(frames)):
        t0 = time.perf_counter()
        mask, label = sd.detect(frames[k].bgr, frames[k].depth, k)
        mask = sd.morph_ellipse(mask, 15, 0)
        t1 = time.perf_counter()
        kps, kdesc = orb.extract(grays[k], mask)
        t2 = time.perf_counter()
        # the "next" rows downstream of the path (not part of pairs_per_s): Frame construction (f2), dense-map cloud (f3)
        un, dep, _, _, _, _ = orb.frame_features(frames[k].depth, cam.fx, cam.fy, cam.cx, cam.cy, (0.0, 0.0, 0.0, 0.0, 0.0), 40.0, 1.0 / cam.depth_factor)
        t3 = time.perf_counter()
        # frame-to-frame descriptor matching (f4) against the previous frame's key points as map points
        t_m0 = time.perf_counter()
        if last_pts is not None:
            orb.search_by_projection(last_pts, np.linalg.inv(frames[k].T_wc), np.linalg.inv(frames[k - 1].T_wc), cam.fx, cam.fy, cam.cx, cam.cy,
                                     40.0, 40.0 / cam.fx, 15.0)
        t_m1 = time.perf_counter()
        nk = len(kps)
        z = dep[:nk]
        pc = np.stack([(un[:nk, 0] - cam.cx) * z / cam.fx, (un[:nk, 1] - cam.cy) * z / cam.fy, z, np.ones(nk)], 1)
        last_pts = dict(xyz_w=(frames[k].T_wc @ pc.T).T[:, :3].astype(np.float32), valid=z > 0, desc=kdesc, octave=kps["octave"], angle=kps["angle"],
                        observed=np.zeros(nk, bool))
        t3b = time.perf_counter()
        if prev is not None:
            T_rel = np.linalg.inv(frames[prev[0]].T_wc) @ frames[k].T_wc
            sd.cloud_consistent(frames[k].bgr, frames[k].depth, frames[prev[0]].depth, mask, prev[1], label, T_rel, frames[k].T_wc)
        t4 = time.perf_counter()
        prev = (k, mask)
        if k >= 4:
            acc += sd.stage_ms()
            t_det += t1 - t0
            t_orb += t2 - t1
            t_ff += t3 - t2
            t_cloud += t4 - t3b
            t_match += t_m1 - t_m0
            nkp += len(kps)
            n += 1
    sd.close()
    orb.close()
    acc /= max(n, 1)
    names = ["upload_gray_resize", "brox", "largemotion_refine_upsample", "homography", "residual_masks", "kmeans", "depth_edges",
             "plane_edge_filter", "recluster", "decide", "total_device"]
    return {"workload": "sindyn_detect + 15x15 dilation + masked ORB (1500 features, 8 levels) per frame, host buffers",
            "pairs_per_s": n / (t_det + t_orb), "detect_ms_wall": 1e3 * t_det / n, "orb_ms_wall": 1e3 * t_orb / n,
            "keypoints_per_frame": nkp / n, "stage_ms_device": {nm: float(acc[i]) for i, nm in enumerate(names)},
            "plane_edges": True,
            "next_rows_ms_wall": {"f2_frame_features": 1e3 * t_ff / n, "f3_cloud_consistent": 1e3 * t_cloud / n,
                                  "f4_search_by_projection": 1e3 * t_match / n,
                                  "note": "host C-ABI calls with pageable numpy buffers (H2D + D2H inside), not counted in pairs_per_s"}}


def multi_sequence_stats(cam, device, refine, n_seq=8, steps=24):
    """Batched throughput ('report both single-frame latency and batched throughput'): n_seq independent
    sequences on ONE GPU, one handle + stream + host thread each, device-resident frames.  A single sequence leaves most
    of the chip idle (the coarse pyramid levels run on a handful of SMs), so concurrent sequences overlap."""
    import torch
    from sindslam_b200 import synth
    from sindslam_b200.capi import SinDyn
    handles = []
    for s in range(n_seq):
        _, fr = synth.make_sequence(N_FRAMES, cam, seq=100 + s, kind="box", start=8)
        sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=device, refine=refine)
        for i, f in enumerate(fr):
            sd.upload_frame(i, f.bgr, f.depth)
        sd.set_prev_frames(fr[1].bgr, fr[0].bgr)
        for i in range(3):
            sd.flow_residual_resident(2 + i, roll=True)
        sd.synchronize()
        handles.append(sd)

    def work(sd):
        for i in range(steps):
            sd.flow_residual_resident(5 + i % (N_FRAMES - 5), roll=True)
        sd.synchronize()

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(sd,)) for sd in handles]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for sd in handles:
        sd.close()
    return {"sequences_per_gpu": n_seq, "pairs_per_s": n_seq * steps / dt, "steps_per_sequence": steps,
            "note": "same flow + residual workload as the headline, host wall clock, n_seq concurrent handles on one GPU"}


def run_ours(args):
    import torch
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    from sindslam_b200 import replicas
    rank, world, local = replicas.rank_info()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sindslam_b200.capi import SinDyn

    cam, frames = make_frames(replicas.sequence_seed_index(rank))
    refine = 0 if args.no_refine else 1
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=local, refine=refine)
    # one explicit (non-default) stream for the handle, the L2 flush and the timing events: torch's default stream has handle
    # 0, which sindyn_set_stream reads as "use the handle's own stream" -- events recorded there would not bracket the work
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    sd.set_stream(stream.cuda_stream)
    for i, f in enumerate(frames):
        sd.upload_frame(i, f.bgr, f.depth)
    sd.set_prev_frames(frames[1].bgr, frames[0].bgr)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    order = [(2 + i) % N_FRAMES for i in range(args.warmup + args.steps)]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.6)          # let nvidia-smi come up so that short runs still get samples under load
        clocks.begin_window()    # window = warm-up steps + both timed regions: one stretch of continuous load
    # ---------------- device-resident throughput
    for i in range(args.warmup):
        flush.fill_(i & 255)
        sd.flow_residual_resident(order[i], roll=True)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = sd.launches
    n_large = 0
    for i in range(args.steps):
        flush.fill_(i & 255)
        ev[i][0].record(stream)
        sd.flow_residual_resident(order[args.warmup + i], roll=True)
        ev[i][1].record(stream)
    barrier()
    gpu_launches = sd.launches - l0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # ---------------- end to end through the host C-ABI (pinned host buffers)
    pin_in = [torch.from_numpy(np.ascontiguousarray(f.bgr)).pin_memory() for f in frames]
    pin_lo = torch.empty((cam.height, cam.width), dtype=torch.uint8).pin_memory()
    pin_hi = torch.empty((cam.height, cam.width), dtype=torch.uint8).pin_memory()
    lo_np, hi_np = pin_lo.numpy(), pin_hi.numpy()
    in_np = [t.numpy() for t in pin_in]
    for i in range(args.warmup):
        sd.flow_residual(in_np[order[i]], True, lo_np, hi_np)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lm_flag = ctypes.c_int(0)
    e0.record(stream)
    for i in range(args.steps):
        sd.flow_residual(in_np[order[args.warmup + i]], True, lo_np, hi_np)
        sd.lib.sindyn_get_flow_results(sd.h, None, None, None, None, None, ctypes.byref(lm_flag))
        n_large += lm_flag.value
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None

    dev_ms, e2e_ms = replicas.max_over_ranks([dev_ms, e2e_ms], dist, "cuda")
    if rank == 0:
        value = replicas.aggregate_throughput(world, args.steps, dev_ms)
        e2e = replicas.aggregate_throughput(world, args.steps, e2e_ms)
        prof = sd.brox_profile()
        prof = sd.brox_profile()  # second call: warm
        peak, which = peaks()
        algo_bytes = ALGO_BYTES_PER_PX_SWEEP * prof["pixel_sweeps"]
        ach = algo_bytes / (prof["sor_ms"] * 1e-3) / 1e9
        stage = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor, device=local, refine=refine, stage_timing=1)
        stage.set_prev_frames(frames[1].bgr, frames[0].bgr)
        acc = np.zeros(16)
        for k in range(2, 10):
            stage.flow_residual(frames[k].bgr, True)
            if k >= 4:
                acc += stage.stage_ms()
        acc /= 6
        stage.close()
        if args.headline_only:
            full, multi, (cpu_v, cpu_n, cpu_dt) = None, None, (None, 0, 0.0)
        else:
            full = full_pipeline_stats(cam, frames, local, refine)
            multi = multi_sequence_stats(cam, local, refine)
            cpu_v, cpu_n, cpu_dt = cpu_pairs_per_s(frames, "brox", budget_s=12.0, max_pairs=200)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(cam, extra={"refine": bool(refine), "large_motion_steps": n_large}),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": cam.width * cam.height * 3,
                    "d2h_bytes_per_step": 2 * cam.width * cam.height, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "kernel": "k_brox_sor (temporally blocked red-black SOR, 5 sweeps per launch; the 9 finest pyramid levels)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic_from_profile(),
                         "peak_source": which, "algorithmic_bytes_per_launch": algo_bytes / max(prof["sor_launches"], 1),
                         "launches_per_solve": prof["sor_launches"], "avg_launch_us": 1e3 * prof["sor_ms"] / max(prof["sor_launches"], 1),
                         "sor_share_of_solve": prof["sor_ms"] / prof["solve_ms"]},
            "stage_ms": {"prep": float(acc[0]), "brox": float(acc[1]), "largemotion_refine_upsample": float(acc[2]),
                         "homography": float(acc[3]), "residual_masks": float(acc[4]), "total": float(acc[10])},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{cpu_n} frame pairs in {cpu_dt:.1f} s: oracle/brox_cpu.c (OpenMP) + cv2 refinement/RHO/thresholds"},
            "full_pipeline": full,
            "multi_sequence": multi,
            "clocks": clk,
        }
        print(json.dumps(line))
    sd.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-refine", action="store_true", help="skip the VariationalRefinement-equivalent pass (diagnostics only)")
    ap.add_argument("--headline-only", action="store_true", help="skip the full-pipeline / multi-sequence / CPU-baseline extras (ncu runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
