"""The N > 1 path of bench.py (replicas: one sequence per rank, barrier + max-over-ranks timing, whole-job aggregate)
exercised with gloo, world size 2, on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sindslam_b200 import replicas, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, l = replicas.rank_info()
    assert (r, w, l) == (rank, world, rank)
    _, frames = synth.make_sequence(2, synth.TUM3, seq=replicas.sequence_seed_index(r), kind="box", start=8)
    checksum = int(frames[0].bgr.astype(np.int64).sum())
    fake_ms = [100.0 + 50.0 * rank, 80.0 - 10.0 * rank]       # rank 1 is slower on the first timing, rank 0 on the second
    dist.barrier()
    mx = replicas.max_over_ranks(fake_ms, dist)
    sums = torch.tensor([checksum], dtype=torch.int64)
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    if rank == 0:
        out.put((mx, [int(g[0]) for g in gathered], replicas.aggregate_throughput(world, 30, mx[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicas_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    mx, sums, agg = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mx == [150.0, 80.0]                     # max over ranks, element-wise
    assert sums[0] != sums[1]                      # each rank streams its own sequence (seed + rank)
    assert abs(agg - 2 * 30 / 0.150) < 1e-9        # whole-job pairs/s = all ranks' pairs / slowest rank's time
