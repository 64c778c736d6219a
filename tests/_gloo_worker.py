"""Worker for tests/test_multi_gpu_host.py: one process per (pretend) GPU over gloo."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def broadcast_mode(out_dir):
    """rank 0 owns the stream and broadcasts each block (gloo here, NCCL on GPUs); each rank
    computes ITS clients with the oracle; rank 0 gathers and compares with computing all
    clients itself."""
    import torch
    from oracle import pyoracle as po
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    fs, block_elems, n_clients, n_blocks = 2016000, 8192, 12, 4
    taps = po.lpf_design(1.0, fs, 24000, 16400)
    centers = [-900000 + 150000 * c for c in range(n_clients)]
    mine = bench.shard_clients(n_clients, rank, world)
    filters = {c: po.OracleFilter(42, taps, centers[c], fs, block_elems) for c in mine}
    src = bench.synth_blocks("cu8", n_blocks, block_elems, seed=bench.stream_seed(0)) if rank == 0 else None
    outs = {c: [] for c in mine}
    for b in range(n_blocks):
        buf = torch.from_numpy(src[b].copy()) if rank == 0 else torch.empty(block_elems, dtype=torch.uint8)
        dist.broadcast(buf, src=0)
        x = buf.numpy()
        for c in mine:
            outs[c].append(filters[c].process_cf32("cu8", x).tobytes().hex())
    gathered = [None] * world
    dist.all_gather_object(gathered, outs)
    if rank == 0:
        allf = [po.OracleFilter(42, taps, centers[c], fs, block_elems) for c in range(n_clients)]
        same = True
        merged = {}
        for g in gathered:
            merged.update(g)
        for b in range(n_blocks):
            for c in range(n_clients):
                same &= merged[c][b] == allf[c].process_cf32("cu8", src[b]).tobytes().hex()
        with open(os.path.join(out_dir, "broadcast.json"), "w") as f:
            json.dump({"clients_checked": len(merged), "bit_identical": bool(same)}, f)
    dist.barrier()
    dist.destroy_process_group()


def main():
    out_dir = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] == "broadcast":
        return broadcast_mode(out_dir)
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    wl = bench.workload("cfg2", "default")
    # every rank owns an independent stream: different seed -> different samples
    blocks = bench.synth_blocks(wl["fmt"], 2, 4096, seed=bench.stream_seed(rank))
    digest = int(np.frombuffer(blocks.tobytes(), dtype=np.uint64).sum() % (1 << 62))
    # pretend device times: rank r took (10 + r) ms for 100 steps
    local_ms = 10.0 + rank
    dist.barrier()
    ms_max = bench.max_over_ranks(local_ms, world)
    value = bench.job_throughput_msps(wl["block_samples"], 100, world, ms_max)
    gathered = [None] * world
    dist.all_gather_object(gathered, digest)
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "ms_max": ms_max, "value": value, "digests": gathered,
                   "clients": len(wl["plan"])}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
