"""Worker for tests/test_gpu_multi.py: BASELINE configs[4]'s data path on `world` GPUs.
Rank 0 owns the wideband stream and NCCL-broadcasts every block; every rank decimates ITS
clients (client c -> rank c mod world, bench.shard_clients) straight from the NCCL receive
buffer (xlg_wait_stream + XLG_INPUT_DEVICE) and compares every one of them with the oracle.
usage (under torchrun): _nccl_cfg5_worker.py <out_dir> [backend]"""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (checker)
from util import assert_cf32_close, oracle_stream  # noqa: E402


def main():
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("sdr-server_b200")
    fs, block_elems, n_clients, n_blocks = 61440000, 131072, 48, 6
    taps = pkg.create_low_pass_filter(1.0, fs, 24000, 9600)  # 15419 taps: the long-filter kernel
    centers = [int(-30000000 + c * 1200000) for c in range(n_clients)]
    mine = bench.shard_clients(n_clients, rank, world)
    g = pkg.Group(fs, block_elems, device=local)
    ids = [g.add_client(1280, taps, centers[c]) for c in mine]
    oracles = [po.OracleFilter(1280, taps, centers[c], fs, block_elems) for c in mine]
    blocks = bench.synth_blocks("cs16", n_blocks, block_elems, seed=bench.stream_seed(0))  # same on every rank (checker)
    refs = oracle_stream(oracles, "cs16", list(blocks))
    stream = torch.cuda.current_stream()
    ring = [torch.empty(blocks[0].nbytes, dtype=torch.uint8, device="cuda") for _ in range(pkg.XLG_SLOTS)]
    worst = 0.0
    for b in range(n_blocks):
        buf = ring[b % len(ring)]
        if rank == 0:
            buf.copy_(torch.from_numpy(blocks[b].view(np.uint8)), non_blocking=True)
        dist.broadcast(buf, src=0)
        g.wait_stream(stream.cuda_stream)
        t = g.submit_ptr(pkg.FMT["cs16"], buf.data_ptr(), block_elems, pkg.XLG_INPUT_DEVICE)
        g.wait(t)
        for i, cid in enumerate(ids):
            worst = max(worst, assert_cf32_close(g.output(t, cid), refs[i][b], f"rank {rank} client {mine[i]} block {b}"))
    kinds = sorted({g.client_info(c)[1] for c in ids})
    g.close()
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "clients": mine, "worst": worst, "kinds": kinds}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
