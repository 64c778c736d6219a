"""Worker for tests/test_multi_gpu_host.py: one process per (pretend) GPU over gloo."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    out_dir = sys.argv[1]
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
