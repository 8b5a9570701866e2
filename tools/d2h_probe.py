"""Concurrent device->host bandwidth of the box: what bounds iSS::generate_samples() end to end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29541 tools/d2h_probe.py [--gb 2.0]

Every rank copies a `--gb` GB device buffer into pinned host memory (the size of one C4 step's
hadron list, 2.2 GB), all ranks at the same time, five timed repetitions after one warm-up.
Rank 0 prints one JSON line: per-rank and aggregate GB/s (time = max over ranks per repetition)."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(args.gb*1e9)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    dev.fill_(1)
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(dev)                              # warm-up (touches every host page)
    torch.cuda.synchronize()
    mine, agg = [], []
    for _ in range(args.reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        host.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mine.append(n/dt/1e9)
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        agg.append(world*n/float(t.item())/1e9)
    per_rank = torch.tensor([max(mine)], dtype=torch.float64, device="cuda")
    lo, hi = per_rank.clone(), per_rank.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"probe": "concurrent_d2h", "n_gpus": world, "gb_per_rank": args.gb,
                          "aggregate_gbs_best": max(agg), "aggregate_gbs_median": sorted(agg)[len(agg)//2],
                          "per_rank_gbs_best_min": float(lo.item()), "per_rank_gbs_best_max": float(hi.item()),
                          "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
