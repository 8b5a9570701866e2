"""Surface-chunk sharding, the ranks of a `world`-rank run emulated one after the other on ONE GPU
(the kernels and entry points of the multi-GPU run; the all-gather goes through host memory).
Prints what the multi-GPU probe cannot show from rank 0 alone: cells, hadrons and device times of
the sampling phase of EVERY rank, for the even cut and for the cut balanced with
iss_cuda_chunk_block_yields.

    python tools/chunk_emulate.py --cells 1000000 --events 1000 --world 8
"""
import argparse
import json
import os
import shutil
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402

# cost model of the balanced cut (ms, one B200, C4 kernels): yields + prefix per cell, sampling per hadron
COST_PER_CELL = 5.0e-6
COST_PER_HADRON = 2.5e-7


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1000000)
    ap.add_argument("--events", type=int, default=1000)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=12345)
    args = ap.parse_args()

    import torch
    from iss_b200 import capi, sharding

    saved = os.dup(1)
    os.dup2(2, 1)
    work = tempfile.mkdtemp(prefix="iss_chunk_emu_")
    out = {"probe": "surface_chunk_emulation_one_gpu", "world": args.world, "events": args.events}
    try:
        bench.make_case(work, args.cells)
        s = capi.Sampler(work, bench.PARAM, "surface.dat",
                         **dict(bench.OVERRIDES, number_of_repeated_sampling=args.events))
        s.read_in_FO_surface()
        s.set_random_seed(args.seed)
        s.prepare_sampler()
        e = s.engine()
        stream = torch.cuda.current_stream()
        e.set_stream(stream.cuda_stream)
        lrf = s.lrf_surface().copy()
        out["cells"] = len(lrf)

        def ev():
            t = torch.cuda.Event(enable_timing=True)
            t.record(stream)
            return t

        def run(ranges):
            blocks = []
            for b, en in ranges:
                e.upload_surface(lrf[b:en])
                e.set_surface_chunk(b, len(lrf))
                blocks.append(e.chunk_tilesums_host())
            ntiles = [x.shape[1] for x in blocks]
            rows, block_yield = [], None
            for b, en in ranges:
                e.upload_surface(lrf[b:en])
                e.set_surface_chunk(b, len(lrf))
                t_y, t_f, t_s, kern = [], [], [], None
                for step in range(args.steps + 1):
                    torch.cuda.synchronize()
                    t0 = ev()
                    e.chunk_yields_local()
                    t1 = ev()
                    e.chunk_yields_finish(blocks, ntiles, on_device=False)
                    t2 = ev()
                    e.timing(True, reset=True)
                    c = e.sample(args.seed, 0, args.events)
                    t3 = ev()
                    torch.cuda.synchronize()
                    if step > 0:
                        t_y.append(t0.elapsed_time(t1))
                        t_f.append(t1.elapsed_time(t2))
                        t_s.append(t2.elapsed_time(t3))
                        kern = e.timing(True)[0]
                if block_yield is None:
                    block_yield = e.chunk_block_yields(len(lrf))
                rows.append({"cells": en - b, "hadrons": int(c.n_hadrons),
                             "yields_local_ms": float(np.mean(t_y)), "finish_ms": float(np.mean(t_f)),
                             "sample_ms": float(np.mean(t_s)),
                             "kernel_ms": {k: float(v) for k, v in kern.items() if v > 0}})
            return rows, block_yield

        even = sharding.split_cells(len(lrf), args.world)
        rows, by = run(even)
        out["even"] = {"ranks": rows,
                       "max_rank_ms": max(r["yields_local_ms"] + r["finish_ms"] + r["sample_ms"] for r in rows)}
        cost = COST_PER_CELL*sharding.CHUNK_ALIGN + COST_PER_HADRON*args.events*by
        bal = sharding.split_cells_weighted(cost, len(lrf), args.world)
        rows, _ = run(bal)
        out["balanced"] = {"ranks": rows, "ranges": bal,
                           "max_rank_ms": max(r["yields_local_ms"] + r["finish_ms"] + r["sample_ms"] for r in rows)}
        s.close()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
