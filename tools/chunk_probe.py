"""Surface-chunk sharding over NCCL (include/iss_cuda.h): every rank holds
1/N of the cells of ONE surface and samples, for the same events, the hadrons of its cells.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/chunk_probe.py --cells 1000000 --events 1000

Checks, on every run: species totals identical on all ranks and identical to a whole-surface run of
the same handle; sum over ranks of the hadron counts and of the additive QA entries (all-reduced
with NCCL) equal to the whole-surface run.  Prints one JSON line with the device times of the
phases (strong scaling: the total work is fixed, the cells are divided)."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1000000)
    ap.add_argument("--events", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--no-whole", action="store_true", help="skip the whole-surface comparison run")
    ap.add_argument("--collectives", choices=["abi", "torch"], default="abi",
                    help="abi: iss_cuda_chunk_yields_allgather / iss_cuda_histograms_allreduce on the "
                         "handle's stream; torch: torch.distributed around the three-step entry points")
    ap.add_argument("--balance", action="store_true",
                    help="re-cut the chunks with iss_cuda_chunk_block_yields after a first pass")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from iss_b200 import capi, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    os.environ["ISS_CUDA_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    work = tempfile.mkdtemp(prefix="iss_chunk_r%d_" % rank)
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
        ns = e.nspecies
        pids = np.asarray([211, -211, 321, -321, 2212, -2212, 3122, 111], dtype=np.int32)
        qa_n = int(capi.cuda_lib().iss_cuda_qa_size())
        add = np.r_[9:29]

        def ev(name=None):
            t = torch.cuda.Event(enable_timing=True)
            t.record(stream)
            return t

        whole = None
        if not args.no_whole:
            e.upload_surface(lrf)
            dN_w = e.compute_yields().copy()
            c = e.sample(args.seed, 0, args.events)
            qa_w = e.histograms(pids)
            whole = dict(dN=dN_w, hadrons=int(c.n_hadrons), tries=int(c.n_tries),
                         redraws=int(c.n_cell_redraws), qa=qa_w.copy())

        dev = torch.device("cuda", local)
        abi = args.collectives == "abi"
        if abi and world > 1 and not sharding.join_engine_communicator(e):
            raise SystemExit("chunk_probe: no NCCL communicator for the handle")
        ranges = sharding.split_cells(len(lrf), world)
        if args.balance:
            # first pass with the even cut: block yields of the whole surface, known on every rank
            b, en = ranges[rank]
            e.upload_surface(lrf[b:en])
            e.set_surface_chunk(b, len(lrf))
            if abi:
                e.chunk_yields_allgather([sharding.ntiles_of(r) for r in ranges])
            else:
                sharding.chunk_yields(e, lrf, rank, world, dev)
            cost = (5.0e-6*sharding.CHUNK_ALIGN
                    + 2.5e-7*args.events*e.chunk_block_yields(len(lrf)))
            ranges = sharding.split_cells_weighted(cost, len(lrf), world)
        b, en = ranges[rank]
        ntiles = [sharding.ntiles_of(r) for r in ranges]
        e.upload_surface(lrf[b:en])
        e.set_surface_chunk(b, len(lrf))
        if abi:
            ms = {"yields_allgather_finish": [], "sample": [], "qa_allreduce": []}
        else:
            ms = {"yields_local": [], "allgather": [], "finish": [], "sample": [], "qa_allreduce": []}
        hadrons_local = 0
        for step in range(args.steps + 1):          # first pass = warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if abi:
                t0 = ev()
                dN = e.chunk_yields_allgather(ntiles)
                t3 = ev()
                marks = [(t0, t3)]
            else:
                t0 = ev()
                ptr, nt = e.chunk_yields_local()
                t1 = ev()
                local_t = sharding.device_block_as_tensor(ptr, ns*nt, dev).view(ns, nt)
                blocks = sharding.gather_tile_sums(local_t, ranges, ns)
                t2 = ev()
                dN = e.chunk_yields_finish([t.data_ptr() for t in blocks], [t.shape[1] for t in blocks],
                                           on_device=True)
                t3 = ev()
                marks = [(t0, t1), (t1, t2), (t2, t3)]
            c = e.sample(args.seed, 0, args.events)
            t4 = ev()
            e.L.iss_cuda_histograms(e.h, capi._ptr(pids), len(pids), 0)
            qa_t = sharding.device_block_as_tensor(e.qa_device_ptr(), qa_n, dev)
            if abi:
                e.check(e.L.iss_cuda_histograms_allreduce(e.h, None), "histograms_allreduce")
            else:
                sharding.allreduce_sum_(qa_t)
            t5 = ev()
            torch.cuda.synchronize()
            if step > 0:
                for k, (a, z) in zip(ms, marks + [(t3, t4), (t4, t5)]):
                    ms[k].append(a.elapsed_time(z))
            hadrons_local = int(c.n_hadrons)
        per_rank = torch.zeros(world, 3, dtype=torch.float64, device=dev)
        per_rank[rank, 0] = en - b
        per_rank[rank, 1] = hadrons_local
        per_rank[rank, 2] = float(np.mean(ms["sample"]))
        if world > 1:
            dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        qa_sum = qa_t.cpu().numpy().copy()
        tot = torch.tensor([float(hadrons_local), float(c.n_tries), float(c.n_cell_redraws)],
                           dtype=torch.float64, device=dev)
        step_ms = torch.tensor([sum(float(np.mean(v)) for v in ms.values())], dtype=torch.float64,
                               device=dev)
        dn_min = torch.from_numpy(dN).to(dev)
        dn_max = dn_min.clone()
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(dn_min, op=dist.ReduceOp.MIN)
            dist.all_reduce(dn_max, op=dist.ReduceOp.MAX)
        checks = {"totals_identical_on_all_ranks": bool(torch.equal(dn_min, dn_max))}
        if whole is not None:
            checks["totals_equal_whole_surface_run"] = bool(np.array_equal(dN, whole["dN"]))
            checks["hadrons_sum_equals_whole"] = int(tot[0].item()) == whole["hadrons"]
            checks["tries_sum_equals_whole"] = int(tot[1].item()) == whole["tries"]
            checks["qa_additive_equal_whole"] = bool(np.allclose(qa_sum[add], whole["qa"][add],
                                                                 rtol=1e-11, atol=1e-9))
        if abi and world > 1:
            e.L.iss_cuda_nccl_finalize(e.h)
        s.close()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        shutil.rmtree(work, ignore_errors=True)
    if rank == 0:
        line = {"probe": "surface_chunk_sharding", "n_gpus": world, "cells": len(lrf),
                "cells_rank0": en - b, "species": ns, "events": args.events,
                "collectives": "C ABI (ncclAllGather / ncclAllReduce on the handle's stream)" if abi
                else "torch.distributed", "cut": "balanced by block yields" if args.balance else "even",
                "per_rank": {"cells": [int(x) for x in per_rank[:, 0].tolist()],
                             "hadrons": [int(x) for x in per_rank[:, 1].tolist()],
                             "sample_ms": [round(float(x), 3) for x in per_rank[:, 2].tolist()]},
                "hadrons_all_ranks": int(tot[0].item()), "cell_redraws": int(tot[2].item()),
                "phase_ms_rank0": {k: float(np.mean(v)) for k, v in ms.items()},
                "step_ms_max_over_ranks": float(step_ms.item()),
                "hadrons_per_sec": float(tot[0].item())/(float(step_ms.item())*1e-3),
                "scaling": "strong", "checks": checks}
        print(json.dumps(line), flush=True)
        if not all(checks.values()):
            raise SystemExit("chunk_probe: a consistency check failed: %r" % checks)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
