#!/usr/bin/env python3
"""Benchmark of the Cooper-Frye particlization hot path (BASELINE.json metric: sampled hadrons/s
and cell x species yields/s) on N B200s of one node, next to the reference's CPU sampler.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path

Workload (config.workload): BASELINE.json configs[3], "C4": synthetic 3+1D MUSIC-format surface,
10^6 cells (binary, 34 float32 per cell), EOS 14 with net baryon density and baryon diffusion,
Chapman-Enskog delta-f (kind 21) for shear, bulk and diffusion, urqmd_v3.3+ list (321 species).
A "step" is one pass of the hot path over one batch: yields for all cells x species, cell CDFs,
multiplicities, offsets, momentum sampling fused with boost/emit for EVENTS_PER_STEP events, QA
histograms (reduced over ranks with NCCL when N > 1).  The real job computes the yields once per
10^4 events; recomputing them every 1000 events makes the step a conservative 1/10 of C4.

Event sharding: every rank holds the whole surface and samples its own
events, no data-path collective; per-GPU work is fixed -> "scaling": "weak".

Timing: CUDA events on the stream the kernels run on (the handle is bound to torch's current
stream), W >= 3 warm-up steps, barrier + synchronize on both sides, max over ranks.  The per-step
working set (2 x 2.6 GB of yields/CDF + 2.3 GB of hadrons) is far larger than the 126 MB L2.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

PARAM = os.path.join(REPO, "tests", "fixtures", "iSS_parameters_CEdeltaf.dat")
OVERRIDES = dict(afterburner_type=1, include_deltaf_diffusion=1, include_deltaf_shear=1,
                 include_deltaf_bulk=1, bulk_deltaf_kind=21, hydro_mode=2, perform_decays=0,
                 use_OSCAR_format=0, use_gzip_format=0, use_binary_format=0, perform_checks=0,
                 MC_sampling=4, local_charge_conservation=0, sample_upto_desired_particle_number=0)
SURFACE_SEED = 2024          # C4 seed
# BASELINE.json configs: the default (and the configuration the metric is quoted on) is C4; the
# others are selectable with --workload for measurements that are not the headline line.
#   c3        configs[2]: 2+1D boost-invariant surface, 10^5 cells, SMASH list, CE delta-f, decays
#             requested -- FSSW::shell skips the feed-down for SMASH (FSSW.cpp:346), so no decays run
#   c3-decays the same surface with the UrQMD list, for which the reference does decay: times
#             decay_kernel (roofline_decay)
#   c5        configs[4]: C4's generator at 10^7 cells, 100 events per step
WORKLOADS = {
    "c4": dict(cells=1000000, events=1000, gen=dict(eos=14, rhob=1, diffusion=1, binary=1), over={},
               label="C4 (BASELINE.json configs[3]): synthetic 3+1D MUSIC-format surface, %d cells, EOS 14 + "
                     "rho_B + baryon diffusion, CE delta-f shear+bulk+diffusion, urqmd_v3.3+ list (321 "
                     "species)"),
    "c3": dict(cells=100000, events=1000, gen=dict(eos=91, boost_invariant=True, binary=0),
               over=dict(hydro_mode=1, include_deltaf_diffusion=0, perform_decays=1, y_LB=-2.5, y_RB=2.5),
               label="C3 (BASELINE.json configs[2]): synthetic 2+1D boost-invariant surface, %d cells, EOS 91, "
                     "SMASH list (400 species), CE delta-f shear+bulk, |y| < 2.5, perform_decays = 1 (skipped "
                     "for SMASH like FSSW::shell does)"),
    "c3-decays": dict(cells=100000, events=1000, gen=dict(eos=9, boost_invariant=True, binary=0),
                      over=dict(hydro_mode=1, include_deltaf_diffusion=0, perform_decays=1, y_LB=-2.5,
                                y_RB=2.5),
                      label="C3 surface with the UrQMD list: synthetic 2+1D boost-invariant surface, %d cells, "
                            "EOS 9, urqmd_v3.3+ list (321 species), CE delta-f shear+bulk, |y| < 2.5, "
                            "resonance decays on"),
    "c5": dict(cells=10000000, events=100, gen=dict(eos=14, rhob=1, diffusion=1, binary=1), over={},
               label="C5 (BASELINE.json configs[4]): C4's generator at %d cells, EOS 14 + rho_B + baryon "
                     "diffusion, CE delta-f shear+bulk+diffusion, urqmd_v3.3+ list (321 species)"),
}
BYTES_PER_DECAY_IN = 40.0    # one primary record in; out: 40 B per final hadron (measured ratio)
FLOP_PER_DECAY = 250.0
# algorithmic work per unit (restated in DESIGN.md)
BYTES_PER_HADRON = 152.0     # 40 B record out + 112 B cell record in
FLOP_PER_HADRON = 1.0e3
BYTES_PER_YIELD = 8.2        # 8 B FP64 yield out + 64 B of cell fields / 321 species
FLOP_PER_YIELD = 175.0       # CE bulk + diffusion series
NCU_PROPOSE_TRAFFIC_BYTES = 8.14e9    # 4.65 GB read + 3.48 GB written per launch (profiles/r2_ncu_sampler_kernels.json)


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md): ONE
    long-lived `nvidia-smi -lms 200` for the GPUs of the job, started by rank 0 before the warm-up
    steps and stopped after the timed region (a process per sample and per rank perturbs the CUDA
    calls of an 8-rank run)."""

    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        self.indices = [int(i) for i in indices]
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="iss_clocks_", suffix=".csv")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is not None and self.proc.poll() is None:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=5)
            except Exception:
                try:
                    self.proc.kill()
                except Exception:
                    pass

    @staticmethod
    def _epoch(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def summary(self, t_begin=None, t_end=None):
        """samples inside [t_begin, t_end] (epoch seconds of the load: warm-up + timed steps)"""
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=5)
            except Exception:
                try:
                    self.proc.kill()
                except Exception:
                    pass
            try:
                rows = [[x.strip() for x in ln.split(",")] for ln in open(self.path) if ln.strip()]
                os.remove(self.path)
            except Exception:
                rows = []
        rows = [r for r in rows if len(r) >= 9]
        if t_begin is not None and t_end is not None:
            inside = []
            for r in rows:
                t = self._epoch(r[0])
                if t is not None and t_begin - 0.05 <= t <= t_end + 0.05:
                    inside.append(r)
            if inside:
                rows = inside
        rows = [r[1:] for r in rows]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(rows), "gpus": self.indices}


def bind_to_gpu_numa_node(local):
    """Pins this rank (and the threads/pinned buffers it creates later) to the NUMA node of its GPU:
    with one process per GPU, the 2.2 GB/step device->host hadron stream otherwise crosses the
    socket interconnect for half of the ranks.  Returns a short description for the JSON line."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return "numa node unknown for %s" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no allowed cpus on node %d" % node
        os.sched_setaffinity(0, cpus)
        return "gpu %s -> numa node %d (%d cpus)" % (bdf, node, len(cpus))
    except Exception as exc:        # placement is an optimisation, never a failure
        return "not bound (%s)" % type(exc).__name__


def make_case(folder, ncell, workload="c4"):
    from iss_b200 import synthetic
    synthetic.make_case(folder, ncell=ncell, seed=SURFACE_SEED, **WORKLOADS[workload]["gen"])


def overrides_of(workload):
    return dict(OVERRIDES, **WORKLOADS[workload]["over"])


def d2h_link_probe(torch, dist, world, barrier, nbytes=1 << 29, reps=3):
    """Pinned device->host bandwidth of this rank's GPU with all ranks copying at the same time (the
    ceiling of the end-to-end figure), and the PCIe link state nvidia-smi reports."""
    out = {}
    try:
        dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        dev.fill_(1)
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.copy_(dev)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            host.copy_(dev, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, nbytes/(time.perf_counter() - t0)/1e9)
        t = torch.tensor([best], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out["gbs_per_rank_all_ranks_busy"] = float(t.item())
        del dev, host
    except Exception as exc:
        out["error"] = "%s: %s" % (type(exc).__name__, exc)
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current",
                            "--format=csv,noheader", "-i", str(torch.cuda.current_device())],
                           capture_output=True, text=True, timeout=10).stdout.strip()
        out["pcie_gen_width"] = q
    except Exception:
        pass
    return out


def workload_name(ncell, events, workload="c4"):
    return (WORKLOADS[workload]["label"] % ncell
            + "; step = yields of all cells x species + CDF + multiplicities + %d sampled events"
              "%s (+ QA histograms)" % (events, " + resonance decays" if workload == "c3-decays" else ""))


def spectra_leg(cells=5000):
    """Secondary measurement, N = 1 only: the smooth-spectra mode of
    the facade (MC_sampling = 0, calculate_vn = 1) on a `cells`-cell surface of the same generator,
    all species of the list, shear + bulk (kind 1) + diffusion delta f.  Not part of `value`."""
    from iss_b200 import capi
    work = tempfile.mkdtemp(prefix="iss_bench_spectra_")
    try:
        make_case(work, cells)
        over = dict(OVERRIDES, MC_sampling=0, calculate_vn=1, bulk_deltaf_kind=1,
                    calculate_vn_to_order=4)
        s = capi.Sampler(work, PARAM, "surface.dat", **over)
        s.read_in_FO_surface()
        t0 = time.perf_counter()
        s.generate_samples()
        wall = time.perf_counter() - t0
        _, kernel_ms, evals = s.spectra_table(211)
        s.close()
        return {"what": "iSS::generate_samples() with MC_sampling=0, calculate_vn=1: dN/(pT dpT dphi dy) "
                        "and v_n of every species (EmissionFunctionArray::calculate_dN_pTdpTdphidy)",
                "cells": cells, "points": evals, "kernel_ms": kernel_ms,
                "points_per_sec": evals/(kernel_ms*1e-3) if kernel_ms > 0 else None,
                "wall_ms_generate_samples": 1e3*wall, "bound": "fp64",
                "note": "ncu of the shear-only variant (profiles/r1_ncu_spectra_qa.json): 37.3 FP64 "
                        "instructions of 68.7 per point, FP64 pipe 67.6 % of peak; reference CPU "
                        "5.2e7 points/s per core (profiles/r1_spectra_probe.txt)"}
    finally:
        shutil.rmtree(work, ignore_errors=True)


# ------------------------------------------------------------------------------------ engine arm
def run_engine(args):
    import torch
    import torch.distributed as dist
    from iss_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    # Everything below that writes to file descriptor 1 (the facade logs like the reference, NCCL
    # prints its version) goes to stderr; the real stdout is kept for the single JSON line.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    os.environ["ISS_CUDA_DEVICE"] = str(local)
    numa = bind_to_gpu_numa_node(local) if os.environ.get("ISS_BENCH_NUMA", "1") == "1" else "off"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    E = args.events_per_step
    # one nvidia-smi loop for all GPUs of the job, started early so that it samples at full cadence
    # when the load begins; only the samples of the load window (warm-up + timed steps) are kept
    clocks = ClockSampler(range(world)) if rank == 0 else None
    if clocks:
        clocks.start()
    work = tempfile.mkdtemp(prefix="iss_bench_r%d_" % rank)
    make_case(work, args.cells, args.workload)
    decays_on = args.workload == "c3-decays"
    try:
        over = dict(overrides_of(args.workload), number_of_repeated_sampling=E)
        s = capi.Sampler(work, PARAM, "surface.dat", **over)
        s.read_in_FO_surface()
        s.set_random_seed(args.seed)
        s.prepare_sampler()
        e = s.engine()
        stream = torch.cuda.current_stream()
        e.set_stream(stream.cuda_stream)
        ncell, ns = e.ncell, e.nspecies
        qa_pids = [211, -211, 321, -321, 2212, -2212, 3122, 111]

        from iss_b200 import sharding
        qa_n = int(capi.cuda_lib().iss_cuda_qa_size())

        # the one collective of the path goes through the C ABI (iss_cuda_histograms_allreduce on a
        # communicator the handle owns); torch.distributed only carries the NCCL id to the ranks
        qa_via_c_abi = world > 1 and sharding.join_engine_communicator(e)
        step_counter = [0]

        def step():
            k = step_counter[0]
            step_counter[0] += 1
            e.compute_yields()
            ev0, ev1 = sharding.weak_event_range(k, rank, world, E)
            c = e.sample(args.seed, ev0, ev1)
            n_primary = c.n_hadrons
            if decays_on:
                c2 = e.decay(args.seed)
                decay_counts[0] += n_primary
                decay_counts[1] += c2.n_hadrons
            e.L.iss_cuda_histograms(e.h, capi._ptr(np.asarray(qa_pids, dtype=np.int32)),
                                    len(qa_pids), 0)
            if world > 1:       # NCCL: QA histograms only
                if qa_via_c_abi:
                    e.check(e.L.iss_cuda_histograms_allreduce(e.h, None), "iss_cuda_histograms_allreduce")
                else:
                    sharding.allreduce_sum_(sharding.device_block_as_tensor(e.qa_device_ptr(), qa_n, "cuda"))
            return n_primary, c.n_tries

        decay_counts = [0, 0]       # primaries in, final hadrons out (timed steps only)
        t_load_begin = time.time()
        for _ in range(args.warmup):
            step()
        e.timing(enable=True, reset=True)
        decay_counts[0] = decay_counts[1] = 0
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        hadrons = tries = 0
        trace = os.environ.get("ISS_BENCH_TRACE") == "1"    # per-step host wall times on stderr
        for _ in range(args.steps):
            w_a = time.perf_counter()
            h, t = step()
            if trace:
                sys.stderr.write("[bench trace] rank %d step wall %.2f ms\n"
                                 % (rank, 1e3*(time.perf_counter() - w_a)))
            hadrons += h
            tries += t
        t1.record(stream)
        barrier()
        ms = t0.elapsed_time(t1)
        fam_ms, fam_n = e.timing(enable=False)
        clk = clocks.summary(t_load_begin, time.time()) if clocks else None
        # per-rank step time (a slow rank shows here; `ms_per_step` is the max)
        tr = torch.zeros(world, dtype=torch.float64, device="cuda")
        tr[rank] = ms/args.steps
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.SUM)
        ms_per_rank = [round(float(x), 3) for x in tr.tolist()]
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
        th = torch.tensor([float(hadrons)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(th, op=dist.ReduceOp.SUM)
        ms_max, hadrons_all = float(tm.item()), float(th.item())

        if qa_via_c_abi:
            e.L.iss_cuda_nccl_finalize(e.h)
        fp64_peak = e.fp64_peak()
        # (generate_samples() below builds a new device sampler: `e` must not be used after it)

        # ---- end to end through the reference-facing call: class iSS::generate_samples() with the
        # surface in host memory (std::vector<FO_surf_LRF>): H2D of surface and tables, yields,
        # sampling of E events, D2H of the hadron lists into the pinned host buffer.
        s.set_param("number_of_repeated_sampling", E)
        # (a generate_samples() call occasionally takes ~45 ms longer on the shared boxes: five timed
        # calls and two warm-up calls keep one such call from dominating the figure)
        e2e_steps = max(1, min(args.steps, 5))
        # same seed on every rank, disjoint event indices: together the ranks produce the events
        # one process would produce for the whole range
        s.set_param("first_event_index", rank*E)
        s.set_random_seed(args.seed)
        s.generate_samples()                         # warm-up (allocations, pinned buffer)
        s.generate_samples()
        barrier()
        w0 = time.perf_counter()
        e2e_hadrons = 0
        for _ in range(e2e_steps):
            s.generate_samples()
            h_all, off = s.hadrons()
            e2e_hadrons += len(h_all)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        barrier()
        te = torch.tensor([w1 - w0], dtype=torch.float64, device="cuda")
        the = torch.tensor([float(e2e_hadrons)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(the, op=dist.ReduceOp.SUM)
        e2e_value = float(the.item())/float(te.item())
        table_bytes = 200*200*5*8 + 150*100*8 + 7991*12*8
        h2d = ncell*28*4 + table_bytes
        d2h = int(e2e_hadrons/e2e_steps)*40 + (E + 1)*8
        s.close()
        link = d2h_link_probe(torch, dist, world, barrier)
        spectra = None
        if world == 1 and not args.no_spectra:
            try:
                spectra = spectra_leg()
            except Exception as exc:        # secondary figure: never fails the headline line
                spectra = {"error": "%s: %s" % (type(exc).__name__, exc)}
    finally:
        if clocks:
            clocks.stop()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        shutil.rmtree(work, ignore_errors=True)

    peaks, peak_src = load_peaks()
    launches = int(sum(fam_n.values()))
    dom = max(fam_ms, key=fam_ms.get)
    sample_s = fam_ms["sample"]*1e-3
    yields_s = fam_ms["yields"]*1e-3
    n_launch_sample = max(1, int(fam_n["sample"]))
    roof = {
        "kernel": "propose_kernel (persistent momentum sampler fused with boost/emit, tasks streamed with "
                  "cp.async.bulk); the set-up and partition kernels that feed it are timed separately "
                  "(kernel_ms.setup)",
        "bound": "hbm",
        "achieved": BYTES_PER_HADRON*hadrons/sample_s/1e9 if sample_s > 0 else None,
        "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": (BYTES_PER_HADRON*hadrons/sample_s/1e9/peaks["hbm_gbs"]) if sample_s > 0 else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch on this workload, from the
        # ncu --set full capture summarised in profiles/ (static evidence, not measured here)
        "traffic": NCU_PROPOSE_TRAFFIC_BYTES if (args.cells == 1000000 and E == 1000) else None,
        "traffic_source": "profiles/r2_ncu_sampler_kernels.json (ncu --set full, one launch of this workload)",
        "algorithmic_bytes_per_launch": BYTES_PER_HADRON*hadrons/max(1, n_launch_sample),
        "peak_source": peak_src,
        "avg_launch_ms": fam_ms["sample"]/n_launch_sample,
        "share_of_step": fam_ms["sample"]/ms if ms > 0 else None,
        "note": "the sampler is bound by the shared-memory data pipe (78 %: the random 8-byte table loads "
                "of the bisection), instruction issue (69 % of the slots) and divergence (22.6 of 32 lanes), "
                "not by HBM (SURVEY.md 8(d)); see fp64 and profiles/r2_ncu_propose_source_final.txt",
        "fp64": {"achieved_tflops": FLOP_PER_HADRON*hadrons/sample_s/1e12 if sample_s > 0 else None,
                 "peak_tflops": fp64_peak, "peak_source": "DFMA microbenchmark run in this process "
                 "(iss_cuda_fp64_peak)",
                 "frac": FLOP_PER_HADRON*hadrons/sample_s/1e12/fp64_peak if sample_s > 0 else None,
                 "tries_per_hadron": tries/max(1, hadrons)},
    }
    ycs = float(ncell)*ns*args.steps
    roof_y = {
        "kernel": "yields_kernel", "bound": "hbm",
        "achieved": BYTES_PER_YIELD*ycs/yields_s/1e9 if yields_s > 0 else None,
        "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": BYTES_PER_YIELD*ycs/yields_s/1e9/peaks["hbm_gbs"] if yields_s > 0 else None,
        "avg_launch_ms": fam_ms["yields"]/max(1, args.steps),
        "fp64": {"achieved_tflops": FLOP_PER_YIELD*ycs/yields_s/1e12 if yields_s > 0 else None,
                 "peak_tflops": fp64_peak,
                 "frac": FLOP_PER_YIELD*ycs/yields_s/1e12/fp64_peak if yields_s > 0 else None},
    }
    line = {
        "metric": "sampled_hadrons_per_sec", "value": hadrons_all/(ms_max*1e-3),
        "unit": "hadrons/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max/args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.cells, E, args.workload), "workload_key": args.workload,
                   "cells": ncell, "species": ns,
                   "events_per_step_per_gpu": E, "sharding": "events (weak), surface replicated",
                   "qa_allreduce": ("none (1 rank)" if world == 1 else
                                    "iss_cuda_histograms_allreduce (C ABI, NCCL)" if qa_via_c_abi
                                    else "torch.distributed all_reduce (NCCL)"),
                   "l2": "inputs_larger_than_L2", "seed": args.seed, "host_placement": numa},
        "yields_per_sec": ycs/yields_s if yields_s > 0 else None,
        "sampler_hadrons_per_sec_kernel": hadrons/sample_s if sample_s > 0 else None,
        "kernel_ms": {k: v/args.steps for k, v in fam_ms.items()},
        "dominant_kernel": dom,
        "roofline": roof, "roofline_yields": roof_y,
        "e2e": {"value": e2e_value, "unit": "hadrons/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "call": "iSS::generate_samples() (class iSS facade, host surface -> host hadron lists)",
                # what bounds this figure: the device->host copy of the 40-byte records of a step over
                # THIS box's link (measured right after the timed calls; the boxes of the pool differ)
                "d2h_link": dict(link, floor_ms_per_call=(d2h/(link["gbs_per_rank_all_ranks_busy"]*1e9)*1e3
                                                          if link.get("gbs_per_rank_all_ranks_busy") else None),
                                 measured_ms_per_call=float(te.item())/e2e_steps*1e3)},
        "gpu_launches": launches,
        "ms_per_step_per_rank": ms_per_rank,
        "clocks": clk,
        "spectra": spectra,
    }
    if decays_on and fam_ms["decay"] > 0 and decay_counts[0] > 0:
        dec_s = fam_ms["decay"]*1e-3
        per_primary = BYTES_PER_DECAY_IN + 40.0*decay_counts[1]/decay_counts[0]
        line["roofline_decay"] = {
            "kernel": "decay_kernel (count pass + write pass over every primary's decay tree) + offsets scan",
            "bound": "hbm", "achieved": per_primary*decay_counts[0]/dec_s/1e9, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": per_primary*decay_counts[0]/dec_s/1e9/peaks["hbm_gbs"],
            "algorithmic_bytes_per_primary": per_primary,
            "final_hadrons_per_primary": decay_counts[1]/decay_counts[0],
            "avg_ms_per_step": fam_ms["decay"]/args.steps,
            "primaries_per_sec": decay_counts[0]/dec_s,
            "fp64": {"achieved_tflops": FLOP_PER_DECAY*decay_counts[1]/dec_s/1e12, "peak_tflops": fp64_peak,
                     "frac": FLOP_PER_DECAY*decay_counts[1]/dec_s/1e12/fp64_peak},
            "note": "latency and divergence bound (18 of 32 lanes): one thread walks one primary's decay tree, first as a tree of species only (count pass: channel picks and the three-body energy rejection, the other random numbers skipped), then with the kinematics (write pass)"}
    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(args, quick=True)
        line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------- reference arm
def ref_binary():
    p = os.path.join(REPO, "oracle", "_ref", "iSS.e")
    return p if os.path.exists(p) else None


def cpu_reference(args, quick):
    """Times the UNMODIFIED reference binary (oracle/_ref/iSS.e, compiled from /root/reference by
    oracle/Makefile) on a bounded sample of the bench workload: a surface of args.cells/50 cells
    generated with the same generator and seed, and the same events-per-step, so that the ratio
    of yield work to sampling work per hadron is that of the GPU step.  The reference is serial:
    `cores` independent processes run side by side with distinct seeds.  Throughput of one
    process = hadrons / the reference's own timer line "sample_using_dN_dxtdy_4all_particles
    finished in X seconds" (FSSW.cpp:1067-1070, CPU clock(): yields + sampling, no file I/O)."""
    exe = ref_binary()
    if exe is None:
        return {"unavailable": "oracle/_ref/iSS.e not built (needs /root/reference at build time)"}
    full = bool(getattr(args, "full_size", False))
    cores = 1 if full else (os.cpu_count() or 1)
    ncell = args.cells if full else max(1000, args.cells//50)
    events = min(100, args.events_per_step) if full else args.events_per_step
    root = tempfile.mkdtemp(prefix="iss_ref_")
    try:
        make_case(os.path.join(root, "case"), ncell, args.workload)
        os.symlink(os.path.join(REPO, "iSS_tables"), os.path.join(root, "iSS_tables"))
        over = dict(overrides_of(args.workload), number_of_repeated_sampling=events, use_binary_format=0)
        procs = []
        t0 = time.perf_counter()
        for c in range(cores):
            d = os.path.join(root, "p%d" % c)
            os.makedirs(d)
            os.symlink(os.path.join(root, "iSS_tables"), os.path.join(d, "iSS_tables"))
            os.symlink(os.path.join(root, "case"), os.path.join(d, "case"))
            cmd = [exe, PARAM, "case", "surface.dat"] + ["%s=%s" % kv for kv in over.items()]
            cmd.append("randomSeed=%d" % (c + 1))
            procs.append(subprocess.Popen(cmd, cwd=d, stdout=open(os.path.join(d, "log"), "w"),
                                          stderr=subprocess.STDOUT))
        for p in procs:
            p.wait()
        wall = time.perf_counter() - t0
        rate = 0.0
        hadrons = 0.0
        secs = []
        for c in range(cores):
            txt = open(os.path.join(root, "p%d" % c, "log")).read()
            sec = None
            dN = 0.0
            for ln in txt.splitlines():
                if "finished in" in ln and "sample_using_dN_dxtdy_4all_particles" in ln:
                    sec = float(ln.split("finished in")[1].split()[0])
                if "Sampling using dN_dy=" in ln:
                    dN += float(ln.split("dN=")[1].split("...")[0])
            if sec is None or sec <= 0:
                return {"unavailable": "reference run failed: " + txt[-300:].replace("\n", " | ")}
            n = dN*events        # expected hadrons (Poisson mean); the reference does not print counts
            rate += n/sec
            hadrons += n
            secs.append(sec)
        out = {"value": rate, "unit": "hadrons/s", "cores": cores, "kind": "reference",
               "sample": "%d-cell surface (1/%d of the workload, same generator and seed), %d events, "
                         "%d independent single-thread processes of oracle/_ref/iSS.e; hadrons = "
                         "events x sum of species dN; time = reference timer line (yields + sampling)"
                         % (ncell, args.cells//ncell, events, cores),
               "per_core": rate/cores, "per_box": rate, "cpu_seconds_per_process": float(np.mean(secs)),
               "wall_seconds": wall}
        # one recorded run of the reference on the FULL-SIZE surface (bench.py --impl reference
        # --full-size, minutes on one core) calibrates the 1/50 sample
        chk = os.path.join(REPO, "profiles", "r2_reference_fullsize_%s.json" % args.workload)
        if not full and os.path.exists(chk):
            try:
                out["full_size_check"] = json.load(open(chk))
            except Exception:
                pass
        return out
    finally:
        shutil.rmtree(root, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_reference(args, quick=True)
        if "unavailable" in cb:
            print(json.dumps({"impl": "reference", "unavailable": cb["unavailable"]}), flush=True)
            return
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": "sampled_hadrons_per_sec", "value": v, "unit": "hadrons/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.cells, args.events_per_step, args.workload),
                       "workload_key": args.workload},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "hadrons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--cells", type=int, default=None, help="default: the workload's size")
    ap.add_argument("--events-per-step", type=int, default=None, help="default: the workload's")
    ap.add_argument("--full-size", action="store_true",
                    help="--impl reference only: ONE single-thread reference run on the full-size "
                         "surface (calibrates the 1/50 sample; minutes)")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spectra", action="store_true")
    args = ap.parse_args()
    if args.cells is None:
        args.cells = WORKLOADS[args.workload]["cells"]
    if args.events_per_step is None:
        args.events_per_step = WORKLOADS[args.workload]["events"]
    if args.impl == "reference":
        # a reference "step" is a full multi-process run of the binary: keep the count small
        args.steps = max(1, min(args.steps, 2))
        args.warmup = min(args.warmup, 0)
        if args.full_size:
            args.steps = 1
        run_reference(args)
    else:
        args.warmup = max(3, args.warmup)
        run_engine(args)


if __name__ == "__main__":
    main()
