#!/usr/bin/env python
"""bench.py -- batched ksw_extz2 throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU kernel (oracle/_ref)

Workload (config.workload): BASELINE.json configs[1] -- 100k synthetic 1 kbp pairs per GPU, 5 % divergence
(makeSmall event model), soft-masked, band w=100, SEDEF scoring, CIGAR + exact max (flag 0) + fused SD statistics.
A "step" is one pass of the hot path over the whole pair set: DP kernel + traceback/stats kernel.

value  = whole-job in-band GCUPS with inputs resident in HBM (device time by CUDA events, max over ranks).
e2e    = the same metric through the C ABI with HOST buffers (pack + H2D + kernels + D2H + gather/malloc of the
         CIGARs inside the timed region).
Multi-GPU: one process per GPU (torchrun), every rank aligns its own equal-work block of the N x 100k pair set
(weak scaling; pairs are independent, no collective on the data path).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

OPS_PER_CELL = 34            # integer lane-ops per in-band cell in the reference formulation
TB_BYTES_PER_CELL = 0.5      # 4-bit traceback code per cell


def load_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


def peak_int_tlaneops() -> tuple:
    """Measured single-pipe integer peak (IADD3 / VIMNMX lane-ops per second), profiles/r01_int_peak.json."""
    d = load_json(os.path.join(ROOT, "profiles", "r01_int_peak.json"))
    if d:
        r = d["results"]
        return min(r["iadd3_3in"]["glaneops_per_s"], r["vimnmx3"]["glaneops_per_s"]) / 1e3, "measured: profiles/r01_int_peak.json"
    return 148 * 64 * 1.965e9 / 1e12, "nominal 148 SMs x 64 lanes x 1.965 GHz"


def peak_hbm_gbs() -> tuple:
    d = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    if d and "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured: MEASURED_PEAKS.json"
    return 6650.0, "fallback: B200_PROFILING.md"


class ClockSampler:
    """Streams `nvidia-smi -lms 100` during the timed region (clocks + throttle reasons)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev: int):
        self.dev, self.proc = dev, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.35)           # let the first sample land before the timed region starts
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        rows = []
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill(); out = ""
            rows = [[x.strip() for x in ln.split(",")] for ln in out.strip().splitlines() if ln.strip()]
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if len(r) > 1 and num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if len(r) > 2 and num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        busy = [c for c, p_ in zip(sm, pw) if p_ is not None and p_ > 250] or sm
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, reasons=reasons, samples=len(rows))


def make_workload(n_pairs: int, rank: int):
    from sedef_b200 import synth
    # block `rank` of the global N x n_pairs set; all queries are 1 kbp so the blocks are equal-work
    return synth.make_pairs_small(n_pairs, length=1000, div=0.05, seed=0x5EDEF002 + 7919 * rank)


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    return rank, world, local, dist


def barrier_sync(dist, local):
    import torch
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(local)


def allreduce_max(dist, local, x: float) -> float:
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum(dist, local, x: float) -> float:
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_ours(args):
    import torch
    from sedef_b200 import engine, synth
    rank, world, local, dist = dist_setup(args.gpus)
    n_gpus = world
    mat = synth.sedef_matrix()
    W, ZD, FLAG = 100, -1, 0
    engine.init(local, 1)
    host_threads = max(1, (os.cpu_count() or 1) // max(1, world))   # torchrun exports OMP_NUM_THREADS=1; share the cores
    engine.set_host_threads(host_threads)
    torch.cuda.set_device(local)
    ps = make_workload(args.pairs, rank)

    # ---- device-resident arm --------------------------------------------------------------------
    rb = engine.ResidentBatch(ps, mat, synth.SEDEF_GAPO, synth.SEDEF_GAPE, W, ZD, FLAG)
    cells_rank = rb.cells()
    for _ in range(max(args.warmup, 3)):
        rb.run()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier_sync(dist, local)
    t0 = time.perf_counter()
    dev_ms = 0.0; dp_ms = 0.0; tb_ms = 0.0
    for _ in range(args.steps):
        dev_ms += rb.run()                      # CUDA events on the engine's stream around every launch of the step
        k = rb.kernel_ms(); dp_ms += k["dp_ms"]; tb_ms += k["tb_ms"]
    barrier_sync(dist, local)
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = rb.launches() * args.steps
    clocks = sampler.stop() if sampler else None
    dev_ms_max = allreduce_max(dist, local, dev_ms)
    wall_ms_max = allreduce_max(dist, local, wall_ms)
    cells_total = allreduce_sum(dist, local, float(cells_rank))
    pairs_total = allreduce_sum(dist, local, float(ps.n))
    launches_total = int(allreduce_sum(dist, local, float(launches)))
    ms_per_step = dev_ms_max / args.steps
    gcups = cells_total / (ms_per_step * 1e-3) / 1e9
    dp_ms_step = dp_ms / args.steps

    # ---- end-to-end arm: host buffers in, ksw_extz_t + CIGARs + stats out -----------------------------
    def e2e_once():
        # the call a user makes: one batched C-ABI call with host buffers (pack + H2D + kernels + D2H + CIGAR mallocs inside)
        res = engine.extz2_batch(ps, mat, synth.SEDEF_GAPO, synth.SEDEF_GAPE, W, ZD, FLAG, want_stats=True, keep_cigars=False)
        return res, engine.last_call_io()
    for _ in range(3):
        e2e_once()
    barrier_sync(dist, local)
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        res, io = e2e_once()
    barrier_sync(dist, local)
    e2e_ms = allreduce_max(dist, local, (time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_gcups = cells_total / (e2e_ms * 1e-3) / 1e9
    io = (int(allreduce_sum(dist, local, float(io[0]))), int(allreduce_sum(dist, local, float(io[1]))), io[2])
    score_sum = int(res.ez["score"].astype(np.int64).sum())

    # ---- CPU baseline on this box (rank 0, N == 1 only) -------------------------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        cpu = cpu_reference_run(ps, mat, W, ZD, FLAG, cells_rank, best_of=2, sample_pairs=ps.n)

    if rank == 0:
        p_int, p_int_src = peak_int_tlaneops()
        p_hbm, p_hbm_src = peak_hbm_gbs()
        cells_per_gpu = cells_total / n_gpus
        dp_s = dp_ms_step * 1e-3
        achieved_tops = cells_per_gpu * OPS_PER_CELL / dp_s / 1e12
        tb_gbs = cells_per_gpu * TB_BYTES_PER_CELL / dp_s / 1e9
        prof = load_json(os.path.join(ROOT, "profiles", "dp_kernel_ncu_latest.json"), {})
        # dram__bytes_read+write of one ncu --set full capture of the same kernel, scaled by cells to this launch size
        traffic = int(prof["dram_bytes_per_cell"] * cells_per_gpu) if "dram_bytes_per_cell" in prof else None
        line = {
            "metric": "batched ksw_extz2 GCUPS", "value": round(gcups, 2), "unit": "GCUPS",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i8", "data": "synthetic",
            "pairs_per_s": round(pairs_total / (ms_per_step * 1e-3), 1),
            "config": {"workload": "BASELINE.json configs[1]: 100k x 1 kbp pairs per GPU, w=100, 5% divergence, flag=0 "
                                   "(CIGAR + exact max + fused SD stats), SEDEF scoring 5/-4/40/1",
                       "pairs_per_gpu": ps.n, "cells_per_gpu": int(cells_per_gpu), "band_w": W, "zdrop": ZD, "flag": FLAG,
                       "l2_policy": "inputs_larger_than_l2 (0.2 GB sequences + 12.8 GB traceback per step vs 126 MB L2)",
                       "parallelism": f"{n_gpus} x independent shards, no collective"},
            "wall_ms_per_step": round(wall_ms_max / args.steps, 4),
            "kernel_ms_per_step": {"dp": round(dp_ms_step, 4), "traceback_stats": round(tb_ms / args.steps, 4)},
            "gpu_launches": launches_total,
            "e2e": {"value": round(e2e_gcups, 2), "unit": "GCUPS", "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
                    "ms_per_step": round(e2e_ms, 3), "pairs_per_s": round(pairs_total / (e2e_ms * 1e-3), 1),
                    "api": "ksw_extz2_batch_flat (one call, host buffers in, ksw_extz_t + malloc'd CIGARs + sd_stats_t out; chunked upload/launch/fetch pipeline inside)",
                    "checksum_score_sum": score_sum, "host_threads_per_rank": host_threads},
            "roofline": {"bound": "int_alu", "achieved": round(achieved_tops, 3), "peak": round(p_int, 3), "unit": "Tlane-op/s",
                         "frac": round(achieved_tops / p_int, 4), "traffic": traffic,
                         "kernel": "extz_dp16_kernel<4,cigar,left> (packed: 4 lanes x 32 slots per pair, 2 slots per register, 8 pairs per warp)",
                         "ops_per_cell": OPS_PER_CELL, "peak_source": p_int_src,
                         "frac_vs_packed_peak": round(achieved_tops / (2.0 * p_int), 4),
                         "note": "integer min/max DP: the binding unit is the INT ALU pipe, not HBM or tensor cores; achieved = in-band "
                                 "cells/s x 34 reference lane-ops per cell (SURVEY 8d) / DP-kernel device time; peak = measured 32-bit "
                                 "lane-op issue rate of the ALU pipe (64 lanes/clk/SM).  The kernel issues VIADD.16x2 / VIMNMX.16x2, "
                                 "which retire two of those reference ops per lane slot at the same issue rate, so the stricter ceiling "
                                 "for the 28 packable ops is 2 x peak: frac_vs_packed_peak states the fraction of that"},
            "roofline_hbm": {"bound": "hbm", "achieved": round(tb_gbs, 2), "peak": p_hbm, "unit": "GB/s",
                             "frac": round(tb_gbs / p_hbm, 5), "traffic": traffic,
                             "peak_source": p_hbm_src, "note": "traceback write stream, 0.5 B per in-band cell (algorithmic)"},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    rb.free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def host_cpu_threads(lib) -> int:
    """Threads for the CPU reference: ALL host cores.  Launchers such as torchrun export OMP_NUM_THREADS=1, which made the
    reference arm run single-threaded at N > 1 (1.25 instead of 19.9 GCUPS) -- the OpenMP default is therefore not trusted."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    return max(lib.max_threads(), avail, 1)


def cpu_reference_run(ps, mat, W, ZD, FLAG, cells, best_of=2, sample_pairs=None, threads=0):
    """The reference's own ksw_extz2_sse (oracle/_ref, compiled from the untouched source) under an OpenMP
    parallel-for over pairs on all host cores; falls back to the scalar port if _ref is absent."""
    import oracle
    kind = "reference" if oracle.have_ref() else "port"
    lib = oracle.ref() if oracle.have_ref() else oracle.port()
    nthreads = threads or host_cpu_threads(lib)
    best = None
    for _ in range(best_of):
        s = lib.batch(ps, mat, 40, 1, W, ZD, FLAG, nthreads=nthreads, keep=False)
        best = s if best is None else min(best, s)
    return {"value": round(cells / best / 1e9, 3), "unit": "GCUPS", "cores": nthreads, "kind": kind,
            "pairs_per_s": round(ps.n / best, 1), "seconds": round(best, 3),
            "sample": f"{ps.n} of the {sample_pairs or ps.n} pairs of the same workload, best of {best_of}, "
                      f"OpenMP parallel-for over pairs, {nthreads} threads"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from sedef_b200 import synth
    mat = synth.sedef_matrix()
    W, ZD, FLAG = 100, -1, 0
    n = min(args.pairs, args.ref_pairs)
    ps = make_workload(n, 0)
    import oracle
    lib = oracle.ref() if oracle.have_ref() else oracle.port()
    kind = "reference" if oracle.have_ref() else "port"
    nthreads = host_cpu_threads(lib)
    # exact in-band cell count of the sample (the oracle's own counter, oracle/ksw2_extz2_port.c)
    cnt = oracle.port().lib.oracle_count_cells
    cells = float(sum(int(cnt(int(q), int(t), W)) for q, t in zip(ps.qlen, ps.tlen)))
    for _ in range(max(1, min(args.warmup, 2))):
        lib.batch(ps, mat, 40, 1, W, ZD, FLAG, nthreads=nthreads, keep=False)
    secs = 0.0
    for _ in range(args.steps):
        secs += lib.batch(ps, mat, 40, 1, W, ZD, FLAG, nthreads=nthreads, keep=False)
    ms_per_step = secs / args.steps * 1e3
    gcups = cells / (ms_per_step * 1e-3) / 1e9
    line = {"impl": "reference", "metric": "batched ksw_extz2 GCUPS", "value": round(gcups, 3), "unit": "GCUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i8", "data": "synthetic",
            "pairs_per_s": round(ps.n / (ms_per_step * 1e-3), 1),
            "config": {"workload": "BASELINE.json configs[1]: 1 kbp pairs, w=100, 5% divergence, flag=0, SEDEF scoring "
                                   f"(bounded sample: {ps.n} pairs per step)", "band_w": W, "zdrop": ZD, "flag": FLAG},
            "cpu_baseline": {"value": round(gcups, 3), "unit": "GCUPS", "cores": nthreads, "kind": kind,
                             "sample": f"{ps.n} pairs per step x {args.steps} steps, OpenMP parallel-for over pairs calling "
                                       "the unmodified extern/ksw2_extz2_sse.cc (SSE4.1 path)"},
            "e2e": {"value": round(gcups, 3), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=100000, help="pairs per GPU (BASELINE.json configs[1]: 100k)")
    ap.add_argument("--ref-pairs", type=int, default=20000, help="pairs per step of the --impl reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
