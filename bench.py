#!/usr/bin/env python
"""bench.py -- batched ksw_extz2 throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (default workload: BASELINE.json configs[1])
    python bench.py --config 3 ...                            # BASELINE.json configs[2]: 10k pairs of 10-50 kbp, w=500, z-drop 400
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU kernel (oracle/_ref) on this box's cores

Workload (config.workload).  The pair set is GLOBAL: N blocks of 100k synthetic 1 kbp pairs (block b from seed
0x5EDEF002 + 7919 b; 5 % divergence, makeSmall event model, soft-masked), band w=100, SEDEF scoring, CIGAR + exact max
(flag 0) + fused SD statistics.  It is cut into N shards by the length-balanced greedy (LPT) partition of
sedef_b200/shard.py on in-band cell estimates; rank r (one process per GPU) aligns shard r; the results are gathered on
rank 0's host.  Per-GPU work is fixed as N grows ("weak"); pairs are independent, so there is no collective on the data path.
A "step" is one pass of the hot path over the whole pair set.

value  = whole-job in-band GCUPS with the shard's sequences resident in HBM: DP + traceback/statistics kernels, device time
         by CUDA events on the engine's stream, max over ranks.
e2e    = the same metric through the reference-facing call with HOST buffers: `ksw_extz2_batch_arena` takes the shard's
         sequences as original-case bytes in page-locked host memory (SEDEF's Alignment(fa, fb) call shape: align_dna runs on
         the device) and returns ksw_extz_t records, CIGARs and sd_stats_t records on the host; at N > 1 every rank then
         exports its records into ONE shared host segment at their global indices (the gather).  H2D, kernels, D2H and the
         gather are all inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

OPS_PER_CELL = 34            # integer lane-ops per in-band cell in the reference formulation
TB_BYTES_PER_CELL = 0.5      # 4-bit traceback code per cell

CONFIGS = {
    2: dict(name="BASELINE.json configs[1]: 100k x 1 kbp pairs per GPU, w=100, 5% divergence, flag=0 "
                 "(CIGAR + exact max + fused SD stats), SEDEF scoring 5/-4/40/1",
            pairs=100000, w=100, zdrop=-1, flag=0, seed=0x5EDEF002,
            kernel="extz_dp16_kernel<4,cigar,left> (packed: 4 lanes x 32 slots per pair, 2 slots per register, 8 pairs per warp)"),
    3: dict(name="BASELINE.json configs[2]: 10k pairs of 10-50 kbp per GPU, w=500, z-drop 400, 15% divergence with indels, flag=0 "
                 "(CIGAR + exact max + fused SD stats), SEDEF scoring 5/-4/40/1",
            pairs=10000, w=500, zdrop=400, flag=0, seed=0x5EDEF003,
            kernel="extz_dp16_kernel<16,cigar,left> (packed: 16 lanes x 32 slots + a spare block spread over the lanes = 528 live slots, 2 pairs per warp)"),
}


def load_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


def peak_int_tlaneops() -> tuple:
    """Measured issue rate of the integer ALU pipe (lane-instructions per second; 64 lanes/clk/SM), profiles/r02_int_peak.json."""
    for name in ("r02_int_peak.json", "r01_int_peak.json"):
        d = load_json(os.path.join(ROOT, "profiles", name))
        if d:
            r = d["results"]
            return min(r["iadd3_3in"]["glaneops_per_s"], r["vimnmx3"]["glaneops_per_s"]) / 1e3, f"measured: profiles/{name}"
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


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
def make_block(cfg: dict, b: int, n_pairs: int):
    """Block b of the global pair set."""
    from sedef_b200 import synth
    if cfg["w"] == 100:
        return synth.make_pairs_small(n_pairs, length=1000, div=0.05, seed=cfg["seed"] + 7919 * b)
    parts, left, j = [], n_pairs, 0
    while left > 0:                                     # 1000 pairs at a time: the generator holds several int64 arrays per base
        k = min(1000, left)
        parts.append(synth.make_pairs_large(k, min_len=10000, max_len=50000, seed=cfg["seed"] + 7919 * b + 104729 * j))
        left -= k; j += 1
    return synth.concat_pairsets(parts)


def build_shard(cfg: dict, n_pairs: int, rank: int, world: int, dist, tag: str):
    """Returns (shard PairSet, global index of every shard pair, total pairs, block 0 or None)."""
    from sedef_b200 import shard, synth
    mine = make_block(cfg, rank, n_pairs)
    if world == 1:
        return mine, np.arange(mine.n, dtype=np.int64), mine.n, mine
    # every rank publishes its block once (setup, not the data path), then takes its LPT shard of the GLOBAL set
    base = os.path.join(tempfile.gettempdir(), f"sedef_bench_{tag}")
    np.savez(base + f"_b{rank}.tmp.npz", qlen=mine.qlen, qoff=mine.qoff, tlen=mine.tlen, toff=mine.toff, q_raw=mine.q_raw, t_raw=mine.t_raw)
    os.replace(base + f"_b{rank}.tmp.npz", base + f"_b{rank}.npz")
    dist.barrier()
    blocks = []
    for b in range(world):
        z = np.load(base + f"_b{b}.npz")
        blocks.append(z)
    qlen = np.concatenate([z["qlen"] for z in blocks]); tlen = np.concatenate([z["tlen"] for z in blocks])
    first = np.concatenate([[0], np.cumsum([len(z["qlen"]) for z in blocks])]).astype(np.int64)
    idx = np.sort(shard.shard_for_rank(qlen, tlen, cfg["w"], rank, world)).astype(np.int64)       # global indices of this rank's shard
    parts = []
    for b in range(world):
        loc = idx[(idx >= first[b]) & (idx < first[b + 1])] - first[b]
        if len(loc) == 0:
            continue
        z = blocks[b]
        q_raw, t_raw = z["q_raw"], z["t_raw"]
        full = synth.PairSet(z["qlen"], z["qoff"], synth.encode(q_raw), z["tlen"], z["toff"], synth.encode(t_raw), q_raw, t_raw)
        parts.append(full.subset(loc))
    dist.barrier()
    if rank == 0:
        for b in range(world):
            try:
                os.remove(base + f"_b{b}.npz")
            except OSError:
                pass
    return synth.concat_pairsets(parts), idx, int(first[-1]), (mine if rank == 0 else None)


def dist_setup():
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


def allreduce(dist, local, x: float, op: str) -> float:
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


class SharedGather:
    """ONE host segment for the results of all ranks (a file mapping every rank opens): ksw_extz_t records and sd_stats_t
    records at their GLOBAL pair index, CIGAR words in per-rank regions (ez.cigar = word offset into the segment)."""

    def __init__(self, tag: str, rank: int, world: int, n_total: int, cigar_words_per_rank: int, dist):
        from sedef_b200 import engine
        d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
        need = n_total * (56 + 64) + world * cigar_words_per_rank * 4
        try:
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize < need * 1.1:
                d = tempfile.gettempdir()
        except OSError:
            d = tempfile.gettempdir()
        self.path = os.path.join(d, f"sedef_gather_{tag}")
        self.rank, self.world, self.cap = rank, world, cigar_words_per_rank
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(need)
        dist.barrier()
        self.ez = np.memmap(self.path, engine.EZ_DTYPE, "r+", 0, (n_total,))
        self.stats = np.memmap(self.path, engine.STATS_DTYPE, "r+", n_total * 56, (n_total,))
        self.cigar = np.memmap(self.path, np.uint32, "r+", n_total * 120, (world * cigar_words_per_rank,))
        self.mine = self.cigar[rank * cigar_words_per_rank:(rank + 1) * cigar_words_per_rank]
        self.mine[:] = 0; self.ez[rank::world]["score"] = 0       # touch the pages once outside the timed region

    def close(self, dist):
        dist.barrier()
        if self.rank == 0:
            try:
                os.remove(self.path)
            except OSError:
                pass


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import oracle
    from sedef_b200 import engine, synth
    cfg = CONFIGS[args.config]
    rank, world, local, dist = dist_setup()
    n_gpus = world
    mat = synth.sedef_matrix()
    W, ZD, FLAG = cfg["w"], cfg["zdrop"], cfg["flag"]
    n_pairs = args.pairs or cfg["pairs"]
    engine.init(local, 1)
    host_threads = max(1, (os.cpu_count() or 1) // max(1, world))   # torchrun exports OMP_NUM_THREADS=1; share the cores
    engine.set_host_threads(host_threads)
    torch.cuda.set_device(local)
    tag = f"{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}"
    ps, gidx, n_total, block0 = build_shard(cfg, n_pairs, rank, world, dist, tag)
    ps_pinned, keep = engine.pin_pairset(ps)            # the caller's sequences live in page-locked host memory

    # ---- device-resident arm --------------------------------------------------------------------
    rb = engine.ResidentBatch(ps_pinned, mat, synth.SEDEF_GAPO, synth.SEDEF_GAPE, W, ZD, FLAG, raw_only=True)
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
    dev_ms_max = allreduce(dist, local, dev_ms, "max")
    wall_ms_max = allreduce(dist, local, wall_ms, "max")
    cells_total = allreduce(dist, local, float(cells_rank), "sum")
    pairs_total = allreduce(dist, local, float(ps.n), "sum")
    launches_total = int(allreduce(dist, local, float(launches), "sum"))
    cells_max = allreduce(dist, local, float(cells_rank), "max")
    ms_per_step = dev_ms_max / args.steps
    gcups = cells_total / (ms_per_step * 1e-3) / 1e9
    dp_ms_step = allreduce(dist, local, dp_ms, "max") / args.steps
    tb_ms_step = allreduce(dist, local, tb_ms, "max") / args.steps
    rb.free()

    # ---- end-to-end arm: host buffers in, ksw_extz_t + CIGARs + stats on the host (rank 0's, at N > 1) out -----------
    gather = None

    phase = [0.0, 0.0]

    def e2e_once():
        t_a = time.perf_counter()
        res = engine.extz2_batch_arena(ps_pinned, mat, synth.SEDEF_GAPO, synth.SEDEF_GAPE, W, ZD, FLAG, want_stats=True, raw_only=True)
        t_b = time.perf_counter()
        if gather is not None:
            res.export(gather.ez, gather.stats, gather.mine, cigar_base=rank * gather.cap, index=gidx)
        phase[0] += t_b - t_a; phase[1] += time.perf_counter() - t_b
        return res
    warm = e2e_once()
    words = int(warm.ez["n_cigar"].sum())
    if dist is not None:
        warm.free()
        gather = SharedGather(tag, rank, world, n_total, int(allreduce(dist, local, float(words), "max") * 1.25) + 1024, dist)
        warm = e2e_once()
    warm.free()
    e2e_once().free()
    barrier_sync(dist, local)
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    res = None
    phase[0] = phase[1] = 0.0
    for _ in range(e2e_steps):
        if res is not None:
            res.free()
        res = e2e_once()
    barrier_sync(dist, local)
    e2e_ms = allreduce(dist, local, (time.perf_counter() - t0) * 1e3, "max") / e2e_steps
    e2e_gcups = cells_total / (e2e_ms * 1e-3) / 1e9
    call_ms = allreduce(dist, local, phase[0] * 1e3 / e2e_steps, "max")
    gather_ms = allreduce(dist, local, phase[1] * 1e3 / e2e_steps, "max")
    io = res.io()
    io = (int(allreduce(dist, local, float(io[0]), "sum")), int(allreduce(dist, local, float(io[1]), "sum")),
          int(allreduce(dist, local, float(io[2]), "sum")))

    # ---- parity of the e2e results + CPU baseline on this box (rank 0) ---------------------------------------------------
    cpu = None; parity = None
    if rank == 0 and not args.no_cpu:
        ref_ps = block0
        n_ref = ref_ps.n if (args.ref_pairs <= 0) else min(ref_ps.n, args.ref_pairs)
        if n_ref < ref_ps.n:
            ref_ps = ref_ps.subset(np.arange(n_ref))
        kind = "reference" if oracle.have_ref() else "port"
        lib = oracle.ref() if oracle.have_ref() else oracle.port()
        nthreads = host_cpu_threads(lib)
        secs, ref_ez, ref_keep = lib.batch_records(ref_ps, mat, 40, 1, W, ZD, FLAG, nthreads=nthreads)
        if gather is None:
            got_ez, got_cig = res.ez[:n_ref], None
        else:
            got_ez, got_cig = gather.ez[:n_ref], gather.cigar           # block 0 = global indices [0, n_pairs)
        parity = compare_records(got_ez, got_cig, ref_ez)
        lib.free_records(ref_keep)
        ref_cells = sum(engine.count_cells(int(a), int(b), W) for a, b in zip(ref_ps.qlen, ref_ps.tlen)) if n_ref != ps.n or world > 1 else cells_rank
        if n_gpus == 1:
            best = min(secs, lib.batch(ref_ps, mat, 40, 1, W, ZD, FLAG, nthreads=nthreads, keep=False)) if args.config == 2 else secs
            cpu = {"value": round(ref_cells / best / 1e9, 3), "unit": "GCUPS", "cores": nthreads, "kind": kind,
                   "pairs_per_s": round(ref_ps.n / best, 1), "seconds": round(best, 3),
                   "sample": f"{ref_ps.n} of the {ps.n} pairs of the same workload, "
                             f"OpenMP parallel-for over pairs calling the unmodified extern/ksw2_extz2_sse.cc, {nthreads} threads"}
    res.free()
    if gather is not None:
        gather.close(dist)

    if rank == 0:
        p_int, p_int_src = peak_int_tlaneops()
        p_hbm, p_hbm_src = peak_hbm_gbs()
        dp_s = dp_ms_step * 1e-3
        achieved_tops = cells_max * OPS_PER_CELL / dp_s / 1e12
        tb_gbs = cells_max * TB_BYTES_PER_CELL / dp_s / 1e9
        prof = load_json(os.path.join(ROOT, "profiles", "dp_kernel_ncu_latest.json"), {})
        # dram__bytes_read+write of this round's ncu --set full capture of the same kernel, scaled by cells to this launch size
        traffic = int(prof["dram_bytes_per_cell"] * cells_max) if (args.config == 2 and "dram_bytes_per_cell" in prof) else None
        line = {
            "metric": "batched ksw_extz2 GCUPS", "value": round(gcups, 2), "unit": "GCUPS",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i8", "data": "synthetic",
            "pairs_per_s": round(pairs_total / (ms_per_step * 1e-3), 1),
            "config": {"workload": cfg["name"],
                       "pairs_total": int(pairs_total), "cells_total": int(cells_total), "band_w": W, "zdrop": ZD, "flag": FLAG,
                       "l2_policy": "inputs_larger_than_l2 (sequences + GBs of traceback rows per step vs 126 MB L2)",
                       "parallelism": f"one GLOBAL set of {n_gpus} x {n_pairs} pairs, LPT-partitioned into {n_gpus} shards "
                                      "(sedef_b200/shard.py), one process per GPU, results gathered in one host segment; no collective"},
            "wall_ms_per_step": round(wall_ms_max / args.steps, 4),
            "kernel_ms_per_step": {"dp": round(dp_ms_step, 4), "traceback_stats": round(tb_ms_step, 4)},
            "gpu_launches": launches_total,
            "e2e": {"value": round(e2e_gcups, 2), "unit": "GCUPS", "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
                    "ms_per_step": round(e2e_ms, 3), "pairs_per_s": round(pairs_total / (e2e_ms * 1e-3), 1),
                    "gpu_launches_per_step": int(io[2]),
                    "call_ms_max": round(call_ms, 3), "gather_ms_max": round(gather_ms, 3),
                    "api": "ksw_extz2_batch_arena: one call per rank, sequences as original-case bytes in page-locked host memory "
                           "(align_dna on the device), ksw_extz_t + CIGARs + sd_stats_t back in one page-locked arena"
                           + ("; ksw_b200_result_export gathers every rank's records into one shared host segment" if n_gpus > 1 else ""),
                    "host_threads_per_rank": host_threads},
            "roofline": {"bound": "int_alu", "achieved": round(achieved_tops, 3), "peak": round(p_int, 3), "unit": "Tlane-op/s",
                         "frac": round(achieved_tops / p_int, 4), "traffic": traffic,
                         "kernel": cfg["kernel"],
                         "ops_per_cell": OPS_PER_CELL, "peak_source": p_int_src,
                         "frac_vs_packed_peak": round(achieved_tops / (2.0 * p_int), 4),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this round's ncu --set full capture "
                                           "(profiles/dp_kernel_ncu_latest.json), per cell x the cells of one launch" if traffic else None,
                         "note": "integer min/max DP: the binding unit is the INT ALU pipe, not HBM or tensor cores; achieved = in-band "
                                 "cells/s x 34 reference lane-ops per cell (SURVEY 8d) / DP-kernel device time; peak = measured issue rate "
                                 "of the ALU pipe, 64 lanes/clk/SM for EVERY ALU instruction (profiles/r02_int_peak_sass.md: the 127 "
                                 "lanes/clk readings of 2-input add / max are ptxas fusing two of them into one IADD3 / VIMNMX3).  The "
                                 "kernel's VIADD.16x2 / VIMNMX.16x2 retire two reference ops per lane slot, so the ceiling for the 28 "
                                 "packable ops is 2 x peak: frac_vs_packed_peak states the fraction of that"},
            "roofline_hbm": {"bound": "hbm", "achieved": round(tb_gbs, 2), "peak": p_hbm, "unit": "GB/s",
                             "frac": round(tb_gbs / p_hbm, 5), "traffic": traffic,
                             "peak_source": p_hbm_src, "note": "traceback write stream, 0.5 B per in-band cell (algorithmic)"},
            "clocks": clocks,
        }
        if parity is not None:
            line.update(parity)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    for k in keep:
        k.free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def compare_records(got_ez, got_cigar_words, ref_ez) -> dict:
    """Every integer output of every pair against the reference's records: score, max, coordinates, z-drop flag, CIGAR."""
    import ctypes as C
    n = len(ref_ez)
    bad = np.zeros(n, bool)
    for k in ("max_zd", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "m_cigar"):
        bad |= np.asarray(got_ez[k]) != np.asarray(ref_ez[k])
    ok_idx = np.nonzero(~bad)[0]
    for i in ok_idx:
        nc = int(ref_ez["n_cigar"][i])
        if nc == 0:
            continue
        ref_bytes = C.string_at(int(ref_ez["cigar"][i]), nc * 4)
        if got_cigar_words is None:
            mine = C.string_at(int(got_ez["cigar"][i]), nc * 4)
        else:
            o = int(got_ez["cigar"][i])
            mine = got_cigar_words[o:o + nc].tobytes()
        if mine != ref_bytes:
            bad[i] = True
    return {"parity_pairs_checked": int(n), "mismatches": int(bad.sum()),
            "parity_fields": "max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, m_cigar, cigar[] vs the compiled reference"}


def host_cpu_threads(lib) -> int:
    """Threads for the CPU reference: ALL host cores.  Launchers such as torchrun export OMP_NUM_THREADS=1, which made the
    reference arm run single-threaded at N > 1 (1.25 instead of 19.9 GCUPS) -- the OpenMP default is therefore not trusted."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    return max(lib.max_threads(), avail, 1)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, on block 0 of the same
    pair set (all of its pairs every step)."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from sedef_b200 import engine, synth
    cfg = CONFIGS[args.config]
    mat = synth.sedef_matrix()
    W, ZD, FLAG = cfg["w"], cfg["zdrop"], cfg["flag"]
    n_pairs = args.pairs or cfg["pairs"]
    ps = make_block(cfg, 0, n_pairs if args.ref_pairs <= 0 else min(n_pairs, args.ref_pairs))
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
            "config": {"workload": cfg["name"], "pairs_total": ps.n, "cells_total": int(cells), "band_w": W, "zdrop": ZD, "flag": FLAG,
                       "parallelism": f"host CPU, {nthreads} threads; block 0 of the pair set ({ps.n} pairs) every step"},
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
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="2: BASELINE.json configs[1] (default), 3: configs[2]")
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the config's own count)")
    ap.add_argument("--ref-pairs", type=int, default=0, help="bound on the pairs of the CPU legs (0: all pairs of block 0)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    args = ap.parse_args()
    if args.config == 3 and args.ref_pairs == 0:
        args.ref_pairs = 1000          # 10k pairs of 10-50 kbp are 300 G cells: a bounded sample (~15 s on 16 cores) for the CPU legs
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
