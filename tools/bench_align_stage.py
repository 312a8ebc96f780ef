"""Align-stage bench (BASELINE.json configs 1 / 4: `sedef align generate` over the bucket files of a synthetic genome with a
planted SD catalog): the product's `sedef_b200_align_generate` (all regions through fast_align_batch) against the reference
BINARY's own `align generate` (oracle/_ref/sedef_ref, one process per bucket, as many processes at a time as the host has cores --
the way sedef.sh runs it, sedef.sh:190), on the same bucket files, with a byte-level comparison of the outputs.

    python tools/bench_align_stage.py --config 1                      # configs[0]: 2 Mbp, 40 duplications
    python tools/bench_align_stage.py --config 4                      # configs[3]: 50 Mbp chromosome, 1000 duplications
    python tools/bench_align_stage.py --config 5 --scale 1.0          # configs[4]: 24 chromosomes, 3.085 Gbp, 24 680 duplications up to 30 %
    python tools/bench_align_stage.py --config 5 --scale 1.0 --procs 1,2,4,8   # ... and one process per GPU
Prints one JSON line.  `--gpus N`: N in-process devices (ksw_b200_init(0, N); every batched call is LPT-sharded over them)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sedef_ref")


def worker(k: int, n: int, fa: str, bdir: str, wd: str):
    """One process per GPU (the caller set CUDA_VISIBLE_DEVICES): seed hits k, k + n, ... of the schedule.  A warm-up run, then the
    timed run starts when the parent creates the go-file; start / end wall-clock times go to a JSON file."""
    from sedef_b200 import engine
    engine.init(0, 1)
    out = os.path.join(wd, "shard_%d_of_%d.bed" % (k, n))
    engine.align_generate(fa, bdir, out, shard_index=k, shard_count=n)
    open(os.path.join(wd, "ready_%d_of_%d" % (k, n)), "w").close()
    go = os.path.join(wd, "go_%d" % n)
    while not os.path.exists(go):
        time.sleep(0.001)
    t0 = time.time()
    st = engine.align_generate(fa, bdir, out, shard_index=k, shard_count=n)
    t1 = time.time()
    with open(os.path.join(wd, "done_%d_of_%d.json" % (k, n)), "w") as f:
        json.dump(dict(start=t0, end=t1, **st), f)


def run_procs(n: int, fa: str, bdir: str, wd: str, cores: int):
    """n worker processes, one GPU each; returns (makespan seconds, sorted output lines, per-worker seconds)."""
    procs = []
    for k in range(n):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(k), OMP_NUM_THREADS=str(max(1, cores // n)))
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(k), str(n), fa, bdir, wd], env=env))
    while not all(os.path.exists(os.path.join(wd, "ready_%d_of_%d" % (k, n))) for k in range(n)):
        if any(p.poll() not in (None, 0) for p in procs):
            raise SystemExit("a worker failed")
        time.sleep(0.01)
    open(os.path.join(wd, "go_%d" % n), "w").close()
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("a worker failed")
    recs = [json.load(open(os.path.join(wd, "done_%d_of_%d.json" % (k, n)))) for k in range(n)]
    lines = []
    for k in range(n):
        lines += [ln for ln in open(os.path.join(wd, "shard_%d_of_%d.bed" % (k, n))).read().split("\n") if ln]
    return max(r["end"] for r in recs) - min(r["start"] for r in recs), sorted(lines), [round(r["end"] - r["start"], 3) for r in recs]


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        return worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], sys.argv[6])
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1, choices=[1, 4, 5])
    ap.add_argument("--scale", type=float, default=0.02, help="config 5 only: fraction of hg38's 3.1 Gbp (24 chromosomes)")
    ap.add_argument("--dups", type=int, default=0, help="planted duplications (default: the config's own count)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--procs", default="", help="comma-separated process counts, one GPU per process (e.g. 1,2,4,8): every process takes the "
                    "seed hits k mod n of the schedule (align_generate's shard arguments), the outputs are concatenated and sorted")
    ap.add_argument("--buckets", type=int, default=0, help="bucket files (default: host cores)")
    ap.add_argument("--repeat", type=int, default=2, help="timed product runs (best is reported; the first also warms the pools)")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--keep", default="", help="keep the work directory here")
    args = ap.parse_args()
    from sedef_b200 import engine, genome

    cores = len(os.sched_getaffinity(0))
    nb = args.buckets or cores
    cfg = dict(genome.config5(args.scale) if args.config == 5 else genome.CONFIGS[args.config])
    if args.dups:
        cfg["n_dups"] = args.dups
    wd = args.keep or tempfile.mkdtemp(prefix="align_stage_")
    t0 = time.time()
    fa, bed, catalog = genome.write_align_stage_input(wd, **cfg)
    t_gen = time.time() - t0
    bdir = os.path.join(wd, "buckets")
    os.makedirs(bdir, exist_ok=True)
    have_ref = os.path.exists(REF_BIN)
    if not have_ref:
        raise SystemExit("oracle/_ref/sedef_ref is missing (build it where /root/reference exists: make -C oracle ref_full)")
    subprocess.run([REF_BIN, "align", "bucket", "-n", str(nb), bed, bdir, fa], check=True, capture_output=True)
    for f in os.listdir(bdir):                                    # `align generate` on a directory globs *.bed
        if not f.endswith(".bed"):
            os.rename(os.path.join(bdir, f), os.path.join(bdir, f + ".bed"))
    buckets = sorted(os.path.join(bdir, f) for f in os.listdir(bdir))
    n_regions = sum(open(b).read().count("\n") for b in buckets)
    region_bases = 0
    for b in buckets:
        for ln in open(b):
            f = ln.split("\t")
            region_bases += int(f[2]) - int(f[1]) + int(f[5]) - int(f[4])

    engine.init(0, args.gpus)
    out = os.path.join(wd, "ours.aligned.bed")
    best = None
    for _ in range(max(1, args.repeat)):
        t0 = time.time()
        st = engine.align_generate(fa, bdir, out)
        dt = time.time() - t0
        if best is None or dt < best[0]:
            best = (dt, st)
    ours_s, st = best
    ours_lines = sorted(ln for ln in open(out).read().split("\n") if ln)

    line = dict(metric="align-stage regions/s (sedef align generate, all buckets)",
                config="configs[%d]" % (args.config - 1) + (" at scale %g" % args.scale if args.config == 5 else ""),
                genome_bp=sum(cfg["chrom_lengths"].values()), planted=len(catalog), regions=n_regions, region_bases=region_bases,
                buckets=len(buckets), n_gpus=args.gpus, hits=st["hits"], seconds=round(ours_s, 3),
                value=round(n_regions / ours_s, 1), unit="regions/s", hits_per_s=round(st["hits"] / ours_s, 1),
                region_mbp_per_s=round(region_bases / ours_s / 1e6, 2),
                phases_ms=dict(total=round(st["ms_total"], 1), align=round(st["ms_align"], 1), io=round(st["ms_io"], 1)),
                rounds=st["rounds"], batch_calls=st["batch_calls"], ksw_requests=st["ksw_requests"],
                ksw_pairs=st["ksw_pairs"], ksw_cells=st["ksw_cells"], ksw_pairs_per_s=round(st["ksw_pairs"] / ours_s, 1),
                ksw_gcups_over_the_stage=round(st["ksw_cells"] / ours_s / 1e9, 2), genome_gen_s=round(t_gen, 1))
    if args.procs:
        scaling = []
        for n in [int(x) for x in args.procs.split(",")]:
            sec, lines_n, per = run_procs(n, fa, bdir, wd, cores)
            scaling.append(dict(processes=n, seconds=round(sec, 3), regions_per_s=round(n_regions / sec, 1), per_worker_seconds=per,
                                identical_to_one_process=lines_n == ours_lines))
        line["one_process_per_gpu"] = scaling
    if not args.no_ref:
        def one(b):
            t = time.time()
            r = subprocess.run([REF_BIN, "align", "generate", "-k", "11", fa, b], check=True, capture_output=True, text=True)
            return r.stdout, time.time() - t
        t0 = time.time()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            res = list(ex.map(one, buckets))
        ref_wall = time.time() - t0
        ref_cpu = sum(r[1] for r in res)
        ref_lines = sorted(ln for r in res for ln in r[0].split("\n") if ln)
        line.update(cpu_baseline=dict(kind="reference", binary="oracle/_ref/sedef_ref align generate -k 11", cores=cores,
                                      processes=len(buckets), seconds=round(ref_wall, 3), cpu_seconds=round(ref_cpu, 3),
                                      value=round(n_regions / ref_wall, 1), unit="regions/s",
                                      one_core_regions_per_s=round(n_regions / ref_cpu, 2)),
                    speedup_vs_all_cores=round(ref_wall / ours_s, 2), speedup_vs_one_core=round(ref_cpu / ours_s, 1),
                    parity=dict(lines_checked=len(ref_lines), identical=ours_lines == ref_lines,
                                mismatches=len(set(ours_lines) ^ set(ref_lines))))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
