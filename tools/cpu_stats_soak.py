"""CPU-only soak of the statistics port: for random soft-masked pairs with N's, the CIGAR string, the five Alignment error
counters and the ten BEDPE stat-loop integers of oracle/sd_stats_port.c against the reference's own Alignment class (column
strings read through oracle/ref_shim.cc, stat loop transcribed in tests/golden/make_golden.py).  Needs oracle/_ref."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import oracle
from sedef_b200 import synth
import make_golden as mg
slib = mg.sedef_ref()
port = oracle.port(); mat = synth.sedef_matrix()
tot = bad = 0; t0 = time.time()
for seed in range(40):
    ps = synth.make_pairs_mixed(100, seed=9000 + seed, min_len=3, max_len=[60, 300, 900][seed % 3], div=[0.03, 0.12, 0.3][(seed // 3) % 3], n_frac=[0.0, 0.01, 0.05][(seed // 9) % 3])
    for i in range(ps.n):
        qa, ta = ps.raw_pair(i)
        fa = "".join(map(chr, qa)); fb = "".join(map(chr, ta))
        ref_rec = mg.ref_alignment(slib, fa, fb)
        aa, ab = mg.ref_alignment_strings(slib, fa, fb)
        exp = mg.stat_loop(aa, ab)
        q, t = ps.pair(i)
        _, cig = port.extz2(q, t, mat, 40, 1, -1, -1, 0)
        st = oracle.sd_stats(cig, qa, ta)
        ok = oracle.cigar_str(cig, "MDI") == ref_rec["cigar"] and all(st[k] == ref_rec[k] for k in ("span", "matches", "mismatches", "gaps", "gap_bases")) and all(st[k] == v for k, v in exp.items())
        tot += 1; bad += not ok
        if not ok and bad <= 3:
            print("MISMATCH seed", seed, "pair", i, len(fa), len(fb), {k: (st[k], v) for k, v in exp.items() if st[k] != v})
print("stats soak: pairs", tot, "bad", bad, "secs %.0f" % (time.time() - t0))
