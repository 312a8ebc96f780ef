"""Developer aid: look for a compact input on which a build WITHOUT the tie-count clamp differs from the reference."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from sedef_b200 import engine, synth
engine.init(0, 1)
chk = oracle.ref()
m = synth.sedef_matrix(1, -100)
found = 0
for div in (0.02, 0.1, 0.2, 0.4):
    for burst in (None, 40, 100):
        kw = dict(burst=burst) if burst else {}
        ps = synth.make_pairs_mixed(357, seed=668295684, min_len=1, max_len=700, div=div, **kw)
        for (w, zd, flag) in [(1000, 1000, 0x42), (-1, 1000, 0), (-1, 300, 0)]:
            got = engine.extz2_batch(ps, m, 63, 5, w, zd, flag)
            _, fr, cr = chk.batch(ps, m, 63, 5, w, zd, flag, nthreads=8)
            bad = [i for i in range(ps.n) if got.fields(i) != fr[i]]
            if bad:
                found += 1
                print("REPRO div", div, "burst", burst, "w", w, "zd", zd, "flag", hex(flag), "bad pairs", bad[:5], [(int(ps.qlen[i]), int(ps.tlen[i])) for i in bad[:3]])
print("repro configurations found:", found)
