"""Developer aid: which pairs of a fuzz set differ from the reference, and how."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from sedef_b200 import engine, synth
mat = synth.sedef_matrix()
engine.init(0, 1)
chk = oracle.ref()
ps = synth.make_pairs_mixed(400, seed=31337 + 30 + 0x42, min_len=1, max_len=700, div=0.2)
for (w, zd, flag) in [(30, 80, 0x42), (30, 80, 0x40), (30, 80, 0x02), (30, 80, 0), (30, -1, 0), (-1, 80, 0)]:
    got = engine.extz2_batch(ps, mat, 40, 1, w, zd, flag)
    _, fr, cr = chk.batch(ps, mat, 40, 1, w, zd, flag, nthreads=8)
    bad = [i for i in range(ps.n) if got.fields(i) != fr[i] or got.cigars[i].tolist() != cr[i]]
    print("w", w, "zd", zd, "flag", hex(flag), "bad", len(bad))
    for i in bad[:6]:
        g = got.fields(i)
        print("   pair", i, "qlen", int(ps.qlen[i]), "tlen", int(ps.tlen[i]), {k: (g[k], fr[i][k]) for k in g if g[k] != fr[i][k]})
# single-pair reruns of the first bad pair with truncated queries: where does it start to differ?
w, zd, flag = 30, 80, 0x42
got = engine.extz2_batch(ps, mat, 40, 1, w, zd, flag)
_, fr, cr = chk.batch(ps, mat, 40, 1, w, zd, flag, nthreads=8)
bad = [i for i in range(ps.n) if got.fields(i) != fr[i]]
if bad:
    i = bad[0]
    q, t = ps.pair(i)
    for ql in range(max(1, len(q) - 40), len(q) + 1):
        f, c = engine.extz2(q[:ql], t, mat, 40, 1, w, zd, flag)
        fr1, cr1 = chk.extz2(q[:ql], t, mat, 40, 1, w, zd, flag)
        if f != fr1:
            print("   qlen", ql, {k: (f[k], fr1[k]) for k in f if f[k] != fr1[k]})
    print("q", "".join("ACGTN"[x] for x in q))
    print("t", "".join("ACGTN"[x] for x in t))
