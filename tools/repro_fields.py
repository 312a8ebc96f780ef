"""Developer aid: rerun the four configurations of test_warp_traceback_kernel_forced and print every differing pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from sedef_b200 import engine, synth
mat = synth.sedef_matrix()
engine.init(0, 1)
chk = oracle.ref() if oracle.have_ref() else oracle.port()
cfgs = [(dict(min_len=1, max_len=700, div=0.12), -1, -1, 0), (dict(min_len=1, max_len=600, div=0.2), 30, 80, 0x42),
        (dict(min_len=300, max_len=1000, div=0.1), 100, -1, 0x80), (dict(min_len=900, max_len=2500, div=0.1), -1, 200, 0)]
for ci, (kw, w, zd, flag) in enumerate(cfgs):
    ps = synth.make_pairs_mixed(120, seed=777 + w, **kw)
    got = engine.extz2_batch(ps, mat, 40, 1, w, zd, flag)
    _, fr, cr = chk.batch(ps, mat, 40, 1, w, zd, flag, nthreads=8)
    bad = 0
    for i in range(ps.n):
        if got.fields(i) != fr[i] or got.cigars[i].tolist() != cr[i]:
            bad += 1
            if bad <= 3:
                print("  cfg", ci, "pair", i, "qlen", int(ps.qlen[i]), "tlen", int(ps.tlen[i]), "fields_ok", got.fields(i) == fr[i])
                print("     got", got.fields(i)); print("     ref", fr[i])
    print("cfg", ci, "w", w, "zdrop", zd, "flag", hex(flag), "bad", bad, "of", ps.n)
