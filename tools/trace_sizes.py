import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sedef_b200 import engine, synth
mat = synth.sedef_matrix(); engine.init(0, 1)
n, hi = int(sys.argv[1]), int(sys.argv[2])
ps = synth.make_pairs_mixed(n, seed=hi, min_len=max(1, hi // 2), max_len=hi, div=0.1)
for k in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    t0 = time.time(); r = engine.extz2_batch(ps, mat, 40, 1, -1, -1, 0, keep_cigars=False); print("call %d: %.1f ms" % (k, (time.time() - t0) * 1e3), file=sys.stderr)
