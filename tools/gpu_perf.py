"""Quick device-resident throughput probe (developer tool)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sedef_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 100
length = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
flag = int(sys.argv[4], 0) if len(sys.argv) > 4 else 0
mat = synth.sedef_matrix()
engine.init(0, 1)
t0 = time.time(); ps = synth.make_pairs_small(n, length=length, div=0.05); print("gen %.2fs" % (time.time() - t0))
t0 = time.time(); rb = engine.ResidentBatch(ps, mat, 40, 1, w, -1, flag); print("upload %.3fs cells %.3e" % (time.time() - t0, rb.cells()))
for it in range(4):
    ms = rb.run(); k = rb.kernel_ms()
    print("run %d: %.2f ms  -> %.1f GCUPS, %.0f pairs/s  (dp %.2f tb %.2f aux %.2f ms, launches %d)" % (
        it, ms, rb.cells() / ms / 1e6, n / ms * 1e3, k["dp_ms"], k["tb_ms"], k["aux_ms"], rb.launches()))
t0 = time.time(); res = rb.fetch(); print("fetch %.3fs" % (time.time() - t0))
t0 = time.time(); r2 = engine.extz2_batch(ps, mat, 40, 1, w, -1, flag, keep_cigars=False); dt = time.time() - t0
print("e2e one-shot %.3fs -> %.1f GCUPS %.0f pairs/s" % (dt, rb.cells() / dt / 1e9, n / dt))
