"""Developer aid: end-to-end pairs/s of the one-shot arena call (host buffers in page-locked memory, original-case bytes only)
vs the reference on the host cores, by pair size (unbanded, SEDEF's call shape)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from sedef_b200 import engine, synth
mat = synth.sedef_matrix()
engine.init(0, 1)
ref = oracle.ref()
nt = ref.max_threads()
print("host threads", nt)
for (n, hi) in [(400000, 30), (200000, 100), (100000, 250), (40000, 500), (10000, 1000), (2000, 2500)]:
    ps = synth.make_pairs_mixed(n, seed=hi, min_len=max(1, hi // 2), max_len=hi, div=0.1)
    cells = sum(synth.count_cells(int(a), int(b), -1) for a, b in zip(ps.qlen[:2000], ps.tlen[:2000])) / 2000 * n
    pp, keep = engine.pin_pairset(ps)
    for _ in range(2):
        engine.extz2_batch_arena(pp, mat, 40, 1, -1, -1, 0).free()
    tg = 1e9
    for _ in range(3):
        t0 = time.time(); r = engine.extz2_batch_arena(pp, mat, 40, 1, -1, -1, 0); tg = min(tg, time.time() - t0); r.free()
    if os.environ.get("KSW_B200_TRACE"):
        pass
    for k in keep:
        k.free()
    tc = min(ref.batch(ps, mat, 40, 1, -1, -1, 0, nthreads=nt, keep=False) for _ in range(2))
    print("pairs <= %4d bp x %6d: GPU e2e %7.1f ms (%6.2f M pairs/s, %6.1f GCUPS) | reference %d threads %7.1f ms (%6.2f M pairs/s, %5.1f GCUPS) | ratio %.1fx"
          % (hi, n, tg * 1e3, n / tg / 1e6, cells / tg / 1e9, nt, tc * 1e3, n / tc / 1e6, cells / tc / 1e9, tc / tg))
