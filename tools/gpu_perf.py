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
zd = int(os.environ.get("PERF_ZDROP", "-1"))
t0 = time.time()
ps = synth.make_pairs_large(n, min_len=length, max_len=5 * length) if os.environ.get("PERF_LARGE") else synth.make_pairs_small(n, length=length, div=0.05)
print("gen %.2fs" % (time.time() - t0))
t0 = time.time(); rb = engine.ResidentBatch(ps, mat, 40, 1, w, zd, flag); print("upload %.3fs cells %.3e" % (time.time() - t0, rb.cells()))
for it in range(4):
    ms = rb.run(); k = rb.kernel_ms()
    print("run %d: %.2f ms  -> %.1f GCUPS, %.0f pairs/s  (dp %.2f tb %.2f aux %.2f ms, launches %d)" % (
        it, ms, rb.cells() / ms / 1e6, n / ms * 1e3, k["dp_ms"], k["tb_ms"], k["aux_ms"], rb.launches()))
t0 = time.time(); res = rb.fetch(); print("fetch %.3fs" % (time.time() - t0))
if os.environ.get("PERF_LARGE"): sys.exit(0)
for it in range(3):
    t0 = time.time(); b = engine.ResidentBatch(ps, mat, 40, 1, w, -1, flag); t1 = time.time(); b.run(); t2 = time.time()
    r2 = b.fetch(keep_cigars=False); t3 = time.time(); hm = b.host_ms(); b.free(); t4 = time.time()
    print("e2e iter %d: upload %.1f run %.1f fetch %.1f free %.1f ms | host phases %s" % (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, {k: round(v, 1) for k, v in hm.items()}))
dt = t4 - t0
lib = engine.load()
ez = np.zeros(n, engine.EZ_DTYPE); st = np.zeros(n, engine.STATS_DTYPE)
for it in range(4):
    t0 = time.time()
    rc = lib.ksw_extz2_batch_flat(n, ps.qlen.ctypes.data, ps.qoff.ctypes.data, ps.q.ctypes.data, ps.tlen.ctypes.data, ps.toff.ctypes.data,
                                  ps.t.ctypes.data, 5, mat.ctypes.data, 40, 1, w, -1, flag, ez.ctypes.data, st.ctypes.data,
                                  ps.q_raw.ctypes.data, ps.t_raw.ctypes.data)
    t1 = time.time(); lib.ksw_b200_free_cigars(ez.ctypes.data, n); t2 = time.time()
    print("one-shot C call %.1f ms (rc %d) + free_cigars %.1f ms; io %s" % ((t1 - t0) * 1e3, rc, (t2 - t1) * 1e3, engine.last_call_io()))
dt = t2 - t0
print("e2e one-shot %.3fs -> %.1f GCUPS %.0f pairs/s" % (dt, rb.cells() / dt / 1e9, n / dt))
