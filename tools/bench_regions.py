"""Region-level throughput: the reference's fast_align, one region after the other on one host core (how
`sedef align generate` runs a bucket, src/align_main.cc:285-337), against `refine_regions_batch` with ALL regions in flight.
Anchors and chains are the reference's own in both arms; the reference arm's time is split into anchoring + chaining and the rest
(the part the driver replaces).  Developer tool: needs oracle/_ref/libsedef_ref.so (test infrastructure) and a GPU."""
import ctypes as C, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sedef_b200 import synth

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
slib = C.CDLL(os.path.join(root, "oracle", "_ref", "libsedef_ref.so"))
slib.ref_region.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
slib.ref_fast_align.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
slib.ref_chain_guides.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
import numpy as np
rng = np.random.default_rng(7)
text, want, t_total, t_front = [], [], 0.0, 0.0
buf = C.create_string_buffer(1 << 24)
for k in range(n):
    L = int(rng.integers(3000, 20000)); div = float(rng.uniform(0.03, 0.2))
    q, t = synth.make_region_pair(L, div, seed=1000 + k)
    qb, tb = q.encode(), t.encode()
    t0 = time.perf_counter(); nh = slib.ref_fast_align(qb, tb, 11, buf, len(buf)); t_total += time.perf_counter() - t0
    hits = buf.value.decode()
    nr = slib.ref_region(qb, tb, 11, 0, 0, 0, buf, len(buf))
    lines = [ln for ln in buf.value.decode().split("\n") if ln.strip()]
    text.append("R 0 0 0\n%s\n%s\n" % (q, t) + "\n".join(ln for ln in lines if ln[0] in "AC") + "\nE\n")
    want.append([ln for ln in lines if ln[0] == "H"])
drv = os.path.join(root, "tests", "cpp", "align_queue_driver")
mode = sys.argv[2] if len(sys.argv) > 2 else "fastalign"          # "fastalign": the complete path; "regions": the driver on the reference's anchors + chains
out = subprocess.run([drv, mode], input="".join(text), capture_output=True, text=True, env=dict(os.environ, REGIONS_REPS="3"))
assert out.returncode == 0, out.stderr
got, cur = [], None
for ln in out.stdout.split("\n"):
    if ln.startswith("R "):
        cur = []; got.append(cur)
    elif ln.startswith("H "):
        cur.append(ln)
    elif ln.startswith("S "):
        stats = ln
ok = sum(a == b for a, b in zip(got, want))
ms = float(out.stderr.strip().split()[-2])
trace = [ln for ln in out.stderr.split("\n") if ln.startswith("[regions]")]
if trace:
    per_rep = len(trace) // 3
    print("\n".join(trace[-per_rep:]))
print("regions %d, identical to fast_align: %d" % (n, ok))
print("reference fast_align (1 core, incl. anchoring + chaining): %.1f ms total, %.2f ms per region" % (t_total * 1e3, t_total * 1e3 / n))
what = "fast_align_batch (GPU anchors + host chaining + region driver, all regions in flight)" if mode == "fastalign" else \
       "refine_regions_batch (all regions in flight, excl. anchoring + chaining)"
print("%s: %.1f ms total, %.3f ms per region; %s" % (what, ms, ms / n, stats))
print("host cores: %d (the reference runs one region per core; %d cores would need %.1f ms)" % (os.cpu_count(), os.cpu_count(), t_total * 1e3 / os.cpu_count()))
