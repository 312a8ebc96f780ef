"""CPU-only soak: the scalar port (oracle/ksw2_extz2_port.c) against the compiled reference (oracle/_ref) over random
configurations -- lengths up to 20 kbp, every flag incl. APPROX_MAX / APPROX_DROP, random scoring up to the int8 limits,
alphabet sizes 5/6/8, random matrices.  usage: cpu_soak.py [n_configs] [seed]   (needs /root/reference-built oracle/_ref)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from sedef_b200 import synth
ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.Generator(np.random.PCG64(seed))
port, ref = oracle.port(), oracle.ref()
FLAGS = [0, 0, 0, 0x02, 0x01, 0x40, 0x80, 0x42, 0xc2, 0x04, 0x08, 0x18, 0x09]
tot = bad = 0; t0 = time.time()
for ci in range(ncfg):
    mode = rng.choice(["sedef", "rand", "extreme"])
    ma, mi, go, ge = 5, -4, 40, 1
    if mode == "rand":
        ma, mi, go, ge = int(rng.integers(1, 13)), -int(rng.integers(1, 13)), int(rng.integers(1, 61)), int(rng.integers(1, 6))
    elif mode == "extreme":
        ma, mi = int(rng.choice([1, 5, 20, 50, 100, 127])), -int(rng.choice([1, 4, 20, 60, 100, 128]))
        go, ge = int(rng.choice([0, 1, 10, 40, 63, 64, 90, 120, 127])), int(rng.choice([0, 1, 2, 5, 10]))
    flag = int(rng.choice(FLAGS)); m = int(rng.choice([5, 5, 5, 6, 8]))
    if flag & 0x04:
        mm = rng.integers(-12, 13, (m, m)).astype(np.int8); mm[np.arange(m), np.arange(m)] = rng.integers(1, 13, m); go = max(go, 7)
    else:
        mm = np.full((m, m), max(mi, -128), np.int8); mm[np.arange(m), np.arange(m)] = ma
        if m == 5: mm[4, :] = 0; mm[:, 4] = 0
    mat = mm.reshape(-1).copy()
    w = int(rng.choice([-1, -1, 0, 1, 5, 16, 33, 100, 500, 2000])); zd = int(rng.choice([-1, -1, 10, 100, 400, 2000]))
    kind = rng.choice(["mixed", "mixed", "lastrow", "large"])
    s = int(rng.integers(1, 1 << 30))
    if kind == "mixed":
        hi = int(rng.choice([40, 300, 900, 2500]))
        ps = synth.make_pairs_mixed(max(6, min(300, 60000 // hi)), seed=s, min_len=1, max_len=hi, div=float(rng.choice([0.02, 0.15, 0.4])))
    elif kind == "lastrow":
        ps = synth.make_pairs_max_on_last_row([int(rng.choice([96, 480, 992, 2000])) + 16 * int(k) for k in rng.integers(0, 12, 8)],
                                              tail=int(rng.integers(20, 300)), sub=float(rng.choice([0.0, 0.05])), seed=s)
    else:
        ps = synth.make_pairs_large(int(rng.integers(1, 4)), min_len=3000, max_len=int(rng.choice([8000, 20000])), seed=s)
        if w < 0 or w > 2000: w = int(rng.choice([100, 500, 2000]))
    _, fp, cp = port.batch(ps, mat, go, ge, w, zd, flag, m=m, nthreads=8)
    _, fr, cr = ref.batch(ps, mat, go, ge, w, zd, flag, m=m, nthreads=8)
    b = [i for i in range(ps.n) if fp[i] != fr[i] or cp[i] != cr[i]]
    tot += ps.n; bad += len(b)
    if b:
        print("cfg", ci, kind, mode, "scoring", (ma, mi, go, ge), "m", m, "w", w, "zd", zd, "flag", hex(flag), "seed", s, "BAD", len(b), "e.g.", b[0], int(ps.qlen[b[0]]), int(ps.tlen[b[0]]))
        print("   port", fp[b[0]]); print("   ref ", fr[b[0]])
print("CPU SOAK port vs compiled reference: configs", ncfg, "pairs", tot, "BAD", bad, "secs %.0f" % (time.time() - t0), "seed", seed)
