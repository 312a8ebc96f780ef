"""Randomised soak: many seeded random configurations (band, z-drop, flags, length ranges, divergence, bursts, last-row
maxima, class-boundary lengths) through the engine and the reference; every field and CIGAR must agree.
usage: gpu_soak.py [n_configs] [seed] [scoring]   (developer tool; the fixed suites live in tests/)
scoring = "rand": every configuration also draws match 1..12, mismatch -1..-12, gap open 1..60, gap extend 1..5 and sometimes
a non-zero N row (fringe-cell wrap-around depends on the scoring; default is SEDEF's 5/-4/40/1).
scoring = "matrix": as "rand", plus alphabet sizes m in {5, 6, 8} (the wildcard is symbol m-1, so with m > 5 an N is an ordinary
symbol) and, under KSW_EZ_GENERIC_SC, a fully random m x m matrix."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from sedef_b200 import engine, synth

ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
rng = np.random.Generator(np.random.PCG64(seed))
rand_scoring = len(sys.argv) > 3 and sys.argv[3] in ("rand", "matrix", "extreme")
extreme = len(sys.argv) > 3 and sys.argv[3] == "extreme"       # match <= 100, mismatch >= -100, gap open 0..120, gap extend 0..10
rand_matrix = len(sys.argv) > 3 and sys.argv[3] == "matrix"
mat0 = synth.sedef_matrix()
engine.init(0, 1)
chk = oracle.ref() if oracle.have_ref() else oracle.port()
FLAGS = [0, 0, 0, 0, 0x02, 0x01, 0x40, 0x80, 0x42, 0xc2, 0x04, 0x08, 0x18, 0x1a, 0x58]
tot = bad = 0
t0 = time.time()
for ci in range(ncfg):
    kind = rng.choice(["mixed", "mixed", "mixed", "small", "lastrow", "boundary", "large", "spare"])
    w = int(rng.choice([-1, -1, -1, 0, 1, 3, 7, 16, 31, 64, 100, 250, 500, 1000]))
    zd = int(rng.choice([-1, -1, 10, 40, 100, 300, 1000]))
    flag = int(rng.choice(FLAGS))
    s = int(rng.integers(1, 1 << 30))
    mat, go, ge, m = mat0, 40, 1, 5
    if rand_scoring:
        ma, mi = int(rng.integers(1, 13)), -int(rng.integers(1, 13))
        go, ge = int(rng.integers(1, 61)), int(rng.integers(1, 6))
        if extreme:
            ma, mi = int(rng.choice([1, 5, 20, 50, 100, 127])), -int(rng.choice([1, 4, 20, 60, 100, 128]))
            go, ge = int(rng.choice([0, 1, 10, 40, 63, 64, 90, 120, 127])), int(rng.choice([0, 1, 2, 5, 10]))
        mat = synth.sedef_matrix(ma, mi)
        if rng.random() < 0.3:                          # N scores something (ksw2's sc_ambi style) instead of 0
            m5 = mat.reshape(5, 5).copy(); m5[4, :] = m5[:, 4] = -int(rng.integers(0, 4)); mat = m5.reshape(-1).copy()
        if rand_matrix:
            m = int(rng.choice([5, 5, 6, 8]))
            if flag & 0x04:                             # GENERIC_SC: the whole matrix is used (:141)
                mm = rng.integers(-12, 13, (m, m)).astype(np.int8)
                mm[np.arange(m), np.arange(m)] = rng.integers(1, 13, m)
                go = max(go, 7)                          # keep -min_sc <= 2(q+e): the early-out (:81) has its own test
            else:                                       # only mat[0], mat[1] and the wildcard rule matter (:129-136)
                mm = np.full((m, m), mi, np.int8); mm[np.arange(m), np.arange(m)] = ma
            mat = mm.reshape(-1).copy()
    if kind == "mixed":
        hi = int(rng.choice([40, 150, 400, 700, 1100, 2200]))
        n = max(8, min(600, 250000 // hi))
        ps = synth.make_pairs_mixed(n, seed=s, min_len=1, max_len=hi, div=float(rng.choice([0.02, 0.1, 0.2, 0.4])),
                                    **({"burst": int(rng.integers(20, 150))} if rng.random() < 0.3 else {}))
    elif kind == "small":
        L = int(rng.choice([60, 200, 500, 1000, 1500]))
        ps = synth.make_pairs_small(max(8, min(400, 200000 // L)), length=L, div=float(rng.choice([0.01, 0.05, 0.15])), seed=s, len_jitter=int(L * 0.2))
    elif kind == "lastrow":
        base = int(rng.choice([96, 480, 992, 1504, 3008, 6000]))
        ps = synth.make_pairs_max_on_last_row([base + 16 * int(k) for k in rng.integers(0, 12, 10 if base < 3000 else 3)],
                                              tail=int(rng.integers(20, 400)), sub=float(rng.choice([0.0, 0.03, 0.1])), seed=s)
    elif kind == "boundary":
        L = int(rng.choice([32, 64, 128, 256, 512, 1024, 2048, 4096]))
        ps = synth.make_pairs_small(10 if L <= 1024 else 4, length=L + 40, div=0.08, seed=s)
        ps.tlen[:] = np.minimum(ps.tlen, L + rng.integers(-17, 18, ps.n).astype(np.int32)).clip(1)
        ps.qlen[:] = np.minimum(ps.qlen, L + rng.integers(-17, 40, ps.n).astype(np.int32)).clip(1)
    elif kind == "spare":                              # 513..528 live slots: the 16-lane class with its spare block (extz_dp16.cuh Spare16)
        if rng.random() < 0.7:
            w = int(rng.integers(496, 512))
            ps = synth.make_pairs_mixed(int(rng.integers(8, 40)), seed=s, min_len=int(rng.choice([520, 800, 1200])), max_len=int(rng.choice([1300, 2500, 4000])),
                                        div=float(rng.choice([0.03, 0.1, 0.2])), **({"burst": int(rng.integers(20, 250))} if rng.random() < 0.4 else {}))
        else:
            w = -1
            ps = synth.make_pairs_mixed(int(rng.integers(8, 40)), seed=s, min_len=500, max_len=int(rng.choice([528, 540, 1500])), div=float(rng.choice([0.05, 0.2])))
    else:
        ps = synth.make_pairs_large(int(rng.integers(2, 6)), min_len=2000, max_len=int(rng.choice([5000, 9000, 14000])), seed=s)
        if w < 0 or w > 1000:
            w = int(rng.choice([200, 500, 1000, 3000]))
    try:
        got = engine.extz2_batch(ps, mat, go, ge, w, zd, flag, m=m)
    except engine.EngineError as ex:
        if ex.code in (-5, -3):
            continue                                   # too wide / outside the scoring domain: an explicit refusal, not a mismatch
        raise
    _, fr, cr = chk.batch(ps, mat, go, ge, w, zd, flag, m=m, nthreads=8)
    nb = 0
    for i in range(ps.n):
        ok = got.fields(i) == fr[i] and ((flag & 1) or got.cigars[i].tolist() == cr[i])
        if ok and not (flag & 1):
            fwd = cr[i] if not (flag & 0x80) else cr[i][::-1]
            ok = got.stats_dict(i) == oracle.sd_stats(fwd, *ps.raw_pair(i))
        if not ok:
            nb += 1
            if nb <= 2:
                print("  MISMATCH cfg", ci, kind, "w", w, "zd", zd, "flag", hex(flag), "seed", s, "scoring m", m, mat[0], mat[1], go, ge, "pair", i, int(ps.qlen[i]), int(ps.tlen[i]))
                print("     got", got.fields(i)); print("     ref", fr[i])
    tot += ps.n; bad += nb
    if nb:
        print("cfg", ci, kind, "n", ps.n, "w", w, "zd", zd, "flag", hex(flag), "scoring", int(mat[0]), int(mat[1]), go, ge, "BAD", nb)
print("SOAK configs", ncfg, "pairs", tot, "BAD", bad, "secs %.1f" % (time.time() - t0), "seed", seed)
