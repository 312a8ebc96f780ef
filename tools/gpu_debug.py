"""Quick GPU-vs-oracle differential run (developer tool; the real tests live in tests/)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from sedef_b200 import engine, synth

mat = synth.sedef_matrix()
engine.init(0, 1)
checker = oracle.ref() if oracle.have_ref() else oracle.port()
total = bad = 0
t_start = time.time()
configs = [
    # (gen, n, kwargs, w, zdrop, flag)
    ("mixed", 64, dict(min_len=1, max_len=30, div=0.1), -1, -1, 0),
    ("mixed", 200, dict(min_len=1, max_len=120, div=0.1), -1, -1, 0),
    ("mixed", 200, dict(min_len=1, max_len=500, div=0.1), -1, -1, 0),
    ("mixed", 200, dict(min_len=1, max_len=600, div=0.15), 20, -1, 0),
    ("mixed", 200, dict(min_len=1, max_len=600, div=0.15), 50, 100, 0),
    ("small", 300, dict(length=1000, div=0.05), 100, -1, 0),
    ("mixed", 200, dict(min_len=1, max_len=600, div=0.2), 30, 80, 2),
    ("mixed", 200, dict(min_len=1, max_len=600, div=0.2), 30, 80, 1),
    ("mixed", 200, dict(min_len=1, max_len=600, div=0.2), 30, 80, 0x40),
    ("mixed", 200, dict(min_len=1, max_len=600, div=0.2), 30, 80, 0x80),
    ("mixed", 200, dict(min_len=1, max_len=600, div=0.2), 5, -1, 0),
    ("mixed", 200, dict(min_len=1, max_len=300, div=0.3), 1, -1, 0),
    ("mixed", 100, dict(min_len=1, max_len=100, div=0.1), 0, -1, 0),
    ("mixed", 100, dict(min_len=300, max_len=900, div=0.1), -1, -1, 0),
    ("large", 20, dict(min_len=2000, max_len=6000), 200, 400, 0),
    ("mixed", 24, dict(min_len=900, max_len=1400, div=0.1), -1, -1, 0),          # wide (64..128 lanes)
    ("mixed", 12, dict(min_len=1500, max_len=3500, div=0.1), -1, -1, 0),         # wide (128..256 lanes)
    ("large", 12, dict(min_len=4000, max_len=9000), 500, 400, 0),                # w=500: 528 slots
    ("large", 8, dict(min_len=4000, max_len=9000), 1500, -1, 0),                 # w=1500: wide banded
    ("mixed", 12, dict(min_len=1500, max_len=3500, div=0.3), -1, 300, 2),        # wide, z-drop, right-align
    ("small", 6, dict(length=6000, div=0.08), -1, -1, 0),                        # cluster of 2 CTAs (8192 slots)
    ("small", 4, dict(length=10000, div=0.08), -1, -1, 0),                       # cluster of 4 CTAs (16384 slots), SEDEF's MAX_GAP fills
    ("small", 4, dict(length=9000, div=0.3), -1, 500, 0x42),                     # cluster, z-drop, right-align, EXTZ_ONLY
    ("lastrow", 0, dict(lengths=[480, 992, 1504, 3008, 6000, 9008] + [1200 + 16 * k for k in range(12)], tail=200, seed=5), -1, -1, 0),
    ("lastrow", 0, dict(lengths=[992, 1504, 3008, 6000, 9008], tail=150, seed=7), -1, 200, 0),   # maximum in the block that leaves the band
]
only = [int(x) for x in sys.argv[1:]] if len(sys.argv) > 1 else None
for ci, (gen, n, kw, w, zdrop, flag) in enumerate(configs):
    if only and ci not in only:
        continue
    if gen == "lastrow":
        ps = synth.make_pairs_max_on_last_row(**kw); n = ps.n
    elif gen == "mixed":
        ps = synth.make_pairs_mixed(n, seed=1000 + ci, **kw)
    elif gen == "small":
        ps = synth.make_pairs_small(n, seed=1000 + ci, **kw)
    else:
        ps = synth.make_pairs_large(n, seed=1000 + ci, **kw)
    _, fr, cr = checker.batch(ps, mat, 40, 1, w, zdrop, flag, nthreads=8)
    try:
        got = engine.extz2_batch(ps, mat, 40, 1, w, zdrop, flag)
    except Exception as ex:
        print("config", ci, "ENGINE ERROR", ex); bad += n; total += n
        continue
    nb = nbc = nbs = 0
    for i in range(ps.n):
        f = got.fields(i)
        ok_f = f == fr[i]
        ok_c = got.cigars[i].tolist() == cr[i]
        ok_s = True
        if not (flag & 1) and ok_c:
            qa, ta = ps.raw_pair(i)
            ok_s = got.stats_dict(i) == oracle.sd_stats(cr[i] if not (flag & 0x80) else cr[i][::-1], qa, ta)
        if not (ok_f and ok_c and ok_s):
            nb += 1; nbc += (not ok_c); nbs += (not ok_s)
            if nb <= 3:
                print("  MISMATCH cfg", ci, "pair", i, "qlen", ps.qlen[i], "tlen", ps.tlen[i], "fields_ok", ok_f, "cigar_ok", ok_c, "stats_ok", ok_s)
                print("     got", f); print("     ref", fr[i])
                if not ok_c:
                    print("     got cigar", oracle.cigar_str(got.cigars[i].tolist())[:200]); print("     ref cigar", oracle.cigar_str(cr[i])[:200])
                if not ok_s and ok_c:
                    print("     got stats", got.stats_dict(i)); print("     ref stats", oracle.sd_stats(cr[i], qa, ta))
    total += ps.n; bad += nb
    print(f"config {ci} {gen} n={n} w={w} zdrop={zdrop} flag={flag:#x}: mismatches {nb} (cigar {nbc}, stats {nbs})", flush=True)
print("TOTAL", total, "BAD", bad, "secs %.1f" % (time.time() - t_start))
