"""Small inputs through every kernel family, meant to be run under compute-sanitizer (memcheck / racecheck)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sedef_b200 import engine, synth
mat = synth.sedef_matrix()
engine.init(0, 1)
sets = [
    (synth.make_pairs_mixed(40, seed=1, min_len=1, max_len=30, div=0.1), -1, -1, 0),       # (2,16)
    (synth.make_pairs_mixed(24, seed=2, min_len=40, max_len=120, div=0.1), -1, 50, 0),     # (4,16)/(8,16)
    (synth.make_pairs_small(12, length=400, div=0.05, seed=3), 100, -1, 0),                # (8,16) banded
    (synth.make_pairs_small(6, length=500, div=0.1, seed=4), -1, -1, 2),                   # (32,16)
    (synth.make_pairs_large(2, min_len=1500, max_len=2500, seed=5), 500, 400, 0),          # (32,32)
    (synth.make_pairs_small(2, length=1500, div=0.1, seed=6), -1, -1, 0),                  # wide CTA
    (synth.make_pairs_small(1, length=4500, div=0.1, seed=7), -1, -1, 0),                  # cluster x2
]
only = [int(x) for x in sys.argv[1:]] or range(len(sets))
for k in only:
    ps, w, zd, flag = sets[k]
    r = engine.extz2_batch(ps, mat, 40, 1, w, zd, flag)
    print("set", k, "ok", int(r.ez["score"].astype(np.int64).sum()), flush=True)
