"""Small inputs through every kernel family, meant to be run under compute-sanitizer (memcheck / racecheck)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sedef_b200 import engine, synth
mat = synth.sedef_matrix()
engine.init(0, 1)
sets = [
    (synth.make_pairs_mixed(40, seed=1, min_len=1, max_len=30, div=0.1), -1, -1, 0),       # packed, 1 lane x 32 slots
    (synth.make_pairs_mixed(24, seed=2, min_len=40, max_len=120, div=0.1), -1, 50, 0),     # packed, 2-4 lanes
    (synth.make_pairs_small(12, length=400, div=0.05, seed=3), 100, -1, 0),                # packed, 4 lanes, banded
    (synth.make_pairs_small(6, length=500, div=0.1, seed=4), -1, -1, 2),                   # packed, 16 lanes, right-aligned arm
    (synth.make_pairs_large(2, min_len=1500, max_len=2500, seed=5), 500, 400, 0),          # packed, 16 lanes + the spare block (528 slots)
    (synth.make_pairs_large(2, min_len=1500, max_len=2500, seed=10), 700, 400, 0),         # packed, 32 lanes (1024 slots)
    (synth.make_pairs_mixed(6, seed=11, min_len=513, max_len=528, div=0.1), -1, -1, 0),    # packed, 16 lanes, spare block unbanded
    (synth.make_pairs_small(2, length=1500, div=0.1, seed=6), -1, -1, 0),                  # packed CTA-wide, 64 lanes
    (synth.make_pairs_small(1, length=4500, div=0.1, seed=7), -1, -1, 0),                  # packed CTA-wide, 256 lanes (dynamic smem)
    (synth.make_pairs_small(1, length=8500, div=0.1, seed=8), -1, -1, 0),                  # packed cluster of 2 CTAs (DSMEM)
    (synth.make_pairs_large(3, min_len=7000, max_len=9000, seed=9), 300, 400, 0),          # warp-per-pair traceback
]
only = [int(x) for x in sys.argv[1:]] or range(len(sets))
for k in only:
    ps, w, zd, flag = sets[k]
    r = engine.extz2_batch(ps, mat, 40, 1, w, zd, flag)
    print("set", k, "ok", int(r.ez["score"].astype(np.int64).sum()), flush=True)
