"""Developer aid: one cluster-class pair that z-drops after a few hundred anti-diagonals (cheap under racecheck)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sedef_b200 import engine, synth
rng = np.random.Generator(np.random.PCG64(3))
head = "".join("ACGT"[x] for x in rng.integers(0, 4, 300))
q = head + "".join("ACGT"[x] for x in rng.integers(0, 4, 8300))
t = head + "".join("ACGT"[x] for x in rng.integers(0, 4, 8300))     # 300 matching bases, then unrelated: z-drop fires early
ps = synth.pairs_from_strings([(q, t)])
engine.init(0, 1)
r = engine.extz2_batch(ps, synth.sedef_matrix(), 40, 1, -1, 100, 0)
print("ok", r.fields(0))
