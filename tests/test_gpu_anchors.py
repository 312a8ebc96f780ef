"""GPU parity test of SURVEY section 8 f3: `sedef_anchors_batch` (generate_anchors on the GPU) against the reference's own
generate_anchors (src/chain.cc:24-101), through the golden regions (tests/golden/region_golden.json: 'A q r l' lines) and, where
oracle/_ref/libsedef_ref.so is present, live on more regions incl. the same-chromosome diagonal exclusion, N runs, low-complexity
sequence with over-represented k-mers and other k-mer sizes."""
import ctypes as C
import os

import numpy as np
import pytest

from sedef_b200 import engine, synth
from helpers import load_json

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_anchors_match_golden_regions(built, golden_dir):
    engine.init(0, 1)
    g = load_json(golden_dir, "region_golden.json")
    regions, same, oq, orr, want = [], [], [], [], []
    for reg in g["regions"]:
        regions.append(synth.make_region_pair(reg["length"], reg["div"], seed=reg["seed"]))
        same.append(reg["same_chr"]); oq.append(reg["orig_qs"]); orr.append(reg["orig_rs"])
        want.append([tuple(int(x) for x in ln.split()[1:4]) for ln in reg["text"].split("\n") if ln.startswith("A ")])
    got = engine.anchors_batch(regions, 11, same, oq, orr)
    for k in range(len(regions)):
        assert [tuple(int(x) for x in row[:3]) for row in got[k]] == want[k], (k, len(got[k]), len(want[k]))
        assert len(want[k]) > 50


def _ref_anchors(slib, q, r, k, same, qs0, rs0):
    buf = C.create_string_buffer(1 << 24)
    n = slib.ref_region_anchors(q.encode(), r.encode(), k, same, qs0, rs0, buf, len(buf))
    assert n >= 0
    return [tuple(int(x) for x in ln.split()) for ln in buf.value.decode().split("\n") if ln.strip()]


def test_anchors_live_vs_reference(built):
    path = os.path.join(ROOT, "oracle", "_ref", "libsedef_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libsedef_ref.so not built")
    slib = C.CDLL(path)
    if not hasattr(slib, "ref_region_anchors"):
        pytest.skip("oracle/_ref/libsedef_ref.so predates ref_region_anchors")
    slib.ref_region_anchors.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    engine.init(0, 1)
    rng = np.random.default_rng(5)
    cases = []
    for s in range(12):
        q, r = synth.make_region_pair(int(rng.integers(1500, 9000)), float(rng.uniform(0.02, 0.25)), seed=300 + s)
        cases.append((q, r, int(rng.choice([11, 11, 11, 8, 12, 14])), int(s % 3 == 0), int(rng.integers(0, 5000)), int(rng.integers(0, 5000))))
    # a region aligned against ITSELF on the same chromosome at the same coordinates: the main diagonal band is excluded
    q, _ = synth.make_region_pair(3000, 0.05, seed=77)
    cases.append((q, q, 11, 1, 1000, 1000))
    cases.append((q, q, 11, 1, 1000, 1005))
    cases.append((q, q, 11, 0, 0, 0))
    # N runs and lower-case stretches inside matches; a k-mer that occurs more than 1000 times in the reference (homopolymer + tandem repeat)
    core = "ACGTTGCAAGGCTTAACCGGATATCGCGATTACAGGCTTAAGGCCTTAACGT" * 40
    low = core[:500] + "N" * 7 + core[500:900].lower() + "n" + core[900:]
    cases.append((low, core, 11, 0, 0, 0))
    rep = "A" * 1500 + core[:300] + "AC" * 700 + core[300:800]
    rep2 = core[:300] + "A" * 1300 + "AC" * 650 + core[300:800] + "A" * 400
    cases.append((rep, rep2, 11, 0, 0, 0))
    cases.append((rep2, rep, 11, 1, 200, 3000))
    got = engine.anchors_batch([(c[0], c[1]) for c in cases[:12]], 11)   # placeholder call shape check (k differs per case below)
    assert len(got) == 12
    for (q, r, k, same, qs0, rs0) in cases:
        want = _ref_anchors(slib, q, r, k, same, qs0, rs0)
        mine = engine.anchors_batch([(q, r)], k, [same], [qs0], [rs0])[0]
        assert [tuple(int(x) for x in row) for row in mine] == want, (len(q), len(r), k, same, len(mine), len(want))
