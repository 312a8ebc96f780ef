"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/ksw2_b200.h
declares, and fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import numpy as np
import pytest

from sedef_b200 import engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ksw2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(ksw_[a-z0-9_]+|sd_stats_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_t")))


def test_header_symbols_exported(built):
    so = ctypes.CDLL(engine.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(so, n), n
    assert sorted(engine.EXPORTS) == names


def test_struct_sizes(built):
    assert ctypes.sizeof(engine.KswExtz) == 56 and engine.EZ_DTYPE.itemsize == 56
    assert ctypes.sizeof(engine.SdStats) == 64 and engine.STATS_DTYPE.itemsize == 64


def test_count_cells_and_strerror(built):
    lib = engine.load()
    assert engine.count_cells(1000, 1000, 100) == 190900
    assert engine.count_cells(0, 10, -1) == 0
    for (q, t, w) in [(33, 70, 9), (70, 33, 9), (500, 500, -1), (5, 5, 0)]:
        assert engine.count_cells(q, t, w) == synth.count_cells(q, t, w)
    assert b"no CPU fallback" in lib.ksw_b200_strerror(-1)


def test_fp_fields_match_reference_kat(built, golden_dir):
    """Host-side double arithmetic on the KAT integers reproduces SURVEY.md Appendix B.3 (6 significant digits)."""
    from helpers import load_json
    exp = load_json(golden_dir, "ksw2_kat.json")["survey_stat_loop"]
    row = {n: 0 for n in engine.STAT_FIELDS}
    row.update({k: exp[k] for k in ("span", "indel_a", "indel_b", "alnB", "matchB", "mismatchB", "transitionsB",
                                    "transversionsB", "uppercaseA", "uppercaseB", "uppercaseMatches", "gaps", "gap_bases")})
    row["matches"], row["mismatches"] = 1092, 25
    fp = engine.derive_fp(row)
    for k in ("fracMatch", "fracMatchIndel", "jcK", "k2K", "filter_score"):
        assert float("%.6g" % fp[k]) == exp[k], (k, fp[k])
    assert "%.1f" % fp["total_error"] == "4.0"          # BED score column of the KAT


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_device_fails_loudly(built, mat):
    ps = synth.make_pairs_small(4, length=50)
    with pytest.raises(engine.EngineError) as ei:
        engine.extz2_batch(ps, mat, 40, 1)
    assert ei.value.code == -1
