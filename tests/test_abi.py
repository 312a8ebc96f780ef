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
    names = re.findall(r"\b(ksw_[a-z0-9_]+|sd_stats_[a-z0-9_]+|sedef_[a-z0-9_]+)\s*\(", src)
    # the extern "C" entry points of the C++ host layer (include/sedef_align.hpp)
    hpp = open(os.path.join(ROOT, "include", "sedef_align.hpp")).read()
    names += re.findall(r'extern "C"[^;(]*?\b(sedef_b200_[a-z0-9_]+)\s*\(', hpp)
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


def test_fp_fields_are_the_reference_expressions_bit_for_bit(built, golden_dir):
    """The floating-point BEDPE fields are derived on the host from the integer record with the reference's own double-precision
    expressions (src/stats_main.cc:273-283,297-299; src/align.h:84-92).  Over the integers of all golden records (produced by the
    reference's own Alignment class) the C results must equal an independent IEEE-754 double evaluation of the same expressions
    BIT FOR BIT -- not just to the 6 digits the reference prints."""
    import math, struct
    from helpers import load_json
    recs = load_json(golden_dir, "sd_stats_golden.json")["records"]
    assert len(recs) >= 50
    n_checked = 0
    for r in recs:
        row = {n: 0 for n in engine.STAT_FIELDS}
        row.update(r["stat_loop"]); row.update({k: r[k] for k in ("span", "matches", "mismatches", "gaps", "gap_bases")})
        if row["alnB"] == 0:
            continue
        fp = engine.derive_fp(row)
        exp = {}
        exp["fracMatch"] = float(row["matchB"]) / row["alnB"]
        exp["fracMatchIndel"] = float(row["matchB"]) / row["span"]
        jcp = float(row["mismatchB"]) / row["alnB"]
        p = float(row["transitionsB"]) / row["alnB"]; q = float(row["transversionsB"]) / row["alnB"]
        try:
            exp["jcK"] = -0.75 * math.log(1.0 - 4.0 / 3 * jcp)
            exp["k2K"] = 0.5 * math.log(1.0 / (1 - 2.0 * p - q)) + 0.25 * math.log(1.0 / (1 - 2.0 * q))
        except ValueError:
            continue                                    # log of a non-positive number: nan in C, an exception in Python
        exp["errorScaled"] = (row["gaps"] + row["mismatches"]) / float(row["gaps"] + row["mismatches"] + row["matches"])
        exp["filter_score"] = 1 - exp["errorScaled"]
        tot = float(row["matches"] + row["gap_bases"] + row["mismatches"])
        exp["gap_error"] = 100.0 * row["gap_bases"] / tot
        exp["mismatch_error"] = 100.0 * row["mismatches"] / tot
        exp["total_error"] = exp["mismatch_error"] + exp["gap_error"]
        for k, v in exp.items():
            assert struct.pack("<d", fp[k]) == struct.pack("<d", v), (k, fp[k], v)
        n_checked += 1
    assert n_checked >= 50


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
