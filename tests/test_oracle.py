"""CPU tests: the oracle itself is pinned against the reference's golden vectors (tests/golden, generated
from the compiled reference) and -- where oracle/_ref exists -- fuzzed against the compiled reference."""
import numpy as np
import pytest

import oracle
from sedef_b200 import synth
from helpers import FIELD_ORDER, load_json, parse_cigar


def test_struct_layout():
    import ctypes
    assert ctypes.sizeof(oracle.KswExtz) == 56          # sizeof(ksw_extz_t) == 56
    assert ctypes.sizeof(oracle.SdStats) == 64


def test_port_matches_kat(built, mat, golden_dir):
    kat = load_json(golden_dir, "ksw2_kat.json")
    ps = synth.pairs_from_strings([(kat["seq1"], kat["seq2"])])
    q, t = ps.pair(0)
    for row in kat["rows"]:
        f, c = oracle.port().extz2(q, t, mat, 40, 1, row["w"], row["zdrop"], row["flag"])
        assert f == row["fields"], row
        assert oracle.cigar_str(c) == row["cigar"], row
    # the survey's table (Appendix B.3) spot values
    r0 = kat["rows"][0]
    assert r0["fields"]["score"] == 5180 and r0["fields"]["mte_q"] == 1133 and r0["cigar"] == "277M3I15M1I87M1D570M15I168M"


def test_port_matches_golden_vectors(built, mat, golden_dir):
    g = load_json(golden_dir, "ksw2_golden.json")
    assert g["field_order"] == FIELD_ORDER
    n = 0
    for grp in g["groups"]:
        ps = synth.pairs_from_strings([tuple(p) for p in grp["pairs"]])
        for run in grp["runs"]:
            _, fr, cr = oracle.port().batch(ps, mat, 40, 1, run["w"], run["zdrop"], run["flag"], nthreads=2)
            for i in range(ps.n):
                assert [fr[i][k] for k in FIELD_ORDER] == run["fields"][i], (grp["name"], run["w"], run["flag"], i)
                assert oracle.cigar_str(cr[i]) == run["cigars"][i], (grp["name"], run["w"], run["flag"], i)
                n += 1
    assert n > 500


def test_sd_stats_port_matches_reference_alignment_class(built, mat, golden_dir):
    g = load_json(golden_dir, "sd_stats_golden.json")
    for rec in g["records"]:
        ps = synth.pairs_from_strings([(rec["a"], rec["b"])])
        q, t = ps.pair(0)
        _, cig = oracle.port().extz2(q, t, mat, 40, 1, -1, -1, 0)
        assert oracle.cigar_str(cig, "MDI") == rec["cigar"]            # ksw I -> 'D', ksw D -> 'I' (src/align.cc:62)
        st = oracle.sd_stats(cig, ps.q_raw, ps.t_raw)
        for k in ("span", "matches", "mismatches", "gaps", "gap_bases"):
            assert st[k] == rec[k], (k, rec["cigar"])
        # the ten BEDPE integers: src/stats_main.cc:244-271 applied (by tests/golden/make_golden.py, literally) to the column
        # strings the reference's own Alignment built for this pair
        for k, v in rec["stat_loop"].items():
            assert st[k] == v, (k, rec["cigar"])


def test_sd_stats_port_matches_survey_stat_loop(built, mat, golden_dir):
    kat = load_json(golden_dir, "ksw2_kat.json")
    ps = synth.pairs_from_strings([(kat["seq1"], kat["seq2"])])
    q, t = ps.pair(0)
    _, cig = oracle.port().extz2(q, t, mat, 40, 1, -1, -1, 0)
    st = oracle.sd_stats(cig, ps.q_raw, ps.t_raw)
    exp = kat["survey_stat_loop"]
    for k in ("span", "indel_a", "indel_b", "alnB", "matchB", "mismatchB", "transitionsB", "transversionsB",
              "uppercaseA", "uppercaseB", "uppercaseMatches", "gaps", "gap_bases"):
        assert st[k] == exp[k], k
    sa = kat["sedef_alignment"]
    assert (st["matches"], st["mismatches"]) == (sa["matches"], sa["mismatches"])


def test_cell_count_matches_appendix_c(built):
    # SURVEY.md Appendix C
    assert synth.count_cells(1000, 1000, 100) == 190900
    assert synth.count_cells(1000, 1000, -1) == 1000000
    assert synth.count_cells(10000, 10000, 500) == 9759500
    assert synth.count_cells(32, 32, -1) == 1024
    lib = oracle.port().lib
    for (q, t, w) in [(1000, 1000, 100), (7, 300, 5), (300, 7, 50), (1, 1, -1), (50, 60, 0)]:
        assert lib.oracle_count_cells(q, t, w) == synth.count_cells(q, t, w)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("w,zdrop,flag,maxlen,div", [
    (-1, -1, 0, 300, 0.1), (20, -1, 0, 500, 0.1), (50, 100, 0, 500, 0.15), (100, -1, 0, 900, 0.05),
    (30, 80, 0x01, 400, 0.2), (30, 80, 0x02, 400, 0.2), (30, 80, 0x04, 400, 0.2), (30, 80, 0x08, 400, 0.2),
    (30, 80, 0x18, 400, 0.2), (30, 80, 0x40, 400, 0.2), (30, 80, 0x80, 400, 0.2), (30, 80, 0xc2, 400, 0.2),
    (5, -1, 0, 300, 0.3), (1, -1, 0, 100, 0.3), (0, -1, 0, 100, 0.1), (16, 30, 0, 400, 0.4)])
def test_port_vs_compiled_reference_fuzz(built, mat, w, zdrop, flag, maxlen, div):
    ps = synth.make_pairs_mixed(120, seed=31 * (w + 7) + flag, min_len=1, max_len=maxlen, div=div)
    _, fr, cr = oracle.ref().batch(ps, mat, 40, 1, w, zdrop, flag, nthreads=4)
    _, fp, cp = oracle.port().batch(ps, mat, 40, 1, w, zdrop, flag, nthreads=4)
    assert fr == fp
    assert cr == cp


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("ma,mi,go,ge", [(1, -1, 2, 1), (2, -6, 60, 5), (12, -8, 59, 1), (12, -4, 59, 4), (12, -4, 59, 5),
                                         (11, -4, 59, 5), (8, -4, 59, 5), (20, -4, 55, 4), (30, -4, 50, 2), (6, -3, 58, 4)])
def test_port_vs_compiled_reference_scoring_regimes(built, ma, mi, go, ge):
    """Scoring with 2(q+e) + match above 127 puts u / v bytes above 127: the reference reads them as uint8_t in the H update
    (:103,228,255) but carries them through an int8_t + _mm_cvtsi32_si128 at the start of every diagonal (:102,144-145), which
    sign-extends into the next three lanes.  The port must restate both."""
    m = synth.sedef_matrix(ma, mi)
    ps = synth.make_pairs_small(300, length=200, div=0.05, seed=871792603, len_jitter=40)
    for (w, flag) in ((100, 0), (-1, 0x02), (16, 0x80)):
        _, fr, cr = oracle.ref().batch(ps, m, go, ge, w, -1, flag, nthreads=4)
        _, fp, cp = oracle.port().batch(ps, m, go, ge, w, -1, flag, nthreads=4)
        assert fr == fp, (w, flag)
        assert cr == cp, (w, flag)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_port_vs_compiled_reference_extreme_scoring_and_matrices(built):
    """The port against the compiled reference where the reference's own arithmetic has overflowed (match up to 127, mismatch
    down to -128, gap open 0..127, gap extend 0..10), with alphabet sizes 5 / 6 / 8 and fully random matrices under GENERIC_SC."""
    rng = np.random.Generator(np.random.PCG64(20261017))
    for k in range(40):
        ma, mi = int(rng.choice([1, 5, 20, 50, 100, 127])), -int(rng.choice([1, 4, 20, 60, 100, 128]))
        go, ge = int(rng.choice([0, 1, 10, 40, 63, 64, 90, 120, 127])), int(rng.choice([0, 1, 2, 5, 10]))
        w, zd = int(rng.choice([-1, 16, 100, 1000])), int(rng.choice([-1, 100, 1000]))
        flag = int(rng.choice([0, 0x02, 0x42, 0x40, 0x80, 0x04, 0x01]))
        m = int(rng.choice([5, 5, 6, 8]))
        if flag & 0x04:
            mm = rng.integers(-12, 13, (m, m)).astype(np.int8); mm[np.arange(m), np.arange(m)] = rng.integers(1, 13, m); go = max(go, 7)
        else:
            mm = np.full((m, m), max(mi, -128), np.int8); mm[np.arange(m), np.arange(m)] = ma
        mat = mm.reshape(-1).copy()
        ps = synth.make_pairs_mixed(80, seed=int(rng.integers(1, 1 << 30)), min_len=1, max_len=int(rng.choice([150, 600])),
                                    div=float(rng.choice([0.05, 0.2, 0.4])))
        _, fr, cr = oracle.ref().batch(ps, mat, go, ge, w, zd, flag, m=m, nthreads=4)
        _, fp, cp = oracle.port().batch(ps, mat, go, ge, w, zd, flag, m=m, nthreads=4)
        assert fr == fp, (k, ma, mi, go, ge, w, zd, hex(flag), m)
        assert cr == cp, (k, ma, mi, go, ge, w, zd, hex(flag), m)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_region_golden_reproducible(built, golden_dir):
    """tests/golden/fast_align_golden.json is what the compiled reference align stage produces here (CPU, SSE kernel)."""
    import ctypes as C, os
    path = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libsedef_ref.so")
    if not os.path.exists(path):
        pytest.skip("libsedef_ref.so not built")
    lib = C.CDLL(path)
    lib.ref_fast_align.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    for reg in load_json(golden_dir, "fast_align_golden.json")["regions"]:
        q, t = synth.make_region_pair(reg["length"], reg["div"], seed=reg["seed"])
        buf = C.create_string_buffer(1 << 22)
        n = lib.ref_fast_align(q.encode(), t.encode(), 11, buf, len(buf))
        assert (n, buf.value.decode()) == (reg["n_hits"], reg["hits"])
