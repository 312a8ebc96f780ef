"""GPU parity tests of the host-side data paths added in round 2 (all through the C ABI):
  * arena output (`ksw_extz2_batch_arena`): records written by the device in the caller's order, CIGAR pointers into one
    page-locked arena, no per-pair malloc -- every field incl. m_cigar against the compiled reference;
  * sequences as original-case bytes only (codes = align_dna(bytes) derived on the device);
  * dense uploads (caller's flat buffers copied as they are: pageable through staging, page-locked in place, windows of
    one shared buffer) and the sparse per-pair packing;
  * symbol validation on the device; in-process multi-device sharding of wide pairs."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from sedef_b200 import align, engine, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def checker(built):
    engine.init(0, 1)
    return oracle.ref() if oracle.have_ref() else oracle.port()


def check_arena(res, ps, mat, checker, w, zdrop, flag, q=40, e=1, raw=True):
    _, fr, cr = checker.batch(ps, mat, q, e, w, zdrop, flag, nthreads=8)
    assert res.n == ps.n
    for i in range(ps.n):
        f = engine.BatchResult(res.ez, None, None).fields(i)
        assert f == fr[i], (i, f, fr[i])
        assert res.cigar(i).tolist() == cr[i], i
        if not (flag & engine.KSW_EZ_SCORE_ONLY):
            assert int(res.ez[i]["m_cigar"]) == checker.last_m_cigar[i], (i, int(res.ez[i]["m_cigar"]), checker.last_m_cigar[i])
            if res.stats is not None and not (flag & engine.KSW_EZ_REV_CIGAR):
                qa, ta = ps.raw_pair(i) if raw else (synth.ASCII[ps.pair(i)[0]], synth.ASCII[ps.pair(i)[1]])
                got = {n: int(res.stats[i][n]) for n in engine.STAT_FIELDS}
                assert got == oracle.sd_stats(cr[i], qa, ta), i


@pytest.mark.parametrize("w,zdrop,flag,kw", [
    (-1, -1, 0, dict(min_len=1, max_len=300, div=0.12)),
    (30, 80, 0, dict(min_len=1, max_len=700, div=0.2)),
    (30, 80, 0x42, dict(min_len=1, max_len=700, div=0.2)),
    (100, -1, 0x80, dict(min_len=500, max_len=1300, div=0.08)),
    (50, 100, 0x01, dict(min_len=1, max_len=600, div=0.15)),
])
def test_arena_raw_only_vs_oracle(checker, mat, w, zdrop, flag, kw):
    ps = synth.make_pairs_mixed(400, seed=31337 + w + flag, **kw)
    res = engine.extz2_batch_arena(ps, mat, 40, 1, w, zdrop, flag)            # original-case bytes only, arena out
    check_arena(res, ps, mat, checker, w, zdrop, flag)
    h2d, d2h, launches = res.io()
    assert launches >= 1 and d2h >= ps.n * 56
    # one byte per base crosses PCIe (plus descriptors and the score table)
    assert h2d < int(ps.qlen.sum() + ps.tlen.sum()) * 1.05 + ps.n * 40 + 4096
    res.free()
    # the same through codes + bytes and through codes alone: identical records
    both = engine.extz2_batch_arena(ps, mat, 40, 1, w, zdrop, flag, raw_only=False)
    check_arena(both, ps, mat, checker, w, zdrop, flag)
    both.free()
    codes = engine.extz2_batch_arena(ps, mat, 40, 1, w, zdrop, flag, raw_only=False, use_raw=False)
    check_arena(codes, ps, mat, checker, w, zdrop, flag, raw=False)
    codes.free()


def test_arena_with_empty_pairs_and_all_empty(checker, mat):
    ps = synth.pairs_from_strings([("ACGT", "ACGT"), ("", "ACGT"), ("ACGTACGTAA", "ACGAACGT"), ("ACGT", ""), ("acgtn", "ACGTN")])
    res = engine.extz2_batch_arena(ps, mat, 40, 1)
    check_arena(res, ps, mat, checker, -1, -1, 0)
    assert int(res.ez[1]["cigar"]) == 0 and int(res.ez[3]["m_cigar"]) == 0
    assert all(int(res.stats[1][n]) == 0 for n in engine.STAT_FIELDS)
    res.free()
    ps = synth.pairs_from_strings([("", "ACGT"), ("ACGT", "")])
    res = engine.extz2_batch_arena(ps, mat, 40, 1)
    check_arena(res, ps, mat, checker, -1, -1, 0)
    res.free()
    res = engine.extz2_batch_arena(synth.pairs_from_strings([]), mat, 40, 1)
    assert res.n == 0
    res.free()


def test_pinned_inputs_and_sparse_packing(checker, mat, monkeypatch):
    ps = synth.make_pairs_mixed(800, seed=77, min_len=1, max_len=500, div=0.1)
    pinned, keep = engine.pin_pairset(ps)                                      # DMA straight from the caller's buffers
    res = engine.extz2_batch_arena(pinned, mat, 40, 1, 40, 60, 0)
    check_arena(res, ps, mat, checker, 40, 60, 0)
    res.free()
    monkeypatch.setenv("KSW_B200_FORCE_SPARSE", "1")                           # per-pair packing through the staging buffer
    for p in (ps, pinned):
        res = engine.extz2_batch_arena(p, mat, 40, 1, 40, 60, 0)
        check_arena(res, ps, mat, checker, 40, 60, 0)
        res.free()
    monkeypatch.delenv("KSW_B200_FORCE_SPARSE")
    for k in keep:
        k.free()


def genome_windows(n, glen=60000, seed=3):
    """Pairs that are WINDOWS of one shared buffer per side (overlapping, unordered): the shape of SEDEF's requests, whose
    sequences are substrings of the genome it holds."""
    rng = np.random.default_rng(seed)
    base = synth.make_pairs_small(1, length=glen, div=0.08, seed=seed)
    ql = rng.integers(1, 400, n).astype(np.int32); tl = np.maximum(1, ql + rng.integers(-30, 30, n)).astype(np.int32)
    qo = rng.integers(0, int(base.qlen[0]) - 400, n).astype(np.int64)
    to = np.clip(qo + rng.integers(-40, 40, n), 0, int(base.tlen[0]) - 440).astype(np.int64)
    return synth.PairSet(ql, qo, base.q, tl, to, base.t, base.q_raw, base.t_raw)


def test_windows_of_one_buffer(checker, mat):
    ps = genome_windows(500)                                                   # 500 pairs x ~200 bp over 60 kbp: dense
    res = engine.extz2_batch_arena(ps, mat, 40, 1, -1, -1, 0)
    check_arena(res, ps, mat, checker, -1, -1, 0)
    h2d = res.io()[0]
    assert h2d < 2 * 70000 + ps.n * 40 + 8192                                  # the shared buffers went up once, not per pair
    res.free()
    few = genome_windows(40, glen=400000, seed=4)                              # 40 windows of a 400 kbp buffer: sparse
    res = engine.extz2_batch_arena(few, mat, 40, 1, -1, -1, 0)
    check_arena(res, few, mat, checker, -1, -1, 0)
    assert res.io()[0] < 100000
    res.free()
    got = engine.extz2_batch(ps, mat, 40, 1, -1, -1, 0)                        # the malloc-per-CIGAR form on the same input
    _, fr, cr = checker.batch(ps, mat, 40, 1, -1, -1, 0, nthreads=8)
    for i in range(ps.n):
        assert got.fields(i) == fr[i] and got.cigars[i].tolist() == cr[i], i
        assert int(got.ez[i]["m_cigar"]) == checker.last_m_cigar[i]


def test_symbol_validation(checker, mat):
    ps = synth.make_pairs_mixed(64, seed=5, min_len=20, max_len=200, div=0.1)
    # symbols 5..7 with m = 5 and a 2-value scoring: legal for the reference (it only compares for equality / the wildcard m-1)
    odd = synth.PairSet(ps.qlen, ps.qoff, ps.q.copy(), ps.tlen, ps.toff, ps.t.copy(), ps.q_raw, ps.t_raw)
    odd.q[::7] = 6; odd.t[::5] = 7; odd.t[3::11] = 5
    got = engine.extz2_batch(odd, mat, 40, 1, 30, -1, 0, use_raw=False, want_stats=False)
    _, fr, cr = checker.batch(odd, mat, 40, 1, 30, -1, 0, nthreads=8)
    for i in range(ps.n):
        assert got.fields(i) == fr[i] and got.cigars[i].tolist() == cr[i], i
    # ... but not with KSW_EZ_GENERIC_SC, where the reference would index past mat[]
    with pytest.raises(engine.EngineError) as ei:
        engine.extz2_batch(odd, mat, 40, 1, 30, -1, engine.KSW_EZ_GENERIC_SC, use_raw=False, want_stats=False)
    assert ei.value.code == -7
    bad = synth.PairSet(ps.qlen, ps.qoff, ps.q.copy(), ps.tlen, ps.toff, ps.t.copy(), ps.q_raw, ps.t_raw)
    bad.t[int(bad.toff[40]) + 3] = 9
    with pytest.raises(engine.EngineError) as ei:
        engine.extz2_batch(bad, mat, 40, 1, 30, -1, 0, use_raw=False, want_stats=False)
    assert ei.value.code == -7
    # a byte >= 8 that no pair references does not matter (dense uploads carry such bytes along)
    gap = synth.pairs_from_strings([("ACGTACGT", "ACGTTCGT"), ("GGGG", "GGCG")])
    gap.q = np.concatenate([gap.q[:8], np.full(5, 200, np.uint8), gap.q[8:]]); gap.qoff[1] += 5
    r = engine.extz2_batch(gap, mat, 40, 1, -1, -1, 0, use_raw=False, want_stats=False)
    assert r.fields(0)["score"] == 31 and r.fields(1)["score"] == 11
    # bytes only + an alphabet other than align_dna's
    with pytest.raises(engine.EngineError) as ei:
        engine.extz2_batch_arena(ps, np.zeros(36, np.int8), 40, 1, m=6)
    assert ei.value.code == -7


def test_arena_chunked_pipeline(checker, mat, monkeypatch):
    monkeypatch.setenv("KSW_B200_CHUNK_PAIRS", "500")
    ps = synth.make_pairs_mixed(4000, seed=2024, min_len=1, max_len=260, div=0.12)
    ps.qlen[100] = 0; ps.tlen[3100] = 0                                         # reset records inside chunks
    res = engine.extz2_batch_arena(ps, mat, 40, 1, 40, 60, 0)
    assert res.io()[2] >= 8
    check_arena(res, ps, mat, checker, 40, 60, 0)
    res.free()


def test_resident_fetch_arena(checker, mat):
    ps = synth.make_pairs_mixed(600, seed=11, min_len=1, max_len=400, div=0.1)
    rb = engine.ResidentBatch(ps, mat, 40, 1, 50, 100, 0, raw_only=True)
    rb.run(); rb.run()
    res = rb.fetch_arena()
    check_arena(res, ps, mat, checker, 50, 100, 0)
    res.free()
    assert rb.io_bytes()[0] < int(ps.qlen.sum() + ps.tlen.sum()) * 1.05 + ps.n * 40 + 4096      # per run, not accumulated
    rb.free()


def test_multi_device_wide_pairs(checker, mat, tmp_path):
    """In-process LPT sharding over every visible GPU with pairs that need the dynamic-shared-memory kernels (CTA-wide 256
    lanes, cluster of 2 CTAs): the opt-in function attribute is per device."""
    import subprocess, sys, textwrap, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "wide_md.py"
    script.write_text(textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {root!r})
        import oracle
        from sedef_b200 import engine, synth
        mat = synth.sedef_matrix()
        ndev = engine.init(0, 0)
        chk = oracle.ref() if oracle.have_ref() else oracle.port()
        for (cnt, length) in [(6, 5500), (4, 9000), (40, 700)]:
            ps = synth.make_pairs_small(cnt, length=length, div=0.08, seed=length)
            res = engine.extz2_batch_arena(ps, mat, 40, 1, -1, -1, 0)
            _, fr, cr = chk.batch(ps, mat, 40, 1, -1, -1, 0, nthreads=8)
            for i in range(ps.n):
                assert engine.BatchResult(res.ez, None, None).fields(i) == fr[i], (length, i)
                assert res.cigar(i).tolist() == cr[i], (length, i)
                assert {{n: int(res.stats[i][n]) for n in engine.STAT_FIELDS}} == oracle.sd_stats(cr[i], *ps.raw_pair(i)), i
            res.free()
        print("ok devices", ndev)
    """))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ok devices" in out.stdout, out.stdout + out.stderr
    assert int(out.stdout.split("devices")[1].split()[0]) == torch.cuda.device_count()


def test_from_cigar_other_ops(checker):
    """Op letters other than M / D / I in Alignment(fa, fb, cigar): the reference treats them as not-M (a gap run) that
    consumes BOTH strings (src/align.cc:283-305).  Also runs without a count and ';' separators (src/align.cc:94-103)."""
    path = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libsedef_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libsedef_ref.so not built")
    lib = C.CDLL(path)
    fa, fb = "ACGTNacgtACGTTTGACA", "ACGTAacctACGTTGACAGG"
    for cig in ["4M2X3M1D2=3I4M", "5M;4X2I1D6M", "3MX4M", "2S5M1N3M", "19M", "4M0D5M"]:
        a = align.from_cigars([(fa, fb)], [cig])[0]
        v = [C.c_int(0) for _ in range(5)]
        lib.ref_alignment_from_cigar(fa.encode(), fb.encode(), cig.encode(), *[C.byref(x) for x in v])
        assert [x.value for x in v] == [a.span(), a.matches(), a.mismatches(), a.gaps(), a.gap_bases()], cig


def _trim_scans(cigar, qa, ta, match=5, mismatch=-4, gapo=40, gape=1):
    """Alignment::trim_front / trim_back scans (reference src/align.cc:345-365, 402-420) restated over the columns of one
    alignment: returns (trim_front max_i or -1 when never updated, columns trim_back keeps or -1)."""
    up = lambda c: c - 32 if 97 <= c <= 122 else c
    cols, ia, ib = [], 0, 0                       # 0 = '|', 1 = mismatch, 2 = align_a is '-', 3 = align_b is '-'
    for c in cigar:
        op, ln = c & 0xF, c >> 4
        for _ in range(ln):
            if op == 0:
                a, b = up(int(qa[ia])), up(int(ta[ib])); ia += 1; ib += 1
                cols.append(0 if (a == b and a != ord("N") and b != ord("N")) else 1)
            elif op == 1:
                cols.append(3); ia += 1           # ksw I: query only -> align_b is '-'
            else:
                cols.append(2); ib += 1
    n = len(cols)
    score, best, front = 0, 0, -1
    for i in range(n - 1, -1, -1):
        if cols[i] == 0: score += match
        elif cols[i] == 1: score += mismatch
        else:
            if i == n - 1 or cols[i + 1] != cols[i]: score -= gapo
            score -= gape
        if score >= best: best, front = score, i
    score, best, back = 0, 0, -1
    for i in range(n):
        if cols[i] == 0: score += match
        elif cols[i] == 1: score += mismatch
        else:
            if i == 0 or cols[i - 1] != cols[i]: score -= gapo
            score -= gape
        if score >= best: best, back = score, i
    return front, (back + 1 if back >= 0 else -1)


@pytest.mark.parametrize("kw,w,zd", [(dict(min_len=1, max_len=120, div=0.3), -1, -1), (dict(min_len=200, max_len=520, div=0.15, burst=40), -1, -1),
                                     (dict(min_len=5000, max_len=7000, div=0.1), 300, -1), (dict(min_len=1, max_len=400, div=0.45), 30, 60)])
def test_trim_scans_on_the_traceback_walk(checker, mat, kw, w, zd):
    """the maximum-suffix / maximum-prefix scans of trim_front / trim_back computed on the device during the
    traceback (thread-per-pair and warp-per-pair kernels), against the reference's loops over the same columns."""
    ps = synth.make_pairs_mixed(300 if kw["max_len"] < 2000 else 12, seed=515 + w, **kw)
    res = engine.extz2_batch_arena(ps, mat, 40, 1, w, zd, 0)
    assert res.trims is not None and res.trims.shape == (ps.n, 2)
    n_trimmed = 0
    for i in range(ps.n):
        qa, ta = ps.raw_pair(i)
        f, b = _trim_scans(res.cigar(i).tolist(), qa, ta)
        assert (int(res.trims[i, 0]), int(res.trims[i, 1])) == (f, b), (i, res.trims[i].tolist(), (f, b))
        n_trimmed += (f > 0) + (0 <= b < int(res.stats[i]["span"]))
    assert n_trimmed > 0
    res.free()


def test_result_export_into_caller_memory(checker, mat):
    """`ksw_b200_result_export`: position-independent copy of an arena into caller-owned arrays at given indices (the host-side
    gather of a one-process-per-GPU deployment): records land at index[i], `cigar` becomes a word offset into the caller's buffer."""
    ps = synth.make_pairs_mixed(500, seed=99, min_len=1, max_len=300, div=0.1)
    res = engine.extz2_batch_arena(ps, mat, 40, 1, 30, 50, 0)
    rng = np.random.default_rng(1)
    index = rng.permutation(800)[:ps.n].astype(np.int64)
    ez = np.zeros(800, engine.EZ_DTYPE); st = np.zeros(800, engine.STATS_DTYPE)
    words = int(res.ez["n_cigar"].sum())
    cig = np.zeros(words + 10, np.uint32)
    assert res.export(ez, st, cig, cigar_base=1000, index=index) == words
    for i in range(ps.n):
        d = int(index[i])
        for k in ("max_zd", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "m_cigar"):
            assert ez[d][k] == res.ez[i][k], (i, k)
        n = int(ez[d]["n_cigar"])
        if n:
            o = int(ez[d]["cigar"]) - 1000
            assert np.array_equal(cig[o:o + n], res.cigar(i)), i
        else:
            assert int(ez[d]["cigar"]) == 0
        assert st[d] == res.stats[i]
    with pytest.raises(engine.EngineError):
        res.export(ez, st, cig[:max(1, words // 2)], index=index)
    res.free()
