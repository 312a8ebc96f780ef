"""GPU parity tests: the CUDA path (through the C ABI of include/ksw2_b200.h) against
  (1) the committed golden vectors produced by the compiled reference,
  (2) the oracle (oracle/_ref compiled reference when present, else the scalar port) on seeded inputs,
  (3) size-independent properties at BASELINE.json's full sizes.
Bit-exact on every integer output: score, max/end coordinates, z-drop flag, CIGAR, SD statistics."""
import numpy as np
import pytest

import oracle
from sedef_b200 import align, engine, synth
from helpers import FIELD_ORDER, cigar_consistent, load_json

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def checker(built):
    engine.init(0, 1)
    return oracle.ref() if oracle.have_ref() else oracle.port()


def compare(ps, mat, checker, w, zdrop, flag, use_raw=True):
    got = engine.extz2_batch(ps, mat, 40, 1, w, zdrop, flag, use_raw=use_raw)
    _, fr, cr = checker.batch(ps, mat, 40, 1, w, zdrop, flag, nthreads=8)
    for i in range(ps.n):
        assert got.fields(i) == fr[i], (i, int(ps.qlen[i]), int(ps.tlen[i]), got.fields(i), fr[i])
        assert got.cigars[i].tolist() == cr[i], (i, int(ps.qlen[i]), int(ps.tlen[i]))
        if not (flag & engine.KSW_EZ_SCORE_ONLY):
            fwd = cr[i] if not (flag & engine.KSW_EZ_REV_CIGAR) else cr[i][::-1]
            qa, ta = ps.raw_pair(i)
            if not use_raw:
                qa = synth.ASCII[ps.pair(i)[0]]; ta = synth.ASCII[ps.pair(i)[1]]
            assert got.stats_dict(i) == oracle.sd_stats(fwd, qa, ta), i
    return got


def test_kat_table(checker, mat, golden_dir):
    kat = load_json(golden_dir, "ksw2_kat.json")
    ps = synth.pairs_from_strings([(kat["seq1"], kat["seq2"])])
    q, t = ps.pair(0)
    for row in kat["rows"]:
        f, c = engine.extz2(q, t, mat, 40, 1, row["w"], row["zdrop"], row["flag"])       # ksw2-compatible single-pair entry
        assert f == row["fields"], row
        assert oracle.cigar_str(c) == row["cigar"], row


def test_golden_vectors(checker, mat, golden_dir):
    g = load_json(golden_dir, "ksw2_golden.json")
    for grp in g["groups"]:
        ps = synth.pairs_from_strings([tuple(p) for p in grp["pairs"]])
        for run in grp["runs"]:
            got = engine.extz2_batch(ps, mat, 40, 1, run["w"], run["zdrop"], run["flag"])
            for i in range(ps.n):
                f = got.fields(i)
                assert [f[k] for k in FIELD_ORDER] == run["fields"][i], (grp["name"], run, i)
                assert oracle.cigar_str(got.cigars[i].tolist()) == run["cigars"][i], (grp["name"], i)


def test_sedef_alignment_mirror(checker, golden_dir):
    """Alignment(fa, fb) mirror (sedef_b200/align.py) against the reference's own Alignment class."""
    g = load_json(golden_dir, "sd_stats_golden.json")
    res = align.align_pairs([(r["a"], r["b"]) for r in g["records"]])
    for r, a in zip(g["records"], res):
        assert a.cigar_string() == r["cigar"]
        assert (a.span(), a.matches(), a.mismatches(), a.gaps(), a.gap_bases()) == \
               (r["span"], r["matches"], r["mismatches"], r["gaps"], r["gap_bases"])
        for k, v in r["stat_loop"].items():             # the BEDPE stat loop over the reference's own column strings
            assert a.stats[k] == v, (k, r["cigar"])
    kat = load_json(golden_dir, "ksw2_kat.json")
    a = align.align(kat["seq1"], kat["seq2"])
    exp = kat["survey_stat_loop"]
    for k in ("span", "indel_a", "indel_b", "alnB", "matchB", "mismatchB", "transitionsB", "transversionsB",
              "uppercaseA", "uppercaseB", "uppercaseMatches", "gaps", "gap_bases"):
        assert a.stats[k] == exp[k], k
    fp = a.bedpe_fp()
    for k in ("fracMatch", "fracMatchIndel", "jcK", "k2K", "filter_score"):
        assert float("%.6g" % fp[k]) == exp[k]
    assert "%.1f" % a.total_error() == "4.0"


@pytest.mark.parametrize("w,zdrop,flag,kw", [
    (-1, -1, 0, dict(min_len=1, max_len=40, div=0.15)),            # SEDEF's median call size (<= 32 bp)
    (-1, -1, 0, dict(min_len=1, max_len=250, div=0.1)),
    (-1, -1, 0, dict(min_len=200, max_len=520, div=0.1)),          # the 500x500 side extensions
    (-1, 120, 0, dict(min_len=50, max_len=500, div=0.4)),
    (20, -1, 0, dict(min_len=1, max_len=600, div=0.15)),
    (50, 100, 0, dict(min_len=1, max_len=600, div=0.15)),
    (100, -1, 0, dict(min_len=600, max_len=1300, div=0.08)),
    (30, 80, 0x01, dict(min_len=1, max_len=600, div=0.2)),
    (30, 80, 0x02, dict(min_len=1, max_len=600, div=0.2)),
    (30, 80, 0x04, dict(min_len=1, max_len=600, div=0.2)),
    (30, 80, 0x40, dict(min_len=1, max_len=600, div=0.2)),
    (30, 80, 0x80, dict(min_len=1, max_len=600, div=0.2)),
    (30, 80, 0xc2, dict(min_len=1, max_len=600, div=0.2)),
    (5, -1, 0, dict(min_len=1, max_len=600, div=0.3)),
    (3, 20, 0, dict(min_len=1, max_len=600, div=0.3)),
    (1, -1, 0, dict(min_len=1, max_len=300, div=0.3)),
    (0, -1, 0, dict(min_len=1, max_len=100, div=0.1)),
    (17, 30, 0, dict(min_len=100, max_len=900, div=0.4, burst=80)),
    (33, -1, 0x40, dict(min_len=100, max_len=900, div=0.2, burst=100)),
])
def test_fuzz_vs_oracle(checker, mat, w, zdrop, flag, kw):
    ps = synth.make_pairs_mixed(250, seed=9000 + 13 * (w + 2) + flag, **kw)
    compare(ps, mat, checker, w, zdrop, flag)


@pytest.mark.parametrize("w,zdrop,flag,kw", [
    (-1, -1, 0x08, dict(min_len=1, max_len=300, div=0.12)),           # KSW_EZ_APPROX_MAX: score only from the tracked H0, CIGAR end to end
    (30, 80, 0x08, dict(min_len=1, max_len=700, div=0.2)),
    (30, 80, 0x18, dict(min_len=1, max_len=700, div=0.2)),            # + APPROX_DROP: z-drop on the tracked score
    (30, 80, 0x19, dict(min_len=1, max_len=700, div=0.2)),            # score-only
    (30, 60, 0x5a, dict(min_len=1, max_len=700, div=0.3)),            # + RIGHT + EXTZ_ONLY
    (100, 200, 0x98, dict(min_len=500, max_len=1300, div=0.15)),      # + REV_CIGAR, 128-slot class
    (5, 20, 0x18, dict(min_len=1, max_len=600, div=0.3)),
    (-1, 150, 0x18, dict(min_len=300, max_len=1000, div=0.35, burst=100)),   # 1024-slot class
    (-1, 50, 0x1c, dict(min_len=1, max_len=400, div=0.2)),            # + GENERIC_SC
])
def test_approx_max_vs_oracle(checker, mat, w, zdrop, flag, kw):
    """KSW_EZ_APPROX_MAX / KSW_EZ_APPROX_DROP (extern/ksw2_extz2_sse.cc:268-284): one tracked score instead of the exact
    maximum; no mqe / mte; z-drop only with APPROX_DROP."""
    ps = synth.make_pairs_mixed(250, seed=8800 + 7 * (w + 2) + flag, **kw)
    compare(ps, mat, checker, w, zdrop, flag)


def test_approx_max_wide_kernels(checker, mat):
    """The approx-max variant on the CTA-wide (2048 / 4096 / 8192 slots) and cluster (16384 slots) kernels."""
    for (cnt, length, seed) in [(4, 1500, 21), (3, 3000, 22), (2, 6000, 23), (2, 9500, 24)]:
        ps = synth.make_pairs_small(cnt, length=length, div=0.1, seed=seed)
        compare(ps, mat, checker, -1, -1, 0x08)
        compare(ps, mat, checker, -1, 300, 0x18)
        compare(ps, mat, checker, -1, 300, 0x1a)


def test_config2_shape_vs_oracle(checker, mat):
    """BASELINE.json configs[1] shape (1 kbp pairs, w=100, 5 % divergence) at a size the oracle finishes in seconds."""
    ps = synth.make_pairs_small(3000, length=1000, div=0.05, seed=0x5EDEF002)
    compare(ps, mat, checker, 100, -1, 0)


def test_config3_shape_vs_oracle(checker, mat):
    """BASELINE.json configs[2] shape (long pairs, z-drop on, indels) scaled to the widest kernel (w=400)."""
    ps = synth.make_pairs_large(24, min_len=3000, max_len=12000, seed=0x5EDEF003)
    compare(ps, mat, checker, 400, 400, 0)
    compare(ps, mat, checker, 500, 400, 0)              # the configs[2] band: 528 live slots (32 lanes x 32 slots)
    compare(ps, mat, checker, 1200, 400, 0)             # CTA-wide banded


def test_config3_full_length_pairs_vs_oracle(checker, mat):
    """BASELINE.json configs[2] at its real pair sizes (10-50 kbp, w=500, z-drop 400, 15 % divergence with indels): 160 pairs,
    9 GB of traceback rows, the warp-per-pair traceback -- every field, CIGAR and statistic against the oracle."""
    ps = synth.make_pairs_large(160, min_len=10000, max_len=50000, seed=0x5EDEF003)
    compare(ps, mat, checker, 500, 400, 0)


def test_no_raw_bytes_decodes_codes(checker, mat):
    ps = synth.make_pairs_mixed(100, seed=5, min_len=1, max_len=300, div=0.1)
    compare(ps, mat, checker, -1, -1, 0, use_raw=False)


def test_edge_cases(checker, mat):
    pairs = [("A", "A"), ("A", "C"), ("N", "N"), ("ACGT", "A"), ("A", "ACGT"), ("NNNNNNNNNN", "ACGTACGTAC"),
             ("acgtnACGTN" * 5, "ACGTNacgtn" * 5), ("A" * 16, "A" * 16), ("A" * 17, "A" * 15), ("C" * 15, "C" * 33),
             ("ACGT" * 8, "TGCA" * 8), ("G" * 100, "G"), ("G", "G" * 100)]
    ps = synth.pairs_from_strings(pairs)
    for (w, zd, flag) in [(-1, -1, 0), (2, -1, 0), (0, -1, 0), (4, 5, 0), (-1, -1, 2)]:
        compare(ps, mat, checker, w, zd, flag)


def test_empty_and_degenerate_inputs(checker, mat):
    # n = 0
    ps0 = synth.pairs_from_strings([])
    r0 = engine.extz2_batch(ps0, mat, 40, 1)
    assert r0.ez.shape[0] == 0
    # qlen == 0 or tlen == 0 -> reset record (extern/ksw2_extz2_sse.cc:56-57), others unaffected
    ps = synth.pairs_from_strings([("ACGT", "ACGT"), ("", "ACGT"), ("ACGT", ""), ("ACGTACGT", "ACGAACGT")])
    got = engine.extz2_batch(ps, mat, 40, 1)
    reset = dict(max=0, zdropped=0, max_q=-1, max_t=-1, mqe=engine.KSW_NEG_INF, mqe_t=-1, mte=engine.KSW_NEG_INF,
                 mte_q=-1, score=engine.KSW_NEG_INF, n_cigar=0)
    assert got.fields(1) == reset and got.fields(2) == reset
    assert got.fields(0)["score"] == 20 and got.fields(3)["score"] == 31
    # -min_sc > 2(q+e): silent early return (extern/ksw2_extz2_sse.cc:81)
    bad = synth.sedef_matrix(5, -100)
    got = engine.extz2_batch(ps, bad, 40, 1)
    assert all(got.fields(i) == reset for i in range(ps.n))


def test_widest_pair_and_too_wide(checker, mat):
    """Unbanded pairs of growing size walk through every kernel family: 1000x1000 (SEDEF's largest direct gap fill,
    src/align.cc:233-236) and 4 kbp on the CTA-wide kernels, 6 kbp / 10 kbp (SEDEF's MAX_GAP fills, src/refine.cc:77) on the
    thread-block-cluster kernels (DSMEM); beyond 16384 live slots the engine refuses (no silent fallback)."""
    ps = synth.make_pairs_small(6, length=1000, div=0.1, seed=3)
    compare(ps, mat, checker, -1, -1, 0)
    ps4k = synth.make_pairs_small(3, length=4000, div=0.1, seed=8)
    ps4k.tlen[:] = np.minimum(ps4k.tlen, 4090)
    compare(ps4k, mat, checker, -1, -1, 0)
    compare(synth.make_pairs_small(3, length=6000, div=0.08, seed=9), mat, checker, -1, -1, 0)        # cluster of 2 CTAs
    compare(synth.make_pairs_small(2, length=10000, div=0.08, seed=10), mat, checker, -1, -1, 0)      # cluster of 4 CTAs
    compare(synth.make_pairs_small(2, length=9000, div=0.3, seed=11), mat, checker, -1, 500, 0x42)    # z-drop, right, extz-only
    compare(synth.make_pairs_large(3, min_len=12000, max_len=20000, seed=12), mat, checker, 5000, 600, 0)   # banded, 5 k wide


def test_largest_reference_call_sizes(checker, mat):
    """Unbanded pairs beyond 16384 live slots run on clusters of 4 and 8 CTAs (32768 / 65536 slots, DSMEM): the engine now covers
    the largest call the reference can make (align_helper chunks at 60 000 bases, src/align.cc:46-53).  One pair per cluster
    size against the compiled reference (the 33 kbp pair is 1.1 G cells and 2.2 GB of reference traceback), and the refusal above."""
    assert engine.load().ksw_b200_max_slots() == 65536
    compare(synth.make_pairs_small(2, length=17000, div=0.05, seed=4), mat, checker, -1, -1, 0)          # cluster of 4 CTAs
    compare(synth.make_pairs_small(1, length=33000, div=0.08, seed=5), mat, checker, -1, 2000, 0x02)     # cluster of 8 CTAs, z-drop, right
    compare(synth.make_pairs_large(2, min_len=40000, max_len=60000, seed=6), mat, checker, 20000, -1, 0) # banded but wider than 16384 slots
    big = synth.make_pairs_small(1, length=66000, div=0.02, seed=7)
    with pytest.raises(engine.EngineError) as ei:
        engine.extz2_batch(big, mat, 40, 1, -1, -1, 0)
    assert ei.value.code == -5


def test_packed_class_boundaries(checker, mat):
    """Pairs that fill the live-slot window of every packed class exactly (32 .. 8192 slots: the window wraps with no
    slack), one slot less, and one block more (next class), unbanded and banded, both traceback arms."""
    for L in (32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):        # 2048 .. 8192: the packed CTA-wide kernel
        ps = synth.make_pairs_small(8, length=L + 60, div=0.08, seed=700 + L)
        ps.tlen[:] = np.minimum(ps.tlen, np.array([L, L, L - 1, L - 15, L - 16, L + 1, L + 16, L], np.int32))
        ps.qlen[:] = np.minimum(ps.qlen, np.array([L + 60, L, L + 7, L + 60, L, L + 60, L + 3, L - 9], np.int32))
        for flag in (0, 0x02):
            compare(ps, mat, checker, -1, -1, flag)
        compare(ps, mat, checker, L // 2, 60, 0)


@pytest.mark.parametrize("w,zdrop,flag,kw", [
    (500, 400, 0, dict(min_len=900, max_len=2600, div=0.15, burst=200)),       # BASELINE.json configs[2]'s band: 33 blocks on 1/4 of the diagonals
    (500, -1, 0, dict(min_len=1100, max_len=1800, div=0.05)),
    (496, 300, 0, dict(min_len=900, max_len=2000, div=0.2, burst=150)),        # w % 16 == 0: block entry and block exit on the SAME diagonal
    (511, -1, 0, dict(min_len=520, max_len=1600, div=0.1)),                    # widest band of the class; the top row reaches the spare block
    (505, 200, 0x02, dict(min_len=700, max_len=1500, div=0.2)),                # right-aligned gaps
    (500, 250, 0x01, dict(min_len=700, max_len=1500, div=0.2)),                # score only
    (503, 150, 0xc0, dict(min_len=700, max_len=1500, div=0.25, burst=120)),    # EXTZ_ONLY + REV_CIGAR: traceback from (max_t, max_q)
    (500, 100, 0x04, dict(min_len=700, max_len=1500, div=0.2)),                # GENERIC_SC: the score fill stops at en0
    (-1, -1, 0, dict(min_len=513, max_len=528, div=0.1)),                      # unbanded, 16 * ceil(tlen / 16) in (512, 528]
    (-1, 200, 0, dict(min_len=505, max_len=528, div=0.3, burst=100)),
])
def test_spare_block_class(checker, mat, w, zdrop, flag, kw):
    """The 16-lane packed class holds 33 blocks: 32 in registers + a SPARE block spread over its lanes (extz_dp16.cuh Spare16),
    so pairs needing 513..528 live slots -- every band of w = 496..511 -- no longer fall into the 1024-slot class.  Activation
    (also ahead of the rounded range, by the score fill), band entry, the carries from the block below, H[en0] inside the spare
    block, the arg-max in it, its traceback codes, the hand-over to the owning lane and two spare blocks in a row."""
    ps = synth.make_pairs_mixed(120, seed=7700 + 3 * (w + 2) + flag, **kw)
    compare(ps, mat, checker, w, zdrop, flag)


def test_spare_block_long_lived(checker, mat):
    """Target of 513..528 bases against a much longer query, unbanded: block 0 stays live for thousands of anti-diagonals, and so
    does the spare block (never handed over until the very end)."""
    rng = np.random.default_rng(77)
    pairs = []
    for _ in range(12):
        tl = int(rng.integers(513, 529)); ql = int(rng.integers(900, 2500))
        t = "".join("ACGT"[k] for k in rng.integers(0, 4, tl))
        q = "".join("ACGT"[k] for k in rng.integers(0, 4, ql))
        k = int(rng.integers(0, ql - 300)); q = q[:k] + t[100:400] + q[k + 300:]       # a shared stretch somewhere
        pairs.append((q, t))
    ps = synth.pairs_from_strings(pairs)
    compare(ps, mat, checker, -1, -1, 0)
    compare(ps, mat, checker, -1, 300, 0x40)


def test_maximum_in_the_block_that_leaves_the_band(checker, mat):
    """The arg-max of a diagonal can sit in the lowest 16-slot block, which leaves the band on the next diagonal (a
    maximum on the last query row is at slot st0).  The pipelined CTA-wide / cluster kernels run prepare(r+1) -- which
    slides that block -- before the arg-max pass of diagonal r; they must scan with the slot bases of diagonal r.
    One pair per kernel family: 32-lane narrow, CTA-wide with 64 / 128 / 256 lanes, cluster of 2 CTAs; with and
    without z-drop, both traceback arms."""
    ps = synth.make_pairs_max_on_last_row([480, 992, 1504, 3008, 6000, 9008], tail=200, seed=5)
    for (zd, flag) in [(-1, 0), (200, 0), (-1, 0x02)]:
        got = compare(ps, mat, checker, -1, zd, flag)
    ps = synth.make_pairs_max_on_last_row([1200 + 16 * k for k in range(12)] + [2800 + 16 * k for k in range(6)], tail=120, seed=6)
    compare(ps, mat, checker, -1, -1, 0)
    chk = checker.batch(ps, mat, 40, 1, -1, -1, 0, nthreads=8)[1]
    assert sum((f["max_t"] + 1) % 16 == 0 and f["max_q"] == int(ps.qlen[i]) - 1 for i, f in enumerate(chk)) >= 12   # the case is hit


def test_one_slot_per_register_kernels(checker, mat, tmp_path):
    """KSW_B200_PACKED=0 routes the narrow classes to the one-slot-per-register kernels of extz_dp.cuh (the A/B switch is
    read once per process, hence the subprocess): they must stay bit-exact too."""
    import os, subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "unpacked.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {root!r})
        import oracle
        from sedef_b200 import engine, synth
        mat = synth.sedef_matrix()
        engine.init(0, 1)
        chk = oracle.ref() if oracle.have_ref() else oracle.port()
        n = 0
        for (kw, w, zd, flag) in [(dict(min_len=1, max_len=700, div=0.12), -1, -1, 0), (dict(min_len=1, max_len=600, div=0.2), 30, 80, 2),
                                  (dict(min_len=300, max_len=1000, div=0.1), 100, -1, 0), (dict(min_len=1, max_len=600, div=0.2), 30, 80, 1)]:
            ps = synth.make_pairs_mixed(300, seed=4242 + w, **kw)
            got = engine.extz2_batch(ps, mat, 40, 1, w, zd, flag)
            _, fr, cr = chk.batch(ps, mat, 40, 1, w, zd, flag, nthreads=8)
            for i in range(ps.n):
                assert got.fields(i) == fr[i], i
                if not flag & 1:
                    assert got.cigars[i].tolist() == cr[i], i
                n += 1
        for (length, cnt) in [(1500, 3), (4500, 2), (9000, 1)]:        # one-slot CTA-wide, cluster x2 and cluster x4 kernels
            ps = synth.make_pairs_small(cnt, length=length, div=0.08, seed=length)
            got = engine.extz2_batch(ps, mat, 40, 1, -1, -1, 0)
            _, fr, cr = chk.batch(ps, mat, 40, 1, -1, -1, 0, nthreads=8)
            for i in range(ps.n):
                assert got.fields(i) == fr[i] and got.cigars[i].tolist() == cr[i], (length, i)
                n += 1
        print("ok", n)
    """))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=dict(os.environ, KSW_B200_PACKED="0"), timeout=600)
    assert out.returncode == 0 and "ok 1206" in out.stdout, out.stdout + out.stderr


def test_warp_traceback_kernel_forced(checker, mat, tmp_path):
    """The one-warp-per-pair traceback (staged 32-row tiles) normally serves long pairs only; KSW_B200_TB_WARP_MIN=0 sends
    every pair through it (read once per process, hence the subprocess), in both traceback-row layouts."""
    import os, subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "warp_tb.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {root!r})
        import oracle
        from sedef_b200 import engine, synth
        mat = synth.sedef_matrix()
        engine.init(0, 1)
        chk = oracle.ref() if oracle.have_ref() else oracle.port()
        n = 0
        for (kw, w, zd, flag) in [(dict(min_len=1, max_len=700, div=0.12), -1, -1, 0), (dict(min_len=1, max_len=600, div=0.2), 30, 80, 0x42),
                                  (dict(min_len=300, max_len=1000, div=0.1), 100, -1, 0x80), (dict(min_len=900, max_len=2500, div=0.1), -1, 200, 0)]:
            ps = synth.make_pairs_mixed(120, seed=777 + w, **kw)
            got = engine.extz2_batch(ps, mat, 40, 1, w, zd, flag)
            _, fr, cr = chk.batch(ps, mat, 40, 1, w, zd, flag, nthreads=8)
            for i in range(ps.n):
                assert got.fields(i) == fr[i], i
                assert got.cigars[i].tolist() == cr[i], i
                if not flag & 0x80:
                    assert got.stats_dict(i) == oracle.sd_stats(cr[i], *ps.raw_pair(i)), i
                n += 1
        print("ok", n)
    """))
    for packed in ("1", "0"):
        env = dict(os.environ, KSW_B200_TB_WARP_MIN="0", KSW_B200_PACKED=packed)
        out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=600)
        assert out.returncode == 0 and "ok 480" in out.stdout, out.stdout + out.stderr


def test_other_scoring_parameters(checker):
    """Scoring is not hard-wired: user-supplied --match/--mismatch/--gap-open/--gap-extend (src/align_main.cc:343-373)."""
    ps = synth.make_pairs_small(300, length=200, div=0.05, seed=871792603, len_jitter=40)
    # the last six have 2(q+e) + match > 127: u / v bytes above 127, which the reference reads as uint8_t (:103,228,255)
    for (ma, mi, go, ge) in [(1, -1, 2, 1), (2, -4, 4, 2), (5, -4, 40, 1), (10, -9, 30, 3), (5, -4, 50, 1),
                             (2, -6, 60, 5), (12, -8, 59, 1), (6, -3, 58, 4), (12, -4, 59, 4), (20, -4, 55, 4), (11, -4, 59, 5)]:
        m = synth.sedef_matrix(ma, mi)
        for (w, flag) in ((-1, 0), (25, 0), (16, 0x02)):
            got = engine.extz2_batch(ps, m, go, ge, w, -1, flag)
            _, fr, cr = checker.batch(ps, m, go, ge, w, -1, flag, nthreads=8)
            for i in range(ps.n):
                assert got.fields(i) == fr[i], (ma, mi, go, ge, w, flag, i)
                assert got.cigars[i].tolist() == cr[i]
    wide = synth.make_pairs_small(3, length=1500, div=0.1, seed=78)           # the CTA-wide kernel shares the lane code
    m = synth.sedef_matrix(12, -8)
    got = engine.extz2_batch(wide, m, 59, 1, -1, -1, 0)
    _, fr, cr = checker.batch(wide, m, 59, 1, -1, -1, 0, nthreads=8)
    for i in range(wide.n):
        assert got.fields(i) == fr[i] and got.cigars[i].tolist() == cr[i], i


def test_degenerate_scoring_many_tied_maxima(checker):
    """Match 1 / mismatch -100 with a large gap cost makes long stretches of an anti-diagonal tie for the maximum while the
    z-drop test (which needs the exact arg-max slot) is armed.  The fast arg-max encodes (count << 24) + sum(t): 257 equal maxima
    must not read as one.  Also gap costs of 0 and scores at the int8 limits: deterministic nonsense in the reference, to be matched."""
    for (ma, mi, go, ge, w, zd, flag, seed, hi) in [(1, -100, 63, 5, 1000, 1000, 0x42, 668295684, 700), (1, -100, 63, 5, -1, 1000, 0, 5, 1100),
                                                    (1, -128, 40, 1, -1, 300, 0, 6, 700), (5, -4, 0, 1, 50, 100, 0, 7, 400),
                                                    (5, -4, 40, 0, -1, -1, 0x02, 8, 400), (127, -100, 120, 10, 100, 1000, 0, 9, 400),
                                                    (100, -60, 90, 2, -1, -1, 0, 10, 400)]:
        ps = synth.make_pairs_mixed(200, seed=seed, min_len=1, max_len=hi, div=0.4)
        m = synth.sedef_matrix(ma, mi)
        got = engine.extz2_batch(ps, m, go, ge, w, zd, flag)
        _, fr, cr = checker.batch(ps, m, go, ge, w, zd, flag, nthreads=8)
        for i in range(ps.n):
            assert got.fields(i) == fr[i], (ma, mi, go, ge, w, zd, hex(flag), i)
            assert got.cigars[i].tolist() == cr[i], (ma, mi, go, ge, i)
    # the deterministic trigger: targets of exactly 257 / 513 bases under match 1 / mismatch -100 -- whole anti-diagonals tie, and
    # a tie count of 257 or 513 is 1 modulo 256 (every pair a build without the clamp got wrong had tlen == 257)
    ps = synth.make_pairs_small(8, length=640, div=0.4, seed=11)
    ps.tlen[:] = np.minimum(ps.tlen, np.array([257, 257, 257, 257, 513, 513, 513, 513], np.int32))
    ps.qlen[:] = np.minimum(ps.qlen, np.array([360, 364, 300, 620, 600, 640, 530, 514], np.int32))
    m = synth.sedef_matrix(1, -100)
    for (w, zd, flag) in [(-1, 300, 0), (-1, 1000, 0), (1000, 1000, 0x42)]:
        got = engine.extz2_batch(ps, m, 63, 5, w, zd, flag)
        _, fr, cr = checker.batch(ps, m, 63, 5, w, zd, flag, nthreads=8)
        for i in range(ps.n):
            assert got.fields(i) == fr[i] and got.cigars[i].tolist() == cr[i], (w, zd, hex(flag), i)


def test_full_size_config2_properties(checker, mat):
    """BASELINE.json configs[1] at FULL size (100k x 1 kbp, w=100): size-independent properties for every pair
    (CIGAR consumes both sequences, stats are consistent with the CIGAR, idempotence across runs) and the
    oracle on a seeded sample."""
    n = 100000
    ps = synth.make_pairs_small(n, length=1000, div=0.05, seed=0x5EDEF002)
    rb = engine.ResidentBatch(ps, mat, 40, 1, 100, -1, 0)
    rb.run()
    a = rb.fetch()
    rb.run()
    b = rb.fetch()
    rb.free()
    for k in ("max_zd", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar"):
        assert np.array_equal(a.ez[k], b.ez[k]), k                                 # idempotent
    assert np.array_equal(a.stats, b.stats)
    qsum = np.zeros(n, np.int64); tsum = np.zeros(n, np.int64); gaps = np.zeros(n, np.int64); gapb = np.zeros(n, np.int64)
    for i in range(n):
        c = a.cigars[i]
        assert np.array_equal(c, b.cigars[i])
        op = c & 0xF; ln = (c >> 4).astype(np.int64)
        qsum[i] = ln[op != 2].sum(); tsum[i] = ln[op != 1].sum()
        gaps[i] = (op != 0).sum(); gapb[i] = ln[op != 0].sum()
    zd = (a.ez["max_zd"] >> 31).astype(bool)
    assert not zd.any()                                                             # band 100 never breaks at 5 % / 1 bp indels
    assert np.array_equal(qsum, ps.qlen) and np.array_equal(tsum, ps.tlen)
    st = a.stats
    assert np.array_equal(st["gaps"], gaps) and np.array_equal(st["gap_bases"], gapb)
    assert np.array_equal(st["span"], st["alnB"] + st["gap_bases"])
    assert np.array_equal(st["alnB"], st["matchB"] + st["mismatchB"])
    assert np.array_equal(st["alnB"], st["matches"] + st["mismatches"])
    assert np.array_equal(st["mismatchB"], st["transitionsB"] + st["transversionsB"])
    assert np.array_equal(st["indel_a"] + st["indel_b"], st["gap_bases"])
    assert (a.ez["score"] == a.ez["mqe"]).sum() > 0
    sample = np.random.default_rng(5).choice(n, 1500, replace=False)
    sub = ps.subset(sample)
    _, fr, cr = checker.batch(sub, mat, 40, 1, 100, -1, 0, nthreads=8)
    for k, i in enumerate(sample):
        assert a.fields(int(i)) == fr[k], int(i)
        assert a.cigars[int(i)].tolist() == cr[k], int(i)


def test_pipelined_one_shot_matches_oracle(checker, mat, monkeypatch):
    """Large one-shot batches are cut into chunks that flow through the upload/launch/fetch pipeline
    (ksw_extz2_batch_flat); force small chunks and check every pair, in original order, against the oracle."""
    monkeypatch.setenv("KSW_B200_CHUNK_PAIRS", "700")
    ps = synth.make_pairs_mixed(5000, seed=424242, min_len=1, max_len=260, div=0.12)
    got = compare(ps, mat, checker, 40, 60, 0)
    assert engine.last_call_io()[2] >= 8            # several chunks -> several kernel launches
    monkeypatch.setenv("KSW_B200_CHUNK_PAIRS", "100000000")
    one = engine.extz2_batch(ps, mat, 40, 1, 40, 60, 0)
    for k in ("max_zd", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar"):
        assert np.array_equal(got.ez[k], one.ez[k]), k
    assert np.array_equal(got.stats, one.stats)


def _fast_align(libpath, q, t):
    import ctypes as C
    lib = C.CDLL(libpath)
    lib.ref_fast_align.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(1 << 22)
    n = lib.ref_fast_align(q.encode(), t.encode(), 11, buf, len(buf))
    return n, buf.value.decode(), lib


def test_reference_call_sites_with_cuda_kernel(checker, golden_dir):
    """Drop-in check at the reference's own call sites: the UNMODIFIED reference align stage (fast_align: anchors,
    chaining, guide constructors, refine/merge, side extensions) linked against the product library through the
    one-line binding of INTEGRATION.md (oracle/ksw_redirect.c) must produce the same hits and CIGARs as with its
    own SSE kernel, and as the committed golden."""
    import os
    ref_dir = os.path.join(os.path.dirname(oracle.__file__), "_ref")
    sse, b200 = os.path.join(ref_dir, "libsedef_ref.so"), os.path.join(ref_dir, "libsedef_ref_b200.so")
    if not (os.path.exists(sse) and os.path.exists(b200)):
        pytest.skip("oracle/_ref/libsedef_ref*.so not built")
    g = load_json(golden_dir, "fast_align_golden.json")
    for reg in g["regions"]:
        q, t = synth.make_region_pair(reg["length"], reg["div"], seed=reg["seed"])
        n1, out1, _ = _fast_align(sse, q, t)
        n2, out2, lib2 = _fast_align(b200, q, t)
        assert lib2.ksw_redirect_calls() > 0            # the CUDA kernel really was underneath
        assert (n1, out1) == (reg["n_hits"], reg["hits"])
        assert (n2, out2) == (n1, out1)


def test_from_cigar_statistics(checker, mat, golden_dir):
    """Alignment(fa, fb, cigar) ("from_cigar", src/align.cc:90-105): statistics from EXISTING CIGARs on the GPU, against the
    oracle port, the reference's own class (when oracle/_ref is present) and the golden records."""
    import ctypes as C, os
    g = load_json(golden_dir, "sd_stats_golden.json")
    pairs = [(r["a"], r["b"]) for r in g["records"]]
    res = align.from_cigars(pairs, [r["cigar"] for r in g["records"]])
    path = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libsedef_ref.so")
    lib = C.CDLL(path) if os.path.exists(path) else None
    for r, a in zip(g["records"], res):
        assert a.cigar_string() == r["cigar"]
        assert (a.span(), a.matches(), a.mismatches(), a.gaps(), a.gap_bases()) == (r["span"], r["matches"], r["mismatches"], r["gaps"], r["gap_bases"])
        raw = [(n << 4) | "MDI".index(op) for op, n in a.cigar]
        assert a.stats == oracle.sd_stats(raw, np.frombuffer(r["a"].encode(), np.uint8), np.frombuffer(r["b"].encode(), np.uint8))
        if lib is not None:
            v = [C.c_int(0) for _ in range(5)]
            lib.ref_alignment_from_cigar(r["a"].encode(), r["b"].encode(), r["cigar"].encode(), *[C.byref(x) for x in v])
            assert [x.value for x in v] == [a.span(), a.matches(), a.mismatches(), a.gaps(), a.gap_bases()]
    # a CIGAR that overruns its sequences is reported, not silently accepted
    with pytest.raises(ValueError):
        align.from_cigars([("ACGT", "ACGT")], ["9M"])


def test_cpp_host_layer(checker, golden_dir):
    """The C++ host layer (include/sedef_align.hpp: AlignQueue / from_cigar_batch) driven from C++ (tests/cpp), against the
    golden records produced by the reference's own Alignment class and the oracle's statistics."""
    import os, subprocess
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cpp", "align_queue_driver")
    assert os.path.exists(drv), "build() did not produce tests/cpp/align_queue_driver"
    g = load_json(golden_dir, "sd_stats_golden.json")
    text = "".join(f"{r['a']} {r['b']} {r['cigar']}\n" for r in g["records"])
    for mode in ("queue", "from_cigar"):
        out = subprocess.run([drv, mode], input=text, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        lines = out.stdout.strip().split("\n")
        assert len(lines) == len(g["records"])
        for r, ln in zip(g["records"], lines):
            f = ln.split()
            assert f[0] == r["cigar"], mode
            assert [int(x) for x in f[1:6]] == [r["span"], r["matches"], r["mismatches"], r["gaps"], r["gap_bases"]]
            raw = [(int(n) << 4) | "MDI".index(op) for n, op in __import__("re").findall(r"(\d+)([MDI])", r["cigar"])]
            st = oracle.sd_stats(raw, np.frombuffer(r["a"].encode(), np.uint8), np.frombuffer(r["b"].encode(), np.uint8))
            keys = ["span", "matches", "mismatches", "gaps", "gap_bases", "indel_a", "indel_b", "alnB", "matchB", "mismatchB",
                    "transitionsB", "transversionsB", "uppercaseA", "uppercaseB", "uppercaseMatches"]
            assert [int(x) for x in f[1:16]] == [st[k] for k in keys]
            tot = st["matches"] + st["gap_bases"] + st["mismatches"]
            assert f[16] == "%.1f" % (100.0 * st["mismatches"] / tot + 100.0 * st["gap_bases"] / tot)


def test_traceback_waves_and_multi_device(checker, mat, tmp_path):
    """(1) A tiny traceback budget forces every class into many waves (KSW_B200_TB_BUDGET_MB is read at init, hence the
    subprocess); (2) when more than one GPU is visible, the in-process LPT sharding over all devices gives the same results."""
    import os, subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "waves.py"
    script.write_text(textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {root!r})
        import oracle
        from sedef_b200 import engine, synth
        import torch
        mat = synth.sedef_matrix()
        ndev = engine.init(0, int(sys.argv[1]))
        ps = synth.make_pairs_mixed(600, seed=99, min_len=1, max_len=700, div=0.12)
        chk = oracle.ref() if oracle.have_ref() else oracle.port()
        for (w, zd, flag) in [(-1, -1, 0), (30, 60, 0)]:
            got = engine.extz2_batch(ps, mat, 40, 1, w, zd, flag)
            _, fr, cr = chk.batch(ps, mat, 40, 1, w, zd, flag, nthreads=8)
            for i in range(ps.n):
                assert got.fields(i) == fr[i], i
                assert got.cigars[i].tolist() == cr[i], i
                assert got.stats_dict(i) == oracle.sd_stats(cr[i], *ps.raw_pair(i)), i
        print("ok devices", ndev, "launches", engine.last_call_io()[2])
    """))
    env = dict(os.environ, KSW_B200_TB_BUDGET_MB="4")
    out = subprocess.run([sys.executable, str(script), "1"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert int(out.stdout.split("launches")[1]) > 4, out.stdout           # w=30 needs 2 classes = 4 launches without waves
    import torch
    if torch.cuda.device_count() > 1:
        out = subprocess.run([sys.executable, str(script), "0"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "ok devices" in out.stdout, out.stdout + out.stderr
        assert int(out.stdout.split("devices")[1].split()[0]) == torch.cuda.device_count()


def test_chain_wave_batched(checker, golden_dir):
    """SURVEY section 8 f1, first wave: the C++ `align_chains_batch` (every anchor-gap fill of every chain in ONE batched
    ksw_extz2 call + one statistics-from-CIGAR call) against the reference's own Alignment(query, ref, anchors, guide_idx)
    constructor run on the reference's own anchors and chains (tests/golden/chain_wave_golden.json)."""
    import os, subprocess
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cpp", "align_queue_driver")
    g = load_json(golden_dir, "chain_wave_golden.json")
    total = 0
    for reg in g["regions"]:
        q, t = synth.make_region_pair(reg["length"], reg["div"], seed=reg["seed"])
        lines = [ln for ln in reg["chains"].split("\n") if ln.strip()]
        assert len(lines) == reg["n_chains"]
        text = q + "\n" + t + "\n" + "".join(" ".join(ln.split()[10:]) + "\n" for ln in lines)
        out = subprocess.run([drv, "chains"], input=text, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        got = [ln for ln in out.stdout.split("\n") if ln.strip()]
        assert got == [" ".join(ln.split()[:10]) for ln in lines]
        total += len(lines)
    assert total >= 10


def test_refine_wave_hit_guides_batched(checker, golden_dir):
    """SURVEY section 8 f1, second wave: the C++ `align_hit_guides_batch` (gap fills between guide hits + the two +-side
    extensions in ONE batched ksw_extz2 call, trim_front / trim_back on the host) against the reference's own
    Alignment(qstr, rstr, vector<Hit> guide, side) constructor (tests/golden/hit_guide_golden.json)."""
    import os, subprocess
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cpp", "align_queue_driver")
    g = load_json(golden_dir, "hit_guide_golden.json")
    for reg in g["regions"]:
        q, t = synth.make_region_pair(reg["length"], reg["div"], seed=reg["seed"])
        lines = [ln for ln in reg["text"].split("\n") if ln.strip()]
        assert len(lines) == reg["n_guide"] + 1
        text = q + "\n" + t + "\n" + str(reg["side"]) + "\n" + "\n".join(lines[1:]) + "\n"
        out = subprocess.run([drv, "hitguide"], input=text, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        assert out.stdout.strip() == lines[0], (reg["seed"], reg["side"])


def test_merge_batched(checker, golden_dir):
    """the C++ `merge_batch` (overlap trimming on the host, every gap fill in ONE batched call) against the
    reference's own Alignment::merge on overlapping pairs of its own chain alignments (tests/golden/merge_golden.json)."""
    import os, subprocess
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cpp", "align_queue_driver")
    g = load_json(golden_dir, "merge_golden.json")
    total = 0
    for reg in g["regions"]:
        q, t = synth.make_region_pair(reg["length"], reg["div"], seed=reg["seed"])
        lines = [ln for ln in reg["text"].split("\n") if ln.strip()]
        want = [ln for ln in lines if ln.startswith("M ")]
        assert len(want) == reg["n_merges"]
        if not want:
            continue
        text = q + "\n" + t + "\n" + "\n".join(ln for ln in lines if ln[0] in "PC") + "\n"
        out = subprocess.run([drv, "merge"], input=text, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        assert [ln for ln in out.stdout.split("\n") if ln.strip()] == want, reg["seed"]
        total += len(want)
    assert total >= 10


def test_region_driver_matches_fast_align(checker, golden_dir):
    """the region-level driver `refine_regions_batch` (chain wave -> refine DP on the host -> merge levels ->
    final guide constructors, ALL regions advancing together through batched ksw_extz2 calls) on the reference's own anchors and
    chains must return exactly the hits of the reference's fast_align (src/chain.cc:203-268, src/refine.cc:23-193) -- same
    chromosome and different chromosome seeds (tests/golden/region_golden.json)."""
    import os, subprocess
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cpp", "align_queue_driver")
    g = load_json(golden_dir, "region_golden.json")
    text, want = [], []
    for reg in g["regions"]:
        q, t = synth.make_region_pair(reg["length"], reg["div"], seed=reg["seed"])
        lines = [ln for ln in reg["text"].split("\n") if ln.strip()]
        text.append("R %d %d %d\n%s\n%s\n" % (reg["same_chr"], reg["orig_qs"], reg["orig_rs"], q, t))
        text.append("\n".join(ln for ln in lines if ln[0] in "AC") + "\nE\n")
        want.append([ln for ln in lines if ln[0] == "H"])
        assert len(want[-1]) == reg["n_hits"]
    out = subprocess.run([drv, "regions"], input="".join(text), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    got, cur = [], None
    for ln in out.stdout.split("\n"):
        if ln.startswith("R "):
            cur = []; got.append(cur)
        elif ln.startswith("H "):
            cur.append(ln)
        elif ln.startswith("S "):
            rounds, calls, reqs = map(int, ln.split()[1:])
    assert len(got) == len(want)
    for k, (a, b) in enumerate(zip(got, want)):
        assert a == b, (k, g["regions"][k]["seed"], g["regions"][k]["same_chr"])
    # all 13 regions shared their waves: far fewer batched calls than the ~700 synchronous kernel calls of one region alone
    assert rounds <= 40 and calls <= 2 * rounds + 2 and reqs >= 100, (rounds, calls, reqs)


def test_config1_genome_align_stage(checker):
    """BASELINE.json configs[0] shape at the align stage: a 2 Mbp soft-masked chromosome with 40 planted 5-20 kbp duplications at
    2-10 % divergence; one seed hit per planted copy (windows with slop, same chromosome, real coordinates) -> the reference's own
    anchors and chains (live, through oracle/_ref/libsedef_ref.so) -> `refine_regions_batch` with ALL regions in one call.  The
    hits must equal the reference's fast_align hits region by region, and in genome coordinates recover the planted catalog."""
    import ctypes as C, os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "libsedef_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libsedef_ref.so not built")
    slib = C.CDLL(path)
    slib.ref_region.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    g, catalog = synth.make_genome_with_dups()
    rng = np.random.default_rng(17)
    buf = C.create_string_buffer(1 << 24)
    text, want, origin = [], [], []
    for (s0, s1, d0, d1, div) in catalog:
        qs, qe = max(0, s0 - int(rng.integers(300, 900))), min(len(g), s1 + int(rng.integers(300, 900)))
        rs, re_ = max(0, d0 - int(rng.integers(300, 900))), min(len(g), d1 + int(rng.integers(300, 900)))
        q, r = g[qs:qe].tobytes(), g[rs:re_].tobytes()
        n = slib.ref_region(q, r, 11, 1, qs, rs, buf, len(buf))
        assert n >= 0
        lines = [ln for ln in buf.value.decode().split("\n") if ln.strip()]
        text.append("R 1 %d %d\n%s\n%s\n" % (qs, rs, q.decode(), r.decode()) + "\n".join(ln for ln in lines if ln[0] in "AC") + "\nE\n")
        want.append([ln for ln in lines if ln[0] == "H"])
        origin.append((qs, rs))
    drv = os.path.join(root, "tests", "cpp", "align_queue_driver")
    out = subprocess.run([drv, "regions"], input="".join(text), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    got, cur = [], None
    for ln in out.stdout.split("\n"):
        if ln.startswith("R "):
            cur = []; got.append(cur)
        elif ln.startswith("H "):
            cur.append(ln)
    assert len(got) == len(want) == 40
    recovered = 0
    for k, (a, b) in enumerate(zip(got, want)):
        assert a == b, (k, catalog[k])
        s0, s1, d0, d1, _ = catalog[k]
        for ln in a:                                                    # genome coordinates, as align_main prints them
            f = ln.split()
            hq0, hq1, hr0, hr1 = origin[k][0] + int(f[1]), origin[k][0] + int(f[2]), origin[k][1] + int(f[3]), origin[k][1] + int(f[4])
            if hq0 <= s0 + 50 and hq1 >= s1 - 50 and hr0 <= d0 + 50 and hr1 >= d1 - 50:
                recovered += 1
    assert recovered >= 38, recovered


def _run_fastalign(text):
    import os, subprocess
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cpp", "align_queue_driver")
    out = subprocess.run([drv, "fastalign"], input=text, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    got, cur = [], None
    for ln in out.stdout.split("\n"):
        if ln.startswith("R "):
            cur = []; got.append(cur)
        elif ln.startswith("H "):
            cur.append(ln)
    return got


def test_fast_align_batch_complete_path(checker, golden_dir):
    """The COMPLETE align-stage path without any reference code underneath: `fast_align_batch` = anchors on the GPU
    (sedef_anchors_batch) + chaining on the host (chain_anchors) + the region-level driver, given nothing but the region strings
    and the seed's origin.  Hits must equal the reference's fast_align (src/chain.cc:203-268) on the 13 golden regions."""
    g = load_json(golden_dir, "region_golden.json")
    text, want = [], []
    for reg in g["regions"]:
        q, t = synth.make_region_pair(reg["length"], reg["div"], seed=reg["seed"])
        text.append("R %d %d %d\n%s\n%s\n" % (reg["same_chr"], reg["orig_qs"], reg["orig_rs"], q, t))
        want.append([ln for ln in reg["text"].split("\n") if ln.startswith("H ")])
    got = _run_fastalign("".join(text))
    assert len(got) == len(want)
    for k, (a, b) in enumerate(zip(got, want)):
        assert a == b, (k, g["regions"][k]["seed"], g["regions"][k]["same_chr"])


def test_config1_genome_complete_path(checker):
    """BASELINE.json configs[0] shape through the complete path: 2 Mbp chromosome, 40 planted duplications, one seed window per
    copy -> `fast_align_batch` -> hits; compared with the reference's fast_align on the same windows (live, oracle/_ref) and with
    the planted catalog in genome coordinates."""
    import ctypes as C, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "libsedef_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libsedef_ref.so not built")
    slib = C.CDLL(path)
    slib.ref_region.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    g, catalog = synth.make_genome_with_dups(seed=0x5EDEF011)
    rng = np.random.default_rng(23)
    buf = C.create_string_buffer(1 << 24)
    text, want = [], []
    for (s0, s1, d0, d1, div) in catalog:
        qs, qe = max(0, s0 - int(rng.integers(300, 900))), min(len(g), s1 + int(rng.integers(300, 900)))
        rs, re_ = max(0, d0 - int(rng.integers(300, 900))), min(len(g), d1 + int(rng.integers(300, 900)))
        q, r = g[qs:qe].tobytes(), g[rs:re_].tobytes()
        assert slib.ref_region(q, r, 11, 1, qs, rs, buf, len(buf)) >= 0
        want.append([ln for ln in buf.value.decode().split("\n") if ln.startswith("H ")])
        text.append("R 1 %d %d\n%s\n%s\n" % (qs, rs, q.decode(), r.decode()))
    got = _run_fastalign("".join(text))
    assert len(got) == 40
    for k, (a, b) in enumerate(zip(got, want)):
        assert a == b, (k, catalog[k])
    assert sum(len(a) for a in got) >= 38
