"""The align stage's driver (`sedef align generate`, src/align_main.cc:285-337) against the reference BINARY: byte-level
comparison of *.aligned.bed text (SURVEY.md section 8d "Configs 1/4/5", section 8 f2).

CPU tests: the text / FASTA layer of the driver (get_sequence, BED parse / print, schedule order, rc) against the golden fixture
written by the reference binary and against plain slicing.  GPU tests: whole bucket files through `sedef_b200_align_generate`
(anchors on the GPU, chaining on the host, every alignment wave one batched ksw_extz2 call) -- output bytes must equal the
reference binary's, from the committed fixture and, where oracle/_ref/sedef_ref travelled to the box, from a live run on the
BASELINE.json configs[0] genome (2 Mbp, 40 planted duplications)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from helpers import load_json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sedef_ref")


@pytest.fixture(scope="module")
def small_stage(built, golden_dir, tmp_path_factory):
    """The fixture's genome regenerated from its seed (sha1 pinned) + its bucket files on disk."""
    from sedef_b200 import genome
    g = load_json(golden_dir, "align_stage_golden.json")
    wd = str(tmp_path_factory.mktemp("align_stage"))
    fa, bed, _ = genome.write_align_stage_input(wd, **g["config"])
    assert hashlib.sha1(open(fa, "rb").read()).hexdigest() == g["genome_sha1"], "generator drifted: regenerate the fixture"
    assert open(bed).read() == g["seeds"]
    bdir = os.path.join(wd, "buckets")
    os.makedirs(bdir)
    for name, text in g["buckets"].items():
        with open(os.path.join(bdir, name + ".bed"), "w") as f:
            f.write(text)
    return dict(golden=g, fasta=fa, bdir=bdir, wd=wd)


def test_fasta_fetch_matches_slicing(built, tmp_path):
    """FastaReference::get_sequence (src/fasta.cc:106-143): ranges across line ends, clamping of start < 0 and end > length."""
    from sedef_b200 import engine, genome
    rng = np.random.default_rng(5)
    chroms = {"c1": genome.synth.ASCII[rng.integers(0, 4, 1234)], "c2 extra words": genome.synth.ASCII_LOWER[rng.integers(0, 4, 61)],
              "c3": genome.synth.ASCII[rng.integers(0, 4, 60)]}
    fa = str(tmp_path / "t.fa")
    genome.write_fasta(fa, chroms, line=60)
    # the index is keyed by the first token of the name (src/fasta.cc:40)
    for key, name in (("c1", "c1"), ("c2", "c2 extra words"), ("c3", "c3")):
        seq = chroms[name].tobytes()
        n = len(seq)
        cases = [(0, n), (0, 1), (59, 61), (60, 120), (-5, 10), (n - 3, n + 50), (17, 17), (1, n - 1)]
        cases += [tuple(sorted(rng.integers(0, n + 1, 2).tolist())) for _ in range(60)]
        for s, e in cases:
            got, end = engine.fasta_fetch(fa, key, s, e)
            assert got == seq[max(s, 0):min(e, n)], (key, s, e)
            assert end == min(e, n)
    with pytest.raises(engine.EngineError, match="Chromosome nope does not exist"):
        engine.fasta_fetch(fa, "nope", 0, 10)


def test_reverse_complement(built):
    from sedef_b200 import engine
    assert engine.reverse_complement(b"ACGTacgtNnXx-") == b"NNNNNacgtACGT"      # rev_dna: everything else -> 'N' (src/common.h:72-86)
    assert engine.reverse_complement(b"") == b""


def test_schedule_and_bed_text(small_stage):
    """Hit::from_bed -> Hit::to_bed(false) round trip and the processing order of generate_alignments: the seed columns the reference
    binary printed behind each hit (h.to_bed(0)) appear in the product's schedule in the same order, with identical text up to the
    clamped end coordinates (those are clamped while the regions are cut)."""
    from sedef_b200 import engine
    g = small_stage["golden"]
    for name, text in g["buckets"].items():
        path = os.path.join(small_stage["bdir"], name + ".bed")
        sched = [ln for ln in engine.bed_schedule(path).split("\n") if ln]
        # bucket lines were written by the reference's own to_bed(false): parse + print must reproduce them
        lines = [ln for ln in text.split("\n") if ln]
        assert sorted(sched) == sorted(lines)
        # order: by complexity bin (sqrt(span_q * span_r) / 1000), stable inside a bin
        def cplx(ln):
            f = ln.split("\t")
            return int((float(int(f[2]) - int(f[1])) * float(int(f[5]) - int(f[4]))) ** 0.5) // 1000
        want = [ln for _, _, ln in sorted((cplx(ln), i, ln) for i, ln in enumerate(lines))]
        assert sched == want
        seen = []
        for ln in g["aligned"][name].split("\n"):
            if ln:
                seed_name = ln.split("\t")[14 + 6]
                if not seen or seen[-1] != seed_name:
                    seen.append(seed_name)
        order = [ln.split("\t")[6] for ln in sched]
        it = iter(order)
        assert all(s in it for s in seen), (seen, order)        # subsequence, same relative order
    # a directory of *.bed files is read as one schedule
    assert len([ln for ln in engine.bed_schedule(small_stage["bdir"]).split("\n") if ln]) == sum(t.count("\n") for t in g["buckets"].values())
    with pytest.raises(engine.EngineError, match="neither file nor directory"):
        engine.bed_schedule(os.path.join(small_stage["wd"], "missing"))


def test_driver_error_paths_and_empty_inputs(built, tmp_path):
    """What the two drivers do before any alignment is needed (no device involved): an empty bucket file gives an empty
    *.aligned.bed / a header-only report, like the reference's loops over zero hits; unreadable inputs fail with the reference's
    messages (src/fasta.cc:70-72, src/align_main.cc:228-230, src/hit.cc:31)."""
    from sedef_b200 import engine, genome
    rng = np.random.default_rng(3)
    fa = str(tmp_path / "g.fa")
    genome.write_fasta(fa, {"c1": genome.synth.ASCII[rng.integers(0, 4, 5000)]})
    empty = str(tmp_path / "empty.bed")
    open(empty, "w").close()
    out = str(tmp_path / "o.bed")
    st = engine.align_generate(fa, empty, out)
    assert st["regions"] == 0 and st["hits"] == 0 and open(out).read() == ""
    cnt = engine.stats_generate(fa, empty, out)
    assert cnt == dict(hits=0, pieces=0, lines=0)
    text = open(out).read()
    assert text.startswith("#chr1\tstart1\tend1\tchr2") and text.count("\n") == 1 and text.rstrip("\n").count("\t") == 33
    with pytest.raises(engine.EngineError, match="Cannot open file"):
        engine.align_generate(str(tmp_path / "missing.fa"), empty, out)
    with pytest.raises(engine.EngineError, match="neither file nor directory"):
        engine.align_generate(fa, str(tmp_path / "missing.bed"), out)
    short = str(tmp_path / "short.bed")
    with open(short, "w") as f:
        f.write("c1\t0\t100\tc1\t200\t300\n")
    with pytest.raises(engine.EngineError, match="fewer than 10 columns"):
        engine.align_generate(fa, short, out)
    with pytest.raises(engine.EngineError, match="does not exist"):
        engine.stats_generate(fa, str(tmp_path / "missing.bed"), out)


def test_stats_pieces_match_reference_report(built, golden_dir, tmp_path):
    """The host-only half of `stats generate` (no device): Alignment(fa, fb, cigar), the split at assembly gaps -- and, with
    --max-ok-gap, the recursive split at large gaps -- and the re-trimming of every piece (subhit: cigar_from_alignment, trim_back,
    trim_front, src/stats_main.cc:32-211).  Every line of the reference binary's report must be one of the pieces: same
    coordinates, strands, alignment length and CIGAR (tests/golden/stats_golden.json; the pieces the report leaves out are the ones
    its filters reject)."""
    from sedef_b200 import engine, genome
    g = load_json(golden_dir, "stats_golden.json")
    wd = str(tmp_path)
    fa, _, _ = genome.write_align_stage_input(wd, **g["config"])
    assert hashlib.sha1(open(fa, "rb").read()).hexdigest() == g["genome_sha1"], "generator drifted: regenerate the fixture"
    ab = os.path.join(wd, "aligned.bed")
    with open(ab, "w") as f:
        f.write(g["aligned"])
    n_split = 0
    for name, extra in g["variants"].items():
        kw = {}
        for k, v in zip(extra[::2], extra[1::2]):
            kw[{"--max-ok-gap": "max_ok_gap", "--min-split": "min_split"}[k]] = int(v)
        pieces = engine.stats_pieces(fa, ab, **kw)
        have = set(pieces)
        lines = [ln.split("\t") for ln in g["reports"][name].split("\n")[1:] if ln]
        assert lines
        for f in lines:
            key = (f[0], int(f[1]), int(f[2]), f[3], int(f[4]), int(f[5]), f[8], f[9], int(f[11]), f[32])
            assert key in have, (name, key[:9])
        assert len(pieces) >= len(lines)
        n_split += len(pieces) - g["aligned"].count("\n")
    assert n_split > 0                                              # the fixture does split hits


@pytest.mark.gpu
def test_align_generate_matches_reference_binary_golden(small_stage):
    """Whole bucket files: output bytes == the reference binary's *.aligned.bed (tests/golden/align_stage_golden.json: two
    chromosomes, both strands, regions clamped at chromosome ends, same- and different-chromosome seeds)."""
    from sedef_b200 import engine
    g = small_stage["golden"]
    total = 0
    for name in sorted(g["buckets"]):
        out = os.path.join(small_stage["wd"], name + ".aligned.bed")
        st = engine.align_generate(small_stage["fasta"], os.path.join(small_stage["bdir"], name + ".bed"), out)
        got = open(out).read()
        assert got == g["aligned"][name], name
        assert st["regions"] == g["buckets"][name].count("\n") and st["hits"] == got.count("\n")
        total += st["hits"]
    assert total == g["n_lines"]
    # one process per GPU: the shards' outputs together are the same lines
    parts = []
    for k in range(2):
        out = os.path.join(small_stage["wd"], "shard%d.bed" % k)
        engine.align_generate(small_stage["fasta"], small_stage["bdir"], out, shard_index=k, shard_count=2)
        parts += [ln for ln in open(out).read().split("\n") if ln]
    want = [ln for t in g["aligned"].values() for ln in t.split("\n") if ln]
    assert sorted(parts) == sorted(want)
    with pytest.raises(engine.EngineError, match="Chromosome"):
        bad = os.path.join(small_stage["wd"], "bad.bed")
        with open(bad, "w") as f:
            f.write("chrZ\t0\t5000\tchrA\t0\t5000\tx\t\t+\t+\n")
        engine.align_generate(small_stage["fasta"], bad, os.path.join(small_stage["wd"], "bad.out"))


@pytest.mark.gpu
def test_align_generate_degenerate_regions(built, tmp_path):
    """Seed hits whose regions are shorter than a k-mer, all N, or unrelated sequence: no hits and no output for them, like the
    reference's fast_align on such regions, next to an ordinary region that does align.  (Coordinates past the chromosome end are not
    in the comparison: the reference binary crashes on them, this driver cuts an empty region and finds nothing.)"""
    from sedef_b200 import engine, genome
    rng = np.random.default_rng(11)
    core = genome.synth.ASCII[rng.integers(0, 4, 6000)]
    c1 = np.concatenate([genome.synth.ASCII[rng.integers(0, 4, 3000)], core, genome.synth.ASCII[rng.integers(0, 4, 3000)]])
    c2 = np.concatenate([genome.synth.ASCII[rng.integers(0, 4, 2000)], core, genome.synth.ASCII[rng.integers(0, 4, 2000)], np.full(3000, ord("N"), np.uint8)])
    fa = str(tmp_path / "d.fa")
    genome.write_fasta(fa, {"c1": c1, "c2": c2})
    bed = str(tmp_path / "d.bed")
    with open(bed, "w") as f:
        f.write("c1\t2500\t9500\tc2\t1500\t8500\tgood\t\t+\t+\n")          # the shared 6 kbp core
        f.write("c1\t100\t105\tc2\t100\t105\ttiny\t\t+\t+\n")               # shorter than k = 11
        f.write("c1\t100\t3000\tc2\t10000\t13000\tall_n\t\t+\t+\n")          # reference region of N only
        f.write("c1\t0\t2900\tc2\t0\t1900\tunrelated\t\t+\t-\n")              # random against random, reverse strand
    out = str(tmp_path / "d.out")
    st = engine.align_generate(fa, bed, out)
    lines = [ln for ln in open(out).read().split("\n") if ln]
    assert st["regions"] == 4
    assert len(lines) >= 1 and all(ln.split("\t")[20] == "good" for ln in lines), [ln.split("\t")[20] for ln in lines]
    f = lines[0].split("\t")
    assert int(f[1]) <= 3050 and int(f[2]) >= 8950 and int(f[4]) <= 2050 and int(f[5]) >= 7950      # the planted core, in genome coordinates
    if os.path.exists(REF_BIN):
        ref = subprocess.run([REF_BIN, "align", "generate", "-k", "11", fa, bed], check=True, capture_output=True, text=True, timeout=600).stdout
        assert open(out).read() == ref
    past = str(tmp_path / "past.bed")
    with open(past, "w") as f:
        f.write("c1\t50000\t60000\tc2\t100\t5000\tpast_the_end\t\t+\t+\n")
    st = engine.align_generate(fa, past, out)
    assert st["regions"] == 1 and st["hits"] == 0 and open(out).read() == ""


@pytest.mark.gpu
def test_align_generate_config1_live_reference(built, tmp_path):
    """BASELINE.json configs[0] at the align stage, end to end at the file level: 2 Mbp soft-masked chromosome, 40 planted 5-20 kbp
    duplications at 2-10 % -> seed BED -> the reference binary's `align bucket` -> per bucket `align generate -k 11` with the
    reference binary (live, on the host) and with the product: identical bytes.  Also recovers the planted catalog."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/sedef_ref not built")
    from sedef_b200 import engine, genome
    wd = str(tmp_path)
    fa, bed, catalog = genome.write_align_stage_input(wd, **genome.CONFIGS[1])
    bdir = os.path.join(wd, "buckets")
    os.makedirs(bdir)
    subprocess.run([REF_BIN, "align", "bucket", "-n", "4", bed, bdir, fa], check=True, capture_output=True, timeout=600)
    lines = []
    for b in sorted(os.listdir(bdir)):
        ref = subprocess.run([REF_BIN, "align", "generate", "-k", "11", fa, os.path.join(bdir, b)], check=True, capture_output=True,
                             text=True, timeout=1800).stdout
        out = os.path.join(wd, b + ".aligned.bed")
        engine.align_generate(fa, os.path.join(bdir, b), out)
        assert open(out).read() == ref, b
        lines += [ln for ln in ref.split("\n") if ln]
    recovered = 0
    for c in catalog:
        for ln in lines:
            f = ln.split("\t")
            q0, q1, r0, r1 = int(f[1]), int(f[2]), int(f[4]), int(f[5])
            a = (c["s0"], c["s1"], c["d0"], c["d1"])
            if (q0 <= a[0] + 60 and q1 >= a[1] - 60 and r0 <= a[2] + 60 and r1 >= a[3] - 60) or \
               (r0 <= a[0] + 60 and r1 >= a[1] - 60 and q0 <= a[2] + 60 and q1 >= a[3] - 60):
                recovered += 1
                break
    assert recovered >= 36, recovered


@pytest.mark.gpu
def test_stats_generate_matches_reference_binary_golden(built, golden_dir, tmp_path):
    """`sedef stats generate` (src/stats_main.cc:213-395): the 34-column SD report of an aligned.bed -- Alignment(fa, fb, cigar), the
    split at assembly gaps (runs of >= 100 N) with re-trimmed pieces, the BEDPE stat loop and populate_nice_alignment's counters of
    ALL pieces in one GPU call, the floating-point columns in the reference's "%g" text, the filters -- against the reference
    binary's own output (tests/golden/stats_golden.json), with the default parameters and with gap splitting switched on
    (--max-ok-gap 1 --min-split 500: recursive cuts at the largest gaps)."""
    from sedef_b200 import engine, genome
    g = load_json(golden_dir, "stats_golden.json")
    wd = str(tmp_path)
    fa, _, _ = genome.write_align_stage_input(wd, **g["config"])
    assert hashlib.sha1(open(fa, "rb").read()).hexdigest() == g["genome_sha1"], "generator drifted: regenerate the fixture"
    ab = os.path.join(wd, "aligned.bed")
    with open(ab, "w") as f:
        f.write(g["aligned"])
    for name, extra in g["variants"].items():
        kw = {}
        for k, v in zip(extra[::2], extra[1::2]):
            kw[{"--max-ok-gap": "max_ok_gap", "--min-split": "min_split", "--uppercase": "min_uppercase"}[k]] = int(v)
        out = os.path.join(wd, name + ".final.bed")
        cnt = engine.stats_generate(fa, ab, out, **kw)
        got = open(out).read()
        assert got == g["reports"][name], name
        assert cnt["hits"] == g["aligned"].count("\n") and cnt["lines"] == got.count("\n") - 1
    assert g["reports"]["default"].count("\n") - 1 > g["aligned"].count("\n")          # the assembly gaps did split hits
    assert g["reports"]["gap_split"].count("\n") > g["reports"]["default"].count("\n")
