"""Regenerates the golden fixtures in tests/golden/ from the REFERENCE itself.

Run in the build container (needs /root/reference and `make -C oracle ref_full`):
    python tests/golden/make_golden.py
Outputs (committed):
  ksw2_kat.json        the reference's only fixed test pair (python/simulations.py:6-7) through the compiled
                       reference kernel for the parameter table of SURVEY.md Appendix B.3, plus the
                       SEDEF-level record produced by the reference's own Alignment class
  ksw2_golden.json     self-contained fuzz vectors: ASCII pairs + expected ksw_extz_t fields + CIGARs
                       from oracle/_ref/libksw2_ref.so (unmodified extern/ksw2_extz2_sse.cc)
  sd_stats_golden.json soft-masked pairs + the reference Alignment(fa, fb) CIGAR string and error counters
                       from oracle/_ref/libsedef_ref.so (unmodified src/align.cc)
"""
import ctypes as C
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from sedef_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
mat = synth.sedef_matrix()


def kat_pair():
    src = open(os.path.join(REF, "python", "simulations.py")).read()
    m = re.findall(r"'([ACGTNacgtn]{500,})'", src)
    return m[0], m[1]


def sedef_ref():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsedef_ref.so"))
    lib.ref_alignment.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int] + [C.POINTER(C.c_int)] * 5
    return lib


def ref_alignment(lib, fa: str, fb: str):
    buf = C.create_string_buffer(4 * (len(fa) + len(fb)) + 64)
    v = [C.c_int(0) for _ in range(5)]
    rc = lib.ref_alignment(fa.encode(), fb.encode(), buf, len(buf), *[C.byref(x) for x in v])
    assert rc == 0
    return dict(cigar=buf.value.decode(), span=v[0].value, matches=v[1].value, mismatches=v[2].value,
                gaps=v[3].value, gap_bases=v[4].value)


def ref_alignment_strings(lib, fa: str, fb: str):
    lib.ref_alignment_strings.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    cap = 2 * (len(fa) + len(fb)) + 64
    oa, ob = C.create_string_buffer(cap), C.create_string_buffer(cap)
    n = lib.ref_alignment_strings(fa.encode(), fb.encode(), oa, ob, cap)
    assert n >= 0
    return oa.value.decode(), ob.value.decode()


def stat_loop(align_a: str, align_b: str) -> dict:
    """LITERAL transcription of the BEDPE stat loop (reference src/stats_main.cc:244-271), run over the reference's own
    align_a / align_b strings.  Deliberately not shared with any product or port code."""
    d = dict(indel_a=0, indel_b=0, alnB=0, matchB=0, mismatchB=0, transitionsB=0, transversionsB=0,
             uppercaseA=0, uppercaseB=0, uppercaseMatches=0)
    assert len(align_a) == len(align_b)
    for ca, cb in zip(align_a, align_b):
        a, b = ca.upper(), cb.upper()
        d["indel_a"] += a == "-"
        d["indel_b"] += b == "-"
        d["matchB"] += a != "-" and a == b
        d["uppercaseA"] += ca != "-" and ca.upper() != "N" and ca.isupper()
        d["uppercaseB"] += cb != "-" and cb.upper() != "N" and cb.isupper()
        if a != "-" and b != "-":
            d["alnB"] += 1
            if a != b:
                d["mismatchB"] += 1
                if a == "A" or a == "G":
                    d["transitionsB"] += b == "A" or b == "G"
                    d["transversionsB"] += not (b == "A" or b == "G")
                else:
                    d["transitionsB"] += b == "C" or b == "T"
                    d["transversionsB"] += not (b == "C" or b == "T")
            elif ca.isupper() and cb.isupper():
                d["uppercaseMatches"] += 1
    return {k: int(v) for k, v in d.items()}


def main():
    ref = oracle.ref()
    slib = sedef_ref()
    s1, s2 = kat_pair()
    ps = synth.pairs_from_strings([(s1, s2)])
    q, t = ps.pair(0)
    rows = []
    for w, zd, flag in [(-1, -1, 0), (100, -1, 0), (20, -1, 0), (10, -1, 0), (-1, -1, 0x02), (-1, -1, 0x01),
                        (-1, -1, 0x80), (100, 50, 0), (100, 50, 0x40), (10, -1, 0x40), (10, -1, 0x42),
                        (-1, -1, 0x08), (100, 50, 0x18), (10, -1, 0x08), (100, 50, 0x19)]:      # SURVEY App. B.3 approx-max rows
        f, c = ref.extz2(q, t, mat, 40, 1, w, zd, flag)
        rows.append(dict(w=w, zdrop=zd, flag=flag, fields=f, cigar=oracle.cigar_str(c)))
    kat = dict(source="reference python/simulations.py:6-7 (seq1, seq2); outputs of the compiled reference",
               seq1=s1, seq2=s2, scoring=dict(m=5, match=5, mismatch=-4, gapo=40, gape=1), rows=rows,
               sedef_alignment=ref_alignment(slib, s1, s2),
               # stat loop of src/stats_main.cc:244-283 on Alignment(seq1, seq2)
               survey_stat_loop=dict(span=1137, indel_a=1, indel_b=19, alnB=1117, matchB=1092, mismatchB=25,
                                     transitionsB=18, transversionsB=7, uppercaseA=355, uppercaseB=348,
                                     uppercaseMatches=338, gaps=4, gap_bases=20,
                                     fracMatch=0.977619, fracMatchIndel=0.960422, jcK=0.0227221, k2K=0.0227815,
                                     filter_score=0.97413))
    json.dump(kat, open(os.path.join(HERE, "ksw2_kat.json"), "w"), indent=1)

    # ---- kernel-level golden vectors ----
    groups = []
    specs = [
        ("tiny", dict(n=60, min_len=1, max_len=40, div=0.15), [(-1, -1, 0), (3, -1, 0), (8, 10, 0)]),
        ("ragged", dict(n=40, min_len=1, max_len=400, div=0.12), [(-1, -1, 0), (20, -1, 0), (50, 100, 0), (30, 80, 0x02),
                                                                     (30, 80, 0x01), (30, 80, 0x40), (30, 80, 0x80), (5, -1, 0)]),
        ("divergent", dict(n=25, min_len=200, max_len=700, div=0.4, burst=60), [(16, 30, 0), (100, 200, 0), (-1, 150, 0)]),
    ]
    for gi, (name, kw, params) in enumerate(specs):
        pset = synth.make_pairs_mixed(seed=4242 + gi, **kw)
        pairs = [("".join(map(chr, pset.raw_pair(i)[0])), "".join(map(chr, pset.raw_pair(i)[1]))) for i in range(pset.n)]
        runs = []
        for (w, zd, flag) in params:
            _, fr, cr = ref.batch(pset, mat, 40, 1, w, zd, flag, nthreads=4)
            runs.append(dict(w=w, zdrop=zd, flag=flag,
                             fields=[[f[k] for k in ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar")] for f in fr],
                             cigars=[oracle.cigar_str(c) for c in cr]))
        groups.append(dict(name=name, pairs=pairs, runs=runs))
    json.dump(dict(source="oracle/_ref/libksw2_ref.so (unmodified reference extern/ksw2_extz2_sse.cc), SEDEF scoring",
                   field_order=["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar"],
                   groups=groups), open(os.path.join(HERE, "ksw2_golden.json"), "w"))

    # ---- SEDEF-level golden records from the reference's Alignment class ----
    pset = synth.make_pairs_mixed(60, seed=777, min_len=5, max_len=500, div=0.1, n_frac=0.01)
    recs = []
    for i in range(pset.n):
        fa = "".join(map(chr, pset.raw_pair(i)[0])); fb = "".join(map(chr, pset.raw_pair(i)[1]))
        rec = dict(a=fa, b=fb, **ref_alignment(slib, fa, fb))
        aa, ab = ref_alignment_strings(slib, fa, fb)
        assert len(aa) == rec["span"]
        rec["stat_loop"] = stat_loop(aa, ab)             # the ten BEDPE integers from the reference's own column strings
        recs.append(rec)
    json.dump(dict(source="oracle/_ref/libsedef_ref.so: reference Alignment(fa, fb) (src/align.cc:76-88,274-315); stat_loop = "
                          "src/stats_main.cc:244-271 applied to the reference's align_a / align_b",
                   records=recs), open(os.path.join(HERE, "sd_stats_golden.json"), "w"))
    # ---- region-level golden: the reference's whole fast_align() (every call site of the hot path) ----
    slib.ref_fast_align.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    regs = []
    for (L, div, seed) in [(2500, 0.03, 11), (6000, 0.06, 12), (9000, 0.10, 13)]:
        qs, ts = synth.make_region_pair(L, div, seed=seed)
        buf = C.create_string_buffer(1 << 22)
        n = slib.ref_fast_align(qs.encode(), ts.encode(), 11, buf, len(buf))
        regs.append(dict(length=L, div=div, seed=seed, n_hits=n, hits=buf.value.decode()))
    json.dump(dict(source="oracle/_ref/libsedef_ref.so: reference fast_align(query, ref, orig, 11) (src/chain.cc:203-268) on "
                          "synth.make_region_pair(length, div, seed=seed); one line per hit: qs qe rs re cigar span matches mismatches gaps gap_bases",
                   regions=regs), open(os.path.join(HERE, "fast_align_golden.json"), "w"))
    # ---- chain-wave golden: anchors -> chains -> Alignment(query, ref, anchors, guide_idx) of the reference ----
    slib.ref_chain_guides.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    cregs = []
    for (L, div, seed) in [(2500, 0.03, 11), (6000, 0.06, 12), (9000, 0.10, 13), (5000, 0.2, 14)]:
        qs, ts = synth.make_region_pair(L, div, seed=seed)
        buf = C.create_string_buffer(1 << 22)
        n = slib.ref_chain_guides(qs.encode(), ts.encode(), 11, buf, len(buf))
        cregs.append(dict(length=L, div=div, seed=seed, n_chains=n, chains=buf.value.decode()))
    json.dump(dict(source="oracle/_ref/libsedef_ref.so: reference generate_anchors + chain_anchors + Alignment(query, ref, anchors, guide_idx) "
                          "(src/chain.cc:211-258, src/align.cc:199-270); one line per chain: start_a end_a start_b end_b cigar span matches "
                          "mismatches gaps gap_bases n  q r l ...",
                   regions=cregs), open(os.path.join(HERE, "chain_wave_golden.json"), "w"))
    # ---- refine-wave golden: Alignment(qstr, rstr, vector<Hit> guide, side) of the reference (gap fills + side extensions + trims) ----
    slib.ref_hit_guide.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    hregs = []
    for (L, div, seed, side) in [(2500, 0.03, 11, 500), (6000, 0.06, 12, 500), (9000, 0.10, 13, 500), (5000, 0.2, 14, 500),
                                 (4000, 0.05, 15, 0), (7000, 0.3, 16, 500), (3000, 0.02, 17, 200)]:
        qs, ts = synth.make_region_pair(L, div, seed=seed)
        buf = C.create_string_buffer(1 << 22)
        n = slib.ref_hit_guide(qs.encode(), ts.encode(), 11, side, buf, len(buf))
        hregs.append(dict(length=L, div=div, seed=seed, side=side, n_guide=n, text=buf.value.decode()))
    json.dump(dict(source="oracle/_ref/libsedef_ref.so: reference Alignment(qstr, rstr, vector<Hit> guide, side) (src/align.cc:107-197) on a "
                          "co-linear guide of the reference's own chain alignments; line 1 = result, then the guide hits",
                   regions=hregs), open(os.path.join(HERE, "hit_guide_golden.json"), "w"))
    # ---- merge golden: Alignment::merge of the reference on overlapping pairs of its own chain alignments ----
    slib.ref_merge_pairs.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    mregs = []
    for (L, div, seed) in [(2500, 0.03, 11), (6000, 0.06, 12), (9000, 0.10, 13), (5000, 0.2, 14), (7000, 0.3, 16), (12000, 0.15, 18)]:
        qs, ts = synth.make_region_pair(L, div, seed=seed)
        buf = C.create_string_buffer(1 << 22)
        n = slib.ref_merge_pairs(qs.encode(), ts.encode(), 11, buf, len(buf))
        mregs.append(dict(length=L, div=div, seed=seed, n_merges=n, text=buf.value.decode()))
    json.dump(dict(source="oracle/_ref/libsedef_ref.so: reference Alignment::merge (src/align.cc:505-610); per pair: P prev, C cur, M merged",
                   regions=mregs), open(os.path.join(HERE, "merge_golden.json"), "w"))
    # ---- region golden: the reference's anchors + filtered chains (the region-level driver's input) and its fast_align hits ----
    slib.ref_region.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    rregs = []
    for (L, div, seed, same, qs0, rs0) in [(2500, 0.03, 11, 0, 0, 0), (6000, 0.06, 12, 0, 0, 0), (9000, 0.10, 13, 0, 0, 0), (5000, 0.2, 14, 0, 0, 0),
                                           (7000, 0.3, 16, 0, 0, 0), (12000, 0.15, 18, 0, 0, 0), (20000, 0.08, 19, 0, 0, 0),
                                           (6000, 0.06, 12, 1, 100000, 140000), (9000, 0.10, 13, 1, 5000, 9000), (12000, 0.15, 18, 1, 0, 6000),
                                           (8000, 0.05, 21, 1, 1000, 1400), (15000, 0.12, 22, 0, 0, 0), (4000, 0.25, 23, 0, 0, 0)]:
        qs, ts = synth.make_region_pair(L, div, seed=seed)
        buf = C.create_string_buffer(1 << 24)
        n = slib.ref_region(qs.encode(), ts.encode(), 11, same, qs0, rs0, buf, len(buf))
        assert n >= 0
        rregs.append(dict(length=L, div=div, seed=seed, same_chr=same, orig_qs=qs0, orig_rs=rs0, n_hits=n, text=buf.value.decode()))
    json.dump(dict(source="oracle/_ref/libsedef_ref.so: reference generate_anchors + chain_anchors (+ the chain filter of src/chain.cc:222-247) "
                          "and fast_align (src/chain.cc:203-268) on synth.make_region_pair(length, div, seed=seed) for a seed hit with the given "
                          "same_chr / origin; lines: 'A q r l' anchors, 'C n i...' chains (anchor indices), 'H ...' hits",
                   regions=rregs), open(os.path.join(HERE, "region_golden.json"), "w"))
    for f in ("region_golden.json", "ksw2_kat.json", "ksw2_golden.json", "sd_stats_golden.json", "fast_align_golden.json", "chain_wave_golden.json",
              "hit_guide_golden.json", "merge_golden.json"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
