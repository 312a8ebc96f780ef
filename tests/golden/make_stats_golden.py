"""Generates tests/golden/stats_golden.json with the REFERENCE BINARY (oracle/_ref/sedef_ref): a small genome whose planted copies
carry assembly gaps (runs of N) and long indels -> `align bucket` -> `align generate` -> sort | uniq (sedef.sh:219-220) ->
`sedef stats generate` with the default parameters and with --max-ok-gap / --min-split (gap splitting on).  The fixture holds the
aligned.bed and the reference's reports; the genome is regenerated from its seed (sha1 pinned).
Run here (needs /root/reference for the oracle build):  python tests/golden/make_stats_golden.py"""
import hashlib, json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sedef_b200 import genome  # noqa: E402

SMALL = dict(chrom_lengths={"chrA": 500_000, "chrB": 300_000}, n_dups=12, min_len=3000, max_len=9000, min_div=0.02, max_div=0.10,
             seed=0x5EDEF0B2, rc_frac=0.4, large_indels=3, assembly_gaps=5)
VARIANTS = {"default": [], "gap_split": ["--max-ok-gap", "1", "--min-split", "500"]}


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "sedef_ref")
    with tempfile.TemporaryDirectory() as wd:
        fa, bed, catalog = genome.write_align_stage_input(wd, **SMALL)
        bdir = os.path.join(wd, "buckets"); os.makedirs(bdir)
        subprocess.run([ref, "align", "bucket", "-n", "2", bed, bdir, fa], check=True, capture_output=True)
        lines = set()
        for b in sorted(os.listdir(bdir)):
            r = subprocess.run([ref, "align", "generate", "-k", "11", fa, os.path.join(bdir, b)], check=True, capture_output=True, text=True)
            lines.update(ln for ln in r.stdout.split("\n") if ln)
        aligned = "\n".join(sorted(lines)) + "\n"
        ab = os.path.join(wd, "aligned.bed")
        open(ab, "w").write(aligned)
        reports = {}
        for name, extra in VARIANTS.items():
            r = subprocess.run([ref, "stats", "generate"] + extra + [fa, ab], check=True, capture_output=True, text=True)
            reports[name] = r.stdout
            print(name, r.stdout.count("\n") - 1, "report lines from", len(lines), "aligned hits")
        out = dict(config=SMALL, variants=VARIANTS, genome_sha1=hashlib.sha1(open(fa, "rb").read()).hexdigest(), aligned=aligned, reports=reports)
    with open(os.path.join(ROOT, "tests", "golden", "stats_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
