"""Generates tests/golden/align_stage_golden.json with the REFERENCE BINARY (oracle/_ref/sedef_ref, the unmodified reference
without its search stage, built by oracle/Makefile): a small two-chromosome genome with planted duplications on both strands ->
`sedef align bucket` (extension, merging, binning) -> `sedef align generate -k 11` per bucket.  The fixture holds the bucket
files and the reference's *.aligned.bed text; the genome itself is regenerated from its seed (sha1 pinned).
Run here (needs /root/reference for the oracle build):  python tests/golden/make_align_stage_golden.py"""
import hashlib, json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sedef_b200 import genome  # noqa: E402

SMALL = dict(chrom_lengths={"chrA": 420_000, "chrB": 260_000}, n_dups=9, min_len=3000, max_len=9000, min_div=0.02, max_div=0.12,
             seed=0x5EDEF0A1, rc_frac=0.4)


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "sedef_ref")
    with tempfile.TemporaryDirectory() as wd:
        fa, bed, catalog = genome.write_align_stage_input(wd, **SMALL)
        bdir = os.path.join(wd, "buckets"); os.makedirs(bdir)
        subprocess.run([ref, "align", "bucket", "-n", "3", bed, bdir, fa], check=True, capture_output=True)
        buckets, expect = {}, {}
        for b in sorted(os.listdir(bdir)):
            buckets[b] = open(os.path.join(bdir, b)).read()
            r = subprocess.run([ref, "align", "generate", "-k", "11", fa, os.path.join(bdir, b)], check=True, capture_output=True, text=True)
            expect[b] = r.stdout
        out = dict(config={k: v for k, v in SMALL.items()}, genome_sha1=hashlib.sha1(open(fa, "rb").read()).hexdigest(),
                   seeds=open(bed).read(), buckets=buckets, aligned=expect,
                   n_lines=sum(v.count("\n") for v in expect.values()))
    with open(os.path.join(ROOT, "tests", "golden", "align_stage_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("regions", sum(v.count("\n") for v in buckets.values()), "hit lines", out["n_lines"])


if __name__ == "__main__":
    main()
