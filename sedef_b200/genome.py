"""Synthetic genomes for the align-stage configurations (BASELINE.json configs 1/4/5, SURVEY.md section 8d "Configs 1/4/5"):
iid ACGT chromosomes with soft-masked runs and sparse N, a planted segmental-duplication catalog (a source segment copied to a
disjoint place, on either strand, with the makeSmall divergence model plus a few longer indels), written as FASTA + .fai, and the
seed BED lines the align stage starts from (`sedef search` cannot be built here -- SURVEY section 8c -- so the seeds are the planted
intervals themselves, as the survey prescribes).  Test / bench input only."""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from . import synth

_RC = np.full(256, ord("N"), np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _RC[_a] = _b


def revcomp(a: np.ndarray) -> np.ndarray:
    return _RC[a[::-1]]


def make_genome(chrom_lengths: Dict[str, int], n_dups: int, min_len: int = 5000, max_len: int = 20000, min_div: float = 0.02,
                max_div: float = 0.10, seed: int = 0x5EDEF001, rc_frac: float = 0.3, large_indels: int = 2, assembly_gaps: int = 0):
    """Returns ({name: np.uint8 ASCII}, catalog); catalog rows are dicts with src/dst chromosome and [start, end), divergence and
    strand (rc = True: the copy is the reverse complement of the mutated source)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    chroms: Dict[str, np.ndarray] = {}
    for name, length in chrom_lengths.items():
        codes = rng.integers(0, 4, length, dtype=np.uint8)
        codes[rng.random(length) < 0.0005] = 4
        chroms[name] = synth._to_ascii(codes, synth._softmask(rng, length, mean_run=400))
    # disjoint windows on a grid over all chromosomes; sources and destinations are drawn from a shuffled list of windows
    win = max_len + 2500
    windows: List[Tuple[str, int]] = []
    for name, length in chrom_lengths.items():
        for k in range(length // win):
            windows.append((name, k * win))
    assert len(windows) >= 2 * n_dups, "genome too short for the catalog"
    order = rng.permutation(len(windows))
    catalog = []
    for k in range(n_dups):
        (sc, sw), (dc, dw) = windows[int(order[2 * k])], windows[int(order[2 * k + 1])]
        L = int(rng.integers(min_len, max_len + 1)); div = float(rng.uniform(min_div, max_div))
        s0 = sw + int(rng.integers(300, win - max_len - 900)); d0 = dw + int(rng.integers(300, win - max_len - 900))
        src = chroms[sc][s0:s0 + L]
        codes = synth.encode(src)
        cp, cl, _ = synth._mutate_small(rng, codes, src >= ord("a"), div)
        copy = synth._to_ascii(cp, cl)
        for _ in range(large_indels):                            # a few longer indels (the makeLarge idea)
            p = int(rng.integers(200, max(201, len(copy) - 200))); kk = int(rng.integers(20, 120))
            copy = (np.concatenate([copy[:p], copy[p + kk:]]) if rng.random() < 0.5
                    else np.concatenate([copy[:p], synth.ASCII[rng.integers(0, 4, kk)], copy[p:]]))
        copy = copy[:max_len + 400]
        if k < assembly_gaps:                                    # an assembly gap inside the copy: a run of N (`sedef stats` splits there)
            p = len(copy) // 2 + int(rng.integers(-200, 200)); copy = copy.copy(); copy[p:p + int(rng.integers(120, 260))] = ord("N")
        is_rc = bool(rng.random() < rc_frac)
        chroms[dc][d0:d0 + len(copy)] = revcomp(copy) if is_rc else copy
        catalog.append(dict(src_chr=sc, s0=s0, s1=s0 + L, dst_chr=dc, d0=d0, d1=d0 + len(copy), div=div, rc=is_rc))
    return chroms, catalog


def write_fasta(path: str, chroms: Dict[str, np.ndarray], line: int = 60) -> None:
    """FASTA with fixed line length + the samtools-style .fai next to it (format: src/fasta.cc:33-47)."""
    fai = []
    with open(path, "wb") as f:
        for name, seq in chroms.items():
            header = (">%s\n" % name).encode()
            f.write(header)
            off = f.tell()
            n = len(seq)
            full = n // line
            body = np.empty((full, line + 1), np.uint8)
            body[:, :line] = seq[:full * line].reshape(full, line)
            body[:, line] = 10
            f.write(body.tobytes())
            if n % line:
                f.write(seq[full * line:].tobytes() + b"\n")
            fai.append("%s\t%d\t%d\t%d\t%d\n" % (name, n, off, line, line + 1))
    with open(path + ".fai", "w") as f:
        f.write("".join(fai))


def seed_bed_lines(catalog, slop: int = 0) -> List[str]:
    """One seed hit per planted copy, 10 BED columns (Hit::from_bed, src/hit.cc:29-48); the reference's own `sedef align bucket`
    extends (Hit::extend, src/hit.cc:200-207), orders, merges and bins them."""
    out = []
    for k, c in enumerate(catalog):
        out.append("%s\t%d\t%d\t%s\t%d\t%d\tseed%d\t0\t+\t%s" % (c["src_chr"], max(0, c["s0"] - slop), c["s1"] + slop, c["dst_chr"],
                                                                   max(0, c["d0"] - slop), c["d1"] + slop, k, "-" if c["rc"] else "+"))
    return out


def write_align_stage_input(workdir: str, chrom_lengths: Dict[str, int], n_dups: int, **kw):
    """genome.fa (+ .fai) and seeds.bed under `workdir`; returns (fasta path, seed path, catalog)."""
    os.makedirs(workdir, exist_ok=True)
    chroms, catalog = make_genome(chrom_lengths, n_dups, **kw)
    fa = os.path.join(workdir, "genome.fa")
    write_fasta(fa, chroms)
    bed = os.path.join(workdir, "seeds.bed")
    with open(bed, "w") as f:
        f.write("\n".join(seed_bed_lines(catalog)) + "\n")
    return fa, bed, catalog


# hg38 chromosome lengths (Mbp, rounded): the shape of BASELINE.json configs[4]
HG38_MBP = dict(chr1=248, chr2=242, chr3=198, chr4=190, chr5=182, chr6=171, chr7=159, chr8=145, chr9=138, chr10=134, chr11=135,
                chr12=133, chr13=114, chr14=107, chr15=102, chr16=90, chr17=83, chr18=80, chr19=59, chr20=64, chr21=47, chr22=51,
                chrX=156, chrY=57)


def config5(scale: float = 1.0, dups_per_mbp: float = 8.0):
    """configs[4]: 24 chromosomes with hg38's proportions (3.1 Gbp at scale 1), planted SD catalog up to 30 % divergence."""
    lengths = {k: max(200_000, int(v * 1_000_000 * scale)) for k, v in HG38_MBP.items()}
    total = sum(lengths.values())
    return dict(chrom_lengths=lengths, n_dups=int(total / 1e6 * dups_per_mbp), min_len=2000, max_len=20000, min_div=0.02,
                max_div=0.30, seed=0x5EDEF005, rc_frac=0.4)


CONFIGS = {
    # BASELINE.json configs[0]: 2 Mbp, 40 planted 5-20 kbp duplications at 2-10 %
    1: dict(chrom_lengths={"chr1": 2_000_000}, n_dups=40, min_len=5000, max_len=20000, min_div=0.02, max_div=0.10, seed=0x5EDEF001),
    # configs[3]: a 50 Mbp chromosome with a planted SD catalog
    4: dict(chrom_lengths={"chr1": 50_000_000}, n_dups=1000, min_len=5000, max_len=20000, min_div=0.02, max_div=0.15, seed=0x5EDEF004),
}
