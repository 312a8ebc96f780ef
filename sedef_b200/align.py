"""Host-side mirror of SEDEF's `Alignment(fa, fb)` front end for the ksw_extz2 hot path.

Reference behaviour mirrored here (argument meaning and results, not code):
  * `Alignment::Alignment(fa, fb)`                       src/align.cc:76-88
      both strings are encoded with `align_dna` (ACGT/acgt -> 0..3, anything else -> 4) and aligned
      with match 5 / mismatch -4 / N 0, gap open 40, gap extend 1, unbanded, no z-drop, flag 0
      (src/globals.cc:25-28);
  * `align_helper`                                        src/align.cc:39-68
      60 kbp chunking with the same offset on both strings, ksw op -> "MDI" remap
      (ksw I, query only -> 'D'; ksw D, target only -> 'I');
  * `populate_nice_alignment` + getters                    src/align.cc:274-315, src/align.h:79-92
  * the BEDPE stat loop and fp fields of `process()`       src/stats_main.cc:231-283,297-299.
The difference is that requests are *batched*: `align_pairs` sends every pair of a wave through one
`ksw_extz2_batch_flat` call; the statistics come back as compact integer records computed on the GPU.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Sequence, Tuple

import numpy as np

from . import engine, synth

MAX_KSW_SEQ_LEN = 60 * 1000        # src/globals.h:18,54: 60 * KB, and KB is 1000


@dataclasses.dataclass
class Alignment:
    """Result of one `Alignment(fa, fb)` request."""
    a: str
    b: str
    cigar: List[Tuple[str, int]]            # SEDEF alphabet: M, D (a only), I (b only)
    stats: dict                              # integer fields of sd_stats_t

    def cigar_string(self) -> str:           # src/align.cc:614-621
        return "".join(f"{n}{op}" for op, n in self.cigar if n)     # zero-length runs are not printed (src/align.cc:617)

    def span(self) -> int: return self.stats["span"]
    def matches(self) -> int: return self.stats["matches"]
    def mismatches(self) -> int: return self.stats["mismatches"]
    def gaps(self) -> int: return self.stats["gaps"]
    def gap_bases(self) -> int: return self.stats["gap_bases"]

    def _tot(self) -> float:
        return self.matches() + self.gap_bases() + self.mismatches()

    def gap_error(self) -> float:            # src/align.h:84-87, pct() of src/common.h:99
        return 100.0 * self.gap_bases() / self._tot()

    def mismatch_error(self) -> float:       # src/align.h:88-91
        return 100.0 * self.mismatches() / self._tot()

    def total_error(self) -> float:          # src/align.h:92
        return self.mismatch_error() + self.gap_error()

    def bedpe_fp(self) -> dict:
        """fracMatch, fracMatchIndel, jcK, k2K, errorScaled, filter_score (src/stats_main.cc:273-283,297-299)."""
        return engine.derive_fp(self.stats)


def _merge_stats(parts: Sequence[dict]) -> dict:
    out = dict(parts[0])
    for p in parts[1:]:
        for k, v in p.items():
            out[k] += v
    return out


def align_pairs(pairs: Sequence[Tuple[str, str]], match: int = synth.SEDEF_MATCH, mismatch: int = synth.SEDEF_MISMATCH,
                gap_open: int = synth.SEDEF_GAPO, gap_extend: int = synth.SEDEF_GAPE, bandwidth: int = -1) -> List[Alignment]:
    """Batched `Alignment(fa, fb)` for every (fa, fb) in `pairs`."""
    mat = synth.sedef_matrix(match, mismatch)
    chunks, owner = [], []
    for idx, (fa, fb) in enumerate(pairs):
        n = min(len(fa), len(fb))
        sp = 0
        while sp < n:                                            # src/align.cc:46-53
            chunks.append((fa[sp:sp + MAX_KSW_SEQ_LEN], fb[sp:sp + MAX_KSW_SEQ_LEN]))
            owner.append(idx)
            sp += MAX_KSW_SEQ_LEN
    out: List[Alignment] = []
    per_pair_cigar = [[] for _ in pairs]
    per_pair_stats = [[] for _ in pairs]
    if chunks:
        ps = synth.pairs_from_strings(chunks)
        res = engine.extz2_batch(ps, mat, gap_open, gap_extend, bandwidth, -1, 0)
        for k, idx in enumerate(owner):
            for c in res.cigars[k].tolist():
                op, ln = c & 0xF, c >> 4
                if op < 3:                                       # src/align.cc:61
                    per_pair_cigar[idx].append(("MDI"[op], ln))  # src/align.cc:62
            per_pair_stats[idx].append(res.stats_dict(k))
    zero = {n: 0 for n in engine.STAT_FIELDS}
    for idx, (fa, fb) in enumerate(pairs):
        st = _merge_stats(per_pair_stats[idx]) if per_pair_stats[idx] else dict(zero)
        out.append(Alignment(fa, fb, per_pair_cigar[idx], st))
    return out


def align(fa: str, fb: str) -> Alignment:
    """Single `Alignment(fa, fb)` (a batch of one)."""
    return align_pairs([(fa, fb)])[0]


def from_cigars(pairs: Sequence[Tuple[str, str]], cigar_strings: Sequence[str]) -> List[Alignment]:
    """Batched `Alignment(fa, fb, cigar_string)` (src/align.cc:90-105): parse "\\d+[MID]" (';' skipped), statistics on the GPU."""
    parsed, raw = [], []
    for cs in cigar_strings:
        ops, num = [], 0
        for ch in cs:                                            # the reference's own parser, src/align.cc:94-103
            if ch.isdigit():
                num = 10 * num + int(ch)
            elif ch != ";":
                ops.append((ch, num)); num = 0
        parsed.append(ops)
        # SEDEF 'D' = a only = ksw I; any other letter (code 3) consumes both strings and still counts as a gap run
        raw.append(np.array([(n << 4) | {"M": 0, "D": 1, "I": 2}.get(op, 3) for op, n in ops], np.uint32))
    a_list = [np.frombuffer(fa.encode(), np.uint8) for fa, _ in pairs]
    b_list = [np.frombuffer(fb.encode(), np.uint8) for _, fb in pairs]
    st, status = engine.stats_from_cigars(raw, a_list, b_list)
    out = []
    for i, (fa, fb) in enumerate(pairs):
        if status[i]:
            raise ValueError(f"CIGAR {i} overruns a sequence (the reference asserts, src/align.cc:281-282)")
        out.append(Alignment(fa, fb, parsed[i], {n: int(st[i][n]) for n in engine.STAT_FIELDS}))
    return out
