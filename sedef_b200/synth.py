"""Deterministic synthetic workloads for the ksw_extz2 hot path.

The reference ships no benchmark inputs for this path; the
workloads follow BASELINE.json:configs and SURVEY.md section 8(d):

* config 2: N x 1 kbp pairs, 5 % divergence with the `makeSmall` event model of the reference's
  simulator (python/simulations.py:53-75: per-base events, 2/3 substitutions, 1/6 one-base
  insertions, 1/6 one-base deletions), soft-masked (lower-case runs, ~50 %), 0.1 % N.
* config 3: pairs of 10-50 kbp, 10 % substitutions + 5 % indel events of length U[1,30] plus up to
  three large indels U[50,400] (the `makeLarge` idea, python/simulations.py:77-137).

Everything is vectorised numpy (PCG64, fixed seeds) so that 100k pairs generate in seconds.
Outputs are *flat*: one uint8 buffer per side plus int64 offsets / int32 lengths, which is the
layout `ksw_extz2_batch_flat` (include/ksw2_b200.h) takes without any gather.
"""
from __future__ import annotations

import dataclasses
import numpy as np

ASCII = np.frombuffer(b"ACGTN", dtype=np.uint8)
ASCII_LOWER = np.frombuffer(b"acgtn", dtype=np.uint8)

# SEDEF scoring (reference src/globals.cc:25-28, src/align.cc:41-44)
SEDEF_M = 5
SEDEF_MATCH, SEDEF_MISMATCH, SEDEF_GAPO, SEDEF_GAPE = 5, -4, 40, 1


def sedef_matrix(match: int = SEDEF_MATCH, mismatch: int = SEDEF_MISMATCH) -> np.ndarray:
    """5x5 int8 matrix exactly as align_helper builds it (src/align.cc:41-44)."""
    a, b = match, (mismatch if mismatch < 0 else -mismatch)
    mat = np.full((5, 5), b, dtype=np.int8)
    np.fill_diagonal(mat, a)
    mat[4, :] = 0
    mat[:, 4] = 0
    return mat.reshape(-1).copy()


@dataclasses.dataclass
class PairSet:
    """n pairs in flat layout. `q`/`t` hold codes 0..4; `q_raw`/`t_raw` the original-case ASCII."""
    qlen: np.ndarray      # int32 [n]
    qoff: np.ndarray      # int64 [n]
    q: np.ndarray         # uint8 [sum qlen]
    tlen: np.ndarray
    toff: np.ndarray
    t: np.ndarray
    q_raw: np.ndarray
    t_raw: np.ndarray

    @property
    def n(self) -> int:
        return int(self.qlen.shape[0])

    def pair(self, i: int):
        return (self.q[self.qoff[i]:self.qoff[i] + self.qlen[i]],
                self.t[self.toff[i]:self.toff[i] + self.tlen[i]])

    def raw_pair(self, i: int):
        return (self.q_raw[self.qoff[i]:self.qoff[i] + self.qlen[i]],
                self.t_raw[self.toff[i]:self.toff[i] + self.tlen[i]])

    def subset(self, idx) -> "PairSet":
        idx = np.asarray(idx, dtype=np.int64)
        ql, tl = self.qlen[idx], self.tlen[idx]
        qo = np.zeros(len(idx), np.int64); to = np.zeros(len(idx), np.int64)
        if len(idx):
            qo[1:] = np.cumsum(ql[:-1]); to[1:] = np.cumsum(tl[:-1])
        def gather(buf, off, ln, noff):
            # vectorised: source position of every output byte
            tot = int(ln.sum())
            if tot == 0:
                return np.empty(0, np.uint8)
            src = np.repeat(off[idx] - noff, ln) + np.arange(tot, dtype=np.int64)
            return buf[src]
        return PairSet(ql.copy(), qo, gather(self.q, self.qoff, ql, qo), tl.copy(), to,
                       gather(self.t, self.toff, tl, to),
                       gather(self.q_raw, self.qoff, ql, qo), gather(self.t_raw, self.toff, tl, to))


def concat_pairsets(parts) -> "PairSet":
    """Concatenation of PairSets (pairs keep their order; offsets are rebased)."""
    parts = list(parts)
    if len(parts) == 1:
        return parts[0]
    qlen = np.concatenate([p.qlen for p in parts]).astype(np.int32)
    tlen = np.concatenate([p.tlen for p in parts]).astype(np.int32)
    qbase = np.cumsum([0] + [p.q.shape[0] for p in parts[:-1]]).astype(np.int64)
    tbase = np.cumsum([0] + [p.t.shape[0] for p in parts[:-1]]).astype(np.int64)
    return PairSet(qlen, np.concatenate([p.qoff + b for p, b in zip(parts, qbase)]), np.concatenate([p.q for p in parts]),
                   tlen, np.concatenate([p.toff + b for p, b in zip(parts, tbase)]), np.concatenate([p.t for p in parts]),
                   np.concatenate([p.q_raw for p in parts]), np.concatenate([p.t_raw for p in parts]))


def encode(ascii_bytes: np.ndarray) -> np.ndarray:
    """align_dna (src/common.h:58-70,91): ACGT/acgt -> 0..3, everything else -> 4."""
    lut = np.full(256, 4, np.uint8)
    for i, (u, l) in enumerate(zip(b"ACGT", b"acgt")):
        lut[u] = i; lut[l] = i
    return lut[ascii_bytes]


def _offsets(lens: np.ndarray) -> np.ndarray:
    off = np.zeros(len(lens), np.int64)
    if len(lens) > 1:
        off[1:] = np.cumsum(lens[:-1].astype(np.int64))
    return off


def _softmask(rng: np.random.Generator, n: int, mean_run: int = 300, frac: float = 0.5) -> np.ndarray:
    """Boolean lower-case mask: alternating geometric runs, ~frac masked."""
    if n == 0:
        return np.zeros(0, bool)
    nruns = max(4, int(2 * n / mean_run) + 8)
    while True:
        up = rng.geometric(1.0 / (mean_run * (1 - frac) * 2), nruns)
        lo = rng.geometric(1.0 / (mean_run * frac * 2), nruns)
        runs = np.empty(2 * nruns, np.int64); runs[0::2] = up; runs[1::2] = lo
        if runs.sum() >= n:
            break
        nruns *= 2
    flag = np.zeros(2 * nruns, bool); flag[1::2] = True
    return np.repeat(flag, runs)[:n]


def _to_ascii(codes: np.ndarray, lower: np.ndarray) -> np.ndarray:
    return np.where(lower, ASCII_LOWER[codes], ASCII[codes])


def _mutate_small(rng, q_codes, q_lower, rate, seg_ids=None):
    """makeSmall model on a flat buffer: each query base independently suffers an event with
    probability `rate`: 2/3 substitution, 1/6 deletion (base dropped), 1/6 insertion (a random base
    is inserted after it).  Returns (t_codes, t_lower, t_seg_ids)."""
    n = q_codes.shape[0]
    ev = rng.random(n) < rate
    kind = rng.random(n)
    sub = ev & (kind < 2 / 3)
    dele = ev & (kind >= 2 / 3) & (kind < 5 / 6)
    ins = ev & (kind >= 5 / 6)
    base = q_codes.copy()
    acgt = base < 4
    shift = rng.integers(1, 4, n).astype(np.uint8)
    base = np.where(sub & acgt, (base + shift) & 3, base).astype(np.uint8)
    reps = np.ones(n, np.int64); reps[dele] = 0; reps[ins] = 2
    t_codes = np.repeat(base, reps)
    t_lower = np.repeat(q_lower, reps)
    # second copy of an inserted base becomes a fresh random base
    pos = np.cumsum(reps) - 1            # index of the LAST emitted copy of each query base
    ins_pos = pos[ins]
    t_codes[ins_pos] = rng.integers(0, 4, ins_pos.shape[0]).astype(np.uint8)
    seg = None if seg_ids is None else np.repeat(seg_ids, reps)
    return t_codes, t_lower, seg


def make_pairs_small(n: int, length: int = 1000, div: float = 0.05, seed: int = 0x5EDEF002,
                     n_frac: float = 0.001, len_jitter: int = 0) -> PairSet:
    """BASELINE.json configs[1]: n pairs of ~`length` bp at `div` divergence (makeSmall)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if len_jitter:
        qlen = rng.integers(max(1, length - len_jitter), length + len_jitter + 1, n).astype(np.int32)
    else:
        qlen = np.full(n, length, np.int32)
    tot = int(qlen.sum())
    q = rng.integers(0, 4, tot).astype(np.uint8)
    q[rng.random(tot) < n_frac] = 4
    lower = _softmask(rng, tot)
    seg = np.repeat(np.arange(n, dtype=np.int64), qlen)
    t, t_lower, tseg = _mutate_small(rng, q, lower, div, seg)
    tlen = np.bincount(tseg, minlength=n).astype(np.int32)
    # a pair whose target mutated away entirely cannot be aligned; give it one base
    if (tlen == 0).any():
        raise ValueError("empty target generated; use a longer length")
    return PairSet(qlen, _offsets(qlen), q, tlen, _offsets(tlen), t,
                   _to_ascii(q, lower), _to_ascii(t, t_lower))


def make_pairs_large(n: int, min_len: int = 10000, max_len: int = 50000, sub: float = 0.10,
                     indel: float = 0.05, seed: int = 0x5EDEF003, max_big: int = 3,
                     n_frac: float = 0.001) -> PairSet:
    """BASELINE.json configs[2]: long pairs, substitutions + indels of U[1,30] + a few big indels."""
    rng = np.random.Generator(np.random.PCG64(seed))
    qlen = rng.integers(min_len, max_len + 1, n).astype(np.int32)
    tot = int(qlen.sum())
    q = rng.integers(0, 4, tot).astype(np.uint8)
    q[rng.random(tot) < n_frac] = 4
    lower = _softmask(rng, tot)
    seg = np.repeat(np.arange(n, dtype=np.int64), qlen)
    # substitutions
    base = q.copy()
    s = (rng.random(tot) < sub) & (base < 4)
    base = np.where(s, (base + rng.integers(1, 4, tot).astype(np.uint8)) & 3, base).astype(np.uint8)
    # indel events: an event at a base with prob indel/15.5 (mean length 15.5) so that ~`indel` of
    # bases are touched; half insertions (extra random bases after), half deletions (run dropped)
    ev = rng.random(tot) < (indel / 15.5)
    ln = rng.integers(1, 31, tot)
    is_ins = rng.random(tot) < 0.5
    # big indels: up to max_big per pair
    for _ in range(max_big):
        pos = (rng.random(n) * (qlen - 1)).astype(np.int64) + _offsets(qlen)
        use = rng.random(n) < 0.7
        ev[pos[use]] = True
        ln[pos[use]] = rng.integers(50, 401, int(use.sum()))
    reps = np.ones(tot, np.int64)
    ins_ev = ev & is_ins
    reps[ins_ev] = 1 + ln[ins_ev]
    # deletions: drop ln bases starting here (clipped to the pair)
    del_idx = np.nonzero(ev & ~is_ins)[0]
    if del_idx.size:
        end_of_pair = (_offsets(qlen) + qlen)[seg[del_idx]]
        stop = np.minimum(del_idx + ln[del_idx], end_of_pair - 1)   # keep at least the last base
        delta = np.zeros(tot + 1, np.int64)
        np.add.at(delta, del_idx, 1)
        np.add.at(delta, stop, -1)
        dropped = np.cumsum(delta[:-1]) > 0
        reps[dropped & ~ins_ev] = 0
    t = np.repeat(base, reps)
    t_lower = np.repeat(lower, reps)
    tseg = np.repeat(seg, reps)
    # inserted copies (all but the first copy of an insertion event) become random bases
    first = np.cumsum(reps) - reps
    is_first = np.zeros(t.shape[0], bool)
    is_first[first[reps > 0]] = True
    rnd = rng.integers(0, 4, t.shape[0]).astype(np.uint8)
    t = np.where(is_first, t, rnd).astype(np.uint8)
    tlen = np.bincount(tseg, minlength=n).astype(np.int32)
    return PairSet(qlen, _offsets(qlen), q, tlen, _offsets(tlen), t,
                   _to_ascii(q, lower), _to_ascii(t, t_lower))


def make_pairs_mixed(n: int, seed: int = 7, min_len: int = 1, max_len: int = 600, div: float = 0.1,
                     n_frac: float = 0.002, burst: int = 30) -> PairSet:
    """Ragged fuzz set: random lengths, substitutions, indel bursts, unrelated tails."""
    rng = np.random.Generator(np.random.PCG64(seed))
    qs, ts, qr, tr = [], [], [], []
    for _ in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        q = rng.integers(0, 4, L).astype(np.uint8)
        q[rng.random(L) < n_frac] = 4
        lower = _softmask(rng, L, mean_run=40)
        out, olow = [], []
        i = 0
        while i < L:
            x = rng.random()
            if x < div * 0.6:
                out.append((int(q[i]) + int(rng.integers(1, 4))) & 3 if q[i] < 4 else 4); olow.append(lower[i]); i += 1
            elif x < div * 0.8:
                i += int(rng.integers(1, burst + 1))
            elif x < div:
                k = int(rng.integers(1, burst + 1))
                out.extend(rng.integers(0, 4, k).tolist()); olow.extend([bool(rng.random() < 0.5)] * k)
            else:
                out.append(int(q[i])); olow.append(lower[i]); i += 1
        if not out:
            out, olow = [int(rng.integers(0, 4))], [False]
        t = np.array(out, np.uint8); tl = np.array(olow, bool)
        qs.append(q); ts.append(t); qr.append(_to_ascii(q, lower)); tr.append(_to_ascii(t, tl))
    qlen = np.array([len(x) for x in qs], np.int32); tlen = np.array([len(x) for x in ts], np.int32)
    return PairSet(qlen, _offsets(qlen), np.concatenate(qs), tlen, _offsets(tlen), np.concatenate(ts),
                   np.concatenate(qr), np.concatenate(tr))


def pairs_from_strings(pairs) -> PairSet:
    """Build a PairSet from [(query_ascii, target_ascii), ...] (str or bytes)."""
    qr = [np.frombuffer(a.encode() if isinstance(a, str) else a, np.uint8) for a, _ in pairs]
    tr = [np.frombuffer(b.encode() if isinstance(b, str) else b, np.uint8) for _, b in pairs]
    qlen = np.array([len(x) for x in qr], np.int32); tlen = np.array([len(x) for x in tr], np.int32)
    qraw = np.concatenate(qr) if qr else np.zeros(0, np.uint8)
    traw = np.concatenate(tr) if tr else np.zeros(0, np.uint8)
    return PairSet(qlen, _offsets(qlen), encode(qraw), tlen, _offsets(tlen), encode(traw), qraw.copy(), traw.copy())


def make_pairs_max_on_last_row(lengths, tail: int = 200, sub: float = 0.04, seed: int = 1) -> PairSet:
    """Pairs whose best score ends on the LAST QUERY ROW at a slot t with (t + 1) % 16 == 0 while the target goes on:
    target = substitution-only copy of the query (no indels, so the copy ends at t = qlen - 1) + `tail` unrelated bases,
    qlen a multiple of 16.  On the anti-diagonal after the maximum the band's lowest 16-slot block leaves the band; a
    kernel that slides that block before it has looked for the arg-max of the previous diagonal loses max_t / max_q."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pairs = []
    for L in lengths:
        L = (int(L) // 16) * 16
        q = rng.integers(0, 4, L)
        t = q.copy()
        hit = rng.random(L) < sub
        t[hit] = (t[hit] + rng.integers(1, 4, int(hit.sum()))) % 4
        t[-24:] = q[-24:]                                      # a clean end: the copy scores best exactly at its last base
        t = np.concatenate([t, rng.integers(0, 4, tail)])
        pairs.append(("".join("ACGT"[x] for x in q), "".join("ACGT"[x] for x in t)))
    return pairs_from_strings(pairs)


def count_cells(qlen: int, tlen: int, w: int) -> int:
    """In-band cells over all anti-diagonals: sum_r (en0 - st0 + 1)
    (reference extern/ksw2_extz2_sse.cc:105-109)."""
    if qlen <= 0 or tlen <= 0:
        return 0
    if w < 0:
        w = max(qlen, tlen)
    r = np.arange(qlen + tlen - 1, dtype=np.int64)
    st = np.maximum(np.maximum(0, r - qlen + 1), (r - w + 1) >> 1)
    en = np.minimum(np.minimum(tlen - 1, r), (r + w) >> 1)
    width = en - st + 1
    bad = np.nonzero(width <= 0)[0]
    if bad.size:
        width = width[:bad[0]]
    return int(width.sum())


def make_region_pair(length: int = 8000, div: float = 0.06, flank: int = 1500, seed: int = 0x5EDEF004):
    """One candidate region pair as the align stage sees it (BASELINE.json configs[0]/[3] shape): a soft-masked
    query region and a reference region holding a diverged copy (makeSmall substitutions/1-bp indels plus a few
    longer indels) between unrelated flanks.  Returns (query_ascii, ref_ascii) as str."""
    rng = np.random.Generator(np.random.PCG64(seed))
    core = make_pairs_small(1, length=length, div=div, seed=seed, n_frac=0.0005)
    q = core.q_raw.copy(); t = core.t_raw.copy()
    # a few longer indels in the copy (the makeLarge idea)
    for _ in range(3):
        p = int(rng.integers(200, max(201, len(t) - 200)))
        k = int(rng.integers(20, 120))
        if rng.random() < 0.5:
            t = np.concatenate([t[:p], t[p + k:]])
        else:
            ins = ASCII[rng.integers(0, 4, k)]
            t = np.concatenate([t[:p], ins, t[p:]])
    def flank_seq(n):
        codes = rng.integers(0, 4, n).astype(np.uint8)
        return _to_ascii(codes, _softmask(rng, n, mean_run=200))
    qs = np.concatenate([flank_seq(flank), q, flank_seq(flank)])
    ts = np.concatenate([flank_seq(flank), t, flank_seq(flank)])
    return qs.tobytes().decode(), ts.tobytes().decode()


def make_genome_with_dups(length: int = 2_000_000, n_dups: int = 40, min_len: int = 5000, max_len: int = 20000,
                          min_div: float = 0.02, max_div: float = 0.10, seed: int = 0x5EDEF001):
    """BASELINE.json configs[0] shape: one soft-masked chromosome with planted segmental duplications (a source segment copied to
    a disjoint place with makeSmall divergence plus a few longer indels).  Returns (genome bytes as np.uint8 ASCII,
    [(src_start, src_end, dst_start, dst_end, div)]) -- the planted catalog, in genome coordinates."""
    rng = np.random.Generator(np.random.PCG64(seed))
    codes = rng.integers(0, 4, length).astype(np.uint8)
    codes[rng.random(length) < 0.0005] = 4
    g = _to_ascii(codes, _softmask(rng, length, mean_run=400))
    # disjoint slots: 2 * n_dups windows of max_len on a grid, shuffled; sources from the first half, destinations from the second
    grid = length // (2 * n_dups + 1)
    assert grid > max_len + 2000, "genome too short for the catalog"
    slots = rng.permutation(2 * n_dups)
    catalog = []
    for k in range(n_dups):
        L = int(rng.integers(min_len, max_len + 1)); div = float(rng.uniform(min_div, max_div))
        s0 = int(slots[2 * k]) * grid + int(rng.integers(500, grid - max_len - 500))
        d0 = int(slots[2 * k + 1]) * grid + int(rng.integers(500, grid - max_len - 500))
        src = g[s0:s0 + L]
        pair = pairs_from_strings([(src.tobytes().decode(), "A")])
        cp, cl, _ = _mutate_small(rng, pair.q[:L], src >= ord("a"), div)
        copy = _to_ascii(cp, cl)
        for _ in range(2):                                   # a few longer indels (the makeLarge idea)
            p = int(rng.integers(200, max(201, len(copy) - 200))); kk = int(rng.integers(20, 120))
            copy = np.concatenate([copy[:p], copy[p + kk:]]) if rng.random() < 0.5 else np.concatenate([copy[:p], ASCII[rng.integers(0, 4, kk)], copy[p:]])
        copy = copy[:max_len + 400]
        g[d0:d0 + len(copy)] = copy
        catalog.append((s0, s0 + L, d0, d0 + len(copy), div))
    return g, catalog
