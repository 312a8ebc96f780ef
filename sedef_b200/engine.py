"""ctypes binding of the C ABI in include/ksw2_b200.h (libsedef_b200.so).

This is the host-side mirror of the reference's ksw2 call surface for the hot path
(`ksw_extz2_sse`, reference extern/ksw2.h:50): same argument meaning (query/target as byte codes
0..m-1, `mat` m*m int8, gap open/extend as positive int8, band `w`, `zdrop`, `flag`) and the same
result record (`ksw_extz_t`).  The library needs a CUDA device; there is no CPU fallback --
`load()` raises if the shared object is missing and every call raises `EngineError` when the
device path fails.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEDEF_B200_LIB", os.path.join(HERE, "libsedef_b200.so"))   # env override: A/B builds

KSW_NEG_INF = -0x40000000
KSW_EZ_SCORE_ONLY, KSW_EZ_RIGHT, KSW_EZ_GENERIC_SC, KSW_EZ_APPROX_MAX = 0x01, 0x02, 0x04, 0x08
KSW_EZ_APPROX_DROP, KSW_EZ_EXTZ_ONLY, KSW_EZ_REV_CIGAR = 0x10, 0x40, 0x80


class KswExtz(C.Structure):
    """ksw_extz_t (reference extern/ksw2.h:22-30), 56 bytes."""
    _fields_ = [("max_zd", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int),
                ("mqe_t", C.c_int), ("mte", C.c_int), ("mte_q", C.c_int), ("score", C.c_int),
                ("cigar", C.POINTER(C.c_uint32)), ("m_cigar", C.c_int64), ("n_cigar", C.c_int64)]


class SdStats(C.Structure):
    """sd_stats_t (include/ksw2_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "span", "gaps", "gap_bases", "matches", "mismatches", "indel_a", "indel_b", "alnB", "matchB",
        "mismatchB", "transitionsB", "transversionsB", "uppercaseA", "uppercaseB", "uppercaseMatches",
        "reserved")]


class SdStatsFp(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fracMatch", "fracMatchIndel", "jcK", "k2K", "errorScaled",
                                          "filter_score", "gap_error", "mismatch_error", "total_error")]


STAT_FIELDS = [n for n, _ in SdStats._fields_ if n != "reserved"]
EZ_DTYPE = np.dtype([("max_zd", "<u4"), ("max_q", "<i4"), ("max_t", "<i4"), ("mqe", "<i4"), ("mqe_t", "<i4"),
                     ("mte", "<i4"), ("mte_q", "<i4"), ("score", "<i4"), ("cigar", "<u8"),
                     ("m_cigar", "<i8"), ("n_cigar", "<i8")])
assert EZ_DTYPE.itemsize == C.sizeof(KswExtz) == 56
STATS_DTYPE = np.dtype([(n, "<i4") for n, _ in SdStats._fields_])


class EngineError(RuntimeError):
    def __init__(self, code, detail):
        super().__init__(f"ksw_b200 error {code}: {detail}")
        self.code = code


_lib = None
EXPORTS = ["ksw_b200_strerror", "ksw_b200_last_error", "ksw_b200_init", "ksw_b200_destroy",
           "ksw_b200_num_devices", "ksw_b200_max_slots", "ksw_extz2_b200", "ksw_extz2_batch",
           "ksw_extz2_batch_flat", "ksw_b200_batch_upload", "ksw_b200_batch_run", "ksw_b200_batch_fetch",
           "ksw_b200_batch_launches", "ksw_b200_batch_kernel_ms", "ksw_b200_batch_cells",
           "ksw_b200_batch_free", "ksw_b200_count_cells", "sd_stats_derive_fp",
           "ksw_b200_batch_io_bytes", "ksw_b200_batch_set_stats", "ksw_b200_free_cigars", "ksw_b200_batch_host_ms",
           "ksw_b200_last_call_io", "ksw_b200_set_host_threads", "sd_stats_from_cigar_batch_flat",
           "ksw_b200_host_alloc", "ksw_b200_host_free", "ksw_b200_host_register", "ksw_b200_host_unregister",
           "ksw_b200_set_fatal_handler", "ksw_extz2_batch_arena", "ksw_b200_result_ez", "ksw_b200_result_stats",
           "ksw_b200_result_count", "ksw_b200_result_io", "ksw_b200_result_free", "ksw_b200_batch_fetch_arena",
           "ksw_b200_result_export", "ksw_b200_result_trims", "sedef_anchors_batch", "sedef_b200_chain_anchors",
           "sedef_b200_chunk_plan", "sedef_b200_align_generate", "sedef_b200_align_generate_error",
           "sedef_b200_fasta_fetch", "sedef_b200_bed_schedule", "sedef_b200_reverse_complement",
           "sedef_b200_stats_generate", "sedef_b200_stats_generate_error", "sedef_b200_stats_pieces"]


def load():
    """Load libsedef_b200.so (built by __graft_entry__.build()). Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.ksw_b200_strerror.restype = C.c_char_p
    lib.ksw_b200_strerror.argtypes = [i32]
    lib.ksw_b200_last_error.restype = C.c_char_p
    lib.ksw_b200_init.argtypes = [i32, i32]
    lib.ksw_b200_init.restype = i32
    lib.ksw_b200_count_cells.argtypes = [i32, i32, i32]
    lib.ksw_b200_count_cells.restype = i64
    lib.ksw_extz2_b200.argtypes = [vp, i32, vp, i32, vp, C.c_int8, vp, C.c_int8, C.c_int8, i32, i32, i32, vp]
    lib.ksw_extz2_b200.restype = None
    flat = [i32, vp, vp, vp, vp, vp, vp, C.c_int8, vp, C.c_int8, C.c_int8, i32, i32, i32]
    lib.ksw_extz2_batch_flat.argtypes = flat + [vp, vp, vp, vp]
    lib.ksw_extz2_batch_flat.restype = i32
    lib.ksw_extz2_batch.argtypes = [i32, vp, vp, vp, vp, C.c_int8, vp, C.c_int8, C.c_int8, i32, i32, i32, vp, vp, vp, vp]
    lib.ksw_extz2_batch.restype = i32
    lib.ksw_b200_batch_upload.argtypes = flat + [vp, vp, C.POINTER(i32)]
    lib.ksw_b200_batch_upload.restype = vp
    lib.ksw_b200_batch_run.argtypes = [vp, C.POINTER(C.c_float)]
    lib.ksw_b200_batch_run.restype = i32
    lib.ksw_b200_batch_fetch.argtypes = [vp, vp, vp]
    lib.ksw_b200_batch_fetch.restype = i32
    lib.ksw_b200_batch_launches.argtypes = [vp]
    lib.ksw_b200_batch_launches.restype = i32
    lib.ksw_b200_batch_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.ksw_b200_batch_cells.argtypes = [vp]
    lib.ksw_b200_batch_cells.restype = i64
    lib.ksw_b200_batch_free.argtypes = [vp]
    lib.ksw_b200_batch_io_bytes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.ksw_b200_batch_set_stats.argtypes = [vp, i32]
    lib.ksw_b200_free_cigars.argtypes = [vp, i32]
    lib.ksw_b200_batch_host_ms.argtypes = [vp, C.POINTER(C.c_double)]
    lib.ksw_b200_set_host_threads.argtypes = [i32]
    lib.sd_stats_from_cigar_batch_flat.argtypes = [i32] + [vp] * 11
    lib.sd_stats_from_cigar_batch_flat.restype = i32
    lib.ksw_b200_last_call_io.argtypes = [C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    lib.ksw_b200_host_alloc.argtypes = [C.c_size_t]
    lib.ksw_b200_host_alloc.restype = vp
    lib.ksw_b200_host_free.argtypes = [vp]
    lib.ksw_b200_host_free.restype = None
    lib.ksw_b200_host_register.argtypes = [vp, C.c_size_t]
    lib.ksw_b200_host_unregister.argtypes = [vp]
    lib.ksw_b200_set_fatal_handler.argtypes = [vp]
    lib.ksw_b200_set_fatal_handler.restype = None
    lib.ksw_extz2_batch_arena.argtypes = flat + [i32, vp, vp, C.POINTER(vp)]
    lib.ksw_extz2_batch_arena.restype = i32
    lib.ksw_b200_batch_fetch_arena.argtypes = [vp, i32, C.POINTER(vp)]
    lib.ksw_b200_batch_fetch_arena.restype = i32
    lib.ksw_b200_result_ez.argtypes = [vp]
    lib.ksw_b200_result_ez.restype = vp
    lib.ksw_b200_result_stats.argtypes = [vp]
    lib.ksw_b200_result_stats.restype = vp
    lib.ksw_b200_result_trims.argtypes = [vp]
    lib.ksw_b200_result_trims.restype = vp
    lib.ksw_b200_result_count.argtypes = [vp]
    lib.ksw_b200_result_io.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    lib.ksw_b200_result_io.restype = None
    lib.ksw_b200_result_free.argtypes = [vp]
    lib.ksw_b200_result_free.restype = None
    lib.ksw_b200_result_export.argtypes = [vp, vp, vp, vp, i64, i64, vp]
    lib.ksw_b200_result_export.restype = i64
    lib.sedef_anchors_batch.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, C.POINTER(vp), vp]
    lib.sedef_anchors_batch.restype = i32
    lib.sedef_b200_align_generate.argtypes = [C.c_char_p, C.c_char_p, i32, C.c_char_p, i32, i32, vp, vp]
    lib.sedef_b200_align_generate.restype = i32
    lib.sedef_b200_align_generate_error.restype = C.c_char_p
    lib.sedef_b200_stats_generate.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, i32, i32, i32, C.c_double, vp]
    lib.sedef_b200_stats_generate.restype = i32
    lib.sedef_b200_stats_generate_error.restype = C.c_char_p
    lib.sedef_b200_stats_pieces.argtypes = [C.c_char_p, C.c_char_p, i32, i32, C.c_char_p, C.c_longlong]
    lib.sedef_b200_stats_pieces.restype = C.c_longlong
    lib.sedef_b200_fasta_fetch.argtypes = [C.c_char_p, C.c_char_p, i32, C.POINTER(i32), C.c_char_p, C.c_longlong]
    lib.sedef_b200_fasta_fetch.restype = C.c_longlong
    lib.sedef_b200_bed_schedule.argtypes = [C.c_char_p, C.c_char_p, C.c_longlong]
    lib.sedef_b200_bed_schedule.restype = C.c_longlong
    lib.sedef_b200_reverse_complement.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p]
    lib.sd_stats_derive_fp.argtypes = [C.POINTER(SdStats), C.POINTER(SdStatsFp)]
    lib.free = C.CDLL(None).free
    lib.free.argtypes = [vp]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        lib = load()
        raise EngineError(rc, f"{lib.ksw_b200_strerror(rc).decode()} -- {lib.ksw_b200_last_error().decode()}")


def init(first_dev: int = 0, ndev: int = 0) -> int:
    n = load().ksw_b200_init(first_dev, ndev)
    if n <= 0:
        _check(n if n < 0 else -1)
    return n


def set_host_threads(n: int) -> None:
    load().ksw_b200_set_host_threads(int(n))


def count_cells(qlen: int, tlen: int, w: int) -> int:
    return int(load().ksw_b200_count_cells(qlen, tlen, w))


class BatchResult:
    """Results of a batch: `ez` structured array (EZ_DTYPE), per-pair CIGAR lists, optional stats array."""

    def __init__(self, ez: np.ndarray, cigars, stats):
        self.ez, self.cigars, self.stats = ez, cigars, stats

    def fields(self, i: int) -> dict:
        e = self.ez[i]
        return dict(max=int(e["max_zd"]) & 0x7FFFFFFF, zdropped=int(e["max_zd"]) >> 31, max_q=int(e["max_q"]),
                    max_t=int(e["max_t"]), mqe=int(e["mqe"]), mqe_t=int(e["mqe_t"]), mte=int(e["mte"]),
                    mte_q=int(e["mte_q"]), score=int(e["score"]), n_cigar=int(e["n_cigar"]))

    def stats_dict(self, i: int) -> dict:
        return {n: int(self.stats[i][n]) for n in STAT_FIELDS}


def _collect(lib, ez: np.ndarray, keep_cigars: bool):
    cigs = None
    if keep_cigars:
        cigs = []
        for i in range(ez.shape[0]):
            n, p = int(ez[i]["n_cigar"]), int(ez[i]["cigar"])
            cigs.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,)).copy() if n else np.zeros(0, np.uint32))
    n_keep, m_keep = ez["n_cigar"].copy(), ez["m_cigar"].copy()
    lib.ksw_b200_free_cigars(ez.ctypes.data, ez.shape[0])         # frees every ez[i].cigar and clears the fields
    ez["n_cigar"], ez["m_cigar"] = n_keep, m_keep                  # keep the counts visible to the Python caller
    return cigs


def _ptr(a):
    return None if a is None else a.ctypes.data


def extz2_batch(ps, mat, q: int, e: int, w: int = -1, zdrop: int = -1, flag: int = 0, m: int = 5,
                want_stats: bool = True, use_raw: bool = True, keep_cigars: bool = True, raw_only: bool = False) -> BatchResult:
    """One-shot batch through `ksw_extz2_batch_flat` with HOST buffers (H2D + kernels + D2H); every CIGAR is malloc()'d by
    the library like ksw2 does.  raw_only: pass the original-case bytes alone (codes derived on the device)."""
    lib = load()
    mat = np.ascontiguousarray(mat, np.int8)
    n = ps.n
    ez = np.zeros(n, EZ_DTYPE)
    stats = np.zeros(n, STATS_DTYPE) if want_stats else None
    use_raw = use_raw or raw_only
    rc = lib.ksw_extz2_batch_flat(n, _ptr(ps.qlen), _ptr(ps.qoff), None if raw_only else _ptr(ps.q), _ptr(ps.tlen), _ptr(ps.toff),
                                  None if raw_only else _ptr(ps.t),
                                  m, _ptr(mat), q, e, w, zdrop, flag, _ptr(ez), _ptr(stats),
                                  _ptr(ps.q_raw) if use_raw else None, _ptr(ps.t_raw) if use_raw else None)
    _check(rc)
    cigs = _collect(lib, ez, keep_cigars)
    return BatchResult(ez, cigs, stats)


class ArenaResult:
    """Results of `ksw_extz2_batch_arena` / `ksw_b200_batch_fetch_arena`: numpy VIEWS of the library's page-locked arena
    (no copies); `ez[i]["cigar"]` is a pointer into the arena.  Valid until free()."""

    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle
        n = int(lib.ksw_b200_result_count(handle))
        self.n = n
        pe = lib.ksw_b200_result_ez(handle)
        self.ez = np.frombuffer((C.c_char * (n * EZ_DTYPE.itemsize)).from_address(pe), EZ_DTYPE) if n else np.zeros(0, EZ_DTYPE)
        pst = lib.ksw_b200_result_stats(handle)
        self.stats = (np.frombuffer((C.c_char * (n * STATS_DTYPE.itemsize)).from_address(pst), STATS_DTYPE) if (pst and n) else None)
        ptr = lib.ksw_b200_result_trims(handle)
        self.trims = (np.frombuffer((C.c_char * (n * 8)).from_address(ptr), np.int32).reshape(n, 2) if (ptr and n) else None)

    def cigar(self, i: int) -> np.ndarray:
        n, p = int(self.ez[i]["n_cigar"]), int(self.ez[i]["cigar"])
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,)).copy() if n else np.zeros(0, np.uint32)

    def io(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int(0)
        self.lib.ksw_b200_result_io(self.h, C.byref(a), C.byref(b), C.byref(c))
        return int(a.value), int(b.value), int(c.value)

    def export(self, ez_dst: np.ndarray, stats_dst, cigar_dst: np.ndarray, cigar_base: int = 0, index=None) -> int:
        """`ksw_b200_result_export`: position-independent copy into caller-owned (e.g. shared-memory) arrays."""
        n = self.lib.ksw_b200_result_export(self.h, _ptr(ez_dst), _ptr(stats_dst), _ptr(cigar_dst), int(cigar_dst.shape[0]),
                                            int(cigar_base), _ptr(index))
        if n < 0:
            _check(int(n))
        return int(n)

    def to_batch_result(self) -> BatchResult:
        """Deep copy into the BatchResult form the tests compare (the arena can be freed afterwards)."""
        ez = self.ez.copy()
        cigs = [self.cigar(i) for i in range(self.n)]
        return BatchResult(ez, cigs, None if self.stats is None else self.stats.copy())

    def free(self):
        if self.h:
            self.ez = self.stats = self.trims = None
            self.lib.ksw_b200_result_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def extz2_batch_arena(ps, mat, q: int, e: int, w: int = -1, zdrop: int = -1, flag: int = 0, m: int = 5,
                      want_stats: bool = True, raw_only: bool = True, use_raw: bool = True) -> ArenaResult:
    """One-shot batch through `ksw_extz2_batch_arena`: host buffers in, one result arena out (no per-pair malloc).
    raw_only (default): the sequences go in as original-case bytes alone, SEDEF's Alignment(fa, fb) call shape."""
    lib = load()
    mat = np.ascontiguousarray(mat, np.int8)
    out = C.c_void_p(None)
    use_raw = use_raw or raw_only
    rc = lib.ksw_extz2_batch_arena(ps.n, _ptr(ps.qlen), _ptr(ps.qoff), None if raw_only else _ptr(ps.q), _ptr(ps.tlen), _ptr(ps.toff),
                                   None if raw_only else _ptr(ps.t), m, _ptr(mat), q, e, w, zdrop, flag, int(want_stats),
                                   _ptr(ps.q_raw) if use_raw else None, _ptr(ps.t_raw) if use_raw else None, C.byref(out))
    _check(rc)
    return ArenaResult(lib, out.value)


class PinnedArray:
    """A numpy view of page-locked host memory from `ksw_b200_host_alloc` (inputs there are DMA-copied in place)."""

    def __init__(self, src: np.ndarray):
        lib = load()
        self.lib = lib
        nbytes = max(1, src.nbytes)
        self.ptr = lib.ksw_b200_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("ksw_b200_host_alloc failed")
        self.array = np.frombuffer((C.c_char * nbytes).from_address(self.ptr), src.dtype, count=src.size).reshape(src.shape)
        self.array[...] = src

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.ksw_b200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pin_pairset(ps):
    """Copy of a PairSet whose sequence buffers live in page-locked memory.  Returns (pinned PairSet, keepalive list)."""
    import dataclasses
    keep = {k: PinnedArray(getattr(ps, k)) for k in ("q", "t", "q_raw", "t_raw")}
    return dataclasses.replace(ps, **{k: v.array for k, v in keep.items()}), list(keep.values())


def last_call_io():
    """(h2d_bytes, d2h_bytes, kernel_launches) of the last one-shot batch call of this thread."""
    a, b, c = C.c_int64(0), C.c_int64(0), C.c_int(0)
    load().ksw_b200_last_call_io(C.byref(a), C.byref(b), C.byref(c))
    return int(a.value), int(b.value), int(c.value)


def extz2(query, target, mat, q: int, e: int, w: int = -1, zdrop: int = -1, flag: int = 0, m: int = 5):
    """Single pair through the ksw2-compatible entry point `ksw_extz2_b200` -> (fields, cigar list)."""
    lib = load()
    query = np.ascontiguousarray(query, np.uint8); target = np.ascontiguousarray(target, np.uint8)
    mat = np.ascontiguousarray(mat, np.int8)
    ez = np.zeros(1, EZ_DTYPE)
    lib.ksw_extz2_b200(None, len(query), _ptr(query), len(target), _ptr(target), m, _ptr(mat), q, e, w, zdrop, flag, _ptr(ez))
    cigs = _collect(lib, ez, True)
    r = BatchResult(ez, cigs, None)
    return r.fields(0), cigs[0].tolist()


class ResidentBatch:
    """Inputs resident in HBM (`ksw_b200_batch_upload`); `run()` times the device path only."""

    def __init__(self, ps, mat, q, e, w=-1, zdrop=-1, flag=0, m=5, use_raw=True, raw_only=False):
        lib = load()
        self.lib, self.n = lib, ps.n
        self._keep = (ps, np.ascontiguousarray(mat, np.int8))
        err = C.c_int(0)
        use_raw = use_raw or raw_only
        self.h = lib.ksw_b200_batch_upload(ps.n, _ptr(ps.qlen), _ptr(ps.qoff), None if raw_only else _ptr(ps.q), _ptr(ps.tlen), _ptr(ps.toff),
                                           None if raw_only else _ptr(ps.t), m, _ptr(self._keep[1]), q, e, w, zdrop, flag,
                                           _ptr(ps.q_raw) if use_raw else None, _ptr(ps.t_raw) if use_raw else None,
                                           C.byref(err))
        if not self.h:
            _check(err.value)

    def run(self) -> float:
        ms = C.c_float(0)
        _check(self.lib.ksw_b200_batch_run(self.h, C.byref(ms)))
        return float(ms.value)

    def kernel_ms(self):
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        self.lib.ksw_b200_batch_kernel_ms(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(dp_ms=a.value, tb_ms=b.value, aux_ms=c.value)

    def launches(self) -> int:
        return int(self.lib.ksw_b200_batch_launches(self.h))

    def cells(self) -> int:
        return int(self.lib.ksw_b200_batch_cells(self.h))

    def io_bytes(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self.lib.ksw_b200_batch_io_bytes(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def host_ms(self) -> dict:
        a = (C.c_double * 5)()
        self.lib.ksw_b200_batch_host_ms(self.h, a)
        return dict(zip(("plan", "pack", "h2d", "d2h", "gather"), [float(x) for x in a]))

    def set_stats(self, on: bool):
        self.lib.ksw_b200_batch_set_stats(self.h, int(on))

    def fetch_arena(self, want_stats=True) -> ArenaResult:
        out = C.c_void_p(None)
        _check(self.lib.ksw_b200_batch_fetch_arena(self.h, int(want_stats), C.byref(out)))
        return ArenaResult(self.lib, out.value)

    def fetch(self, want_stats=True, keep_cigars=True) -> BatchResult:
        ez = np.zeros(self.n, EZ_DTYPE)
        stats = np.zeros(self.n, STATS_DTYPE) if want_stats else None
        _check(self.lib.ksw_b200_batch_fetch(self.h, _ptr(ez), _ptr(stats)))
        cigs = _collect(self.lib, ez, keep_cigars)
        return BatchResult(ez, cigs, stats)

    def free(self):
        if self.h:
            self.lib.ksw_b200_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def stats_from_cigars(cigars, a_list, b_list):
    """Batched `Alignment(fa, fb, cigar)` statistics (sd_stats_from_cigar_batch_flat): cigars = raw ksw ops per alignment,
    a_list/b_list = original-case byte arrays.  Returns (stats structured array, status int32 array)."""
    lib = load()
    n = len(cigars)
    cn = np.array([len(c) for c in cigars], np.int64); coff = np.zeros(n, np.int64); coff[1:] = np.cumsum(cn[:-1])
    cbuf = np.concatenate([np.asarray(c, np.uint32) for c in cigars] + [np.zeros(1, np.uint32)])
    al = np.array([len(x) for x in a_list], np.int32); ao = np.zeros(n, np.int64); ao[1:] = np.cumsum(al[:-1])
    bl = np.array([len(x) for x in b_list], np.int32); bo = np.zeros(n, np.int64); bo[1:] = np.cumsum(bl[:-1])
    ab = np.concatenate([np.asarray(x, np.uint8) for x in a_list] + [np.zeros(1, np.uint8)])
    bb = np.concatenate([np.asarray(x, np.uint8) for x in b_list] + [np.zeros(1, np.uint8)])
    out = np.zeros(n, STATS_DTYPE); status = np.zeros(n, np.int32)
    _check(lib.sd_stats_from_cigar_batch_flat(n, _ptr(coff), _ptr(cn), _ptr(cbuf), _ptr(al), _ptr(ao), _ptr(ab),
                                              _ptr(bl), _ptr(bo), _ptr(bb), _ptr(out), _ptr(status)))
    return out, status


def anchors_batch(regions, kmer_size: int = 11, same_chr=None, orig_query_start=None, orig_ref_start=None):
    """`sedef_anchors_batch`: generate_anchors (src/chain.cc:24-101) for a list of (query, ref) original-case str / bytes pairs on
    the GPU.  Returns one int32 array [k, 4] = (q, r, l, has_u) per region, in the reference's order."""
    lib = load()
    n = len(regions)
    qs = [x[0].encode() if isinstance(x[0], str) else bytes(x[0]) for x in regions]
    rs = [x[1].encode() if isinstance(x[1], str) else bytes(x[1]) for x in regions]
    ql = np.array([len(x) for x in qs], np.int32); rl = np.array([len(x) for x in rs], np.int32)
    qo = np.zeros(n, np.int64); ro = np.zeros(n, np.int64)
    if n > 1:
        qo[1:] = np.cumsum(ql[:-1]); ro[1:] = np.cumsum(rl[:-1])
    qb = np.frombuffer(b"".join(qs) + b"\0", np.uint8); rb = np.frombuffer(b"".join(rs) + b"\0", np.uint8)
    sc = None if same_chr is None else np.ascontiguousarray(same_chr, np.uint8)
    oq = None if orig_query_start is None else np.ascontiguousarray(orig_query_start, np.int64)
    orr = None if orig_ref_start is None else np.ascontiguousarray(orig_ref_start, np.int64)
    out = C.c_void_p(None); off = np.zeros(n + 1, np.int64)
    _check(lib.sedef_anchors_batch(n, _ptr(ql), _ptr(qo), _ptr(qb), _ptr(rl), _ptr(ro), _ptr(rb), kmer_size, _ptr(sc), _ptr(oq), _ptr(orr),
                                   C.byref(out), _ptr(off)))
    tot = int(off[n])
    flat = np.ctypeslib.as_array(C.cast(out.value, C.POINTER(C.c_int32)), shape=(max(tot, 1), 4))[:tot].copy() if out.value else np.zeros((0, 4), np.int32)
    if out.value:
        lib.free(out.value)
    return [flat[int(off[i]):int(off[i + 1])] for i in range(n)]


def align_generate(ref_path: str, bed_path: str, out_path: str, kmer_size: int = 11, shard_index: int = 0, shard_count: int = 1) -> dict:
    """`sedef align generate -k kmer_size ref_path bed_path > out_path` (src/align_main.cc:285-337) through `fast_align_batch`:
    every seed hit of the bucket file(s) at once.  Returns the run's counters and phase times."""
    lib = load()
    st = np.zeros(9, np.int64); ms = np.zeros(3, np.float64)
    rc = lib.sedef_b200_align_generate(os.fsencode(ref_path), os.fsencode(bed_path), int(kmer_size), os.fsencode(out_path),
                                       int(shard_index), int(shard_count), _ptr(st), _ptr(ms))
    if rc != 0:
        raise EngineError(rc, lib.sedef_b200_align_generate_error().decode())
    keys = ["regions", "hits", "groups", "rounds", "batch_calls", "ksw_requests", "region_bytes", "ksw_pairs", "ksw_cells"]
    out = {k: int(v) for k, v in zip(keys, st)}
    out.update(ms_total=float(ms[0]), ms_align=float(ms[1]), ms_io=float(ms[2]))
    return out


def stats_generate(ref_path: str, bed_path: str, out_path: str, max_ok_gap: int = -1, min_split: int = 1000, min_uppercase: int = 100,
                   max_scaled_error: float = 0.5) -> dict:
    """`sedef stats generate` (src/stats_main.cc:338-395,513-537): the SD report of an aligned.bed; statistics of all pieces in one
    GPU call."""
    lib = load()
    cnt = np.zeros(3, np.int64)
    rc = lib.sedef_b200_stats_generate(os.fsencode(ref_path), os.fsencode(bed_path), os.fsencode(out_path), int(max_ok_gap), int(min_split),
                                       int(min_uppercase), float(max_scaled_error), _ptr(cnt))
    if rc != 0:
        raise EngineError(rc, lib.sedef_b200_stats_generate_error().decode())
    return dict(hits=int(cnt[0]), pieces=int(cnt[1]), lines=int(cnt[2]))


def stats_pieces(ref_path: str, bed_path: str, max_ok_gap: int = -1, min_split: int = 1000):
    """Host-only half of `stats generate`: the pieces it would measure, as (qname, qs, qe, rname, rs, re, strand_q, strand_r, span,
    cigar) tuples."""
    lib = load()
    n = lib.sedef_b200_stats_pieces(os.fsencode(ref_path), os.fsencode(bed_path), int(max_ok_gap), int(min_split), None, 0)
    if n < 0:
        raise EngineError(-1, lib.sedef_b200_stats_generate_error().decode())
    buf = C.create_string_buffer(int(n) + 1)
    lib.sedef_b200_stats_pieces(os.fsencode(ref_path), os.fsencode(bed_path), int(max_ok_gap), int(min_split), buf, n)
    out = []
    for ln in buf.raw[:n].decode().split("\n"):
        if ln:
            f = ln.split("\t")
            out.append((f[0], int(f[1]), int(f[2]), f[3], int(f[4]), int(f[5]), f[6], f[7], int(f[8]), f[9] if len(f) > 9 else ""))
    return out


def fasta_fetch(ref_path: str, name: str, start: int, end: int):
    """FastaReference::get_sequence (src/fasta.cc:106-143): (bases, clamped end)."""
    lib = load()
    e = C.c_int(end)
    cap = max(0, end - max(start, 0)) + 16
    buf = C.create_string_buffer(cap)
    n = lib.sedef_b200_fasta_fetch(os.fsencode(ref_path), name.encode(), int(start), C.byref(e), buf, cap)
    if n < 0:
        raise EngineError(-1, lib.sedef_b200_align_generate_error().decode())
    return buf.raw[:n], e.value


def bed_schedule(bed_path: str) -> str:
    """The seed hits of a bucket file / directory in the order `align generate` processes them (src/align_main.cc:211-283)."""
    lib = load()
    n = lib.sedef_b200_bed_schedule(os.fsencode(bed_path), None, 0)
    if n < 0:
        raise EngineError(-1, lib.sedef_b200_align_generate_error().decode())
    buf = C.create_string_buffer(int(n) + 1)
    lib.sedef_b200_bed_schedule(os.fsencode(bed_path), buf, n)
    return buf.raw[:n].decode()


def reverse_complement(s: bytes) -> bytes:
    out = C.create_string_buffer(len(s) + 1)
    load().sedef_b200_reverse_complement(s, len(s), out)
    return out.raw[:len(s)]


def derive_fp(stats_row) -> dict:
    """Floating-point BEDPE fields from the integer record (host-side, reference src/stats_main.cc:273-283)."""
    lib = load()
    s = SdStats(**{n: int(stats_row[n]) for n in STAT_FIELDS})
    o = SdStatsFp()
    lib.sd_stats_derive_fp(C.byref(s), C.byref(o))
    return {n: getattr(o, n) for n, _ in SdStatsFp._fields_}
