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
           "ksw_b200_batch_io_bytes", "ksw_b200_batch_set_stats", "ksw_b200_free_cigars", "ksw_b200_batch_host_ms", "ksw_b200_last_call_io", "ksw_b200_set_host_threads", "sd_stats_from_cigar_batch_flat"]


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
                want_stats: bool = True, use_raw: bool = True, keep_cigars: bool = True) -> BatchResult:
    """One-shot batch through `ksw_extz2_batch_flat` with HOST buffers (H2D + kernels + D2H)."""
    lib = load()
    mat = np.ascontiguousarray(mat, np.int8)
    n = ps.n
    ez = np.zeros(n, EZ_DTYPE)
    stats = np.zeros(n, STATS_DTYPE) if want_stats else None
    rc = lib.ksw_extz2_batch_flat(n, _ptr(ps.qlen), _ptr(ps.qoff), _ptr(ps.q), _ptr(ps.tlen), _ptr(ps.toff), _ptr(ps.t),
                                  m, _ptr(mat), q, e, w, zdrop, flag, _ptr(ez), _ptr(stats),
                                  _ptr(ps.q_raw) if use_raw else None, _ptr(ps.t_raw) if use_raw else None)
    _check(rc)
    cigs = _collect(lib, ez, keep_cigars)
    return BatchResult(ez, cigs, stats)


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

    def __init__(self, ps, mat, q, e, w=-1, zdrop=-1, flag=0, m=5, use_raw=True):
        lib = load()
        self.lib, self.n = lib, ps.n
        self._keep = (ps, np.ascontiguousarray(mat, np.int8))
        err = C.c_int(0)
        self.h = lib.ksw_b200_batch_upload(ps.n, _ptr(ps.qlen), _ptr(ps.qoff), _ptr(ps.q), _ptr(ps.tlen), _ptr(ps.toff),
                                           _ptr(ps.t), m, _ptr(self._keep[1]), q, e, w, zdrop, flag,
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


def derive_fp(stats_row) -> dict:
    """Floating-point BEDPE fields from the integer record (host-side, reference src/stats_main.cc:273-283)."""
    lib = load()
    s = SdStats(**{n: int(stats_row[n]) for n in STAT_FIELDS})
    o = SdStatsFp()
    lib.sd_stats_derive_fp(C.byref(s), C.byref(o))
    return {n: getattr(o, n) for n, _ in SdStatsFp._fields_}
