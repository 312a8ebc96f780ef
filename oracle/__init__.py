"""oracle -- TEST INFRASTRUCTURE ONLY (see oracle/ksw2_extz2_port.c header).

ctypes loaders for
  * `port`: oracle/liboracle_port.so  -- scalar C restatement of ksw_extz2_sse + SD statistics
  * `ref` : oracle/_ref/libksw2_ref.so -- the reference's own ksw_extz2_sse compiled unmodified
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class KswExtz(C.Structure):
    """Layout of ksw_extz_t (reference extern/ksw2.h:22-30)."""
    _fields_ = [("max_zd", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int),
                ("mqe_t", C.c_int), ("mte", C.c_int), ("mte_q", C.c_int), ("score", C.c_int),
                ("cigar", C.POINTER(C.c_uint32)), ("m_cigar", C.c_int64), ("n_cigar", C.c_int64)]

    @property
    def max(self):
        return self.max_zd & 0x7FFFFFFF

    @property
    def zdropped(self):
        return self.max_zd >> 31

    def cigar_list(self):
        return [int(self.cigar[i]) for i in range(self.n_cigar)]

    def fields(self):
        return dict(max=self.max, zdropped=self.zdropped, max_q=self.max_q, max_t=self.max_t,
                    mqe=self.mqe, mqe_t=self.mqe_t, mte=self.mte, mte_q=self.mte_q, score=self.score,
                    n_cigar=int(self.n_cigar))


assert C.sizeof(KswExtz) == 56


class SdStats(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "span", "gaps", "gap_bases", "matches", "mismatches", "indel_a", "indel_b", "alnB", "matchB",
        "mismatchB", "transitionsB", "transversionsB", "uppercaseA", "uppercaseB", "uppercaseMatches",
        "reserved")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


class OracleDiag(C.Structure):
    _fields_ = [("n_clamp", C.c_int64), ("n_wrap", C.c_int64), ("n_diag", C.c_int64), ("n_cells", C.c_int64)]


def build(force: bool = False) -> None:
    """Compile the checkers (oracle/Makefile). `_ref/` is rebuilt only where /root/reference exists."""
    args = ["make", "-C", HERE, "-s"] + (["-B"] if force else [])
    subprocess.run(args + ["liboracle_port.so"], check=True)
    subprocess.run(args + ["ref_full"], check=False)   # needs sedef_b200/libsedef_b200.so for the redirect variant


_SIG = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int8, C.c_void_p, C.c_int8, C.c_int8,
        C.c_int, C.c_int, C.c_int, C.POINTER(KswExtz)]
_BATCH_SIG = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
              C.c_int8, C.c_void_p, C.c_int8, C.c_int8, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]


class _Lib:
    def __init__(self, path: str, fn: str):
        self.path = path
        self.lib = C.CDLL(path)
        self.fn = getattr(self.lib, fn)
        self.fn.argtypes = _SIG
        self.fn.restype = None
        self.lib.cpu_batch_run.argtypes = _BATCH_SIG
        self.lib.cpu_batch_run.restype = C.c_double
        self.lib.cpu_batch_max_threads.restype = C.c_int
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]

    def extz2(self, query, target, mat, q, e, w=-1, zdrop=-1, flag=0, m=5):
        """One pair -> (fields dict, cigar list)."""
        query = np.ascontiguousarray(query, np.uint8); target = np.ascontiguousarray(target, np.uint8)
        mat = np.ascontiguousarray(mat, np.int8)
        ez = KswExtz()
        self.fn(None, len(query), query.ctypes.data, len(target), target.ctypes.data, m, mat.ctypes.data,
                q, e, w, zdrop, flag, C.byref(ez))
        out = ez.fields(); cig = ez.cigar_list()
        if ez.cigar:
            self.libc.free(ez.cigar)
        return out, cig

    def batch(self, ps, mat, q, e, w=-1, zdrop=-1, flag=0, m=5, nthreads=0, keep=True):
        """All pairs of a synth.PairSet. keep=True -> (seconds, [fields], [cigars]); else seconds."""
        mat = np.ascontiguousarray(mat, np.int8)
        n = ps.n
        ez = (KswExtz * n)() if keep else None
        secs = self.lib.cpu_batch_run(n, ps.qlen.ctypes.data, ps.qoff.ctypes.data, ps.q.ctypes.data,
                                      ps.tlen.ctypes.data, ps.toff.ctypes.data, ps.t.ctypes.data,
                                      m, mat.ctypes.data, q, e, w, zdrop, flag,
                                      C.cast(ez, C.c_void_p) if keep else None, nthreads)
        if not keep:
            return secs
        fields = [ez[i].fields() for i in range(n)]
        self.last_m_cigar = [int(ez[i].m_cigar) for i in range(n)]     # capacity ksw_push_cigar grew the block to
        cigs = [np.ctypeslib.as_array(ez[i].cigar, shape=(int(ez[i].n_cigar),)).copy().tolist()
                if ez[i].n_cigar else [] for i in range(n)]
        for i in range(n):
            if ez[i].cigar:
                self.libc.free(ez[i].cigar)
        return secs, fields, cigs

    def batch_records(self, ps, mat, q, e, w=-1, zdrop=-1, flag=0, m=5, nthreads=0):
        """All pairs -> (seconds, numpy structured view of the ksw_extz_t records incl. live `cigar` pointers, keepalive).
        For whole-benchmark parity checks (no per-pair Python objects); release with free_records(keepalive)."""
        mat = np.ascontiguousarray(mat, np.int8)
        n = ps.n
        ez = (KswExtz * max(1, n))()
        secs = self.lib.cpu_batch_run(n, ps.qlen.ctypes.data, ps.qoff.ctypes.data, ps.q.ctypes.data,
                                      ps.tlen.ctypes.data, ps.toff.ctypes.data, ps.t.ctypes.data,
                                      m, mat.ctypes.data, q, e, w, zdrop, flag, C.cast(ez, C.c_void_p), nthreads)
        dt = np.dtype([("max_zd", "<u4"), ("max_q", "<i4"), ("max_t", "<i4"), ("mqe", "<i4"), ("mqe_t", "<i4"), ("mte", "<i4"),
                       ("mte_q", "<i4"), ("score", "<i4"), ("cigar", "<u8"), ("m_cigar", "<i8"), ("n_cigar", "<i8")])
        return secs, np.frombuffer(ez, dt, count=n), (ez, n)

    def free_records(self, keepalive) -> None:
        ez, n = keepalive
        self.lib.cpu_batch_free_cigars.argtypes = [C.c_int, C.c_void_p]
        self.lib.cpu_batch_free_cigars(n, C.cast(ez, C.c_void_p))

    def max_threads(self) -> int:
        return int(self.lib.cpu_batch_max_threads())


_port = None
_ref = None


def port() -> _Lib:
    global _port
    if _port is None:
        p = os.path.join(HERE, "liboracle_port.so")
        if not os.path.exists(p):
            build()
        _port = _Lib(p, "oracle_ksw_extz2")
        _port.lib.oracle_sd_stats.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                              C.POINTER(SdStats)]
        _port.lib.oracle_sd_stats.restype = C.c_int
        _port.lib.oracle_last_diag.argtypes = [C.POINTER(OracleDiag)]
        _port.lib.oracle_count_cells.argtypes = [C.c_int, C.c_int, C.c_int]
        _port.lib.oracle_count_cells.restype = C.c_int64
    return _port


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libksw2_ref.so"))


def ref() -> _Lib:
    """The compiled reference kernel (oracle/_ref). Raises FileNotFoundError if it was never built."""
    global _ref
    if _ref is None:
        p = os.path.join(HERE, "_ref", "libksw2_ref.so")
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        _ref = _Lib(p, "ksw_extz2_sse")
    return _ref


def sd_stats(cigar, a_raw, b_raw) -> dict:
    """SD statistics of one alignment from raw ksw ops + original-case bytes (oracle/sd_stats_port.c)."""
    cig = np.ascontiguousarray(cigar, np.uint32)
    a = np.ascontiguousarray(a_raw, np.uint8); b = np.ascontiguousarray(b_raw, np.uint8)
    s = SdStats()
    rc = port().lib.oracle_sd_stats(cig.ctypes.data, len(cig), a.ctypes.data, len(a), b.ctypes.data, len(b), C.byref(s))
    if rc != 0:
        raise ValueError("CIGAR overruns a sequence")
    return s.as_dict()


def last_diag() -> dict:
    d = OracleDiag()
    port().lib.oracle_last_diag(C.byref(d))
    return dict(n_clamp=d.n_clamp, n_wrap=d.n_wrap, n_diag=d.n_diag, n_cells=d.n_cells)


def cigar_str(cigar, ops="MID") -> str:
    return "".join(f"{c >> 4}{ops[c & 0xf]}" for c in cigar)
