"""CPU tests of the Alignment(fa, fb) front end's chunk arithmetic (align_helper, reference src/align.cc:46-53;
MAX_KSW_SEQ_LEN = 60 * KB with KB = 1000, src/globals.h:18,54).

The oracle is the UNMODIFIED reference: oracle/_ref/libsedef_ref_rec.so is the reference's own align.cc & co. linked over a
RECORDING stand-in for ksw_extz2_sse (oracle/ksw_record.c), so `Alignment(fa, fb)` runs its real chunk loop and the test
reads back the (qlen, tlen, offset) of every kernel call it made -- no 60 000 x 60 000 DP needed."""
import ctypes as C
import os
import signal
import subprocess
import sys

import numpy as np
import pytest

from sedef_b200 import align, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REC = os.path.join(ROOT, "oracle", "_ref", "libsedef_ref_rec.so")

LENGTHS = [(59999, 59999), (60000, 60000), (60001, 60001), (61440, 61440), (61441, 61440), (60000, 75000), (75000, 60001),
           (120000, 120000), (120001, 130000), (180500, 121000), (1, 1), (5, 70000), (0, 10)]


def reference_calls(alen, blen):
    lib = C.CDLL(REC)
    lib.ref_alignment.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int] + [C.POINTER(C.c_int)] * 5
    lib.ksw_record_get.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ksw_record_reset()
    buf = C.create_string_buffer(64)
    v = [C.c_int(0) for _ in range(5)]
    assert lib.ref_alignment(b"A" * alen, b"C" * blen, buf, 64, *[C.byref(x) for x in v]) == 0
    out = []
    for k in range(lib.ksw_record_count()):
        ql, tl, w, zd, fl = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        qo, to = C.c_int64(), C.c_int64()
        assert lib.ksw_record_get(k, C.byref(ql), C.byref(tl), C.byref(qo), C.byref(to), C.byref(w), C.byref(zd), C.byref(fl)) == 0
        assert qo.value == to.value                       # the same offset on both strings (src/align.cc:50-53)
        assert (w.value, zd.value, fl.value) == (-1, -1, 0)   # Alignment(fa, fb): unbanded, no z-drop, flag 0 (src/align.cc:84-86)
        out.append((qo.value, ql.value, tl.value))
    return out


@pytest.mark.skipif(not os.path.exists(REC), reason="oracle/_ref/libsedef_ref_rec.so not built (needs /root/reference once)")
@pytest.mark.parametrize("alen,blen", LENGTHS)
def test_chunk_plan_matches_reference_align_helper(built, alen, blen):
    exp = reference_calls(alen, blen)
    so = C.CDLL(engine.LIB_PATH)
    so.sedef_b200_chunk_plan.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    sp = np.zeros(16, np.int64); ql = np.zeros(16, np.int32); tl = np.zeros(16, np.int32)
    n = so.sedef_b200_chunk_plan(alen, blen, 16, sp.ctypes.data, ql.ctypes.data, tl.ctypes.data)
    got = [(int(sp[k]), int(ql[k]), int(tl[k])) for k in range(n)]
    assert got == exp, (alen, blen)
    # the Python mirror chunks the same way
    py, s = [], 0
    while s < min(alen, blen):
        py.append((s, min(align.MAX_KSW_SEQ_LEN, alen - s), min(align.MAX_KSW_SEQ_LEN, blen - s)))
        s += align.MAX_KSW_SEQ_LEN
    assert py == exp
    assert align.MAX_KSW_SEQ_LEN == 60000


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_single_pair_api_is_fatal_without_device(built):
    """ksw_extz2_b200 has no error channel (like ksw_extz2_sse): a request it cannot serve must not look like 'no alignment'.
    Default: message + abort(); with a handler installed: the handler sees the code, `ez` holds the reset record."""
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from sedef_b200 import engine, synth\n"
            "engine.extz2(np.zeros(10, np.uint8), np.zeros(10, np.uint8), synth.sedef_matrix(), 40, 1)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == -signal.SIGABRT, (out.returncode, out.stderr)
    assert "no usable CUDA device" in out.stderr
    lib = engine.load()
    seen = []
    HANDLER = C.CFUNCTYPE(None, C.c_int, C.c_char_p)
    cb = HANDLER(lambda c, msg: seen.append((c, msg)))
    lib.ksw_b200_set_fatal_handler(C.cast(cb, C.c_void_p))
    try:
        from sedef_b200 import synth
        fields, cig = engine.extz2(np.zeros(10, np.uint8), np.zeros(10, np.uint8), synth.sedef_matrix(), 40, 1)
    finally:
        lib.ksw_b200_set_fatal_handler(None)
    assert seen and seen[0][0] == -1
    assert fields["score"] == engine.KSW_NEG_INF and fields["max"] == 0 and fields["max_q"] == -1 and cig == []


def test_chain_anchors_matches_reference(built, golden_dir):
    """The host chaining DP (`chain_anchors`, mirror of src/chain.cc:103-199 incl. the tie rules of its range structure,
    src/segment.tpp) and the chain filter of src/chain.cc:222-247: on the reference's own anchors of 13 regions it must return
    exactly the reference's chains (tests/golden/region_golden.json, 'A q r l has_u' -> 'C n i0 i1 ...')."""
    from helpers import load_json
    so = C.CDLL(engine.LIB_PATH)
    so.sedef_b200_chain_anchors.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    g = load_json(golden_dir, "region_golden.json")
    total = 0
    for reg in g["regions"]:
        lines = [ln for ln in reg["text"].split("\n") if ln.strip()]
        anchors = np.array([[int(x) for x in ln.split()[1:5]] for ln in lines if ln[0] == "A"], np.int32)
        want = [[int(x) for x in ln.split()[2:]] for ln in lines if ln[0] == "C"]
        clen = np.zeros(4096, np.int32); cidx = np.zeros(len(anchors) + 16, np.int32)
        n = so.sedef_b200_chain_anchors(len(anchors), anchors.ctypes.data, len(clen), clen.ctypes.data, len(cidx), cidx.ctypes.data)
        got, pos = [], 0
        for k in range(n):
            got.append(cidx[pos:pos + int(clen[k])].tolist()); pos += int(clen[k])
        assert got == want, (reg["seed"], reg["same_chr"], n, len(want))
        total += n
    assert total >= 80
