"""sedef_b200 -- B200-native batched ksw_extz2 engine for SEDEF's alignment hot path.

Only the pieces the hot path needs live here: `csrc/` (CUDA kernels + the C-ABI library
`libsedef_b200.so`, declared in include/ksw2_b200.h), `engine.py` (ctypes binding of that C ABI:
the host-side mirror of ksw2's call surface), `align.py` (mirror of SEDEF's `Alignment(fa, fb)`
front end for this path) and `synth.py` (deterministic synthetic workloads).
"""
__all__ = ["engine", "align", "synth"]
