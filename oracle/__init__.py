"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's alignment hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
(``iterative-pseudo-forced-alignment-ctc_b200``) never does.

Layout
------
``ctc_oracle.c``      standard-CTC alpha (torch ``ctc_loss``) and CTC Viterbi
                      (torchaudio ``forced_align``) in plain C, fp32.
``ctcseg_oracle.c``   ``cython_fill_table`` of ctc-segmentation 1.7.1 in plain C.
``ctcseg.py``         the interpreted half of ctc-segmentation 1.7.1 (token list
                      preparation, tolerance-based backtrace, utterance scoring)
                      and the speechbrain task formatting.
``anchor.py``         the accept/shrink/revert state machine of
                      /root/reference/src/iterative_utterance_alignment.py:203-379.

Pinning status
--------------
* ``ctc`` lattice: pinned against the installed torch / torchaudio through
  ``tests/golden/ctc_golden.npz`` (generator committed beside it).
* ``ctcseg`` lattice: PARITY UNPINNED -- ctc-segmentation and speechbrain are
  neither vendored under /root/reference nor installed, and the reference's
  tests only print.  Hand-checkable known-answer cases live in
  ``tests/test_oracle_ctcseg.py``.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_SOURCES = ["ctc_oracle.c", "ctcseg_oracle.c"]


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (idempotent)."""
    srcs = [os.path.join(_HERE, s) for s in _SOURCES]
    if not force and os.path.exists(_LIB_PATH):
        if all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
            return _LIB_PATH
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-fno-fast-math", "-o", _LIB_PATH] + srcs + ["-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c_f = ctypes.POINTER(ctypes.c_float)
        c_i32 = ctypes.POINTER(ctypes.c_int32)
        c_i64 = ctypes.POINTER(ctypes.c_int64)
        L.oracle_ctc_alpha_batch.argtypes = [
            c_f, ctypes.c_int64, ctypes.c_int64, c_i32, ctypes.c_int64, c_i32, c_i32,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f]
        L.oracle_ctc_alpha_batch.restype = None
        L.oracle_ctc_viterbi_batch.argtypes = [
            c_f, ctypes.c_int64, ctypes.c_int64, c_i32, ctypes.c_int64, c_i32, c_i32,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i32, c_f, c_i32]
        L.oracle_ctc_viterbi_batch.restype = None
        L.oracle_ctcseg_fill.argtypes = [
            c_f, ctypes.c_int, ctypes.c_int, c_f, ctypes.c_int, ctypes.c_int, c_i64,
            ctypes.c_int, c_i64, ctypes.c_int, ctypes.c_int, c_i32]
        L.oracle_ctcseg_fill.restype = ctypes.c_int
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib
