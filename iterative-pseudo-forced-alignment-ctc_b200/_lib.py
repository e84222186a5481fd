"""ctypes binding of libipfa_b200.so (the C ABI declared in include/ipfa_b200.h).

The product path has no CPU fallback: if the shared library cannot be built or
loaded, importing this module raises."""
import ctypes
import os
import re

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "ipfa_b200.h")

c_void = ctypes.c_void_p
c_i = ctypes.c_int
c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t
c_d = ctypes.c_double

_SIGNATURES = {
    "ipfa_version": (c_i, []),
    "ipfa_status_string": (ctypes.c_char_p, [c_i]),
    "ipfa_last_cuda_error": (ctypes.c_char_p, []),
    "ipfa_launch_count": (ctypes.c_uint64, []),
    "ipfa_device_count": (c_i, []),
    "ipfa_tuning_reload": (None, []),
    "ipfa_profile_kernels": (None, [c_i]),
    "ipfa_profile_read_ms": (ctypes.c_float, []),
    "ipfa_ctc_alpha_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i]),
    "ipfa_ctc_alpha_device": (c_i, [c_void, c_i64, c_i64, c_void, c_i64, c_void, c_void,
                                    c_i, c_i, c_i, c_i, c_i, c_void, c_void, c_sz, c_void]),
    "ipfa_ctc_alpha_host": (c_i, [c_void, c_i64, c_i64, c_void, c_i64, c_void, c_void,
                                  c_i, c_i, c_i, c_i, c_i, c_void]),
    "ipfa_ctc_viterbi_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i]),
    "ipfa_ctc_viterbi_device": (c_i, [c_void, c_i64, c_i64, c_void, c_i64, c_void, c_void,
                                      c_i, c_i, c_i, c_i, c_i,
                                      c_void, c_void, c_void, c_void, c_void, c_void, c_void,
                                      c_void, c_sz, c_void]),
    "ipfa_ctc_alpha_strided_device": (c_i, [c_void, c_i64, c_i64, c_i64, c_void, c_i64, c_void, c_void,
                                            c_i, c_i, c_i, c_i, c_i, c_void, c_void, c_sz, c_void]),
    "ipfa_ctc_viterbi_strided_device": (c_i, [c_void, c_i64, c_i64, c_i64, c_void, c_i64, c_void, c_void,
                                              c_i, c_i, c_i, c_i, c_i,
                                              c_void, c_void, c_void, c_void, c_void, c_void, c_void,
                                              c_void, c_sz, c_void]),
    "ipfa_ctc_viterbi_host": (c_i, [c_void, c_i64, c_i64, c_void, c_i64, c_void, c_void,
                                    c_i, c_i, c_i, c_i, c_i,
                                    c_void, c_void, c_void, c_void, c_void, c_void, c_void]),
    "ipfa_ctcseg_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i]),
    "ipfa_ctcseg_device": (c_i, [c_void, c_i64, c_i64, c_void, c_void, c_i64, c_void, c_void, c_void,
                                 c_i, c_i, c_i, c_i, c_i, c_i, c_d, c_i, c_i,
                                 c_void, c_void, c_void, c_void, c_void, c_void,
                                 c_void, c_sz, c_void]),
    "ipfa_ctcseg_host": (c_i, [c_void, c_i64, c_i64, c_void, c_void, c_i64, c_void, c_void, c_void,
                               c_i, c_i, c_i, c_i, c_i, c_i, c_d, c_i, c_i,
                               c_void, c_void, c_void, c_void, c_void, c_void]),
    "ipfa_anchor_select_device": (c_i, [c_void, c_void, c_void, c_void, c_i, c_i, c_d, c_i,
                                        c_void, c_void, c_void]),
    "ipfa_text_round_device": (c_i, [c_void, c_i64, c_i, c_void, c_void]),
    "ipfa_ctcseg_windows_device": (c_i, [c_void, c_void, c_i64, c_void, c_void, c_i64, c_void, c_void, c_void,
                                         c_i, c_i, c_i, c_i, c_i, c_i, c_d, c_i, c_i,
                                         c_void, c_void, c_void, c_void, c_void, c_void,
                                         c_void, c_sz, c_void]),
    "ipfa_ctcseg_windowed_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i, c_i, c_i]),
    "ipfa_ctcseg_windowed_device": (c_i, [c_void, c_void, c_i64, c_i64, c_void, c_void, c_i64, c_void, c_void,
                                          c_void, c_i, c_i, c_i, c_i, c_i, c_i, c_d, c_i, c_i, c_i, c_i,
                                          c_void, c_void, c_void, c_void, c_void, c_void,
                                          c_void, c_sz, c_void]),
    "ipfa_sweep_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i]),
    "ipfa_sweep_step_device": (c_i, [c_void, c_void, c_void, c_void, c_void, c_i, c_i, c_i, c_i, c_i,
                                     c_void, c_sz, c_void]),
    "ipfa_sweep_resident_workspace_bytes": (c_sz, [c_void, c_void, c_i, c_i, c_i]),
    "ipfa_sweep_resident_device": (c_i, [c_void, c_void, c_void, c_void, c_void, c_i, c_i, c_i,
                                         c_void, c_sz, c_void]),
}


def declared_symbols():
    """Every function include/ipfa_b200.h declares (used by the CPU-side ABI test)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipfa_[a-z0-9_]+)\s*\(", text)))


ABI_VERSION = 202  # ipfa_version(): bumped whenever a signature or struct in include/ipfa_b200.h changes
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB
        if not _build.up_to_date():
            try:
                path = _build.build()
            except Exception as exc:
                # no nvcc on the box.  A shipped .so older than its sources may have another ABI than the
                # signatures below: loading it is an explicit decision, not a fallback.
                if not os.path.exists(path):
                    raise ImportError(f"libipfa_b200.so is missing and could not be built: {exc}")
                if os.environ.get("IPFA_ALLOW_STALE_LIB") != "1":
                    raise ImportError(f"libipfa_b200.so is older than csrc/ or include/ and could not be rebuilt "
                                      f"({exc}); set IPFA_ALLOW_STALE_LIB=1 to load it anyway")
        L = ctypes.CDLL(path)
        L.ipfa_version.restype = ctypes.c_int
        if L.ipfa_version() != ABI_VERSION:
            raise ImportError(f"{path} reports ABI version {L.ipfa_version()}, this package binds {ABI_VERSION}")
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class IpfaError(RuntimeError):
    def __init__(self, status, where):
        L = lib()
        msg = L.ipfa_status_string(status).decode()
        if status == 4:
            msg += ": " + L.ipfa_last_cuda_error().decode()
        super().__init__(f"{where}: {msg} (status {status})")
        self.status = status


def check(status, where):
    if status == 5:
        # the reference's callers catch AssertionError for this condition
        # (/root/reference/src/iterative_utterance_alignment.py:390)
        raise AssertionError("Audio is shorter than text!")
    if status == 6:
        # ctc-segmentation gives up with IndexError when doubling the table window does not help
        raise IndexError("Maximum window size reached. Check data for large repetitions or noise.")
    if status != 0:
        raise IpfaError(status, where)
