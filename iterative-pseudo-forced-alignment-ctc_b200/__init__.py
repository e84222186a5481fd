"""B200-native alignment hot path of ferugit/iterative-pseudo-forced-alignment-ctc.

Hand-written sm_100a CUDA kernels behind a C ABI (include/ipfa_b200.h,
libipfa_b200.so) plus the host-side mirror of the reference's aligner interface.
The directory name carries hyphens, so import it with
``importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200")`` or through
the ``ipfa_b200`` alias module at the repository root.
"""
from . import _lib  # noqa: F401
from .ops import (ctc_alpha_nll, ctc_alpha_nll_host, ctc_forced_align, ctc_forced_align_host,  # noqa: F401
                  ctcseg_align, ctcseg_align_host, anchor_select, launch_count, ctc_alpha_redo_count,
                  ctc_alpha_redo_reasons, text_round, tuning)

__all__ = ["ctc_alpha_nll", "ctc_alpha_nll_host", "ctc_forced_align", "ctc_forced_align_host",
           "ctcseg_align", "ctcseg_align_host", "anchor_select", "launch_count", "ctc_alpha_redo_count", "ctc_alpha_redo_reasons", "text_round", "tuning"]
