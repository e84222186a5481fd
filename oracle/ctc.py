"""oracle/ctc.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy-facing wrappers over ``ctc_oracle.c`` (standard 2L+1 CTC lattice):
``ctc_alpha_nll`` restates torch ``F.ctc_loss(..., reduction='none')`` (SURVEY.md
section 8(a) row A9) and ``ctc_viterbi`` restates ``torchaudio.functional.forced_align``
(row A8), batched.  Both are pinned against the installed libraries through
``tests/golden/ctc_golden.npz``.
"""
import ctypes

import numpy as np

from . import lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def ctc_alpha_nll(lp, targets, in_len, tgt_len, blank=0):
    """lp: float32 [N, T, V]; targets int [N, Lmax]; returns float32 [N] nll."""
    lp = np.ascontiguousarray(lp, dtype=np.float32)
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    if targets.ndim == 1:
        targets = targets[None]
    in_len = np.ascontiguousarray(in_len, dtype=np.int32)
    tgt_len = np.ascontiguousarray(tgt_len, dtype=np.int32)
    n, t, v = lp.shape
    out = np.empty(n, dtype=np.float32)
    lib().oracle_ctc_alpha_batch(
        _p(lp, ctypes.c_float), t * v, v, _p(targets, ctypes.c_int32),
        targets.shape[1] if targets.size else 0, _p(in_len, ctypes.c_int32),
        _p(tgt_len, ctypes.c_int32), n, v, blank, _p(out, ctypes.c_float))
    return out


def ctc_viterbi(lp, targets, in_len, tgt_len, blank=0):
    """Returns (paths int32 [N, T], scores float32 [N, T], status int32 [N]).

    status 1 marks windows torchaudio would reject (T < L + repeats)."""
    lp = np.ascontiguousarray(lp, dtype=np.float32)
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    in_len = np.ascontiguousarray(in_len, dtype=np.int32)
    tgt_len = np.ascontiguousarray(tgt_len, dtype=np.int32)
    n, t, v = lp.shape
    paths = np.full((n, t), -1, dtype=np.int32)
    scores = np.zeros((n, t), dtype=np.float32)
    status = np.zeros(n, dtype=np.int32)
    lib().oracle_ctc_viterbi_batch(
        _p(lp, ctypes.c_float), t * v, v, _p(targets, ctypes.c_int32),
        targets.shape[1] if targets.size else 0, _p(in_len, ctypes.c_int32),
        _p(tgt_len, ctypes.c_int32), n, t, v, blank, _p(paths, ctypes.c_int32),
        _p(scores, ctypes.c_float), _p(status, ctypes.c_int32))
    return paths, scores, status


def merge_tokens(path, scores, blank=0):
    """Token spans of one alignment path, after torchaudio.functional.merge_tokens
    (_alignment.py:94-127): consecutive equal non-blank frames form one span;
    returns list of (token, start, end, mean score)."""
    path = np.asarray(path)
    scores = np.asarray(scores, dtype=np.float64)
    t = len(path)
    if t == 0:
        return []
    change = np.nonzero(np.diff(path, prepend=-1, append=-1))[0]
    spans = []
    for a, b in zip(change[:-1], change[1:]):
        tok = int(path[a])
        if tok != blank:
            spans.append((tok, int(a), int(b), float(scores[a:b].mean())))
    return spans


def num_threads():
    return int(lib().oracle_num_threads())
