"""Torch-facing operators over the C ABI (include/ipfa_b200.h).

PyTorch is plumbing only: it owns device memory and the CUDA stream; every
operator hands raw device pointers to libipfa_b200.so.  There is no CPU
fallback -- CPU tensors are rejected by the ``*_device`` operators; the
``*_host`` variants take host (NumPy / pinned torch) buffers and let the library
do H2D -> kernels -> D2H itself.

Comparators these operators replace:
  ctc_alpha_nll     torch.nn.functional.ctc_loss(..., reduction='none')       row A9
  ctc_forced_align  torchaudio.functional.forced_align (+ merge_tokens)        row A8
  ctcseg_align      ctc_segmentation.ctc_segmentation + determine_utterance_segments
                    behind speechbrain CTCSegmentation.get_segments            rows A4-A6
                    (/root/reference/src/iterative_utterance_alignment.py:216)
  anchor_select     the accept/shrink/revert state machine                     row A10
                    (/root/reference/src/iterative_utterance_alignment.py:221-379)
"""
import contextlib
import ctypes
import os

import numpy as np
import torch

from ._lib import check, lib

SEG_BLANK_COST_ZERO = 1
SEG_PREAMBLE_COST_ZERO = 2
SEG_ROUND_NEAREST = 4
SEG_ALL_PREFIXES = 8
SEG_WINDOW_STEP_CEIL = 16
SEG_OFFSET_SHIFT = 32
SEG_SPREAD_2, SEG_SPREAD_4 = 64, 128
WIN_WINDOW_TOO_SMALL = 8


def launch_count():
    return int(lib().ipfa_launch_count())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _need_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (this path has no CPU fallback)")


def _i32(t, device):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t, dtype=np.int32))
    return t.to(device=device, dtype=torch.int32).contiguous()


def _lp_strides(lp, batch_first):
    """(tensor, N, T, V, stride_n, stride_t, stride_v) in elements.  Any strided view is taken as it is --
    [N, T, V], [T, N, V] (batch_first=False) and vocabulary-major storage ([N, V, T] viewed through
    ``.transpose(1, 2)``: stride_v = T) all work without a copy."""
    if lp.dim() != 3:
        raise ValueError("lp must be [N, T, V] (or [T, N, V] with batch_first=False)")
    if lp.dtype != torch.float32:
        raise ValueError("lp must be float32 log-probabilities")
    if min(lp.stride()) < 1:
        lp = lp.contiguous()
    if batch_first:
        n, t, v = lp.shape
        sn, st = lp.stride(0), lp.stride(1)
    else:
        t, n, v = lp.shape
        st, sn = lp.stride(0), lp.stride(1)
    return lp, n, t, v, sn, st, lp.stride(2)


_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# --------------------------------------------------------------------------- kernel (1)
def ctc_alpha_nll(lp, targets, in_len, tgt_len, blank=0, batch_first=True):
    """Negative log-likelihood of every window, fp32 [N] (inf when infeasible)."""
    _need_cuda(lp, "lp")
    lp, n, t, v, sn, st, sv = _lp_strides(lp, batch_first)
    dev = lp.device
    targets = _i32(targets, dev)
    if targets.dim() == 1:
        targets = targets[None]
    lmax = targets.shape[1] if targets.numel() else 0
    in_len, tgt_len = _i32(in_len, dev), _i32(tgt_len, dev)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    L = lib()
    with torch.cuda.device(dev):
        ws_bytes = L.ipfa_ctc_alpha_workspace_bytes(n, t, lmax, v)
        ws = _workspace(ws_bytes, dev)
        rc = L.ipfa_ctc_alpha_strided_device(_ptr(lp), sn, st, sv, _ptr(targets), targets.stride(0) if lmax else 0,
                                             _ptr(in_len), _ptr(tgt_len), n, t, lmax, v, blank, _ptr(out),
                                             _ptr(ws), ws.numel(), _stream(dev))
    check(rc, "ipfa_ctc_alpha_strided_device")
    return out


def ctc_alpha_redo_count(n, device=None):
    """How many of the `n` windows of the last :func:`ctc_alpha_nll` call on this device / stream the
    linear-domain instance handed to the log-domain instance (its exactness guard fired, or the
    targets were infeasible).  Diagnostic: reads the counter kept behind the `n` arrival counters at
    the head of the workspace (csrc/ctc_alpha.cu); synchronises."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    buf = _ws_cache.get((dev.index, torch.cuda.current_stream(dev).cuda_stream))
    if buf is None:
        return 0
    return int(buf[4 * n:4 * n + 4].view(torch.int32).item())


def ctc_alpha_redo_reasons(n, device=None):
    """OR of the reasons behind :func:`ctc_alpha_redo_count` (1 blank named by a target, 2 emission
    ratio out of fp32 range, 4 a reachable state too small, 8 a state too large, 16 / 32 a scale
    step / neighbouring scales too far apart); totals of zero (infeasible targets) set no bit."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    buf = _ws_cache.get((dev.index, torch.cuda.current_stream(dev).cuda_stream))
    if buf is None:
        return 0
    return int(buf[4 * n + 4:4 * n + 8].view(torch.int32).item())


def _np(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.numpy()
    return np.ascontiguousarray(a, dtype=dtype)


def _hp(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def ctc_alpha_nll_host(lp, targets, in_len, tgt_len, blank=0, out=None):
    """Host buffers in, host buffer out (NumPy arrays or pinned CPU tensors)."""
    lp = _np(lp, np.float32)
    n, t, v = lp.shape
    targets = _np(targets, np.int32)
    if targets.ndim == 1:
        targets = targets[None]
    lmax = targets.shape[1] if targets.size else 0
    in_len, tgt_len = _np(in_len, np.int32), _np(tgt_len, np.int32)
    if out is None:
        out = np.empty(n, dtype=np.float32)
    rc = lib().ipfa_ctc_alpha_host(_hp(lp), t * v, v, _hp(targets), lmax, _hp(in_len), _hp(tgt_len),
                                   n, t, lmax, v, blank, _hp(out))
    check(rc, "ipfa_ctc_alpha_host")
    return out


# --------------------------------------------------------------------------- kernel (2a)
class ForcedAlignment:
    """Result of :func:`ctc_forced_align` (device tensors, or NumPy for the host variant)."""

    def __init__(self, paths, scores, tok_start, tok_end, tok_score, total, status):
        self.paths, self.scores = paths, scores
        self.tok_start, self.tok_end, self.tok_score = tok_start, tok_end, tok_score
        self.total, self.status = total, status


def ctc_forced_align(lp, targets, in_len, tgt_len, blank=0, batch_first=True, tokens=True):
    _need_cuda(lp, "lp")
    lp, n, t, v, sn, st, sv = _lp_strides(lp, batch_first)
    dev = lp.device
    targets = _i32(targets, dev)
    if targets.dim() == 1:
        targets = targets[None]
    lmax = targets.shape[1] if targets.numel() else 0
    in_len, tgt_len = _i32(in_len, dev), _i32(tgt_len, dev)
    paths = torch.empty((n, t), dtype=torch.int32, device=dev)
    scores = torch.empty((n, t), dtype=torch.float32, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    total = torch.empty(n, dtype=torch.float32, device=dev)
    if tokens:
        tok_start = torch.empty((n, max(lmax, 1)), dtype=torch.int32, device=dev)
        tok_end = torch.empty_like(tok_start)
        tok_score = torch.empty((n, max(lmax, 1)), dtype=torch.float32, device=dev)
    else:
        tok_start = tok_end = tok_score = None
    L = lib()
    with torch.cuda.device(dev):
        ws_bytes = L.ipfa_ctc_viterbi_workspace_bytes(n, t, lmax, v)
        ws = _workspace(ws_bytes, dev)
        rc = L.ipfa_ctc_viterbi_strided_device(_ptr(lp), sn, st, sv, _ptr(targets), targets.stride(0) if lmax else 0,
                                               _ptr(in_len), _ptr(tgt_len), n, t, lmax, v, blank,
                                               _ptr(paths), _ptr(scores), _ptr(tok_start), _ptr(tok_end),
                                               _ptr(tok_score), _ptr(total), _ptr(status),
                                               _ptr(ws), ws.numel(), _stream(dev))
    check(rc, "ipfa_ctc_viterbi_strided_device")
    return ForcedAlignment(paths, scores, tok_start, tok_end, tok_score, total, status)


def ctc_forced_align_host(lp, targets, in_len, tgt_len, blank=0, tokens=True, out=None):
    """Host buffers in / out.  ``out`` may carry preallocated (pinned) ``paths``, ``scores``,
    ``total`` and ``status`` arrays."""
    out = out or {}
    lp = _np(lp, np.float32)
    n, t, v = lp.shape
    targets = _np(targets, np.int32)
    if targets.ndim == 1:
        targets = targets[None]
    lmax = targets.shape[1] if targets.size else 0
    in_len, tgt_len = _np(in_len, np.int32), _np(tgt_len, np.int32)
    paths = out.get("paths") if out.get("paths") is not None else np.empty((n, t), np.int32)
    scores = out.get("scores") if out.get("scores") is not None else np.empty((n, t), np.float32)
    status = out.get("status") if out.get("status") is not None else np.empty(n, np.int32)
    total = out.get("total") if out.get("total") is not None else np.empty(n, np.float32)
    if tokens:
        tok_start = np.empty((n, max(lmax, 1)), np.int32)
        tok_end = np.empty_like(tok_start)
        tok_score = np.empty((n, max(lmax, 1)), np.float32)
    else:
        tok_start = tok_end = tok_score = None
    rc = lib().ipfa_ctc_viterbi_host(_hp(lp), t * v, v, _hp(targets), lmax, _hp(in_len), _hp(tgt_len),
                                     n, t, lmax, v, blank, _hp(paths), _hp(scores), _hp(tok_start),
                                     _hp(tok_end), _hp(tok_score), _hp(total), _hp(status))
    check(rc, "ipfa_ctc_viterbi_host")
    return ForcedAlignment(paths, scores, tok_start, tok_end, tok_score, total, status)


# --------------------------------------------------------------------------- kernel (2b)
class SegAlignment:
    """Result of :func:`ctcseg_align`.

    ``seg[w, k-1, u]`` = (start s, end s, score) of utterance ``u`` when the
    first ``k`` utterances of window ``w`` are aligned; ``term_t[w, k-1]`` the
    terminal frame; ``timing[w, k-1, c]`` the frame at which column ``c`` was
    entered; ``char_prob[w, k-1, t]``; ``state[w, k-1, t]``."""

    def __init__(self, seg, term_t, timing, char_prob, state, status):
        self.seg, self.term_t, self.timing = seg, term_t, timing
        self.char_prob, self.state, self.status = char_prob, state, status


def ctcseg_align(lp, in_len, gt, n_cols, utt_begin, n_utts, index_duration, blank=0, score_len=30,
                 flags=SEG_PREAMBLE_COST_ZERO, details=True, batch_first=True, window=None):
    """``window``: table rows of ctc-segmentation's windowed mode (``config.min_window_size``).
    ``None`` = full-table kernels (T <= 8000 frames).  In windowed mode ``status`` carries bit 8
    (``WIN_WINDOW_TOO_SMALL``) where the reference would raise IndexError and double the window."""
    _need_cuda(lp, "lp")
    if lp.dim() == 3 and lp.stride(2) != 1:  # the segmentation kernels read whole rows
        lp = lp.contiguous()
    lp, n, t, v, sn, st, sv = _lp_strides(lp, batch_first)
    dev = lp.device
    gt = _i32(gt, dev)
    utt_begin = _i32(utt_begin, dev)
    if gt.dim() == 1:
        gt = gt[None]
    if utt_begin.dim() == 1:
        utt_begin = utt_begin[None]
    cmax = gt.shape[1]
    # [n, cmax, G]: multi-column ground truth of the `classic` text converter -> the general
    # (column-serial) kernels; the full table is the window = T case
    gt_cols = gt.shape[2] if gt.dim() == 3 else 1
    if gt_cols > 1 and window is None:
        window = t
    gt = gt.contiguous()
    kmax = utt_begin.shape[1] - 1
    in_len, n_cols, n_utts = _i32(in_len, dev), _i32(n_cols, dev), _i32(n_utts, dev)
    seg = torch.full((n, kmax, kmax, 3), float("nan"), dtype=torch.float64, device=dev)
    term_t = torch.empty((n, kmax), dtype=torch.int32, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    if details:
        timing = torch.empty((n, kmax, cmax), dtype=torch.int32, device=dev)
        char_prob = torch.empty((n, kmax, t), dtype=torch.float32, device=dev)
        state = torch.empty((n, kmax, t), dtype=torch.int32, device=dev)
    else:
        timing = char_prob = state = None
    L = lib()
    if window is not None:
        with torch.cuda.device(dev):
            ws_bytes = L.ipfa_ctcseg_windowed_workspace_bytes(n, t, cmax, kmax, int(window), gt_cols, int(flags))
            ws = _workspace(ws_bytes, dev)
            rc = L.ipfa_ctcseg_windowed_device(
                _ptr(lp), None, sn, st, _ptr(in_len), _ptr(gt), gt.stride(0), _ptr(n_cols), _ptr(utt_begin),
                _ptr(n_utts), n, t, cmax, kmax, v, blank, float(index_duration), int(score_len), int(flags),
                int(window), gt_cols, _ptr(seg), _ptr(term_t), _ptr(timing), _ptr(char_prob), _ptr(state),
                _ptr(status), _ptr(ws), ws.numel(), _stream(dev))
        check(rc, "ipfa_ctcseg_windowed_device")
        return SegAlignment(seg, term_t, timing, char_prob, state, status)
    with torch.cuda.device(dev):
        ws_bytes = L.ipfa_ctcseg_workspace_bytes(n, t, cmax, kmax, v)
        ws = _workspace(ws_bytes, dev)
        rc = L.ipfa_ctcseg_device(_ptr(lp), sn, st, _ptr(in_len), _ptr(gt), gt.stride(0), _ptr(n_cols),
                                  _ptr(utt_begin), _ptr(n_utts), n, t, cmax, kmax, v, blank,
                                  float(index_duration), int(score_len), int(flags),
                                  _ptr(seg), _ptr(term_t), _ptr(timing), _ptr(char_prob), _ptr(state),
                                  _ptr(status), _ptr(ws), ws.numel(), _stream(dev))
    check(rc, "ipfa_ctcseg_device")
    return SegAlignment(seg, term_t, timing, char_prob, state, status)


def ctcseg_align_host(lp, in_len, gt, n_cols, utt_begin, n_utts, index_duration, blank=0, score_len=30,
                      flags=SEG_PREAMBLE_COST_ZERO, details=True):
    lp = _np(lp, np.float32)
    n, t, v = lp.shape
    gt = _np(gt, np.int32)
    utt_begin = _np(utt_begin, np.int32)
    if gt.ndim == 1:
        gt = gt[None]
    if utt_begin.ndim == 1:
        utt_begin = utt_begin[None]
    cmax, kmax = gt.shape[1], utt_begin.shape[1] - 1
    in_len, n_cols, n_utts = _np(in_len, np.int32), _np(n_cols, np.int32), _np(n_utts, np.int32)
    seg = np.empty((n, kmax, kmax, 3), np.float64)
    term_t = np.empty((n, kmax), np.int32)
    status = np.empty(n, np.int32)
    if details:
        timing = np.empty((n, kmax, cmax), np.int32)
        char_prob = np.empty((n, kmax, t), np.float32)
        state = np.empty((n, kmax, t), np.int32)
    else:
        timing = char_prob = state = None
    rc = lib().ipfa_ctcseg_host(_hp(lp), t * v, v, _hp(in_len), _hp(gt), cmax, _hp(n_cols), _hp(utt_begin),
                                _hp(n_utts), n, t, cmax, kmax, v, blank, float(index_duration),
                                int(score_len), int(flags), _hp(seg), _hp(term_t), _hp(timing),
                                _hp(char_prob), _hp(state), _hp(status))
    check(rc, "ipfa_ctcseg_host")
    return SegAlignment(seg, term_t, timing, char_prob, state, status)


# --------------------------------------------------------------------------- kernel (3)
def anchor_select(seg, n_utts, text_len, is_last, threshold=-2.0, short_len=30):
    """Returns (decision int32 [N, 4], anchor float64 [N]); see include/ipfa_b200.h."""
    _need_cuda(seg, "seg")
    dev = seg.device
    n, kmax = seg.shape[0], seg.shape[1]
    seg = seg.contiguous()
    n_utts, text_len, is_last = _i32(n_utts, dev), _i32(text_len, dev), _i32(is_last, dev)
    decision = torch.empty((n, 4), dtype=torch.int32, device=dev)
    anchor = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib().ipfa_anchor_select_device(_ptr(seg), _ptr(n_utts), _ptr(text_len), _ptr(is_last), n,
                                             kmax, float(threshold), int(short_len), _ptr(decision),
                                             _ptr(anchor), _stream(dev))
    check(rc, "ipfa_anchor_select_device")
    return decision, anchor


def text_round(x, decimals):
    """``float(f"{v:.{decimals}f}")`` of every element of a CUDA float64 tensor (see ipfa_b200.h)."""
    _need_cuda(x, "x")
    x = x.contiguous().double()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib().ipfa_text_round_device(_ptr(x), x.numel(), int(decimals), _ptr(out), _stream(x.device))
    check(rc, "ipfa_text_round_device")
    return out


@contextlib.contextmanager
def tuning(**switches):
    """``with tuning(IPFA_ALPHA_LOG="1"): ...`` -- set IPFA_* tuning switches for the block (tools and
    tests; value ``None`` unsets).  The library reads the environment once per process, so the table
    is re-read on entry and on exit (``ipfa_tuning_reload``)."""
    old = {k: os.environ.get(k) for k in switches}
    try:
        for k, v in switches.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        lib().ipfa_tuning_reload()
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        lib().ipfa_tuning_reload()


def dominant_kernel_ms(fn, repeats=5):
    """Mean duration (ms) of the longest lattice kernel `fn()` launches, timed alone with CUDA events on
    the launching stream (``ipfa_profile_kernels``); None when `fn` launched none outside a graph."""
    L = lib()
    total, n = 0.0, 0
    try:
        for _ in range(repeats):
            L.ipfa_profile_kernels(1)
            fn()
            ms = float(L.ipfa_profile_read_ms())
            if ms >= 0:
                total += ms
                n += 1
    finally:
        L.ipfa_profile_kernels(0)
    return total / n if n else None
