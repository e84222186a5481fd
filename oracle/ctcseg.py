"""oracle/ctcseg.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the interpreted half of the reference's alignment call
``aligner.get_segments(task)`` (/root/reference/src/iterative_utterance_alignment.py:216,
/root/reference/src/word_level_alignment.py:100, /root/reference/src/search_on_speech.py:85).

The arithmetic lives in two third-party packages pinned by
/root/reference/requirements.txt -- ``ctc-segmentation==1.7.1`` (line 13) and
``speechbrain==0.5.11`` (line 87) -- that are NOT vendored and NOT installed here.
PARITY UNPINNED: this module restates their published algorithm (SURVEY.md
section 8(a) rows A3-A7) and keeps the interpreted cost structure of the original
(Python ``while`` backtrace, one ``ndarray.mean`` per frame when scoring), which
is what ``bench.py``'s ``cpu_baseline`` times.  Spots that could not be re-checked
against package source carry ``[verify]``.
"""
import ctypes
import math

import numpy as np

from . import lib


class CtcSegmentationParameters:
    """Defaults of ctc_segmentation.CtcSegmentationParameters (row A4-A6)."""

    max_prob = -10000000000.0
    skip_prob = -10000000000.0
    min_window_size = 8000
    max_window_size = 100000
    index_duration = 0.025
    score_min_mean_over_L = 30
    space = "·"
    blank = 0
    replace_spaces_with_blanks = False
    blank_transition_cost_zero = False
    preamble_transition_cost_zero = True
    backtrack_from_max_t = False
    self_transition = "ε"
    start_of_ground_truth = "#"
    excluded_characters = ".,»«•❍·"
    tokenized_meta_symbol = "▁"
    char_list = None
    # the two [verify] spots of the windowed table mode (T > min_window_size), one switch each:
    window_step_rule = "int+1"       # or "ceil": largest per-column window step
    offset_cascade = "ascending"     # or "shift": how cur_offset[s] is carried to the next column

    def __init__(self, **kwargs):
        self.set(**kwargs)

    def set(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def index_duration_in_seconds(self):
        return self.index_duration

    @property
    def flags(self):
        return (int(self.blank_transition_cost_zero) + 2 * int(self.preamble_transition_cost_zero)
                + 16 * int(self.window_step_rule == "ceil") + 32 * int(self.offset_cascade == "shift"))


def prepare_token_list(config, text):
    """Row A3: ``[-1] + (blank + tokens)* + blank``; single blank between utterances."""
    ground_truth = [-1]
    utt_begin_indices = []
    for utt in text:
        if ground_truth[-1] != config.blank:
            ground_truth += [config.blank]
        utt_begin_indices.append(len(ground_truth) - 1)
        ground_truth += np.asarray(utt).tolist()
    if ground_truth[-1] != config.blank:
        ground_truth += [config.blank]
    utt_begin_indices.append(len(ground_truth) - 1)
    ground_truth_mat = np.array(ground_truth, dtype=np.int64).reshape(-1, 1)
    return ground_truth_mat, utt_begin_indices


def prepare_text(config, text, char_list=None):
    """The ``classic`` text converter of ctc-segmentation:
    every character position gets up to ``max_char_len`` candidate tokens, the s-th
    column holding the token that spans the last s+1 characters.  [verify]"""
    if char_list is not None:
        config.char_list = char_list
    blank = config.char_list[config.blank]
    ground_truth = config.start_of_ground_truth
    utt_begin_indices = []
    for utt in text:
        if not ground_truth.endswith(config.space):
            ground_truth += config.space
        utt_begin_indices.append(len(ground_truth) - 1)
        for char in utt:
            if char.isspace() and config.replace_spaces_with_blanks:
                if not ground_truth.endswith(config.space):
                    ground_truth += config.space
            elif char in config.char_list and char not in config.excluded_characters:
                ground_truth += char
    if not ground_truth.endswith(config.space):
        ground_truth += config.space
    utt_begin_indices.append(len(ground_truth) - 1)
    max_char_len = max(len(c) for c in config.char_list)
    ground_truth_mat = np.ones([len(ground_truth), max_char_len], np.int64) * -1
    for i in range(len(ground_truth)):
        for s in range(max_char_len):
            if i - s < 0:
                continue
            span = ground_truth[i - s:i + 1]
            span = span.replace(config.space, blank)
            if span in config.char_list:
                ground_truth_mat[i, s] = config.char_list.index(span)
    return ground_truth_mat, utt_begin_indices


def fill_table(config, lpz, ground_truth, window_size):
    """Run the C restatement of ``cython_fill_table`` (row A4)."""
    lpz = np.ascontiguousarray(lpz, dtype=np.float32)
    gt = np.ascontiguousarray(ground_truth, dtype=np.int64)
    T, V = lpz.shape
    N, G = gt.shape
    W = min(window_size, T)
    table = np.zeros([W, N], dtype=np.float32)
    table.fill(config.max_prob)
    offsets = np.zeros([N], dtype=np.int64)
    argmax = np.zeros([N], dtype=np.int32)
    f = ctypes.POINTER(ctypes.c_float)
    t = lib().oracle_ctcseg_fill(
        table.ctypes.data_as(f), W, N, lpz.ctypes.data_as(f), T, V,
        gt.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), G,
        offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), int(config.blank),
        int(config.flags), argmax.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return table, offsets, int(t), N - 1, argmax


def backtrace(config, lpz, ground_truth, table, offsets, t, c):
    """Row A5: the tolerance-based backtrace of ``ctc_segmentation()``.

    Re-derives the transition from fp32 table differences instead of stored
    backpointers; an exact tie falls to *stay*.  Returns
    ``(timings f64 [N], char_probs f64 [T], state_list [T])``."""
    blank = config.blank
    offset = 0
    timings = np.zeros([len(ground_truth)])
    char_probs = np.zeros([lpz.shape[0]])
    state_list = [""] * lpz.shape[0]
    while t != 0 or c != 0:
        min_s = None
        min_switch_prob_delta = np.inf
        max_lpz_prob = config.max_prob
        for s in range(ground_truth.shape[1]):
            if ground_truth[c, s] != -1:
                offset = offsets[c] - (offsets[c - 1 - s] if c - s > 0 else 0)
                switch_prob = lpz[t + offsets[c], ground_truth[c, s]] if c > 0 else config.max_prob
                est_switch_prob = table[t, c] - table[t - 1 + offset, c - 1 - s]
                if abs(switch_prob - est_switch_prob) < min_switch_prob_delta:
                    min_switch_prob_delta = abs(switch_prob - est_switch_prob)
                    min_s = s
                max_lpz_prob = max(max_lpz_prob, switch_prob)
        stay_prob = max(lpz[t + offsets[c], blank], max_lpz_prob) if t > 0 else config.max_prob
        est_stay_prob = table[t, c] - table[t - 1, c]
        if abs(stay_prob - est_stay_prob) > min_switch_prob_delta:
            if c > 0:
                for s in range(0, min_s + 1):
                    timings[c - s] = (offsets[c] + t) * config.index_duration_in_seconds
                char_probs[offsets[c] + t] = max_lpz_prob
                char_index = ground_truth[c, min_s]
                state_list[offsets[c] + t] = (
                    config.char_list[char_index] if config.char_list is not None else int(char_index))
            c -= 1 + min_s
            t -= 1 - offset
        else:
            char_probs[offsets[c] + t] = stay_prob
            state_list[offsets[c] + t] = config.self_transition
            t -= 1
    return timings, char_probs, state_list


def ctc_segmentation(config, lpz, ground_truth, return_table=False):
    """Rows A4+A5.  Raises ``AssertionError`` when the text is longer than the
    audio (caught at /root/reference/src/iterative_utterance_alignment.py:390)."""
    lpz = np.asarray(lpz)
    if len(ground_truth) > lpz.shape[0] and config.skip_prob <= config.max_prob:
        raise AssertionError("Audio is shorter than text!")
    window_size = config.min_window_size
    while True:
        table, offsets, t, c, argmax = fill_table(config, lpz, ground_truth, window_size)
        if config.backtrack_from_max_t:
            t = table.shape[0] - 1
        try:
            timings, char_probs, state_list = backtrace(
                config, lpz.astype(np.float32, copy=False), ground_truth, table, offsets, t, c)
        except IndexError:
            window_size *= 2
            if window_size < config.max_window_size:
                continue
            raise
        break
    if return_table:
        return timings, char_probs, state_list, table, argmax
    return timings, char_probs, state_list


# Index rounding of the scoring window, row A6.  SURVEY.md recalls ``floor``; older
# releases of the package used ``int(round(x))``.  One switch, mirrored by the CUDA
# path's ``IPFA_SEG_ROUND_*`` flag.  [verify]
SEG_INDEX_ROUNDING = "floor"


def _frame_index(x):
    if SEG_INDEX_ROUNDING == "floor":
        return int(math.floor(x))
    return int(round(x))


def determine_utterance_segments(config, utt_begin_indices, char_probs, timings, text):
    """Row A6: utterance start/end and the min-of-windowed-mean confidence."""

    def compute_time(index, align_type):
        middle = (timings[index] + timings[index - 1]) / 2
        if align_type == "begin":
            return max(timings[index + 1] - 0.5, middle)
        return min(timings[index - 1] + 0.5, middle)

    segments = []
    min_prob = np.float64(-10000000000.0)
    for i in range(len(text)):
        start = compute_time(utt_begin_indices[i], "begin")
        end = compute_time(utt_begin_indices[i + 1], "end")
        start_t = _frame_index(start / config.index_duration_in_seconds)
        end_t = _frame_index(end / config.index_duration_in_seconds)
        n = config.score_min_mean_over_L
        if end_t <= start_t:
            min_avg = min_prob
        elif end_t - start_t <= n:
            min_avg = char_probs[start_t:end_t].mean()
        else:
            min_avg = np.float64(0.0)
            for t in range(start_t, end_t - n):
                min_avg = min(min_avg, char_probs[t:t + n].mean())
        segments.append((start, end, min_avg))
    return segments


def task_str(name, text, segments, utt_ids=None):
    """Row A7: ``CTCSegmentationTask.__str__`` -- one line per utterance,
    ``"{name}_{i:04} {name} {start:.2f} {end:.2f} {score:3.4f} {text}"``."""
    out = ""
    n = len(segments)
    names = [f"{name}_{i:04}" for i in range(n)] if utt_ids is None else utt_ids
    for i, b in enumerate(segments):
        out += f"{names[i]} {name} {b[0]:.2f} {b[1]:.2f} {b[2]:3.4f} {text[i]}\n"
    return out


def get_segments(config, lpz, ground_truth_mat, utt_begin_indices, text):
    """Rows A5+A6 as called by ``CTCSegmentation.get_segments`` (row A7)."""
    timings, char_probs, state_list = ctc_segmentation(config, lpz, ground_truth_mat)
    segments = determine_utterance_segments(config, utt_begin_indices, char_probs, timings, text)
    return {"timings": timings, "char_probs": char_probs, "state_list": state_list,
            "segments": segments}
