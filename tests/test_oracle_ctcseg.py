"""Known-answer and property tests of the ctcseg oracle (oracle/ctcseg*.{c,py}).

The reference pins nothing at this boundary: these are
hand-checkable cases plus structural properties of the published algorithm."""
import numpy as np
import pytest

from oracle import ctcseg as oseg


def peaked_lpz(frames, v, hi=0.9):
    """frames: token id per frame -> log-probs with `hi` mass on that token."""
    t = len(frames)
    p = np.full((t, v), (1.0 - hi) / (v - 1), dtype=np.float64)
    p[np.arange(t), frames] = hi
    return np.log(p).astype(np.float32)


def naive_table(lpz, gt, blank=0):
    """Independent pure-Python statement of SURVEY 8(a) A4 (default flags)."""
    T, N = lpz.shape[0], len(gt)
    NEG = np.float32(-1e9)
    tab = np.full((T, N), np.float32(-1e10), dtype=np.float32)
    tab[0, 0] = 0
    for c in range(N):
        for t in range(1 if c == 0 else 0, T):
            if c == 0:
                sw, st = NEG, (np.float32(0) if t >= 1 else NEG)
            else:
                g = gt[c]
                sw = NEG if t == 0 else np.float32(tab[t - 1, c - 1] + lpz[t, g])
                st = NEG if t == 0 else np.float32(tab[t - 1, c] + max(lpz[t, blank], lpz[t, g]))
            tab[t, c] = max(sw, st)
    return tab


def test_prepare_token_list_prefix_property():
    cfg = oseg.CtcSegmentationParameters()
    utts = [np.array([3, 4]), np.array([5]), np.array([6, 7, 8])]
    gt, ub = oseg.prepare_token_list(cfg, utts)
    assert gt[:, 0].tolist() == [-1, 0, 3, 4, 0, 5, 0, 6, 7, 8, 0]
    assert ub == [1, 4, 6, 10]
    for k in range(1, 4):
        gk, ubk = oseg.prepare_token_list(cfg, utts[:k])
        assert gk[:, 0].tolist() == gt[:ub[k] + 1, 0].tolist()
        assert ubk == ub[:k + 1]


def test_fill_matches_naive():
    rng = np.random.default_rng(0)
    cfg = oseg.CtcSegmentationParameters()
    lpz = np.log(rng.dirichlet(np.ones(6), size=40)).astype(np.float32)
    gt, _ = oseg.prepare_token_list(cfg, [rng.integers(1, 6, 5), rng.integers(1, 6, 4)])
    table, offsets, t, c, argmax = oseg.fill_table(cfg, lpz, gt, cfg.min_window_size)
    ref = naive_table(lpz, gt[:, 0])
    assert np.array_equal(table, ref)
    assert c == len(gt) - 1 and offsets.sum() == 0
    for col in range(1, len(gt)):
        assert argmax[col] == int(np.argmax(ref[:, col]))  # first max
    assert t == argmax[-1]


def test_known_answer_two_utterances():
    # tokens: 0 blank, 1 'a', 2 'b', 3 'c'
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02, score_min_mean_over_L=30,
                                         char_list=["_", "a", "b", "c"])
    frames = [0, 0, 1, 1, 2, 0, 0, 3, 3, 0, 0, 0]
    lpz = peaked_lpz(frames, 4)
    gt, ub = oseg.prepare_token_list(cfg, [np.array([1, 2]), np.array([3])])
    assert gt[:, 0].tolist() == [-1, 0, 1, 2, 0, 3, 0]
    timings, char_probs, state_list = oseg.ctc_segmentation(cfg, lpz, gt)
    hi, lo = np.float32(np.log(0.9)), np.float32(np.log(0.1 / 3))
    # the best path enters each column at the first frame that emits its token
    # col: 1(blank) 2(a) 3(b) 4(blank) 5(c) 6(blank)
    assert timings.tolist() == pytest.approx([0.0, 0.02, 0.04, 0.08, 0.10, 0.14, 0.18])
    assert state_list[2] == "a" and state_list[4] == "b" and state_list[7] == "c"
    assert state_list[3] == "ε" and state_list[0] == ""
    assert char_probs[0] == 0.0                    # frame 0 is never visited
    assert np.all(char_probs[1:10] == hi)
    segs = oseg.determine_utterance_segments(cfg, ub, char_probs, timings, ["ab", "c"])
    # utt 0: begin=max(t[2]-.5,(t[1]+t[0])/2)=0.01, end=min(t[3]+.5,(t[4]+t[3])/2)=0.09
    assert segs[0][0] == pytest.approx(0.01) and segs[0][1] == pytest.approx(0.09)
    # utt 1: begin=max(t[5]-.5,(t[4]+t[3])/2)=0.09, end=min(t[5]+.5,(t[6]+t[5])/2)=0.16
    assert segs[1][0] == pytest.approx(0.09) and segs[1][1] == pytest.approx(0.16)
    # short segments (<= 30 frames): plain mean over [floor(start/dur), floor(end/dur))
    assert segs[0][2] == pytest.approx(char_probs[0:4].mean())
    s = oseg.task_str("utt", ["ab", "c"], segs)
    lines = s.strip().split("\n")
    assert lines[0].split(" ", 5)[:4] == ["utt_0000", "utt", "0.01", "0.09"]
    assert lines[1].endswith(" c")


def test_audio_shorter_than_text_raises():
    cfg = oseg.CtcSegmentationParameters()
    lpz = peaked_lpz([0, 1, 0], 3)
    gt, _ = oseg.prepare_token_list(cfg, [np.array([1, 2, 1, 2])])
    with pytest.raises(AssertionError, match="Audio is shorter than text"):
        oseg.ctc_segmentation(cfg, lpz, gt)


def test_long_segment_min_of_windowed_mean():
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02, score_min_mean_over_L=5)
    rng = np.random.default_rng(3)
    frames = [0] * 3 + [1] * 8 + [0] * 4 + [2] * 8 + [0] * 3 + [1] * 6 + [0] * 5 + [2] * 6 + [0] * 5
    lpz = peaked_lpz(frames, 3, hi=0.8) + rng.normal(0, 0.05, (len(frames), 3)).astype(np.float32)
    gt, ub = oseg.prepare_token_list(cfg, [np.array([1, 2, 1]), np.array([2])])
    timings, char_probs, _ = oseg.ctc_segmentation(cfg, lpz, gt)
    segs = oseg.determine_utterance_segments(cfg, ub, char_probs, timings, ["aba", "b"])
    start, end, score = segs[0]
    s, f = int(np.floor(start / 0.02)), int(np.floor(end / 0.02))
    assert f - s > 5
    ref = min(0.0, min(char_probs[t:t + 5].mean() for t in range(s, f - 5)))
    assert score == pytest.approx(ref)
    for start, end, score in segs:
        s, f = int(np.floor(start / 0.02)), int(np.floor(end / 0.02))
        if f <= s:
            ref = -1e10
        elif f - s <= 5:
            ref = char_probs[s:f].mean()
        else:
            ref = min(0.0, min(char_probs[t:t + 5].mean() for t in range(s, f - 5)))
        assert score == pytest.approx(ref)


def test_windowed_table_equals_full_when_it_fits():
    """T > min_window_size triggers the sliding-window fill; with a window that
    still covers the path the segmentation is unchanged."""
    rng = np.random.default_rng(5)
    cfg_full = oseg.CtcSegmentationParameters(index_duration=0.02, min_window_size=8000)
    cfg_win = oseg.CtcSegmentationParameters(index_duration=0.02, min_window_size=60)
    frames = []
    toks = rng.integers(1, 5, 10)
    for tok in toks:
        frames += [0] * 6 + [int(tok)] * 4
    frames += [0] * 10
    lpz = peaked_lpz(frames, 5)
    gt, ub = oseg.prepare_token_list(cfg_full, [toks[:5], toks[5:]])
    a = oseg.ctc_segmentation(cfg_full, lpz, gt)
    b = oseg.ctc_segmentation(cfg_win, lpz, gt)
    assert np.allclose(a[0], b[0]) and np.allclose(a[1], b[1])


def test_prepare_text_product_mirror_equals_oracle():
    """The `classic` text converter: product mirror and oracle restatement build the same matrix."""
    import importlib
    from oracle import ctcseg as oseg
    cs = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.ctc_segmentation")
    chars = ["<blank>", "<unk>", "a", "b", "c", "ab", "bc", "abc", "ca"]
    utts = ["abcab", "ca bc", "b.a,c", "zzz"]
    for spaces in (False, True):
        a = oseg.CtcSegmentationParameters(char_list=list(chars), replace_spaces_with_blanks=spaces)
        b = cs.CtcSegmentationParameters(char_list=list(chars), replace_spaces_with_blanks=spaces)
        ga, ua = oseg.prepare_text(a, utts)
        gb, ub = cs.prepare_text(b, utts)
        assert np.array_equal(ga, gb) and list(ua) == list(ub)
        assert ga.shape[1] == len("<blank>") and (ga[0] == -1).all()  # max_char_len counts '<blank>' too
        # a multi-character token sits in the column of its length - 1
        assert (ga[:, 2] >= 0).sum() >= 1


def test_windowed_switches_change_only_what_they_name():
    """The two [verify] switches of the windowed mode: the window-step rule matters only when
    (T - W) / N is an integer, the cascade order only with a multi-column ground truth."""
    rng = np.random.default_rng(9)
    toks = [rng.integers(1, 5, 12), rng.integers(1, 5, 9)]
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02, min_window_size=64)
    gt, ub = oseg.prepare_token_list(cfg, toks)
    n = len(gt)
    for t_len, differs in ((64 + 2 * n, True), (64 + 2 * n + 5, False)):
        lpz = rng.standard_normal((t_len, 5)).astype(np.float32)
        offs = {}
        for rule in ("int+1", "ceil"):
            c = oseg.CtcSegmentationParameters(min_window_size=64, window_step_rule=rule)
            offs[rule] = oseg.fill_table(c, lpz, gt, 64)[1]
            assert offs[rule][-1] <= t_len - 64 and np.all(np.diff(offs[rule]) >= 0)
        assert (not np.array_equal(offs["int+1"], offs["ceil"])) == differs
        # single-column ground truth: the cascade order cannot matter
        a = oseg.fill_table(oseg.CtcSegmentationParameters(min_window_size=64, offset_cascade="shift"), lpz, gt, 64)
        b = oseg.fill_table(oseg.CtcSegmentationParameters(min_window_size=64), lpz, gt, 64)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
