"""GPU parity of kernel (2b) (CTC-segmentation fill + backtrace + utterance scoring),
kernel (3) (anchor selection) and the CTCSegmentation mirror against the oracle
(oracle/ctcseg*.{c,py}, oracle/anchor.py), through the C ABI.

Bar: timings (frame indices), state lists and per-frame path probabilities
bit-exact; segment start/end bit-exact (fp64); scores within 1e-12 relative
(np.mean order is mirrored; they are printed with 4 decimals at the boundary)."""
import numpy as np
import pytest

from cases import seg_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ipfa():
    import ipfa_b200
    return ipfa_b200


@pytest.fixture(scope="module")
def cs():
    import importlib
    return importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.ctc_segmentation")


def _oracle_window(cfg, lpz, utts):
    from oracle import ctcseg as oseg
    gt, ub = oseg.prepare_token_list(cfg, utts)
    timings, char_probs, state_list = oseg.ctc_segmentation(cfg, lpz, gt)
    segs = oseg.determine_utterance_segments(cfg, ub, char_probs, timings, [""] * len(utts))
    return gt, ub, timings, char_probs, state_list, segs


def _pack(cfg, utts_per_window):
    from oracle import ctcseg as oseg
    packed = [oseg.prepare_token_list(cfg, u) for u in utts_per_window]
    n = len(packed)
    cmax = max(len(g) for g, _ in packed)
    kmax = max(len(ub) - 1 for _, ub in packed)
    gt = np.full((n, cmax), -1, np.int32)
    ubs = np.zeros((n, kmax + 1), np.int32)
    n_cols = np.zeros(n, np.int32)
    n_utts = np.zeros(n, np.int32)
    for i, (g, ub) in enumerate(packed):
        gt[i, :len(g)] = g[:, 0]
        ubs[i, :len(ub)] = ub
        ubs[i, len(ub):] = ub[-1]
        n_cols[i], n_utts[i] = len(g), len(ub) - 1
    return gt, ubs, n_cols, n_utts


def _compare_window(cfg, res, i, k, lpz, utts, to_np):
    """Prefix k of window i against a fresh oracle run on utts[:k]."""
    from oracle import ctcseg as oseg
    t_len = lpz.shape[0]
    gt, ub, timings, char_probs, state_list, segs = _oracle_window(cfg, lpz, utts[:k])
    timing = to_np(res.timing)[i, k - 1, :len(gt)]
    got_timings = np.where(timing < 0, 0.0, timing.astype(np.float64) * cfg.index_duration)
    assert np.array_equal(got_timings, timings), (i, k)
    assert np.array_equal(to_np(res.char_prob)[i, k - 1, :t_len].astype(np.float64), char_probs), (i, k)
    state = to_np(res.state)[i, k - 1, :t_len]
    exp_state = np.array([-2 if s == "" else (-1 if s == cfg.self_transition else 0) for s in state_list])
    assert np.array_equal(np.minimum(state, 0), exp_state), (i, k)
    on = state >= 0
    assert [int(gt[c, 0]) for c in state[on]] == [s for s in state_list if s not in ("", cfg.self_transition)]
    seg = to_np(res.seg)[i, k - 1, :k]
    for u in range(k):
        assert seg[u, 0] == segs[u][0] and seg[u, 1] == segs[u][1], (i, k, u, seg[u], segs[u])
        np.testing.assert_allclose(seg[u, 2], segs[u][2], rtol=1e-12, atol=0)


SEG_SHAPES = [
    # (n, t, v, k_utts, tok_lo, tok_hi, score_len)
    (6, 60, 8, 2, 2, 4, 30),       # KC=1, short-utterance mean branch
    (5, 200, 32, 4, 3, 9, 5),      # KC=2, sliding min-of-mean branch
    (4, 400, 32, 6, 8, 16, 30),    # KC=4
    (3, 900, 32, 6, 20, 40, 30),   # 2 warps
    (2, 1500, 40, 8, 40, 60, 30),  # 4 warps
    (2, 2500, 32, 8, 80, 120, 30), # 8 warps
    (2, 600, 3000, 4, 10, 40, 30), # gather panel
]


@pytest.mark.parametrize("shape", SEG_SHAPES)
def test_all_prefixes_vs_oracle(ipfa, shape):
    """One fill + K backtraces == K independent oracle alignments of the shrinking text."""
    import torch
    from oracle import ctcseg as oseg
    n, t, v, k_utts, lo, hi, score_len = shape
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02, score_min_mean_over_L=score_len)
    lp, in_len, utts = seg_case(41, n, t, v, k_utts, lo, hi)
    gt, ubs, n_cols, n_utts = _pack(cfg, utts)
    flags = cfg.flags | 8
    res = ipfa.ctcseg_align(torch.from_numpy(lp).cuda(), in_len, gt, n_cols, ubs, n_utts, cfg.index_duration,
                            score_len=score_len, flags=flags)
    assert int(res.status.abs().sum()) == 0
    to_np = lambda x: x.cpu().numpy()
    for i in range(n):
        for k in range(1, int(n_utts[i]) + 1):
            _compare_window(cfg, res, i, k, lp[i, :in_len[i]], utts[i], to_np)
        term = to_np(res.term_t)[i]
        table, _, t_end, _, argmax = oseg.fill_table(cfg, lp[i, :in_len[i]], oseg.prepare_token_list(cfg, utts[i])[0], 8000)
        for k in range(1, int(n_utts[i]) + 1):
            assert term[k - 1] == argmax[ubs[i, k]]
    # host entry point: identical bytes
    res_h = ipfa.ctcseg_align_host(lp, in_len, gt, n_cols, ubs, n_utts, cfg.index_duration,
                                   score_len=score_len, flags=flags)
    for i in range(n):
        for k in range(1, int(n_utts[i]) + 1):
            assert np.array_equal(res_h.seg[i, k - 1, :k], to_np(res.seg)[i, k - 1, :k])
            assert np.array_equal(res_h.timing[i, k - 1], to_np(res.timing)[i, k - 1])
            assert np.array_equal(res_h.char_prob[i, k - 1], to_np(res.char_prob)[i, k - 1])


def test_flags_and_unpeaked(ipfa):
    """gratis_blank / preamble cost flags and flat (non-peaked) emissions."""
    import torch
    from oracle import ctcseg as oseg
    for bz, pz in [(True, True), (False, False), (True, False)]:
        cfg = oseg.CtcSegmentationParameters(index_duration=0.02, blank_transition_cost_zero=bz,
                                             preamble_transition_cost_zero=pz)
        lp, in_len, utts = seg_case(43, 4, 150, 16, 3, 3, 8, peaked=False)
        gt, ubs, n_cols, n_utts = _pack(cfg, utts)
        res = ipfa.ctcseg_align(torch.from_numpy(lp).cuda(), in_len, gt, n_cols, ubs, n_utts,
                                cfg.index_duration, flags=cfg.flags)
        for i in range(4):
            _compare_window(cfg, res, i, int(n_utts[i]), lp[i, :in_len[i]], utts[i], lambda x: x.cpu().numpy())


def test_text_longer_than_audio(ipfa, cs):
    import torch
    from oracle import ctcseg as oseg
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02)
    lp, in_len, utts = seg_case(44, 3, 30, 8, 2, 4, 6, peaked=False)
    in_len[1] = 5
    gt, ubs, n_cols, n_utts = _pack(cfg, utts)
    res = ipfa.ctcseg_align(torch.from_numpy(lp).cuda(), in_len, gt, n_cols, ubs, n_utts, cfg.index_duration,
                            flags=cfg.flags | 8)
    status = res.status.cpu().numpy()
    assert status[1] == 4 and status[0] == 0 and status[2] == 0
    with pytest.raises(AssertionError, match="Audio is shorter than text"):
        oseg.ctc_segmentation(cfg, lp[1, :5], oseg.prepare_token_list(cfg, utts[1])[0])


def test_ctcsegmentation_mirror_matches_oracle(cs):
    """The four-call sequence of the reference's entry points
    (/root/reference/src/iterative_utterance_alignment.py:201-219) on the stub emitter."""
    import importlib
    import torch
    from oracle import ctcseg as oseg
    stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    asr = stub.StubEncoderASR(device="cuda")
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    aligner.samples_to_frames_ratio = aligner.estimate_samples_to_frames_ratio()
    assert aligner.samples_to_frames_ratio == pytest.approx(320.0, rel=1e-2)
    g = torch.Generator().manual_seed(3)
    audio = torch.randn(16000 * 8, generator=g) * 0.1
    transcript = ["hola que tal", "esto es una prueba de alineamiento forzado", "·", "adios"]
    lpz = aligner.get_lpz(audio)
    assert lpz.is_cuda
    task = aligner.prepare_segmentation_task(transcript, lpz, "sample_0", audio.shape[0])
    segments = aligner.get_segments(task)
    task.set(**segments)
    lines = str(task).strip().split("\n")
    fields = [ln.split(" ", 5) for ln in lines]
    assert len(fields) == 4 and all(len(f) == 6 for f in fields)
    # oracle on the same emissions
    cfg = oseg.CtcSegmentationParameters(index_duration=task.config.index_duration, score_min_mean_over_L=30,
                                         char_list=task.config.char_list)
    lp_host = lpz.cpu().numpy()
    ref = oseg.get_segments(cfg, lp_host, task.ground_truth_mat, task.utt_begin_indices, task.text)
    assert np.array_equal(segments["timings"], ref["timings"])
    assert np.array_equal(segments["char_probs"], ref["char_probs"])
    assert segments["state_list"] == ref["state_list"]
    assert str(task) == oseg.task_str("sample_0", task.text, ref["segments"])
    # text longer than audio -> AssertionError, like the reference
    short = aligner.get_lpz(audio[:3200])
    with pytest.raises(AssertionError, match="Audio is shorter than text"):
        aligner.get_segments(aligner.prepare_segmentation_task(transcript, short, "s", 3200))


def test_anchor_select_matches_oracle_state_machine(ipfa):
    """Kernel (3) against oracle/anchor.py on random per-prefix scores."""
    import torch
    from oracle.anchor import anchor_window
    rng = np.random.default_rng(7)
    n, kmax = 400, 6
    seg = np.full((n, kmax, kmax, 3), np.nan)
    n_utts = rng.integers(1, kmax + 1, n).astype(np.int32)
    text_len = rng.choice([10, 45], size=(n, kmax), p=[0.25, 0.75]).astype(np.int32)
    is_last = (rng.random(n) < 0.1).astype(np.int32)
    for w in range(n):
        for k in range(int(n_utts[w]), 0, -1):
            ends = np.sort(rng.uniform(0.5, 60.0, k))
            starts = np.concatenate([[0.1], ends[:-1]])
            scores = -np.abs(rng.normal(0.0, 1.6, k)) - 0.05
            if rng.random() < 0.2:
                scores[-1] = seg[w, k, k - 1, 2] if k < n_utts[w] else scores[-1]  # provoke equal scores
            seg[w, k - 1, :k, 0], seg[w, k - 1, :k, 1], seg[w, k - 1, :k, 2] = starts, ends, scores
    dec, anchor = ipfa.anchor_select(torch.from_numpy(seg).cuda(), n_utts, text_len, is_last)
    dec, anchor = dec.cpu().numpy(), anchor.cpu().numpy()
    for w in range(n):
        K = int(n_utts[w])
        texts = ["x" * int(text_len[w, u]) for u in range(K)]

        def align_fn(transcript, w=w):
            k = len(transcript)
            return [[f"u_{u:04}", "u", f"{seg[w, k - 1, u, 0]:.2f}", f"{seg[w, k - 1, u, 1]:.2f}",
                     f"{seg[w, k - 1, u, 2]:3.4f}", transcript[u]] for u in range(k)]

        rows, nss, disc, n_iter = anchor_window(texts, align_fn, 100.0, bool(is_last[w]), None, [])
        assert dec[w, 0] == len(rows), (w, dec[w], len(rows))
        assert dec[w, 1] == n_iter
        assert len(disc) == K - dec[w, 0]
        if dec[w, 3] == -1:
            assert nss == 100.0
        elif dec[w, 3] == -2:
            assert nss is None
        else:
            assert nss == 100.0 + float(f"{anchor[w]:.2f}") and abs(anchor[w] - (nss - 100.0)) < 1e-9
            # rows kept are those of the accepted prefix
            k = dec[w, 0]
            assert [r[5] for r in rows] == [100.0 + float(f"{seg[w, k - 1, u, 1]:.2f}") for u in range(k)]


# --------------------------------------------------------------------------- windowed table mode
def _aligned_case(seed, n, t, v, k_utts, lo, hi):
    """Peaked windows whose text spans the whole audio (so the table window has to slide)."""
    return seg_case(seed, n, t, v, k_utts, lo, hi, peaked=True, ragged=False)


def test_windowed_mode_equals_full_table_when_the_audio_fits(ipfa):
    import torch
    from oracle import ctcseg as oseg
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02)
    lp, in_len, utts = seg_case(51, 4, 300, 32, 4, 3, 9)
    gt, ubs, n_cols, n_utts = _pack(cfg, utts)
    dev = torch.from_numpy(lp).cuda()
    full = ipfa.ctcseg_align(dev, in_len, gt, n_cols, ubs, n_utts, cfg.index_duration, flags=cfg.flags | 8)
    win = ipfa.ctcseg_align(dev, in_len, gt, n_cols, ubs, n_utts, cfg.index_duration, flags=cfg.flags | 8,
                            window=512)
    assert int(win.status.sum()) == 0
    for i in range(4):
        for k in range(1, int(n_utts[i]) + 1):
            for name in ("timing", "char_prob", "state"):
                assert torch.equal(getattr(full, name)[i, k - 1], getattr(win, name)[i, k - 1]), (name, i, k)
            assert torch.equal(full.seg[i, k - 1, :k], win.seg[i, k - 1, :k]), (i, k)
            assert int(full.term_t[i, k - 1]) == int(win.term_t[i, k - 1])


@pytest.mark.parametrize("step_rule", ["int+1", "ceil"])
@pytest.mark.parametrize("t,window,k_utts,lo,hi", [(700, 256, 3, 10, 20), (1000, 300, 5, 8, 16),
                                                   (2300, 2100, 4, 20, 40)])
def test_windowed_mode_vs_oracle(ipfa, t, window, k_utts, lo, hi, step_rule):
    """T > window: per-column sliding offsets, one fill per prefix; the oracle doubles its window on
    IndexError exactly where the CUDA path reports WIN_WINDOW_TOO_SMALL."""
    import torch
    from oracle import ctcseg as oseg
    n = 3
    lp, in_len, utts = _aligned_case(61, n, t, 32, k_utts, lo, hi)
    gt, ubs, n_cols, n_utts = _pack(oseg.CtcSegmentationParameters(), utts)
    dev = torch.from_numpy(lp).cuda()
    to_np = lambda x: x.cpu().numpy()
    checked = 0
    for i in range(n):
        for k in range(1, int(n_utts[i]) + 1):
            # one (window, prefix) at a time so each one can double its own window like the reference
            w = window
            while True:
                res = ipfa.ctcseg_align(dev[i:i + 1], in_len[i:i + 1], gt[i:i + 1, :ubs[i, k] + 1],
                                        np.array([ubs[i, k] + 1], np.int32), ubs[i:i + 1, :k + 1],
                                        np.array([k], np.int32), 0.02, window=w,
                                        flags=2 | (16 if step_rule == "ceil" else 0))
                if not int(res.status[0]) & 8:
                    break
                w *= 2
                assert w < 100000
            cfg = oseg.CtcSegmentationParameters(index_duration=0.02, min_window_size=window,
                                                 window_step_rule=step_rule)
            _compare_window(cfg, res, 0, k, lp[i, :in_len[i]], utts[i], to_np)
            checked += 1
            # the offsets really moved
            table, offsets, _, _, _ = oseg.fill_table(cfg, lp[i, :in_len[i]],
                                                      oseg.prepare_token_list(cfg, utts[i][:k])[0], w)
            if w < in_len[i] and k == int(n_utts[i]):
                assert offsets[-1] > 0  # the full text spans the audio: the window had to slide
    assert checked >= n


def _integral_step_case(seed, window, step):
    """A window whose mean offset (T - W) / N is the integer ``step``: the one situation in which
    the two readings of the largest window step (int(mean) + 1, ceil(mean)) differ."""
    rng = np.random.default_rng(seed)
    utts = [rng.integers(1, 32, int(rng.integers(12, 24))).astype(np.int64) for _ in range(4)]
    flat = []
    for u in utts:
        flat += [0] + u.tolist()
    flat += [0]
    n_cols = len(flat) + 1
    t = window + step * n_cols
    lp = rng.standard_normal((t, 32)).astype(np.float32)
    pos = np.sort(rng.permutation(t)[:len(flat)])
    for j, (a, b) in enumerate(zip(pos, list(pos[1:]) + [t])):
        lp[a:b, flat[j]] += 4.0
    lp = lp - np.log(np.exp(lp.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    return lp, utts, n_cols


@pytest.mark.parametrize("step_rule", ["int+1", "ceil"])
def test_windowed_mode_integral_mean_offset(ipfa, step_rule):
    import torch
    from oracle import ctcseg as oseg
    window = 256
    lp, utts, n_cols = _integral_step_case(63, window, 3)
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02, min_window_size=window, window_step_rule=step_rule)
    gt, ub = oseg.prepare_token_list(cfg, utts)
    assert len(gt) == n_cols and (lp.shape[0] - window) % n_cols == 0
    k = len(utts)
    w = window
    while True:
        res = ipfa.ctcseg_align(torch.from_numpy(lp).cuda()[None], [lp.shape[0]], gt[:, 0].astype(np.int32)[None],
                                [n_cols], np.asarray(ub, np.int32)[None], [k], 0.02, window=w,
                                flags=cfg.flags)
        if not int(res.status[0]) & 8:
            break
        w *= 2
    _compare_window(cfg, res, 0, k, lp, utts, lambda x: x.cpu().numpy())
    # and the two rules are not the same algorithm on this input
    other = oseg.CtcSegmentationParameters(min_window_size=window,
                                           window_step_rule="ceil" if step_rule == "int+1" else "int+1")
    off_a = oseg.fill_table(cfg, lp, gt, window)[1]
    off_b = oseg.fill_table(other, lp, gt, window)[1]
    assert not np.array_equal(off_a, off_b)


def test_long_audio_through_the_ctcsegmentation_mirror(cs):
    """> 8000 frames through CTCSegmentation.get_segments (the reference's default min_window_size)."""
    import torch
    from oracle import ctcseg as oseg
    import importlib
    stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    rng = np.random.default_rng(5)
    tok = stub.CharTokenizer()
    words = "uno dos tres cuatro cinco seis siete ocho nueve diez".split()
    utts = [" ".join(rng.choice(words, size=6)) for _ in range(8)]
    t_len = 9000
    flat = []
    for u in utts:
        flat += [0] + tok.encode_as_ids(u)
    flat += [0]
    lp = rng.standard_normal((t_len, tok.vocab_size())).astype(np.float32)
    pos = np.sort(rng.permutation(t_len)[:len(flat)])
    for j, (a, b) in enumerate(zip(pos, list(pos[1:]) + [t_len])):
        lp[a:b, flat[j]] += 5.0
    lp = lp - np.log(np.exp(lp.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    asr = stub.StubEncoderASR(device="cuda")
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    aligner.samples_to_frames_ratio = 320.0
    task = aligner.prepare_segmentation_task(utts, torch.from_numpy(lp).cuda(), "long", t_len * 320)
    got = cs.CTCSegmentation.get_segments(task)
    cfg = oseg.CtcSegmentationParameters(index_duration=task.config.index_duration,
                                         score_min_mean_over_L=30, char_list=task.config.char_list)
    ref = oseg.get_segments(cfg, lp, task.ground_truth_mat, task.utt_begin_indices, task.text)
    assert np.array_equal(got["timings"], ref["timings"])
    assert np.array_equal(got["char_probs"], ref["char_probs"])
    assert got["state_list"] == ref["state_list"]
    for a, b in zip(got["segments"], ref["segments"]):
        assert a[0] == b[0] and a[1] == b[1]
        np.testing.assert_allclose(a[2], b[2], rtol=1e-12)


# --------------------------------------------------------------------------- classic text converter
BPE_LIST = ["<blank>", "<unk>", "a", "b", "c", "d", "ab", "bc", "cd", "abc", "da", "·x"]


def _classic_case(seed, t_len, n_utts, v=None):
    """Character strings over {a,b,c,d} + a vocabulary with multi-character tokens: the ground
    truth matrix has up to 3 candidate tokens per position."""
    from oracle import ctcseg as oseg
    rng = np.random.default_rng(seed)
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02, char_list=list(BPE_LIST))
    utts = ["".join(rng.choice(list("abcd"), size=int(rng.integers(4, 12)))) for _ in range(n_utts)]
    gt, ub = oseg.prepare_text(cfg, utts)
    v = len(BPE_LIST)
    lp = rng.standard_normal((t_len, v)).astype(np.float32)
    # make some multi-character tokens attractive so that s > 0 transitions are really taken
    pos = np.sort(rng.permutation(t_len)[:len(gt)])
    for c, (a, b) in enumerate(zip(pos, list(pos[1:]) + [t_len])):
        cand = [int(g) for g in gt[c] if g >= 0]
        if cand:
            lp[a:b, cand[-1]] += 3.0
    lp = lp - np.log(np.exp(lp.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    return cfg, utts, gt, ub, lp


@pytest.mark.parametrize("cascade", ["ascending", "shift"])
@pytest.mark.parametrize("t_len,window", [(300, None), (500, 160)])
def test_multi_column_ground_truth_vs_oracle(ipfa, t_len, window, cascade):
    """ctc-segmentation's `classic` converter: up to G switch transitions per cell,
    full table (window = T) and sliding window."""
    import torch
    from oracle import ctcseg as oseg
    cfg, utts, gt, ub, lp = _classic_case(71, t_len, 4)
    cfg.offset_cascade = cascade
    assert gt.shape[1] >= 3 and (gt[:, 1:3] >= 0).any() and (gt[:, 3:] < 0).all()
    n_cols, k = len(gt), len(ub) - 1
    w = window
    while True:
        res = ipfa.ctcseg_align(torch.from_numpy(lp).cuda()[None], [t_len], gt.astype(np.int32)[None],
                                [n_cols], np.asarray(ub, np.int32)[None], [k], 0.02, window=w,
                                flags=2 | (32 if cascade == "shift" else 0))
        if not int(res.status[0]) & 8:
            break
        w *= 2
    if window is not None:
        cfg.min_window_size = window
    timings, char_probs, state_list = oseg.ctc_segmentation(cfg, lp, gt)
    segs = oseg.determine_utterance_segments(cfg, ub, char_probs, timings, utts)
    timing = res.timing[0, k - 1, :n_cols].cpu().numpy()
    assert np.array_equal(np.where(timing < 0, 0.0, timing * 0.02), timings)
    assert np.array_equal(res.char_prob[0, k - 1, :t_len].cpu().numpy().astype(np.float64), char_probs)
    state = res.state[0, k - 1, :t_len].cpu().numpy()
    got_states = ["" if s == -2 else (cfg.self_transition if s == -1 else
                                      cfg.char_list[int(gt[s & 0xffffff, s >> 24])]) for s in state]
    assert got_states == state_list
    assert any(len(s) > 1 for s in state_list)   # a multi-character token was taken
    seg = res.seg[0, k - 1, :k].cpu().numpy()
    for u in range(k):
        assert seg[u, 0] == segs[u][0] and seg[u, 1] == segs[u][1]
        np.testing.assert_allclose(seg[u, 2], segs[u][2], rtol=1e-12)


def test_classic_text_converter_through_the_mirror(cs):
    import torch
    from oracle import ctcseg as oseg
    import types

    class PieceTokenizer:
        pieces = list(BPE_LIST)

        def vocab_size(self):
            return len(self.pieces)

        def id_to_piece(self, i):
            return self.pieces[i]

        def unk_id(self):
            return 1

        def encode_as_pieces(self, text):
            return list(text)

        def encode_as_ids(self, text):
            return [self.pieces.index(c) for c in text]

    asr = types.SimpleNamespace(tokenizer=PieceTokenizer(), encode_batch=lambda *a: None, device="cuda",
                                hparams=types.SimpleNamespace(sample_rate=16000, log_softmax=lambda x: x))
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", text_converter="classic")
    aligner.samples_to_frames_ratio = 320.0
    cfg, utts, gt, ub, lp = _classic_case(72, 350, 3)
    task = aligner.prepare_segmentation_task(utts, torch.from_numpy(lp).cuda(), "classic", 350 * 320)
    assert np.array_equal(task.ground_truth_mat, gt) and list(task.utt_begin_indices) == list(ub)
    got = cs.CTCSegmentation.get_segments(task)
    ref = oseg.get_segments(cfg, lp, gt, ub, utts)
    assert np.array_equal(got["timings"], ref["timings"]) and got["state_list"] == ref["state_list"]
    assert np.array_equal(got["char_probs"], ref["char_probs"])
    for a, b in zip(got["segments"], ref["segments"]):
        assert a[0] == b[0] and a[1] == b[1]
        np.testing.assert_allclose(a[2], b[2], rtol=1e-12)


def test_host_entry_point_long_audio(ipfa):
    """ipfa_ctcseg_host with more than 8000 frames: the windowed kernels, window doubling inside."""
    from oracle import ctcseg as oseg
    lp, in_len, utts = _aligned_case(81, 1, 8300, 32, 3, 30, 60)
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02)
    gt, ubs, n_cols, n_utts = _pack(cfg, utts)
    res = ipfa.ctcseg_align_host(lp, in_len, gt, n_cols, ubs, n_utts, 0.02, flags=2)
    assert int(res.status[0]) == 0
    _compare_window(cfg, res, 0, int(n_utts[0]), lp[0, :in_len[0]], utts[0], lambda x: x)


@pytest.mark.parametrize("spread", [64, 128])
def test_window_columns_spread_over_a_cluster(ipfa, spread):
    """IPFA_SEG_SPREAD_2 / _4: a window's columns over 2 / 4 SMs (thread-block cluster, boundary values handed
    over a chunk at a time through distributed shared memory) -- same bits as one SM."""
    import torch
    from oracle import ctcseg as oseg
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02)
    for shape in ((3, 900, 32, 6, 20, 40), (2, 1500, 32, 8, 40, 60), (2, 2500, 32, 8, 80, 120), (5, 400, 32, 6, 8, 16)):
        n, t, v, k_utts, lo, hi = shape
        lp, in_len, utts = seg_case(43, n, t, v, k_utts, lo, hi)
        gt, ubs, n_cols, n_utts = _pack(cfg, utts)
        dev = torch.from_numpy(lp).cuda()
        one = ipfa.ctcseg_align(dev, in_len, gt, n_cols, ubs, n_utts, 0.02, flags=cfg.flags | 8)
        many = ipfa.ctcseg_align(dev, in_len, gt, n_cols, ubs, n_utts, 0.02, flags=cfg.flags | 8 | spread)
        for i in range(n):
            for k in range(1, int(n_utts[i]) + 1):
                for name in ("timing", "char_prob", "state"):
                    a, b = getattr(one, name)[i, k - 1], getattr(many, name)[i, k - 1]
                    m = int(in_len[i]) if name != "timing" else int(ubs[i, k]) + 1
                    assert torch.equal(a[:m], b[:m]), (shape, name, i, k)
                assert torch.equal(one.seg[i, k - 1, :k], many.seg[i, k - 1, :k])
                assert int(one.term_t[i, k - 1]) == int(many.term_t[i, k - 1])
        _compare_window(cfg, many, 0, int(n_utts[0]), lp[0, :in_len[0]], utts[0], lambda x: x.cpu().numpy())
