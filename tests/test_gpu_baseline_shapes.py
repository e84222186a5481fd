"""Parity at the EXACT shapes BASELINE.json names -- the kernel instances the bench lines run:

* configs[3]  256 windows x T=3000 x L=400, V=5000 (gather panel, widest instance): window scorer and
              Viterbi + backtrace, a sample of the windows against the oracle, all of them through
              size-independent properties;
* configs[2]  ragged T<=500, L<=40, V=32 with the device-side length buckets (8192 of the 65 536
              utterances): every path, frame score and token span against the oracle;
* configs[4]  the anchor-iteration unit (256 windows x T=3500 x 908 columns x 6 prefixes) and a
              three-file sweep built like the 100 h corpus of the bench.
Inputs are drawn on the GPU by the same generators bench.py uses."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def bench():
    import bench as b
    return b


@pytest.fixture(scope="module")
def ipfa():
    import ipfa_b200
    return ipfa_b200


def test_config4_window_scorer_and_viterbi(ipfa, bench):
    import torch
    from oracle import ctc as octc
    wl = bench.WORKLOADS["c4"]
    assert (wl.n, wl.t, wl.l, wl.v) == (256, 3000, 400, 5000)
    lp, tg, il, tl = wl.make(5, device="cuda")
    nll = ipfa.ctc_alpha_nll(lp, tg, il, tl)
    res = ipfa.ctc_forced_align(lp, tg, il, tl)
    sample = [0, 1, 97, 128, 254, 255]
    lp_s = lp[sample].cpu().numpy()
    tg_s, il_s, tl_s = (x[sample].cpu().numpy() for x in (tg, il, tl))
    ref = octc.ctc_alpha_nll(lp_s, tg_s, il_s, tl_s)
    np.testing.assert_allclose(nll[sample].cpu().numpy(), ref, rtol=1e-4)
    paths, scores, status = octc.ctc_viterbi(lp_s, tg_s, il_s, tl_s)
    assert not status.any() and int(res.status.sum()) == 0
    assert np.array_equal(res.paths[sample].cpu().numpy(), paths)
    assert np.array_equal(res.scores[sample].cpu().numpy(), scores)
    # all 256 windows: the best path cannot beat the sum over paths; its frame scores add up to its
    # score; collapsing it gives the target
    total = res.total.double()
    assert bool((total <= -nll.double() + 1e-3).all())
    assert torch.allclose(res.scores.double().sum(1), total, rtol=1e-5)
    p = res.paths.cpu().numpy()
    tgn = tg.cpu().numpy()
    for w in range(0, wl.n, 17):
        keep = np.concatenate([[True], p[w, 1:] != p[w, :-1]])
        collapsed = p[w][keep]
        assert np.array_equal(collapsed[collapsed != 0], tgn[w])


def test_config3_ragged_utterances_with_length_buckets(ipfa, bench):
    import torch
    from oracle import ctc as octc
    wl = bench.WORKLOADS["c3"]
    assert (wl.t, wl.l, wl.v, wl.ragged) == (500, 40, 32, True)
    n = 8192  # >= 4096: the two length buckets are decided on the device, as for the 65 536 of the bench
    lp, tg, il, tl = wl.make(6, device="cuda", n=n)
    res = ipfa.ctc_forced_align(lp, tg, il, tl, tokens=True)
    lp_h, tg_h, il_h, tl_h = (x.cpu().numpy() for x in (lp, tg, il, tl))
    paths, scores, status = octc.ctc_viterbi(lp_h, tg_h, il_h, tl_h)
    assert np.array_equal(res.status.cpu().numpy() & 1, status)
    gp, gs = res.paths.cpu().numpy(), res.scores.cpu().numpy()
    ts, te, tp = res.tok_start.cpu().numpy(), res.tok_end.cpu().numpy(), res.tok_score.cpu().numpy()
    assert ((tl_h + 1 <= 32).sum() > 1000) and ((tl_h + 1 > 32).sum() > 1000)  # both buckets are populated
    for i in range(n):
        if status[i]:
            continue
        t_i, l_i = int(il_h[i]), int(tl_h[i])
        assert np.array_equal(gp[i, :t_i], paths[i, :t_i]), i
        assert np.array_equal(gs[i, :t_i], scores[i, :t_i]), i
        if i % 64 == 0:  # token spans: torchaudio.merge_tokens over the path
            pth = paths[i, :t_i]
            change = np.flatnonzero(np.concatenate([[True], pth[1:] != pth[:-1], [True]]))
            spans = [(a, b) for a, b in zip(change[:-1], change[1:]) if pth[a] != 0]
            assert len(spans) == l_i
            assert np.array_equal(ts[i, :l_i], [a for a, _ in spans]) and np.array_equal(te[i, :l_i], [b for _, b in spans])
            np.testing.assert_allclose(tp[i, :l_i], [scores[i, a:b].mean() for a, b in spans], rtol=1e-4, atol=1e-6)


def test_anchor_iteration_unit_shape(ipfa, bench):
    """The `seg` bench shape: 256 windows in flight, T=3500, 908 columns, 6 prefixes."""
    from oracle import ctcseg as oseg
    wl = bench.WORKLOADS["seg"]
    assert (wl.n, wl.t, wl.k, wl.cols) == (256, 3500, 6, 908)
    lp, il, gt, nc, ub, nu, tlen, last = wl.make(7, device="cuda")
    res = ipfa.ctcseg_align(lp, il, gt, nc, ub, nu, 0.02, flags=2 | 8, details=True)
    assert int(res.status.sum()) == 0
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02)
    for w in (0, 255):
        lpz = lp[w].cpu().numpy()
        g = gt[w].cpu().numpy().astype(np.int64)
        u = ub[w].cpu().numpy()
        for k in range(1, wl.k + 1):
            n_cols = int(u[k]) + 1
            timings, char_probs, _ = oseg.ctc_segmentation(cfg, lpz, g[:n_cols].reshape(-1, 1))
            segs = oseg.determine_utterance_segments(cfg, u[:k + 1].tolist(), char_probs, timings, [""] * k)
            timing = res.timing[w, k - 1, :n_cols].cpu().numpy()
            assert np.array_equal(np.where(timing < 0, 0.0, timing * 0.02), timings), (w, k)
            assert np.array_equal(res.char_prob[w, k - 1].cpu().numpy().astype(np.float64), char_probs), (w, k)
            seg = res.seg[w, k - 1, :k].cpu().numpy()
            want = np.asarray(segs, np.float64)
            assert np.array_equal(seg[:, :2], want[:, :2]), (w, k)
            np.testing.assert_allclose(seg[:, 2], want[:, 2], rtol=1e-12, atol=0)


def _sweep_cpu(args):
    from oracle import sweep as osweep
    stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    spec, lp = args
    return osweep.sweep_file(spec.file_id, spec.audio_path, lp, spec.n_samples, spec.rows, stub.CharTokenizer())[:2]


def test_three_file_sweep_like_the_100h_corpus(ipfa):
    """configs[4]: files built like bench.py's corpus (6 % of the utterances not what was said, a
    silence every 9 rows), ten minutes in three files, groups + CUDA graphs as in the bench."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sweep_corpus
    from ipfa_b200 import sweep as sw
    stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    specs = [sweep_corpus.make_spec(f"f{i:04d}", m, 7000 + i, corrupt_frac=0.06, non_speech_every=9)
             for i, m in enumerate((5.0, 3.0, 2.0))]
    lps = [sweep_corpus.emissions(s, "cuda", seed=i) for i, s in enumerate(specs)]
    files = [sw.SweepFile(s.file_id, s.audio_path, lp, s.n_samples, s.rows) for s, lp in zip(specs, lps)]
    run = sw.AnchorSweep(sw.SweepCorpus(files, stub.CharTokenizer()), index_duration=0.02,
                         samples_to_frames_ratio=320.0, groups=2, use_graphs=True)
    status = run.run(steps_per_poll=16)
    got = run.file_rows()
    with mp.get_context("fork").Pool(3) as pool:
        ref = pool.map(_sweep_cpu, [(s, lp.cpu().numpy()) for s, lp in zip(specs, lps)])
    for f, (rows, st) in enumerate(ref):
        assert sw.STATUS_NAMES[status[f]] == st
        assert got[f] == rows and len(rows) > 20
