"""Consumer of tests/golden/ctcseg_golden.npz (written by tests/golden/make_ctcseg_golden.py on a
machine that has ``ctc-segmentation==1.7.1``, /root/reference/requirements.txt:13).

With the file present: the oracle (CPU) and the CUDA path (``-m gpu``) must reproduce what the real
package returned -- that is the pin the ``ctcseg`` lattice lacks in this image.  Without it: the CPU
test runs the generator in its ``--selftest`` mode (the repo's oracle standing in for the package) so
that generator and consumer stay in working order; that run pins nothing and says so.
"""
import itertools
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ctcseg_golden.npz")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ctcseg_golden as gen  # noqa: E402

SWITCHES = list(itertools.product(("floor", "round"), ("int+1", "ceil"), ("ascending", "shift")))


def _cases():
    """(case id, lp, ground-truth builder inputs, window, scoring length, char_list, classic?)"""
    for spec in gen.TOKEN_CASES:
        lp, utts, win, score_len = gen.token_case(spec)
        yield spec[0], lp, utts, win, score_len, [str(i) for i in range(spec[3])], False
    lp, utts, win, score_len = gen.integral_case()
    yield "integral_step", lp, utts, win, score_len, [str(i) for i in range(32)], False
    for cid, seed, t_len, win in (("classic_full", 71, 300, 8000), ("classic_slide", 72, 500, 160)):
        lp, utts = gen.classic_case(seed, t_len)
        yield cid, lp, utts, win, 30, list(gen.BPE_LIST), True


def _oracle_outputs(case, rounding, step_rule, cascade):
    from oracle import ctcseg as oseg
    cid, lp, utts, win, score_len, char_list, classic = case
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02, min_window_size=win, score_min_mean_over_L=score_len,
                                         char_list=char_list, window_step_rule=step_rule, offset_cascade=cascade)
    old = oseg.SEG_INDEX_ROUNDING
    oseg.SEG_INDEX_ROUNDING = rounding
    try:
        if classic:
            gt, ub = oseg.prepare_text(cfg, utts)
            text = utts
        else:
            gt, ub = oseg.prepare_token_list(cfg, [np.asarray(u) for u in utts])
            text = [" ".join(map(str, u)) for u in utts]
        timings, char_probs, state_list = oseg.ctc_segmentation(cfg, lp, gt)
        segs = oseg.determine_utterance_segments(cfg, ub, char_probs, timings, text)
    finally:
        oseg.SEG_INDEX_ROUNDING = old
    return {"gt": np.asarray(gt, np.int64), "utt_begin": np.asarray(ub, np.int64), "timings": timings,
            "char_probs": char_probs, "states": gen.encode_states(state_list, char_list, cfg.self_transition),
            "segments": np.asarray(segs, np.float64).reshape(len(text), 3)}


def _equal(got, npz, cid):
    return all(np.array_equal(got[k], npz[f"{cid}/{k}"]) for k in got)


def _check_oracle(npz):
    wrong = {}
    for case in _cases():
        if not _equal(_oracle_outputs(case, *SWITCHES[0]), npz, case[0]):
            wrong[case[0]] = [sw for sw in SWITCHES[1:] if _equal(_oracle_outputs(case, *sw), npz, case[0])]
    assert not wrong, ("the oracle's default reading of ctc-segmentation differs from the package; switch "
                       "combinations (seg_index_rounding, window_step_rule, offset_cascade) that reproduce it, "
                       "per case: %r" % wrong)


def _golden_or_selftest(tmp_path):
    if os.path.isfile(GOLDEN):
        return np.load(GOLDEN)
    out = str(tmp_path / "selftest.npz")
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_ctcseg_golden.py"), "--selftest", out],
                   check=True, capture_output=True)
    npz = np.load(out)
    assert "selftest" in str(npz["package_version"])  # not a pin: the oracle stands in for the package
    return npz


def test_oracle_reproduces_the_package(tmp_path):
    _check_oracle(_golden_or_selftest(tmp_path))


@pytest.mark.gpu
def test_cuda_path_reproduces_the_package(tmp_path):
    import torch
    import ipfa_b200
    npz = _golden_or_selftest(tmp_path)
    for cid, lp, utts, win, score_len, char_list, classic in _cases():
        gt, ub = npz[cid + "/gt"], npz[cid + "/utt_begin"]
        k, n_cols, t_len = len(ub) - 1, len(gt), lp.shape[0]
        w = win if t_len > win else None
        while True:
            res = ipfa_b200.ops.ctcseg_align(torch.from_numpy(lp).cuda()[None], [t_len],
                                             (gt[:, 0] if gt.shape[1] == 1 else gt).astype(np.int32)[None], [n_cols],
                                             ub.astype(np.int32)[None], [k], 0.02, score_len=score_len, flags=2,
                                             window=w)
            if w is None or not int(res.status[0]) & 8:
                break
            w *= 2
        timing = res.timing[0, k - 1, :n_cols].cpu().numpy()
        assert np.array_equal(np.where(timing < 0, 0.0, timing * 0.02), npz[cid + "/timings"]), cid
        assert np.array_equal(res.char_prob[0, k - 1, :t_len].cpu().numpy().astype(np.float64),
                              npz[cid + "/char_probs"]), cid
        state = res.state[0, k - 1, :t_len].cpu().numpy()
        tok = np.array([s if s < 0 else int(gt[s & 0xffffff, s >> 24]) for s in state])
        assert np.array_equal(tok, npz[cid + "/states"]), cid
        seg = res.seg[0, k - 1, :k].cpu().numpy()
        want = npz[cid + "/segments"]
        assert np.array_equal(seg[:, :2], want[:, :2]), cid
        np.testing.assert_allclose(seg[:, 2], want[:, 2], rtol=1e-12, atol=0)
