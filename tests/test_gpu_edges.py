"""Edge cases of the three lattice kernels against the oracle: odd vocabulary sizes
(no 16-byte row copies), non-contiguous emission batches, blank != 0, T = 1, single
states, empty utterances, text exactly as long as the audio, non-default scoring windows."""
import numpy as np
import pytest

from cases import ctc_case, seg_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ipfa():
    import ipfa_b200
    return ipfa_b200


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _nll_ok(got, ref):
    assert np.array_equal(np.isinf(got), np.isinf(ref)), (got, ref)
    fin = np.isfinite(ref)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("v", [3, 5, 7, 29, 33, 63, 65, 130])
def test_odd_vocabulary_sizes(ipfa, v):
    """V % 4 != 0 takes the 4-byte cp.async path; V > 64 the gather panel."""
    from oracle import ctc as octc
    lp, tg, il, tl = ctc_case(50 + v, 5, 37, 9, v, ragged=True, repeats=True)
    _nll_ok(ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), il, tl).cpu().numpy(), octc.ctc_alpha_nll(lp, tg, il, tl))
    rp, rs, rst = octc.ctc_viterbi(lp, tg, il, tl)
    res = ipfa.ctc_forced_align(_dev(lp), _dev(tg), il, tl)
    assert np.array_equal(res.status.cpu().numpy() & 1, rst)
    for i in range(5):
        if rst[i] == 0:
            assert np.array_equal(res.paths[i, :il[i]].cpu().numpy(), rp[i, :il[i]])
            assert np.array_equal(res.scores[i, :il[i]].cpu().numpy(), rs[i, :il[i]])


def test_non_contiguous_batch_and_blank_index(ipfa):
    """Every other window of a larger buffer (stride_n = 2*T*V) and blank = V-1."""
    import torch
    from oracle import ctc as octc
    n, t, l, v = 6, 45, 11, 12
    rng = np.random.default_rng(3)
    big = rng.standard_normal((2 * n, t, v)).astype(np.float32)
    big = big - np.log(np.exp(big.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    lp = big[::2]
    blank = v - 1
    tg = rng.integers(0, v - 1, (n, l)).astype(np.int32)
    il = rng.integers(30, t + 1, n).astype(np.int32)
    tl = rng.integers(3, l + 1, n).astype(np.int32)
    dev = torch.from_numpy(big).cuda()[::2]
    assert not dev.is_contiguous()
    ref = octc.ctc_alpha_nll(np.ascontiguousarray(lp), tg, il, tl, blank=blank)
    _nll_ok(ipfa.ctc_alpha_nll(dev, tg, il, tl, blank=blank).cpu().numpy(), ref)
    rp, rs, rst = octc.ctc_viterbi(np.ascontiguousarray(lp), tg, il, tl, blank=blank)
    res = ipfa.ctc_forced_align(dev, tg, il, tl, blank=blank)
    for i in range(n):
        assert np.array_equal(res.paths[i, :il[i]].cpu().numpy(), rp[i, :il[i]])


def test_tiny_lattices(ipfa):
    """T = 1 and L in {0, 1}: the recursion never runs, only the init frame."""
    from oracle import ctc as octc
    lp, tg, il, tl = ctc_case(7, 4, 3, 1, 4)
    il = np.array([1, 1, 2, 3], np.int32)
    tl = np.array([1, 0, 1, 0], np.int32)
    _nll_ok(ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), il, tl).cpu().numpy(), octc.ctc_alpha_nll(lp, tg, il, tl))
    rp, rs, rst = octc.ctc_viterbi(lp, tg, il, tl)
    res = ipfa.ctc_forced_align(_dev(lp), _dev(tg), il, tl)
    assert np.array_equal(res.status.cpu().numpy() & 1, rst)
    for i in range(4):
        if rst[i] == 0:
            assert np.array_equal(res.paths[i, :il[i]].cpu().numpy(), rp[i, :il[i]])


def test_minus_inf_emissions_do_not_poison(ipfa):
    """log(0) emissions (masked vocabulary): finite answer when a path avoids them, inf otherwise."""
    import torch
    lp, tg, il, tl = ctc_case(9, 3, 20, 4, 6)
    lp[0, :, 5] = -np.inf                 # symbol 5 never emitted
    tg[0] = [1, 2, 3, 4]                  # window 0 avoids it
    tg[1] = [1, 5, 2, 3]
    lp[1, :, 5] = -np.inf                 # window 1 needs it: infeasible
    ref = torch.nn.functional.ctc_loss(torch.from_numpy(lp).transpose(0, 1), torch.from_numpy(tg).long(),
                                       torch.from_numpy(il).long(), torch.from_numpy(tl).long(),
                                       reduction="none").numpy()
    got = ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), il, tl).cpu().numpy()
    assert np.isinf(ref[1]) and np.isinf(got[1]) and np.isfinite(got[0])
    _nll_ok(got, ref)


def _seg_pack(cfg, utts_per_window):
    from oracle import ctcseg as oseg
    packed = [oseg.prepare_token_list(cfg, u) for u in utts_per_window]
    n = len(packed)
    cmax = max(len(g) for g, _ in packed)
    kmax = max(len(ub) - 1 for _, ub in packed)
    gt = np.full((n, cmax), -1, np.int32)
    ubs = np.zeros((n, kmax + 1), np.int32)
    n_cols, n_utts = np.zeros(n, np.int32), np.zeros(n, np.int32)
    for i, (g, ub) in enumerate(packed):
        gt[i, :len(g)] = g[:, 0]
        ubs[i, :len(ub)] = ub
        ubs[i, len(ub):] = ub[-1]
        n_cols[i], n_utts[i] = len(g), len(ub) - 1
    return packed, gt, ubs, n_cols, n_utts


@pytest.mark.parametrize("score_len,rounding", [(30, "floor"), (7, "floor"), (200, "floor"), (30, "round")])
def test_seg_scoring_variants(ipfa, score_len, rounding):
    import torch
    from oracle import ctcseg as oseg
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02, score_min_mean_over_L=score_len)
    oseg.SEG_INDEX_ROUNDING = rounding
    try:
        lp, in_len, utts = seg_case(61, 4, 500, 16, 3, 20, 45)
        packed, gt, ubs, n_cols, n_utts = _seg_pack(cfg, utts)
        flags = cfg.flags | (4 if rounding == "round" else 0)
        res = ipfa.ctcseg_align(torch.from_numpy(lp).cuda(), in_len, gt, n_cols, ubs, n_utts, 0.02,
                                score_len=score_len, flags=flags)
        for i in range(4):
            k = int(n_utts[i])
            ref = oseg.get_segments(cfg, lp[i, :in_len[i]], packed[i][0], packed[i][1], [""] * k)["segments"]
            got = res.seg[i, k - 1, :k].cpu().numpy()
            assert np.array_equal(got[:, :2], np.array(ref)[:, :2])
            np.testing.assert_allclose(got[:, 2], np.array(ref)[:, 2], rtol=1e-12)
    finally:
        oseg.SEG_INDEX_ROUNDING = "floor"


def test_seg_text_exactly_as_long_as_audio_and_single_token(ipfa):
    """N == T (every frame must switch) and one-token utterances."""
    import torch
    from oracle import ctcseg as oseg
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02)
    rng = np.random.default_rng(2)
    utts = [[rng.integers(1, 8, 6), rng.integers(1, 8, 5)], [np.array([3])], [np.array([2]), np.array([5])]]
    packed, gt, ubs, n_cols, n_utts = _seg_pack(cfg, utts)
    t = int(n_cols.max())
    lp = rng.standard_normal((3, t, 8)).astype(np.float32)
    lp = lp - np.log(np.exp(lp.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    in_len = np.array([n_cols[0], t, t], np.int32)  # window 0: N == T
    res = ipfa.ctcseg_align(torch.from_numpy(lp).cuda(), in_len, gt, n_cols, ubs, n_utts, 0.02, flags=cfg.flags | 8)
    assert int(res.status.abs().sum()) == 0
    for i in range(3):
        for k in range(1, int(n_utts[i]) + 1):
            gk, ubk = oseg.prepare_token_list(cfg, utts[i][:k])
            ref = oseg.get_segments(cfg, lp[i, :in_len[i]], gk, ubk, [""] * k)
            timing = res.timing[i, k - 1, :len(gk)].cpu().numpy()
            assert np.array_equal(np.where(timing < 0, 0.0, timing * 0.02), ref["timings"]), (i, k)
            got = res.seg[i, k - 1, :k].cpu().numpy()
            assert np.array_equal(got[:, :2], np.array(ref["segments"])[:, :2]), (i, k)
            np.testing.assert_allclose(got[:, 2], np.array(ref["segments"])[:, 2], rtol=1e-12)


def test_seg_wide_lattice_and_long_window(ipfa):
    """A 70 s window (T = 3500) with ~900 columns: the 16-warp instance, plus its 8000-frame limit."""
    import torch
    from oracle import ctcseg as oseg
    cfg = oseg.CtcSegmentationParameters(index_duration=0.02)
    lp, in_len, utts = seg_case(71, 2, 3500, 32, 6, 120, 150)
    packed, gt, ubs, n_cols, n_utts = _seg_pack(cfg, utts)
    res = ipfa.ctcseg_align(torch.from_numpy(lp).cuda(), in_len, gt, n_cols, ubs, n_utts, 0.02, flags=cfg.flags | 8,
                            details=False)
    for i in range(2):
        for k in (1, int(n_utts[i])):
            gk, ubk = oseg.prepare_token_list(cfg, utts[i][:k])
            ref = oseg.get_segments(cfg, lp[i, :in_len[i]], gk, ubk, [""] * k)["segments"]
            got = res.seg[i, k - 1, :k].cpu().numpy()
            assert np.array_equal(got[:, :2], np.array(ref)[:, :2])
            np.testing.assert_allclose(got[:, 2], np.array(ref)[:, 2], rtol=1e-12)
    with pytest.raises(RuntimeError, match="wider than the widest|unsupported|status 2"):
        big = torch.zeros(1, 8001, 4, device="cuda").log_softmax(-1)
        ipfa.ctcseg_align(big, [8001], [[-1, 0, 1, 0]], [4], [[1, 3]], [1], 0.02)


def test_text_round_equals_the_format_round_trip(ipfa):
    """Device rounding of times (.2f) and scores (.4f) == Python's float(f"{x:.Nf}") -- including the
    values where x * 10^N rounds to an exact half although x itself lies beside it (means of grid-valued
    emissions hit these: -1.08125 is not a binary fraction)."""
    import torch
    rng = np.random.default_rng(0)
    ks = rng.integers(-400000, 400000, 20000)
    halves = np.concatenate([(2 * ks + 1) / 20000.0, (2 * ks[:5000] + 1) / 200.0])
    vals = np.concatenate([halves, np.nextafter(halves, np.inf), np.nextafter(halves, -np.inf),
                           rng.standard_normal(20000) * 3, -rng.random(2000) * 1e-4, ks[:2000] / 30720.0,
                           [0.0, -0.0, -0.00004, 0.00005, -1.08125, 1e10, -1e10, -1e-300]])
    x = torch.as_tensor(vals, dtype=torch.float64, device="cuda")
    for decimals in (2, 4):
        got = ipfa.ops.text_round(x, decimals).cpu().numpy()
        want = np.array([float(f"{v:.{decimals}f}") for v in vals])
        assert np.array_equal(got, want) and np.array_equal(np.signbit(got), np.signbit(want))
