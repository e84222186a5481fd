"""GPU parity of kernels (1) and (2a) against the pinned oracle and the committed
torch / torchaudio golden vectors, through the C ABI (device and host entry points).

Tolerances (BASELINE.json north_star): loss and confidences within 1e-4 relative
in fp32; alignment paths, frame indices and per-frame scores bit-exact."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES
from cases import ctc_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ipfa():
    import ipfa_b200
    return ipfa_b200


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check_nll(got, ref):
    assert np.array_equal(np.isinf(got), np.isinf(ref)), (got, ref)
    fin = np.isfinite(ref)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_alpha_golden(ipfa, golden, name):
    lp, tg = golden[f"{name}/lp"], golden[f"{name}/targets"]
    il, tl = golden[f"{name}/in_len"], golden[f"{name}/tgt_len"]
    nll = ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), _dev(il), _dev(tl)).cpu().numpy()
    _check_nll(nll, golden[f"{name}/nll"])
    nll_h = ipfa.ctc_alpha_nll_host(lp, tg, il, tl)
    assert np.array_equal(nll_h, nll, equal_nan=True)


def test_alpha_empty_target(ipfa, golden):
    lp = golden["empty/lp"]
    n = lp.shape[0]
    nll = ipfa.ctc_alpha_nll(_dev(lp), _dev(np.zeros((n, 0), np.int32)), _dev(golden["empty/in_len"]),
                             _dev(np.zeros(n, np.int32))).cpu().numpy()
    np.testing.assert_allclose(nll, golden["empty/nll"], rtol=1e-5)


# (n, t, l, v, ragged, repeats, peaked): every (P, WARPS) lattice shape and both panel layouts
ALPHA_SHAPES = [
    (5, 40, 10, 32, True, True, False),       # P=1
    (5, 90, 50, 32, True, True, False),       # P=2
    (9, 300, 100, 32, False, False, False),   # P=4 (BASELINE config 2 lattice)
    (3, 500, 200, 32, True, True, True),      # 2 warps
    (3, 700, 400, 40, True, False, True),     # 4 warps
    (2, 1200, 900, 48, True, False, True),    # 8 warps
    (2, 400, 130, 700, True, True, True),     # gather panel, 2 warps
    (3, 150, 60, 5000, True, True, False),    # gather panel, 1 warp (config 4 vocabulary)
    (1, 2500, 1800, 64, False, False, True),  # P=8, 8 warps
]


@pytest.mark.parametrize("shape", ALPHA_SHAPES)
def test_alpha_vs_oracle(ipfa, shape):
    from oracle import ctc as octc
    n, t, l, v, ragged, repeats, peaked = shape
    lp, tg, il, tl = ctc_case(11, n, t, l, v, ragged, repeats, peaked)
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    nll = ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), _dev(il), _dev(tl)).cpu().numpy()
    _check_nll(nll, ref)


def test_alpha_time_major_strides(ipfa):
    """[T, N, V] input (torch ctc_loss layout) goes through strides, not a copy."""
    import torch
    from oracle import ctc as octc
    lp, tg, il, tl = ctc_case(12, 6, 70, 20, 32, True, True)
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    lp_tm = torch.from_numpy(lp).cuda().transpose(0, 1).contiguous()
    nll = ipfa.ctc_alpha_nll(lp_tm, _dev(tg), _dev(il), _dev(tl), batch_first=False).cpu().numpy()
    _check_nll(nll, ref)


def test_alpha_infeasible_and_zero_length(ipfa):
    from oracle import ctc as octc
    lp, tg, il, tl = ctc_case(13, 4, 6, 6, 5, repeats=True)
    il = np.array([6, 3, 0, 6], np.int32)
    tl = np.array([6, 6, 2, 0], np.int32)
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    nll = ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), _dev(il), _dev(tl)).cpu().numpy()
    _check_nll(nll, ref)


def _check_viterbi(res, ref_paths, ref_scores, ref_status, in_len, to_np=lambda x: x.cpu().numpy()):
    paths, scores, status = to_np(res.paths), to_np(res.scores), to_np(res.status)
    assert np.array_equal(status & 1, ref_status), (status, ref_status)
    for i in range(len(in_len)):
        t = int(in_len[i])
        if ref_status[i]:
            assert np.all(paths[i] == -1)
            continue
        assert np.array_equal(paths[i, :t], ref_paths[i, :t]), i      # bit-exact
        assert np.array_equal(scores[i, :t], ref_scores[i, :t]), i    # bit-exact fp32 gather
        assert np.all(paths[i, t:] == -1)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_viterbi_golden(ipfa, golden, name):
    lp, tg = golden[f"{name}/lp"], golden[f"{name}/targets"]
    il, tl = golden[f"{name}/in_len"], golden[f"{name}/tgt_len"]
    res = ipfa.ctc_forced_align(_dev(lp), _dev(tg), _dev(il), _dev(tl))
    _check_viterbi(res, golden[f"{name}/paths"], golden[f"{name}/scores"], golden[f"{name}/fa_status"], il)
    res_h = ipfa.ctc_forced_align_host(lp, tg, il, tl)
    _check_viterbi(res_h, golden[f"{name}/paths"], golden[f"{name}/scores"], golden[f"{name}/fa_status"], il,
                   to_np=lambda x: x)


def test_viterbi_tie_rules(ipfa, golden):
    """Exact ties resolve like torchaudio."""
    for i in range(int(golden["ties/count"])):
        lp = golden[f"ties/{i}/lp"][None]
        tg = golden[f"ties/{i}/targets"][None]
        res = ipfa.ctc_forced_align(_dev(lp), _dev(tg), [lp.shape[1]], [tg.shape[1]])
        assert int(res.status[0]) == 0
        assert np.array_equal(res.paths[0].cpu().numpy(), golden[f"ties/{i}/paths"]), i
        assert np.array_equal(res.scores[0].cpu().numpy(), golden[f"ties/{i}/scores"]), i


@pytest.mark.parametrize("shape", ALPHA_SHAPES)
def test_viterbi_vs_oracle(ipfa, shape):
    from oracle import ctc as octc
    n, t, l, v, ragged, repeats, peaked = shape
    lp, tg, il, tl = ctc_case(21, n, t, l, v, ragged, repeats, peaked)
    ref_paths, ref_scores, ref_status = octc.ctc_viterbi(lp, tg, il, tl)
    res = ipfa.ctc_forced_align(_dev(lp), _dev(tg), _dev(il), _dev(tl))
    _check_viterbi(res, ref_paths, ref_scores, ref_status, il)
    # token spans / confidences against merge_tokens on the oracle's path
    ts, te, tp = res.tok_start.cpu().numpy(), res.tok_end.cpu().numpy(), res.tok_score.cpu().numpy()
    total = res.total.cpu().numpy()
    for i in range(n):
        if ref_status[i]:
            continue
        ti, li = int(il[i]), int(tl[i])
        spans = octc.merge_tokens(ref_paths[i, :ti], ref_scores[i, :ti])
        # merge_tokens fuses a repeated label only across a blank, so spans == tokens
        assert len(spans) == li
        assert [s[1] for s in spans] == ts[i, :li].tolist()
        assert [s[2] for s in spans] == te[i, :li].tolist()
        np.testing.assert_allclose(tp[i, :li], [s[3] for s in spans], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(total[i], ref_scores[i, :ti].astype(np.float64).sum(), rtol=1e-4)


def test_viterbi_tight_and_infeasible(ipfa):
    """T == L + repeats (single feasible path) and T < L + repeats (rejected)."""
    from oracle import ctc as octc
    rng = np.random.default_rng(5)
    n, l, v = 12, 9, 6
    lp, tg, il, tl = ctc_case(31, n, 20, l, v, repeats=True)
    rep = (tg[:, 1:] == tg[:, :-1]).sum(1)
    il = (l + rep + rng.integers(-1, 2, n)).astype(np.int32)
    ref_paths, ref_scores, ref_status = octc.ctc_viterbi(lp, tg, il, tl)
    assert ref_status.any() and not ref_status.all()
    res = ipfa.ctc_forced_align(_dev(lp), _dev(tg), _dev(il), _dev(tl))
    _check_viterbi(res, ref_paths, ref_scores, ref_status, il)


def test_full_size_config2_properties(ipfa):
    """BASELINE config 2 at full size (1024 x T=1000 x L=100, V=32): size-independent
    properties -- (i) the Viterbi path score never exceeds the total log-likelihood,
    (ii) each path collapses to its target, (iii) frame scores sum to the path score,
    (iv) a sample of windows matches the oracle."""
    import torch
    from oracle import ctc as octc
    n, t, l, v = 1024, 1000, 100, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    lp = torch.randn(n, t, v, generator=g, device="cuda").log_softmax(-1)
    tg = torch.randint(1, v, (n, l), generator=g, device="cuda", dtype=torch.int32)
    il = torch.full((n,), t, dtype=torch.int32, device="cuda")
    tl = torch.full((n,), l, dtype=torch.int32, device="cuda")
    nll = ipfa.ctc_alpha_nll(lp, tg, il, tl)
    res = ipfa.ctc_forced_align(lp, tg, il, tl)
    assert torch.isfinite(nll).all() and int(res.status.abs().sum()) == 0
    assert bool((res.total <= -nll + 1e-3).all())
    np.testing.assert_allclose(res.scores.double().sum(1).cpu().numpy(), res.total.cpu().numpy(), rtol=1e-5)
    paths = res.paths.cpu().numpy()
    tgn = tg.cpu().numpy()
    for i in range(0, n, 37):
        p = paths[i]
        keep = np.concatenate([[True], p[1:] != p[:-1]]) & (p != 0)
        assert np.array_equal(p[keep], tgn[i])
    idx = np.arange(0, n, 128)
    lps = lp[idx].cpu().numpy()
    ref = octc.ctc_alpha_nll(lps, tgn[idx], np.full(len(idx), t, np.int32), np.full(len(idx), l, np.int32))
    _check_nll(nll[idx].cpu().numpy(), ref)
    rp, rs, rst = octc.ctc_viterbi(lps, tgn[idx], np.full(len(idx), t, np.int32), np.full(len(idx), l, np.int32))
    assert np.array_equal(paths[idx], rp)
    # host-buffer entry points (chunked H2D / kernel overlap, other lattice instances per chunk):
    # same bytes as the device entry points
    lp_h = lp.cpu().numpy()
    il_h, tl_h = np.full(n, t, np.int32), np.full(n, l, np.int32)
    nll_h = ipfa.ctc_alpha_nll_host(lp_h, tgn, il_h, tl_h)
    np.testing.assert_allclose(nll_h, nll.cpu().numpy(), rtol=2e-6)
    res_h = ipfa.ctc_forced_align_host(lp_h, tgn, il_h, tl_h, tokens=False)
    assert np.array_equal(res_h.paths, paths) and np.array_equal(res_h.scores, res.scores.cpu().numpy())


def test_viterbi_length_buckets(ipfa, monkeypatch):
    """Large ragged batch (>= 4096 windows): windows whose lattice fits half the launch width run on
    a half-width instance, decided on the device.  Same paths / scores / tokens as the oracle and as
    the single-instance launch."""
    import torch
    from oracle import ctc as octc
    n, t, l, v = 4500, 48, 40, 32
    lp, tg, il, tl = ctc_case(31, n, t, l, v, ragged=True, repeats=True, peaked=True)
    tl[:7] = [0, 1, 31, 32, 33, 40, 16]   # both sides of the bucket boundary (32 state pairs)
    il[:7] = t
    ref_paths, ref_scores, ref_status = octc.ctc_viterbi(lp, tg, il, tl)
    dev = [_dev(x) for x in (lp, tg, il, tl)]
    res = ipfa.ctc_forced_align(*dev)
    _check_viterbi(res, ref_paths, ref_scores, ref_status, il)
    with ipfa.tuning(IPFA_NO_BUCKETS="1"):
        one = ipfa.ctc_forced_align(*dev)
    for name in ("paths", "scores", "tok_start", "tok_end", "tok_score", "total", "status"):
        assert torch.equal(getattr(res, name), getattr(one, name)), name
    assert ((tl + 1 <= 32).sum() > 1000) and ((tl + 1 > 32).sum() > 300)


def test_alpha_length_buckets(ipfa, monkeypatch):
    """Kernel (1) on a large ragged batch: two length buckets decided on the device."""
    import torch
    from oracle import ctc as octc
    n, t, l, v = 4300, 40, 40, 32
    lp, tg, il, tl = ctc_case(32, n, t, l, v, ragged=True, repeats=True)
    tl[:6] = [0, 1, 31, 32, 33, 40]
    il[:6] = t
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    dev = [_dev(x) for x in (lp, tg, il, tl)]
    nll = ipfa.ctc_alpha_nll(*dev)
    _check_nll(nll.cpu().numpy(), ref)
    with ipfa.tuning(IPFA_NO_BUCKETS="1"):
        one = ipfa.ctc_alpha_nll(*dev)
    _check_nll(one.cpu().numpy(), ref)
    # the two launch shapes associate the log-sum-exp differently: equal to fp32 rounding
    fin = torch.isfinite(one)
    assert torch.equal(torch.isfinite(nll), fin)
    assert torch.allclose(nll[fin], one[fin], rtol=1e-5, atol=1e-5)


# --- kernel (1), linear-domain (fp64) instance and its hand-over to the log-domain instance --------
LIN_SHAPES = [(64, 200, 40, 32), (16, 400, 31, 32), (16, 400, 32, 32), (16, 400, 63, 8), (8, 900, 128, 32),
              (4, 1200, 255, 64), (16, 17, 3, 32), (16, 15, 3, 32), (32, 257, 64, 33), (8, 30, 0, 32)]


@pytest.mark.parametrize("shape", LIN_SHAPES)
@pytest.mark.parametrize("peaked", [False, True])
def test_alpha_linear_instance(ipfa, monkeypatch, shape, peaked):
    """Dense panels with <= 256 state pairs run the linear-domain instance (csrc/ctc_alpha.cu, LIN):
    same results as the oracle and as the log-domain instance, nothing handed over on ordinary
    emissions."""
    from oracle import ctc as octc
    n, t, l, v = shape
    lp, tg, il, tl = ctc_case(40 + l, n, t, l, v, ragged=True, repeats=True, peaked=peaked)
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    dev = [_dev(x) for x in (lp, tg, il, tl)]
    lin = ipfa.ctc_alpha_nll(*dev).cpu().numpy()
    assert ipfa.ctc_alpha_redo_count(n) == 0
    _check_nll(lin, ref)
    with ipfa.tuning(IPFA_ALPHA_LOG="1"):
        log = ipfa.ctc_alpha_nll(*dev).cpu().numpy()
    _check_nll(log, ref)
    fin = np.isfinite(ref)
    np.testing.assert_allclose(lin[fin], log[fin], rtol=2e-5)


def test_alpha_linear_guard_hands_over(ipfa):
    """Inputs the linear-domain instance cannot vouch for go through the redo list to the
    log-domain instance: emissions sharper than its range, -inf emissions, a target that names the
    blank, infeasible targets.  Results equal the oracle's either way."""
    import torch
    from oracle import ctc as octc
    rng = np.random.default_rng(7)
    n, t, l, v = 32, 300, 30, 32
    tg = rng.integers(1, v, (n, l)).astype(np.int32)
    il, tl = np.full(n, t, np.int32), np.full(n, l, np.int32)
    for scale, expect_redo in ((8.0, False), (40.0, True), (200.0, True)):
        raw = (rng.standard_normal((n, t, v)) * scale).astype(np.float32)
        lp = torch.from_numpy(raw).log_softmax(-1).numpy()
        got = ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), _dev(il), _dev(tl)).cpu().numpy()
        redo = ipfa.ctc_alpha_redo_count(n)
        assert (redo > 0) == expect_redo, (scale, redo)
        _check_nll(got, octc.ctc_alpha_nll(lp, tg, il, tl))
    # -inf emissions: one symbol on a few frames of every other window, the blank everywhere in one
    lp, tg, il, tl = ctc_case(3, 16, 120, 12, 32)
    lp[::2, 5:9, 3] = -np.inf
    lp[1, :, 0] = -np.inf
    got = ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), _dev(il), _dev(tl)).cpu().numpy()
    assert ipfa.ctc_alpha_redo_count(16) >= 9
    _check_nll(got, octc.ctc_alpha_nll(lp, tg, il, tl))
    # infeasible (T < L + repeats): inf from the log-domain instance, like torch
    lp, tg, il, tl = ctc_case(5, 8, 40, 30, 32, repeats=True)
    il[:] = 33
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    assert np.isinf(ref).any()
    got = ipfa.ctc_alpha_nll(_dev(lp), _dev(tg), _dev(il), _dev(tl)).cpu().numpy()
    assert ipfa.ctc_alpha_redo_count(8) == int(np.isinf(ref).sum())
    _check_nll(got, ref)


def test_alpha_linear_blank_in_target(ipfa, monkeypatch):
    """A target that names the blank symbol is handed over (the linear panel keeps raw logs in
    the blank column); both instances agree."""
    lp, tg, il, tl = ctc_case(4, 16, 120, 12, 32)
    tg[::3, 4] = 0
    dev = [_dev(x) for x in (lp, tg, il, tl)]
    lin = ipfa.ctc_alpha_nll(*dev).cpu().numpy()
    assert ipfa.ctc_alpha_redo_count(16) == 6
    with ipfa.tuning(IPFA_ALPHA_LOG="1"):
        log = ipfa.ctc_alpha_nll(*dev).cpu().numpy()
    np.testing.assert_allclose(lin, log, rtol=1e-6)


def test_alpha_linear_large_batch_buckets(ipfa, monkeypatch):
    """Linear-domain instance on a large ragged batch (two length buckets, P and P/2 pairs per lane)."""
    from oracle import ctc as octc
    n, t, l, v = 4300, 64, 60, 32
    lp, tg, il, tl = ctc_case(33, n, t, l, v, ragged=True, repeats=True)
    tl[:6] = [0, 1, 31, 32, 33, 60]
    il[:6] = t
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    dev = [_dev(x) for x in (lp, tg, il, tl)]
    got = ipfa.ctc_alpha_nll(*dev).cpu().numpy()
    n_inf = int(np.isinf(ref).sum())
    assert ipfa.ctc_alpha_redo_count(n) == n_inf
    _check_nll(got, ref)


@pytest.mark.parametrize("shape_env", ["1,2", "2,2", "4,2", "8,1"])
def test_alpha_linear_instance_shapes(ipfa, monkeypatch, shape_env):
    """The other instances of the linear-domain kernel (two warps per half window, 8 pairs per
    lane), forced through IPFA_ALPHA_LIN_SHAPE: same results, nothing handed over."""
    from oracle import ctc as octc
    p, w = (int(x) for x in shape_env.split(","))
    l = min(32 * p * w - 1, 200)
    lp, tg, il, tl = ctc_case(50 + l, 12, 3 * l + 40, l, 32, ragged=True, repeats=True, peaked=(p == 2))
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    with ipfa.tuning(IPFA_ALPHA_LIN_SHAPE=shape_env):
        got = ipfa.ctc_alpha_nll(*[_dev(x) for x in (lp, tg, il, tl)]).cpu().numpy()
    assert ipfa.ctc_alpha_redo_count(12) == int(np.isinf(ref).sum())
    _check_nll(got, ref)


def test_alpha_fp32_tier_is_exact_where_it_answers(ipfa):
    """The optional fp32 tier (IPFA_ALPHA_F32=1) in front of the fp64 one: whatever it cannot vouch for
    goes down the redo lists (fp32 -> fp64 -> log domain), so results do not depend on it."""
    from oracle import ctc as octc
    for seed, n, t, l, peaked in ((1, 48, 300, 100, False), (2, 24, 500, 100, True), (3, 40, 90, 50, False),
                                  (6, 64, 40, 10, False), (8, 33, 200, 0, False)):
        lp, tg, il, tl = ctc_case(seed, n, t, max(l, 1), 32, ragged=True, repeats=True, peaked=peaked)
        if l == 0:
            tl[:] = 0
        ref = octc.ctc_alpha_nll(lp, tg, il, tl)
        dev = [_dev(x) for x in (lp, tg, il, tl)]
        plain = ipfa.ctc_alpha_nll(*dev).cpu().numpy()
        with ipfa.tuning(IPFA_ALPHA_F32="1"):
            tiered = ipfa.ctc_alpha_nll(*dev).cpu().numpy()
        _check_nll(tiered, ref)
        fin = np.isfinite(ref)
        np.testing.assert_allclose(tiered[fin], plain[fin], rtol=2e-5)


@pytest.mark.parametrize("shape", [(3, 150, 60, 5000), (2, 400, 130, 700), (4, 300, 100, 32), (2, 95, 40, 33)])
def test_vocabulary_major_emissions(ipfa, shape):
    """Emissions stored [N, V, T] (stride_v = T, stride_t = 1) through the strided entry points: the
    window scorer and Viterbi read runs of frames per column instead of rows; results do not depend on
    the storage order."""
    import torch
    from oracle import ctc as octc
    n, t, l, v = shape
    lp, tg, il, tl = ctc_case(21, n, t, l, v, ragged=True, repeats=True, peaked=True)
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    paths, scores, status = octc.ctc_viterbi(lp, tg, il, tl)
    rows = _dev(lp)
    cols = rows.transpose(1, 2).contiguous().transpose(1, 2)  # [N, T, V] view of [N, V, T] storage
    assert cols.stride(2) == t and cols.stride(1) == 1 and torch.equal(cols, rows)
    a = [_dev(x) for x in (tg, il, tl)]
    nll = ipfa.ctc_alpha_nll(cols, *a).cpu().numpy()
    _check_nll(nll, ref)
    res = ipfa.ctc_forced_align(cols, *a)
    _check_viterbi(res, paths, scores, status, il)
    # and a non-unit stride on a row-major tensor (every other column of a wider matrix)
    wide = torch.zeros((n, t, 2 * v), device="cuda")
    wide[:, :, ::2] = rows
    _check_nll(ipfa.ctc_alpha_nll(wide[:, :, ::2], *a).cpu().numpy(), ref)
