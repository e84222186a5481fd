"""Memory-safety checks that do not need compute-sanitizer (closed on this pool).

* Guard bands: every output tensor and the workspace of an operator call are carved out of larger
  allocations whose margins hold a canary pattern; the workspace is handed over with EXACTLY the
  size `ipfa_*_workspace_bytes` asked for.  After the call every margin must be intact -- an
  out-of-bounds write of a kernel (bulk-copy pipe, backpointer planes, per-prefix outputs, redo
  lists) lands in one.
* Repeatability: the same call twenty times must give the same bits -- the kernels stage emissions
  through shared-memory rings refilled by the copy engine (mbarriers, proxy fences) and exchange
  lattice values between warps; a hazard there shows up as run-to-run differences.
Shapes are chosen to hit every panel layout, the multi-warp instances, the length buckets, the redo
tiers of the window scorer, the windowed table mode and the anchor sweep."""
import importlib

import numpy as np
import pytest

from cases import ctc_case, seg_case

pytestmark = pytest.mark.gpu

GUARD = 4096
PATTERN = 0xA5


class _GuardedTorch:
    """`torch` as ops.py sees it, with CUDA allocations wrapped in canary margins."""

    def __init__(self, torch):
        self._torch = torch
        self.buffers = []

    def __getattr__(self, name):
        return getattr(self._torch, name)

    def _alloc(self, shape, dtype, device):
        torch = self._torch
        shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)))
        nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size()
        pad = (-nbytes) % 256
        raw = torch.full((GUARD + nbytes + pad + GUARD,), PATTERN, dtype=torch.uint8, device=device)
        self.buffers.append((raw, nbytes))
        return raw[GUARD:GUARD + nbytes].view(dtype).reshape(shape)

    def empty(self, *shape, dtype=None, device=None, **kw):
        if device is None or self._torch.device(device).type != "cuda":
            return self._torch.empty(*shape, dtype=dtype, device=device, **kw)
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return self._alloc(shape, dtype or self._torch.float32, device)

    def full(self, shape, value, dtype=None, device=None, **kw):
        if device is None or self._torch.device(device).type != "cuda":
            return self._torch.full(shape, value, dtype=dtype, device=device, **kw)
        out = self._alloc(shape, dtype or self._torch.float32, device)
        out.fill_(value)
        return out

    def zeros(self, *shape, dtype=None, device=None, **kw):
        if device is None or self._torch.device(device).type != "cuda":
            return self._torch.zeros(*shape, dtype=dtype, device=device, **kw)
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        out = self._alloc(shape, dtype or self._torch.float32, device)
        out.zero_()
        return out

    def empty_like(self, x, **kw):
        if not x.is_cuda:
            return self._torch.empty_like(x, **kw)
        return self._alloc(x.shape, x.dtype, x.device)

    def check(self):
        for raw, nbytes in self.buffers:
            head, tail = raw[:GUARD], raw[GUARD + nbytes:]
            assert bool((head == PATTERN).all()), "write BEFORE a buffer of %d bytes" % nbytes
            assert bool((tail == PATTERN).all()), "write BEYOND a buffer of %d bytes" % nbytes
        n = len(self.buffers)
        self.buffers = []
        return n


@pytest.fixture()
def guarded(monkeypatch):
    import torch
    import ipfa_b200
    ops = ipfa_b200.ops
    g = _GuardedTorch(torch)
    monkeypatch.setattr(ops, "torch", g)
    # the workspace with exactly the size asked for, fresh for every call
    monkeypatch.setattr(ops, "_workspace", lambda nbytes, device: g._alloc((max(int(nbytes), 256),), torch.uint8, device))
    return ipfa_b200, g


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


ALPHA_SHAPES = [(5, 40, 10, 32), (9, 300, 100, 32), (3, 700, 400, 40), (2, 400, 130, 700), (3, 150, 60, 5000),
                (4300, 40, 40, 32), (7, 33, 0, 32), (1, 1, 1, 3)]


@pytest.mark.parametrize("shape", ALPHA_SHAPES)
def test_window_scorer_writes_inside_its_buffers(guarded, shape):
    ipfa, g = guarded
    n, t, l, v = shape
    lp, tg, il, tl = ctc_case(7, n, t, max(l, 1), v, ragged=True, repeats=True, peaked=True)
    if l == 0:
        tl[:] = 0
    for switches in ({}, {"IPFA_ALPHA_F32": "1"}, {"IPFA_ALPHA_LOG": "1"}):
        with ipfa.tuning(**switches):
            out = ipfa.ops.ctc_alpha_nll(_dev(lp * 30.0) if switches else _dev(lp), _dev(tg), _dev(il), _dev(tl))
            out.sum().item()
        assert g.check() >= 2


@pytest.mark.parametrize("shape", ALPHA_SHAPES)
def test_viterbi_writes_inside_its_buffers(guarded, shape):
    ipfa, g = guarded
    n, t, l, v = shape
    lp, tg, il, tl = ctc_case(8, n, t, max(l, 1), v, ragged=True, repeats=True, peaked=True)
    if l == 0:
        tl[:] = 0
    res = ipfa.ops.ctc_forced_align(_dev(lp), _dev(tg), _dev(il), _dev(tl))
    res.total.sum().item()
    assert g.check() >= 5


SEG_SHAPES = [(6, 60, 8, 2, 2, 4, None), (4, 400, 32, 6, 8, 16, None), (2, 1500, 40, 8, 40, 60, None),
              (2, 600, 3000, 4, 10, 40, None), (2, 700, 32, 3, 10, 20, 256), (2, 1500, 32, 8, 40, 60, "spread2"),
              (2, 2500, 32, 8, 80, 120, "spread4")]


@pytest.mark.parametrize("shape", SEG_SHAPES)
def test_segmentation_writes_inside_its_buffers(guarded, shape):
    ipfa, g = guarded
    from oracle import ctcseg as oseg
    from test_gpu_ctcseg import _pack
    n, t, v, k_utts, lo, hi, window = shape
    lp, in_len, utts = seg_case(9, n, t, v, k_utts, lo, hi)
    gt, ubs, n_cols, n_utts = _pack(oseg.CtcSegmentationParameters(), utts)
    spread = {"spread2": 64, "spread4": 128}.get(window, 0)  # columns over a 2- / 4-CTA cluster
    res = ipfa.ops.ctcseg_align(_dev(lp), in_len, gt, n_cols, ubs, n_utts, 0.02, flags=2 | 8 | spread,
                                window=None if spread else window)
    res.status.sum().item()
    dec, anchor = ipfa.ops.anchor_select(res.seg, n_utts, np.full((n, res.seg.shape[1]), 40, np.int32),
                                         np.zeros(n, np.int32))
    dec.sum().item()
    assert g.check() >= 6


@pytest.mark.parametrize("mode", ["resident", "lockstep"])
def test_anchor_sweep_writes_inside_its_workspace(guarded, monkeypatch, mode):
    """Both device paths of the anchor sweep: the exactly-sized workspace (descriptors, backpointer words, timing /
    char_probs scratch, ticket) and the state / output arrays keep their canary margins -- also across a capacity
    growth (the first launch gets a workspace far too small for the corpus)."""
    ipfa, g = guarded
    import sweep_corpus
    sw = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.sweep")
    stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    monkeypatch.setattr(sw, "torch", g)  # sweep.py allocates its own workspaces / state / outputs
    specs = [sweep_corpus.make_spec(f"g{i}", 0.6 + 0.3 * i, 40 + i, corrupt_frac=0.3, non_speech_every=3) for i in range(3)]
    files = [sw.SweepFile(s.file_id, s.audio_path, sweep_corpus.emissions(s, "cuda", seed=i), s.n_samples, s.rows)
             for i, s in enumerate(specs)]
    for capacity in (None, (64, 16, 1)):
        run = sw.AnchorSweep(sw.SweepCorpus(files, stub.CharTokenizer()), index_duration=0.02,
                             samples_to_frames_ratio=320.0, groups=2, use_graphs=False, mode=mode, capacity=capacity)
        status = run.run(steps_per_poll=4)
        assert run.mode == mode and (status == sw.DONE).all()
        assert g.check() >= 10


def test_results_are_repeatable_bit_for_bit():
    """Twenty runs of every kernel family on the same inputs: identical bits."""
    import torch
    import ipfa_b200 as ipfa
    from oracle import ctcseg as oseg
    from test_gpu_ctcseg import _pack
    lp, tg, il, tl = ctc_case(3, 96, 500, 100, 32, ragged=True, repeats=True, peaked=True)
    a = [_dev(x) for x in (lp, tg, il, tl)]
    lp2, tg2, il2, tl2 = ctc_case(4, 12, 900, 400, 48, ragged=True, peaked=True)
    b = [_dev(x) for x in (lp2, tg2, il2, tl2)]
    slp, sil, utts = seg_case(5, 6, 1500, 32, 8, 40, 60)
    gt, ubs, n_cols, n_utts = _pack(oseg.CtcSegmentationParameters(), utts)
    sdev = _dev(slp)

    def valid_only(s):
        """The slots the call defines: prefix k < K_w, its own columns / frames / utterances."""
        parts = []
        for w in range(len(utts)):
            for k in range(int(n_utts[w])):
                parts += [s.seg[w, k, :k + 1].reshape(-1), s.term_t[w, k].reshape(1).double(),
                          s.timing[w, k, :int(ubs[w, k + 1]) + 1].double(), s.char_prob[w, k, :int(sil[w])].double(),
                          s.state[w, k, :int(sil[w])].double()]
        return torch.cat(parts)

    def run_all():
        out = [ipfa.ctc_alpha_nll(*a), ipfa.ctc_alpha_nll(*b)]
        with ipfa.tuning(IPFA_ALPHA_F32="1"):
            out.append(ipfa.ctc_alpha_nll(*a))
        with ipfa.tuning(IPFA_ALPHA_LOG="1"):
            out.append(ipfa.ctc_alpha_nll(*a))
        for args in (a, b):
            r = ipfa.ctc_forced_align(*args)
            out += [r.paths, r.scores, r.tok_start, r.tok_score, r.total]
        out.append(valid_only(ipfa.ctcseg_align(sdev, sil, gt, n_cols, ubs, n_utts, 0.02, flags=2 | 8)))
        return [o.clone() for o in out]

    first = run_all()
    for _ in range(19):
        again = run_all()
        for x, y in zip(first, again):
            same = (x == y) | ((x != x) & (y != y)) if x.is_floating_point() else (x == y)
            assert bool(same.all())
