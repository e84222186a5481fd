"""Seeded synthetic inputs shared by the GPU parity tests, smoke() and bench.py
(log-softmax of randn emissions, targets in [1, V))."""
import numpy as np


def ctc_case(seed, n, t, l, v, ragged=False, repeats=False, peaked=False):
    rng = np.random.default_rng(seed)
    lp = rng.standard_normal((n, t, v)).astype(np.float32)
    targets = rng.integers(1, v, (n, max(l, 1))).astype(np.int32)
    if repeats and l > 1:
        m = rng.random((n, l - 1)) < 0.35
        for i in range(1, l):
            targets[:, i] = np.where(m[:, i - 1], targets[:, i - 1], targets[:, i])
    if ragged:
        in_len = rng.integers(max(t // 2, 1), t + 1, n).astype(np.int32)
        tgt_len = rng.integers(max(l // 2, 0), l + 1, n).astype(np.int32)
    else:
        in_len = np.full(n, t, np.int32)
        tgt_len = np.full(n, l, np.int32)
    if peaked:
        for i in range(n):
            ti, li = int(in_len[i]), int(tgt_len[i])
            if li == 0:
                continue
            pos = np.sort(rng.permutation(ti)[:min(li, ti)])
            lp[i, pos, targets[i, :len(pos)]] += 6.0
            lp[i, :, 0] += 1.0
    lp = lp - np.log(np.exp(lp.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    if l == 0:
        targets = targets[:, :0]
    return lp, targets, in_len, tgt_len


def seg_case(seed, n, t, v, k_utts, tok_lo=2, tok_hi=8, peaked=True, ragged=True):
    """Windows for the ctcseg lattice: each window has up to k_utts utterances of
    tok_lo..tok_hi tokens.  Returns (lp, in_len, utts) with utts[w] = list of int arrays."""
    rng = np.random.default_rng(seed)
    lp = rng.standard_normal((n, t, v)).astype(np.float32)
    in_len = (rng.integers(max(t * 2 // 3, 1), t + 1, n) if ragged else np.full(n, t)).astype(np.int32)
    utts = []
    for w in range(n):
        k = int(rng.integers(1, k_utts + 1)) if ragged else k_utts
        us = [rng.integers(1, v, int(rng.integers(tok_lo, tok_hi + 1))).astype(np.int64) for _ in range(k)]
        utts.append(us)
        if peaked:
            flat = []
            for u in us:
                flat += [0] + u.tolist()
            flat += [0]
            ti = int(in_len[w])
            if len(flat) <= ti:
                pos = np.sort(rng.permutation(ti)[:len(flat)])
                bounds = list(pos) + [ti]
                for j, tok in enumerate(flat):
                    lp[w, bounds[j]:bounds[j + 1], tok] += 4.0
    lp = lp - np.log(np.exp(lp.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    return lp, in_len, utts
