"""CPU check of the arithmetic the linear-domain window-scoring instance is built on
(csrc/ctc_alpha.cu, LIN), restated here in NumPy -- not the kernel, which is
tested on the GPU in test_gpu_ctc.py:

* states divided by the blank emission of every frame walked so far (blank needs no multiply),
* emission ratios exp(lp[label] - lp[blank]) taken in fp32 and rounded to 20 mantissa bits (what
  the kernel keeps as the high word of an fp64),
* exact power-of-two re-scaling every 64 frames,
* the meet-in-the-middle join of a forward and a reversed walk.

The result must match the oracle (pinned against torch's ctc_loss) far inside the 1e-4 relative
tolerance of the north star, so the 20-bit ratios are not where the tolerance is spent."""
import numpy as np
import pytest

from cases import ctc_case
from oracle import ctc as octc


def _ratios(lp, blank):
    d = (lp - lp[:, [blank]]).astype(np.float32)
    r32 = np.exp2(d * np.float32(1.4426950408889634)).astype(np.float32)
    bits = r32.view(np.uint32).astype(np.uint64)
    hi = ((bits + 4) >> 3) + 0x38000000                      # high word of the fp64 with that value
    return (hi << 32).astype(np.uint64).view(np.float64)     # low word zero


def _walk(r, tg, frames):
    """Blank-normalised probability-domain recursion over `frames` (row indices, in walking order).
    Returns (blank states [L+1], label states [L], log2 of the scale taken out)."""
    L = len(tg)
    ab, al = np.zeros(L + 1), np.zeros(L)
    ab[0] = 1.0
    if L:
        al[0] = r[frames[0], tg[0]]
    skip = np.zeros(L, bool)
    skip[1:] = tg[1:] != tg[:-1]
    taken = 0
    for k, t in enumerate(frames[1:], 1):
        lm1 = np.concatenate(([0.0], al))                    # label of the pair to the left
        nb = ab + lm1
        x = np.where(skip, nb[:L], ab[:L])
        al = (al + x) * r[t, tg] if L else al
        ab = nb
        if k % 64 == 0:                                      # exact power-of-two re-scaling
            e = int(np.floor(np.log2(max(ab.max(), al.max() if L else 0.0))))
            ab, al, taken = np.ldexp(ab, -e), np.ldexp(al, -e), taken + e
    return ab, al, taken


def lin_nll(lp, tg, blank=0, bidir=True):
    T = lp.shape[0]
    L = len(tg)
    r = _ratios(lp, blank)
    base = float(lp[:, blank].astype(np.float64).sum())
    if not bidir or T < 16:
        ab, al, e = _walk(r, tg, list(range(T)))
        tot = ab[L] + (al[L - 1] if L else 0.0)
        return -(np.log(tot) + e * np.log(2.0) + base) if tot > 0 else np.inf
    m = (T - 1) >> 1
    fb, fl, fe = _walk(r, tg, list(range(m + 1)))                         # alpha_m
    rb, rl, re = _walk(r, tg[::-1], list(range(T - 1, m, -1)))            # b_{m+1}, reversed numbering
    Bb = rb[::-1]                                                         # blank of forward pair j
    Bl = rl[::-1]                                                         # label of forward pair j
    tot = 0.0
    for j in range(L + 1):
        succ = Bb[j] + (Bl[j] if j < L else 0.0)
        tot += fb[j] * succ
        if j < L:
            s2 = Bl[j] + Bb[j + 1]
            if j + 1 < L and tg[j + 1] != tg[j]:
                s2 += Bl[j + 1]
            tot += fl[j] * s2
    return -(np.log(tot) + (fe + re) * np.log(2.0) + base) if tot > 0 else np.inf


@pytest.mark.parametrize("shape", [(6, 200, 40, 32), (4, 1000, 100, 32), (6, 60, 7, 5), (4, 300, 0, 8),
                                   (4, 17, 3, 32), (4, 15, 3, 32)])
@pytest.mark.parametrize("peaked", [False, True])
def test_linear_domain_arithmetic_matches_oracle(shape, peaked):
    n, t, l, v = shape
    lp, tg, il, tl = ctc_case(70 + l, n, t, l, v, ragged=True, repeats=True, peaked=peaked)
    ref = octc.ctc_alpha_nll(lp, tg, il, tl).astype(np.float64)
    for i in range(n):
        for bidir in (False, True):
            got = lin_nll(lp[i, :il[i]], tg[i, :tl[i]].astype(np.int64), bidir=bidir)
            assert np.isfinite(ref[i]) == np.isfinite(got)
            if np.isfinite(got):
                # the oracle's own fp32 output carries ~1e-7 relative; the 20-bit ratios add less
                assert abs(got - ref[i]) <= 3e-6 * abs(ref[i]) + 1e-5, (i, bidir, got, ref[i])


def test_infeasible_target_has_zero_total():
    lp, tg, il, tl = ctc_case(5, 3, 40, 30, 32, repeats=True)
    for i in range(3):
        got = lin_nll(lp[i, :33], tg[i].astype(np.int64))
        ref = octc.ctc_alpha_nll(lp[i:i + 1, :33], tg[i:i + 1], np.array([33], np.int32), tl[i:i + 1])[0]
        assert np.isinf(got) == np.isinf(ref)


def _arrival_formula(tg):
    """First frame (0-based) at which blank_j / label_j can be alive, as csrc/ctc_alpha.cu computes it
    for its exactness guard: pair index + repeated labels so far (a repeat needs a blank in between)."""
    L = len(tg)
    reps = 0
    need_l, need_b = [], [0]
    for j in range(L):
        isrep = 1 if (j >= 1 and tg[j] == tg[j - 1]) else 0
        reps += isrep
        need_l.append(j + reps)
        if j >= 1:
            need_b[j] = need_l[j] - isrep
        need_b.append(0)
    need_b[L] = (need_l[L - 1] + 1) if L else 0
    return need_b[:L + 1], need_l


def _arrival_bruteforce(tg, T):
    L = len(tg)
    S = 2 * L + 1
    alive = np.zeros(S, bool)
    alive[0] = True
    if L:
        alive[1] = True
    first = np.where(alive, 0, -1)
    for t in range(1, T):
        new = alive.copy()
        new[1:] |= alive[:-1]
        for s in range(3, S, 2):              # skip s-2 -> s between different labels
            if tg[s // 2] != tg[s // 2 - 1]:
                new[s] |= alive[s - 2]
        first[(first < 0) & new] = t
        alive = new
    return first


@pytest.mark.parametrize("seed", range(8))
def test_guard_arrival_times_are_exact(seed):
    """The guard of the linear-domain instance only vouches for a window if every state that CAN be
    alive holds a normal number, so its arrival times must be the graph's, not an over-estimate."""
    rng = np.random.default_rng(seed)
    L = int(rng.integers(0, 24))
    tg = rng.integers(1, 4, L)                # small alphabet: many repeats
    need_b, need_l = _arrival_formula(tg)
    first = _arrival_bruteforce(tg, 3 * L + 4)
    for j in range(L + 1):
        assert first[2 * j] == need_b[j], (tg, j)
    for j in range(L):
        assert first[2 * j + 1] == need_l[j], (tg, j)
