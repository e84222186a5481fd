"""Pins the ``ctcseg`` lattice: writes tests/golden/ctcseg_golden.npz from the REAL package.

    pip install ctc-segmentation==1.7.1      # /root/reference/requirements.txt:13
    python tests/golden/make_ctcseg_golden.py

The build image has neither the package nor a network, so ``oracle/ctcseg*.{c,py}`` restate its
algorithm from the published description ("parity unpinned").  On a machine that has the package,
this script runs ``ctc_segmentation.prepare_token_list`` / ``prepare_text`` / ``ctc_segmentation`` /
``determine_utterance_segments`` on the seeded cases below and stores what they return;
``tests/test_ctcseg_golden.py`` then holds the oracle (CPU) and the CUDA path (``-m gpu``) against the
file -- and, where the default reading of a ``[verify]`` spot is the wrong one, reports which
combination of the switches (``seg_index_rounding``, ``window_step_rule``, ``offset_cascade``)
reproduces the package.

``--selftest <out.npz>`` runs the same code with the repo's oracle standing in for the package
(used by the CPU tests to keep this script and its consumer working; it pins nothing).
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (ROOT, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)

from cases import seg_case  # noqa: E402

BPE_LIST = ["<blank>", "a", "b", "c", "d", "ab", "bc", "cd", "abc", "bcd", "·"]

# (case id, seed, T, V, utterances, tokens per utterance lo..hi, min_window_size, scoring length)
TOKEN_CASES = [
    ("short", 41, 60, 8, 2, 2, 4, 8000, 30),
    ("mean_branch", 42, 200, 32, 4, 3, 9, 8000, 5),
    ("window_70s", 43, 3500, 32, 6, 60, 120, 8000, 30),
    ("slide_256", 44, 1000, 32, 5, 8, 16, 256, 30),
    ("slide_doubling", 45, 2300, 32, 4, 20, 40, 300, 30),
]


def token_case(spec):
    cid, seed, t, v, k, lo, hi, win, score_len = spec
    lp, in_len, utts = seg_case(seed, 1, t, v, k, lo, hi, peaked=True, ragged=False)
    return lp[0], utts[0], win, score_len


def integral_case(seed=63, window=256, step=3):
    """(T - W) / N is an integer: the window-step rule matters (see tests/test_gpu_ctcseg.py)."""
    rng = np.random.default_rng(seed)
    utts = [rng.integers(1, 32, int(rng.integers(12, 24))).astype(np.int64) for _ in range(4)]
    flat = []
    for u in utts:
        flat += [0] + u.tolist()
    flat += [0]
    t = window + step * (len(flat) + 1)
    lp = rng.standard_normal((t, 32)).astype(np.float32)
    pos = np.sort(rng.permutation(t)[:len(flat)])
    for j, (a, b) in enumerate(zip(pos, list(pos[1:]) + [t])):
        lp[a:b, flat[j]] += 4.0
    lp = lp - np.log(np.exp(lp.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    return lp, utts, window, 30


def classic_case(seed, t_len, n_utts=4):
    """Strings over {a,b,c,d} with multi-character tokens: ``prepare_text``, up to 3 candidates."""
    rng = np.random.default_rng(seed)
    utts = ["".join(rng.choice(list("abcd"), size=int(rng.integers(4, 12)))) for _ in range(n_utts)]
    lp = rng.standard_normal((t_len, len(BPE_LIST))).astype(np.float32)
    lp = lp - np.log(np.exp(lp.astype(np.float64)).sum(-1, keepdims=True)).astype(np.float32)
    return lp, utts


def encode_states(state_list, char_list, self_transition):
    """state_list (str per frame) -> int32: -2 unset, -1 self transition, else index in char_list."""
    idx = {c: i for i, c in reversed(list(enumerate(char_list)))}
    return np.array([-2 if s == "" else (-1 if s == self_transition else idx[s]) for s in state_list], np.int32)


def run(pkg, out_path):
    store = {}

    def put(cid, config, lp, gt, ub, text):
        timings, char_probs, state_list = pkg.ctc_segmentation(config, lp, gt)
        segments = pkg.determine_utterance_segments(config, ub, char_probs, timings, text)
        store[cid + "/gt"] = np.asarray(gt, np.int64)
        store[cid + "/utt_begin"] = np.asarray(ub, np.int64)
        store[cid + "/timings"] = np.asarray(timings, np.float64)
        store[cid + "/char_probs"] = np.asarray(char_probs, np.float64)
        store[cid + "/states"] = encode_states(state_list, config.char_list, config.self_transition)
        store[cid + "/segments"] = np.asarray(segments, np.float64).reshape(len(text), 3)

    def config_for(char_list, win, score_len):
        config = pkg.CtcSegmentationParameters()
        config.index_duration = 0.02
        config.min_window_size = win
        config.score_min_mean_over_L = score_len
        config.char_list = list(char_list)
        return config

    for spec in TOKEN_CASES:
        lp, utts, win, score_len = token_case(spec)
        config = config_for([str(i) for i in range(spec[3])], win, score_len)
        gt, ub = pkg.prepare_token_list(config, [np.asarray(u) for u in utts])
        put(spec[0], config, lp, gt, ub, [" ".join(map(str, u)) for u in utts])
    lp, utts, win, score_len = integral_case()
    config = config_for([str(i) for i in range(32)], win, score_len)
    gt, ub = pkg.prepare_token_list(config, [np.asarray(u) for u in utts])
    put("integral_step", config, lp, gt, ub, [" ".join(map(str, u)) for u in utts])
    for cid, seed, t_len, win in (("classic_full", 71, 300, 8000), ("classic_slide", 72, 500, 160)):
        lp, utts = classic_case(seed, t_len)
        config = config_for(BPE_LIST, win, 30)
        gt, ub = pkg.prepare_text(config, utts)
        put(cid, config, lp, gt, ub, utts)
    store["package_version"] = np.array(getattr(pkg, "__version__", "unknown"))
    np.savez_compressed(out_path, **store)
    return sorted(k for k in store if k.endswith("/segments"))


def oracle_as_package():
    """The repo's oracle behind the package's function names (selftest only)."""
    import types
    from oracle import ctcseg as oseg
    pkg = types.SimpleNamespace(CtcSegmentationParameters=oseg.CtcSegmentationParameters,
                                ctc_segmentation=oseg.ctc_segmentation, prepare_token_list=oseg.prepare_token_list,
                                prepare_text=oseg.prepare_text,
                                determine_utterance_segments=oseg.determine_utterance_segments,
                                __version__="selftest (repo oracle, pins nothing)")
    return pkg


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--selftest", default=None, help="write here using the repo's oracle as the package")
    a = ap.parse_args()
    if a.selftest:
        print(run(oracle_as_package(), a.selftest))
    else:
        import ctc_segmentation as pkg
        print(run(pkg, os.path.join(HERE, "ctcseg_golden.npz")))
