"""Generates tests/golden/ctc_golden.npz: outputs of the INSTALLED comparators
(torch.nn.functional.ctc_loss, torchaudio.functional.forced_align; the "reference
CPU path" BASELINE.json's north_star names) on
seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The oracle (oracle/ctc_oracle.c) and the CUDA path are both checked against the
committed file, so parity does not depend on the libraries being importable at
test time.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F
import torchaudio.functional as AF

HERE = os.path.dirname(os.path.abspath(__file__))


def rand_case(seed, n, t, l, v, ragged=False, repeats=False, peaked=False):
    g = torch.Generator().manual_seed(seed)
    lp = torch.randn(n, t, v, generator=g)
    targets = torch.randint(1, v, (n, max(l, 1)), generator=g, dtype=torch.int64)
    if l == 0:
        targets = targets[:, :0]
    if repeats and l > 1:
        m = torch.rand(n, l - 1, generator=g) < 0.35
        for i in range(1, l):
            targets[:, i] = torch.where(m[:, i - 1], targets[:, i - 1], targets[:, i])
    if ragged:
        in_len = torch.randint(max(t // 2, 1), t + 1, (n,), generator=g)
        tgt_len = torch.randint(max(l // 2, 0), l + 1, (n,), generator=g)
    else:
        in_len = torch.full((n,), t, dtype=torch.int64)
        tgt_len = torch.full((n,), l, dtype=torch.int64)
    if peaked:
        # a plausible monotone alignment gets +6 on its token, so paths are tie-free
        for i in range(n):
            ti, li = int(in_len[i]), int(tgt_len[i])
            if li == 0:
                continue
            pos = torch.sort(torch.randperm(ti, generator=g)[:min(li, ti)]).values
            for j, p in enumerate(pos.tolist()):
                lp[i, p, int(targets[i, j])] += 6.0
            lp[i, :, 0] += 1.0
    lp = lp.log_softmax(-1)
    return lp, targets, in_len, tgt_len


def run_ctc_loss(lp, targets, in_len, tgt_len):
    return F.ctc_loss(lp.transpose(0, 1).contiguous(), targets, in_len, tgt_len, blank=0,
                      reduction="none", zero_infinity=False)


def run_forced_align(lp, targets, in_len, tgt_len):
    n, t, _ = lp.shape
    paths = np.full((n, t), -1, dtype=np.int32)
    scores = np.zeros((n, t), dtype=np.float32)
    status = np.zeros(n, dtype=np.int32)
    for i in range(n):
        ti, li = int(in_len[i]), int(tgt_len[i])
        try:
            p, s = AF.forced_align(lp[i:i + 1, :ti].contiguous(), targets[i:i + 1, :li].contiguous(),
                                   blank=0)
            paths[i, :ti] = p[0].numpy()
            scores[i, :ti] = s[0].numpy()
        except RuntimeError:
            status[i] = 1
    return paths, scores, status


def main():
    out = {}
    specs = {
        # name: (seed, n, t, l, v, ragged, repeats, peaked)
        "small": (1, 6, 12, 4, 5, False, False, False),
        "repeats": (2, 8, 20, 7, 4, True, True, False),
        "ragged": (3, 8, 50, 12, 32, True, True, False),
        "tight": (4, 6, 9, 8, 6, False, False, False),       # T barely >= L (+repeats may fail)
        "peaked": (5, 6, 80, 20, 32, True, False, True),
        "c2_like": (6, 3, 1000, 100, 32, False, False, False),
        "c4_like": (7, 2, 400, 120, 300, True, True, True),
        "c3_like": (8, 16, 500, 40, 32, True, True, True),
    }
    for name, (seed, n, t, l, v, ragged, repeats, peaked) in specs.items():
        lp, targets, in_len, tgt_len = rand_case(seed, n, t, l, v, ragged, repeats, peaked)
        nll = run_ctc_loss(lp, targets, in_len, tgt_len).numpy()
        paths, scores, status = run_forced_align(lp, targets, in_len, tgt_len)
        out[f"{name}/lp"] = lp.numpy()
        out[f"{name}/targets"] = targets.numpy().astype(np.int32)
        out[f"{name}/in_len"] = in_len.numpy().astype(np.int32)
        out[f"{name}/tgt_len"] = tgt_len.numpy().astype(np.int32)
        out[f"{name}/nll"] = nll
        out[f"{name}/paths"] = paths
        out[f"{name}/scores"] = scores
        out[f"{name}/fa_status"] = status

    # empty target (L = 0): ctc_loss only (forced_align rejects empty targets)
    lp, _, in_len, _ = rand_case(9, 3, 10, 0, 5)
    targets = torch.zeros(3, 0, dtype=torch.int64)
    tgt_len = torch.zeros(3, dtype=torch.int64)
    out["empty/lp"] = lp.numpy()
    out["empty/in_len"] = in_len.numpy().astype(np.int32)
    out["empty/nll"] = run_ctc_loss(lp, targets, in_len, tgt_len).numpy()

    # exact-tie cases documented in SURVEY.md section 8(a) row A8
    ties = []
    # uniform emissions T=6, [1,2] -> [1,2,2,2,2,2]; T=7, [1,1] -> [1,0,1,1,1,1,1]
    for t, tg, v in [(6, [1, 2], 5), (7, [1, 1], 5), (5, [1, 2, 3], 4), (8, [2, 2, 3], 4)]:
        lp = torch.full((1, t, v), 1.0 / v).log()
        p, s = AF.forced_align(lp, torch.tensor([tg]), blank=0)
        ties.append((lp[0].numpy(), np.array(tg, np.int32), p[0].numpy().astype(np.int32),
                     s[0].numpy()))
    # s-1 / s-2 exact tie that beats stay: torchaudio falls to *stay*
    probs = torch.tensor([[.05, .9, .05], [.45, .45, .10], [.05, .05, .9], [.05, .05, .9]])
    lp = probs.log()[None]
    p, s = AF.forced_align(lp, torch.tensor([[1, 2]]), blank=0)
    ties.append((lp[0].numpy(), np.array([1, 2], np.int32), p[0].numpy().astype(np.int32),
                 s[0].numpy()))
    # final-state tie: last frame P(blank) == P(label)
    probs = torch.tensor([[.2, .6, .2], [.4, .4, .2], [.4, .4, .2]])
    lp = probs.log()[None]
    p, s = AF.forced_align(lp, torch.tensor([[1]]), blank=0)
    ties.append((lp[0].numpy(), np.array([1], np.int32), p[0].numpy().astype(np.int32),
                 s[0].numpy()))
    out["ties/count"] = np.array(len(ties))
    for i, (lp_i, tg_i, p_i, s_i) in enumerate(ties):
        out[f"ties/{i}/lp"] = lp_i
        out[f"ties/{i}/targets"] = tg_i
        out[f"ties/{i}/paths"] = p_i
        out[f"ties/{i}/scores"] = s_i

    path = os.path.join(HERE, "ctc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          "torch", torch.__version__)


if __name__ == "__main__":
    main()
