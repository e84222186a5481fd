"""Times the windowed table mode (audio longer than 8000 frames) on one window."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ipfa_b200 as ipfa

for t_len, n_tok, n_utts in ((9000, 400, 8), (30000, 3000, 40)):
    rng = np.random.default_rng(0)
    v = 32
    flat, ub = [-1], []
    per = n_tok // n_utts
    for u in range(n_utts):
        flat.append(0); ub.append(len(flat) - 1)
        flat += rng.integers(1, v, per).tolist()
    flat.append(0); ub.append(len(flat) - 1)
    gt = np.asarray(flat, np.int32)
    lp = rng.standard_normal((t_len, v)).astype(np.float32)
    pos = np.sort(rng.permutation(t_len)[:len(gt) - 1])
    for j, (a, b) in enumerate(zip(pos, list(pos[1:]) + [t_len])):
        lp[a:b, flat[j + 1]] += 5.0
    lp = torch.log_softmax(torch.from_numpy(lp).cuda(), -1)[None].contiguous()
    args = (lp, [t_len], gt[None], [len(gt)], np.asarray(ub, np.int32)[None], [n_utts], 0.02)
    for all_prefixes in (False, True):
        flags = 2 | (8 if all_prefixes else 0)
        w = 8000
        while True:
            res = ipfa.ctcseg_align(*args, flags=flags, window=w, details=False)
            if not int(res.status[0]) & 8:
                break
            w *= 2
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = ipfa.ctcseg_align(*args, flags=flags, window=w, details=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"T={t_len} columns={len(gt)} utterances={n_utts} all_prefixes={all_prefixes} window={w}: "
              f"{dt * 1e3:.1f} ms ({dt / len(gt) * 1e6:.1f} us per column of the full text)")
