"""Sweeps the (units-per-thread, warps-per-window) lattice instances on the bench workloads.
Run on the GPU box:  python tools/tune_shapes.py [c2 c2v c3 c4]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import ipfa_b200 as ipfa  # noqa: E402

SHAPES = [(p, w) for p in (1, 2, 4) for w in (1, 2, 4, 8, 16)] + [(8, 8), (8, 16), (8, 32)]


def time_fn(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def main():
    names = sys.argv[1:] or ["c2", "c2v", "c3", "c4"]
    dev = torch.device("cuda:0")
    for name in names:
        kind, n, t, l, v, ragged = bench.WORKLOADS[name]
        sets = [bench.make_inputs(name, s, device=dev) for s in range(2 if n * t * v * 4 < 1e9 else 1)]
        env = "IPFA_ALPHA_SHAPE" if kind == "alpha" else "IPFA_VITERBI_SHAPE"
        i = [0]

        def fn():
            lp, tg, il, tl = sets[i[0] % len(sets)]
            i[0] += 1
            if kind == "alpha":
                ipfa.ctc_alpha_nll(lp, tg, il, tl)
            else:
                ipfa.ctc_forced_align(lp, tg, il, tl, tokens=False)

        os.environ.pop(env, None)
        base = time_fn(fn, 10)
        print(f"{name}: default shape {base:.3f} ms", flush=True)
        for p, w in SHAPES:
            if 32 * w * p < l + 1 or 32 * w * p > 8 * (l + 1) + 64:
                continue
            os.environ[env] = f"{p},{w}"
            try:
                ms = time_fn(fn, 10)
                print(f"  {name} shape P={p} W={w}: {ms:.3f} ms", flush=True)
            except Exception as exc:
                print(f"  {name} shape P={p} W={w}: {exc}", flush=True)
        os.environ.pop(env, None)
        del sets
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
