"""Sweeps the (units-per-thread, warps-per-window) lattice instances on the bench workloads.
Run on the GPU box:  python tools/tune_shapes.py [c2 c2v c3 c4]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import ipfa_b200 as ipfa  # noqa: E402

SHAPES = [(p, w) for p in (1, 2, 4) for w in (1, 2, 4, 8, 16)] + [(8, 1), (8, 2), (8, 4), (8, 8), (8, 16), (8, 32)]


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
        wl = bench.WORKLOADS[name]
        kind = wl.kind
        l = wl.cols - 1 if kind == "seg" else wl.l
        sets = [wl.make(s, device=dev) for s in range(2 if wl.set_bytes < 1e9 else 1)]
        env = {"alpha": "IPFA_ALPHA_SHAPE", "viterbi": "IPFA_VITERBI_SHAPE", "seg": "IPFA_SEG_SHAPE"}[kind]
        i = [0]

        def fn():
            i[0] += 1
            wl.step(ipfa, sets[i[0] % len(sets)])

        os.environ.pop(env, None)
        ipfa.ops.lib().ipfa_tuning_reload()  # the library reads its switches once per process
        base = time_fn(fn, 10)
        print(f"{name}: default shape {base:.3f} ms", flush=True)
        for p, w in SHAPES:
            if 32 * w * p < l + 1 or 32 * w * p > 8 * (l + 1) + 64:
                continue
            os.environ[env] = f"{p},{w}"
            ipfa.ops.lib().ipfa_tuning_reload()
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
