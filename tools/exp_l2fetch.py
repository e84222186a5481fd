"""Experiment: L2 fetch granularity (cuCtxSetLimit CU_LIMIT_MAX_L2_FETCH_GRANULARITY) vs the
V = 5000 gather panel of the alpha kernel (BASELINE configs[3] shapes).
    python tools/exp_l2fetch.py [granularity ...]"""
import ctypes
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ipfa = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200")

cu = ctypes.CDLL("libcuda.so.1")
CU_LIMIT_MAX_L2_FETCH_GRANULARITY = 0x05


def set_gran(g):
    rc = cu.cuCtxSetLimit(CU_LIMIT_MAX_L2_FETCH_GRANULARITY, ctypes.c_size_t(g))
    v = ctypes.c_size_t(0)
    cu.cuCtxGetLimit(ctypes.byref(v), CU_LIMIT_MAX_L2_FETCH_GRANULARITY)
    return rc, v.value


def main():
    n, t, l, v = 256, 3000, 400, 5000
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    lp = torch.log_softmax(torch.randn(n, t, v, generator=g, device=dev), dim=-1)
    tg = torch.randint(1, v, (n, l), generator=g, device=dev, dtype=torch.int32)
    il = torch.full((n,), t, dtype=torch.int32, device=dev)
    tl = torch.full((n,), l, dtype=torch.int32, device=dev)
    ref = None
    grans = [int(a) for a in sys.argv[1:]] or [0, 32, 64, 128]
    for gran in grans:
        if gran:
            print("set", gran, set_gran(gran))
        else:
            v0 = ctypes.c_size_t(0)
            cu.cuCtxGetLimit(ctypes.byref(v0), CU_LIMIT_MAX_L2_FETCH_GRANULARITY)
            print("default granularity", v0.value)
        for _ in range(2):
            out = ipfa.ctc_alpha_nll(lp, tg, il, tl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = ipfa.ctc_alpha_nll(lp, tg, il, tl)
        e1.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        print(f"granularity {gran}: {e0.elapsed_time(e1) / 5:.3f} ms/launch, equal={bool(torch.equal(ref, out))}")


if __name__ == "__main__":
    main()
