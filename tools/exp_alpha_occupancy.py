"""Window scorer (fp64 linear-domain instance): time against the number of half-window chains per SM
sub-partition -- N = 148*4*k/2 windows put exactly k chains on every sub-partition.
python tools/exp_alpha_occupancy.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ipfa_b200 as ipfa
dev = torch.device("cuda:0")
t, l, v = 1000, 100, 32
for n in (296, 592, 888, 1024, 1184, 1480, 1776):
    sets = []
    for k in range(3):
        g = torch.Generator(device=dev).manual_seed(k)
        sets.append((torch.randn(n, t, v, generator=g, device=dev).log_softmax(-1),
                     torch.randint(1, v, (n, l), generator=g, device=dev, dtype=torch.int32)))
    il = torch.full((n,), t, dtype=torch.int32, device=dev)
    tl = torch.full((n,), l, dtype=torch.int32, device=dev)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(3):
            ipfa.ctc_alpha_nll(*sets[i], il, tl)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            outs = [ipfa.ctc_alpha_nll(*sets[i], il, tl) for i in range(3)]
        for _ in range(3):
            g.replay()
        side.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(side)
        for _ in range(20):
            g.replay()
        b.record(side)
        side.synchronize()
    us = a.elapsed_time(b) / 60 * 1000
    chains = 2 * n / 592
    print(f"N={n:5d}  chains per sub-partition {chains:4.2f}  {us:7.1f} us per call  {us / chains:6.1f} us per chain-slot", flush=True)
    del sets
