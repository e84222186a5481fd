"""Window scorer, BASELINE configs[1]: graph-replay time and parity of each tier
(fp32 linear-domain -> fp64 linear-domain -> log-domain).  python tools/exp_alpha_tiers.py [--quick]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ipfa_b200 as ipfa
from cases import ctc_case
from oracle import ctc as octc
ops = ipfa.ops
dev = torch.device("cuda:0")
D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

def relerr(a, b):
    fin = np.isfinite(b)
    if not np.array_equal(np.isfinite(a), fin):
        return float("inf")
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-6))) if fin.any() else 0.0

TIERS = [("f32", {"IPFA_ALPHA_F32": "1"}), ("lin", {}), ("log", {"IPFA_ALPHA_LOG": "1"})]
# parity on assorted cases
cases = [(1, 64, 300, 100, 32, False, False, False), (2, 32, 500, 100, 32, True, True, True),
         (3, 40, 90, 50, 32, True, True, False), (4, 16, 1000, 100, 32, False, False, True),
         (5, 12, 700, 250, 32, True, False, True), (6, 64, 40, 10, 32, True, True, False),
         (7, 9, 400, 200, 29, True, True, True), (8, 33, 200, 0, 32, True, False, False)]
for seed, n, t, l, v, ragged, repeats, peaked in ([] if "--quick" in sys.argv else cases):
    lp, tg, il, tl = ctc_case(seed, n, t, max(l, 1), v, ragged=ragged, repeats=repeats, peaked=peaked)
    if l == 0:
        tl[:] = 0
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    line = f"case {seed} n={n} t={t} l={l} v={v} peaked={peaked}:"
    for name, env in TIERS:
        with ipfa.tuning(**env):
            got = ipfa.ctc_alpha_nll(D(lp), D(tg), D(il), D(tl)).cpu().numpy()
        buf = ops._ws_cache.get((dev.index, torch.cuda.current_stream(dev).cuda_stream))
        c = buf[4 * n:4 * n + 16].view(torch.int32).tolist()
        line += f"  {name} err {relerr(got, ref):.1e} redo(lin->log {c[0]}, f32->lin {c[2]} why {c[3]})"
    print(line, flush=True)

n, t, l, v = 1024, 1000, 100, 32
sets = []
for k in range(3):
    g = torch.Generator(device=dev).manual_seed(k)
    lp = torch.randn(n, t, v, generator=g, device=dev).log_softmax(-1)
    tg = torch.randint(1, v, (n, l), generator=g, device=dev, dtype=torch.int32)
    sets.append((lp, tg))
il = torch.full((n,), t, dtype=torch.int32, device=dev)
tl = torch.full((n,), l, dtype=torch.int32, device=dev)
ref0 = octc.ctc_alpha_nll(sets[0][0][:32].cpu().numpy(), sets[0][1][:32].cpu().numpy(), il[:32].cpu().numpy(), tl[:32].cpu().numpy())
for name, env in TIERS:
    with ipfa.tuning(**env):
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for i in range(3):
                out = ipfa.ctc_alpha_nll(*sets[i % 3], il, tl)
            side.synchronize()
            err = relerr(ipfa.ctc_alpha_nll(*sets[0], il, tl)[:32].cpu().numpy(), ref0)
            buf = ops._ws_cache.get((dev.index, side.cuda_stream))
            c = buf[4 * n:4 * n + 16].view(torch.int32).tolist()
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
        print(f"c2 {name}: graph replay {a.elapsed_time(b) / 60 * 1000:.1f} us per call; err {err:.1e}; "
              f"redo lin->log {c[0]}, f32->lin {c[2]} (why {c[3]})", flush=True)
