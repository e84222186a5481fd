"""Linear-domain (fp64) instance of the alpha kernel against the log-domain instance and the
oracle: parity over shapes, the redo path on emissions built to break the exactness guard, and
timing on BASELINE configs[1].  Run on a GPU box: python tools/exp_alpha_lin.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ipfa_b200 as ipfa  # noqa: E402
from cases import ctc_case  # noqa: E402
from oracle import ctc as octc  # noqa: E402
import importlib  # noqa: E402
ops = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.ops")

dev = torch.device("cuda:0")


def run(lp, tg, il, tl, log):
    if log:
        os.environ["IPFA_ALPHA_LOG"] = "1"
    else:
        os.environ.pop("IPFA_ALPHA_LOG", None)
    ops.lib().ipfa_tuning_reload()  # the library reads its switches once per process
    out = ipfa.ctc_alpha_nll(torch.from_numpy(lp).to(dev), tg, il, tl).cpu().numpy()
    redo = 0 if log else ops.ctc_alpha_redo_count(len(il), dev)
    if redo and "--why" in sys.argv:
        print("   redo reasons:", ops.ctc_alpha_redo_reasons(len(il), dev))
    os.environ.pop("IPFA_ALPHA_LOG", None)
    return out, redo


def relerr(a, b):
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "inf pattern differs"
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-3)))


TIMING_ONLY = "--timing" in sys.argv
worst = 0.0
cases = []
for seed, (n, t, l, v) in enumerate([(64, 200, 40, 32), (32, 1000, 100, 32), (16, 50, 1, 32), (8, 30, 0, 32),
                                     (16, 400, 31, 32), (16, 400, 32, 32), (16, 400, 63, 8), (8, 900, 127, 32),
                                     (8, 900, 128, 32), (4, 1200, 255, 64), (16, 300, 20, 5), (16, 17, 3, 32),
                                     (16, 15, 3, 32), (8, 600, 200, 48), (32, 257, 64, 33)]):
    for ragged, repeats, peaked in [(False, False, False), (True, True, False), (True, True, True)]:
        cases.append((seed, n, t, l, v, ragged, repeats, peaked))
if TIMING_ONLY:
    cases = []
FORCED = os.environ.get("IPFA_ALPHA_LIN_SHAPE")
for seed, n, t, l, v, ragged, repeats, peaked in cases:
    lp, tg, il, tl = ctc_case(seed, n, t, l, v, ragged=ragged, repeats=repeats, peaked=peaked)
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    lin, redo = run(lp, tg, il, tl, False)
    if not FORCED:  # the one-warp instances with the most pairs per lane as well
        for shape in ("8,1", "4,2"):
            os.environ["IPFA_ALPHA_LIN_SHAPE"] = shape
            lin2, redo2 = run(lp, tg, il, tl, False)
            os.environ.pop("IPFA_ALPHA_LIN_SHAPE")
            assert relerr(lin2, ref) < 1e-4 and redo2 == redo, (shape, relerr(lin2, ref), redo2, redo)
    log, _ = run(lp, tg, il, tl, True)
    e_lin, e_log = relerr(lin, ref), relerr(log, ref)
    worst = max(worst, e_lin)
    print(f"n={n} t={t} l={l} v={v} ragged={int(ragged)} rep={int(repeats)} peaked={int(peaked)}: "
          f"lin {e_lin:.2e} log {e_log:.2e} redo {redo}/{n}")
    assert e_lin < 1e-4, "linear-domain instance off"

if not TIMING_ONLY:
    # emissions that break the guard: very sharp (scaled logits), -inf entries, blank in the target
    rng = np.random.default_rng(7)
    for name, scale in [("sharp x8", 8.0), ("sharp x40", 40.0), ("sharp x200", 200.0)]:
        n, t, l, v = 32, 300, 30, 32
        raw = (rng.standard_normal((n, t, v)) * scale).astype(np.float32)
        lp = torch.from_numpy(raw).log_softmax(-1).numpy()
        tg = rng.integers(1, v, (n, l)).astype(np.int32)
        il, tl = np.full(n, t, np.int32), np.full(n, l, np.int32)
        ref = octc.ctc_alpha_nll(lp, tg, il, tl)
        lin, redo = run(lp, tg, il, tl, False)
        log, _ = run(lp, tg, il, tl, True)
        print(f"{name}: lin {relerr(lin, ref):.2e} log {relerr(log, ref):.2e} redo {redo}/{n}")
        assert relerr(lin, ref) < 1e-4
    lp, tg, il, tl = ctc_case(3, 16, 120, 12, 32)
    lp[::2, 5:9, 3] = -np.inf
    lp[1, :, 0] = -np.inf
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    lin, redo = run(lp, tg, il, tl, False)
    print(f"-inf entries: lin {relerr(lin, ref):.2e} redo {redo}/16", lin[:4], ref[:4])
    assert relerr(lin, ref) < 1e-4
    lp, tg, il, tl = ctc_case(4, 16, 120, 12, 32)
    tg[::3, 4] = 0
    lin, redo = run(lp, tg, il, tl, False)
    log, _ = run(lp, tg, il, tl, True)
    print(f"blank in target: lin-vs-log {relerr(lin, log):.2e} redo {redo}/16")
    assert relerr(lin, log) < 1e-6
    # infeasible: T < L + repeats
    lp, tg, il, tl = ctc_case(5, 8, 40, 30, 32, repeats=True)
    il[:] = 33
    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
    lin, redo = run(lp, tg, il, tl, False)
    print(f"infeasible: redo {redo}/8", lin, ref)
    assert relerr(lin, ref) < 1e-4

# timing, BASELINE configs[1]
n, t, l, v = 1024, 1000, 100, 32
sets = []
for k in range(3):
    g = torch.Generator(device=dev).manual_seed(k)
    lp = torch.randn(n, t, v, generator=g, device=dev).log_softmax(-1)
    tg = torch.randint(1, v, (n, l), generator=g, device=dev, dtype=torch.int32)
    sets.append((lp, tg))
il = torch.full((n,), t, dtype=torch.int32, device=dev)
tl = torch.full((n,), l, dtype=torch.int32, device=dev)
for label, env in [("log", {"IPFA_ALPHA_LOG": "1"}), ("lin 4,1", {"IPFA_ALPHA_LIN_SHAPE": "4,1"}),
                   ("lin 2,2", {"IPFA_ALPHA_LIN_SHAPE": "2,2"}), ("lin", {})]:
    os.environ.pop("IPFA_ALPHA_LOG", None)
    os.environ.pop("IPFA_ALPHA_LIN_SHAPE", None)
    os.environ.update(env)
    ops.lib().ipfa_tuning_reload()
    for i in range(6):
        out = ipfa.ctc_alpha_nll(*sets[i % 3], il, tl)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30):
        out = ipfa.ctc_alpha_nll(*sets[i % 3], il, tl)
    b.record()
    torch.cuda.synchronize()
    print(f"c2 {label}: {a.elapsed_time(b) / 30 * 1000:.1f} us per call; nll[0]={float(out[0]):.4f}"
          f" redo={ops.ctc_alpha_redo_count(n, dev) if label != 'log' else '-'}")
    # host time per call (no synchronisation inside the loop)
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        out = ipfa.ctc_alpha_nll(*sets[i % 3], il, tl)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"c2 {label}: host {1e6 * (t1 - t0) / 200:.1f} us per call issued, {1e6 * (t2 - t0) / 200:.1f} us per call drained")
    # device time: the three calls captured in one CUDA graph, replayed
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(3):
            ipfa.ctc_alpha_nll(*sets[i % 3], il, tl)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            outs = [ipfa.ctc_alpha_nll(*sets[i], il, tl) for i in range(3)]
        for _ in range(3):
            g.replay()
        side.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(side)
        for _ in range(10):
            g.replay()
        b.record(side)
        side.synchronize()
    print(f"c2 {label}: graph replay {a.elapsed_time(b) / 30 * 1000:.1f} us per call; nll[0]={float(outs[0][0]):.4f}")
os.environ.pop("IPFA_ALPHA_LOG", None)
os.environ.pop("IPFA_ALPHA_LIN_SHAPE", None)
print("worst lin rel err", worst)
