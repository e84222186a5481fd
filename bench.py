#!/usr/bin/env python
"""bench.py -- throughput of the alignment hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|seg] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic emissions.
Default workload = BASELINE.json configs[1] ("c2"): batched CTC-loss window
scoring, 1024 candidate windows x T=1000 frames x L=100 labels, V=32, fp32.
Weak scaling: every rank scores its own 1024-window shard (windows are
independent; no data-path collective -- SURVEY.md section 8(e)).

value      whole-job aligned audio-hours/s (20 ms per frame, alignment_utils.py:87
           of the reference) with the emissions already resident in HBM
e2e        same metric through the host-buffer C-ABI call (ipfa_*_host): pinned
           host emissions -> H2D -> kernel -> D2H of the scores, every step
roofline   algorithmic bytes of the dominant kernel / its CUDA-event duration,
           against MEASURED_PEAKS.json's HBM copy bandwidth
cpu_baseline  the CPU restatement (oracle/, OpenMP over windows) and the installed
           torch CPU comparator on a bounded sample of the same workload

--impl reference times the CPU path only (rank 0; other ranks exit 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FRAME_SECONDS = 0.02

WORKLOADS = {
    # name: (kind, windows, T, L, V, ragged)
    "c2": ("alpha", 1024, 1000, 100, 32, False),
    "c3": ("viterbi", 65536, 500, 40, 32, True),
    "c4": ("alpha", 256, 3000, 400, 5000, False),
    "c2v": ("viterbi", 1024, 1000, 100, 32, False),
}
WORKLOAD_TEXT = {
    "c2": "BASELINE configs[1]: batched CTC-loss window scoring, 1024 windows x T=1000 x L=100, V=32, fp32",
    "c3": "BASELINE configs[2]: Viterbi forced align + backtrace, 65536 utterances, T<=500, L<=40, V=32",
    "c4": "BASELINE configs[3]: long-window large-vocab scoring, 256 windows x T=3000 x L=400, V=5000",
    "c2v": "configs[1] shapes through the Viterbi + backtrace path",
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, False, [], set()
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_inputs(workload, seed, device=None):
    """Synthetic log-softmax emissions + targets."""
    import torch
    kind, n, t, l, v, ragged = WORKLOADS[workload]
    g = torch.Generator(device=device or "cpu").manual_seed(seed)
    lp = torch.randn(n, t, v, generator=g, device=device or "cpu", dtype=torch.float32)
    lp = torch.log_softmax(lp, dim=-1)
    tg = torch.randint(1, v, (n, l), generator=g, device=device or "cpu", dtype=torch.int32)
    if ragged:
        il = torch.randint(t // 2, t + 1, (n,), generator=g, device=device or "cpu", dtype=torch.int32)
        tl = torch.randint(l // 2, l + 1, (n,), generator=g, device=device or "cpu", dtype=torch.int32)
    else:
        il = torch.full((n,), t, dtype=torch.int32, device=device or "cpu")
        tl = torch.full((n,), l, dtype=torch.int32, device=device or "cpu")
    return lp, tg, il, tl


def work_units(workload, il, tl):
    """(cells, audio-hours, algorithmic bytes) of one step of one rank."""
    kind, n, t, l, v, _ = WORKLOADS[workload]
    il = il.cpu().numpy().astype(np.int64)
    tl = tl.cpu().numpy().astype(np.int64)
    cells = int((il * (2 * tl + 1)).sum())
    hours = float(il.sum()) * FRAME_SECONDS / 3600.0
    read = int((il * np.minimum(tl + 1, v) * 4 + tl * 4).sum())
    if kind == "alpha":
        write = 4 * n
    else:  # 2-bit backpointers written once and read once + paths + frame scores
        write = int((2 * ((il * (2 * tl + 1) * 2 + 7) // 8) + il * 8).sum())
    return cells, hours, read + write


# ----------------------------------------------------------------------------- CPU legs
def cpu_leg(workload, budget_s, steps=1, warmup=0):
    """Times the CPU restatement (oracle/) and torch's CPU comparator on a bounded
    sample of the workload.  Returns a dict with audio-h/s figures."""
    import torch
    from oracle import ctc as octc
    kind, n, t, l, v, ragged = WORKLOADS[workload]
    lp, tg, il, tl = make_inputs(workload, 1234) if n * t * v <= (1 << 26) else (None,) * 4
    if lp is None:  # keep host memory bounded for the big workloads
        g = torch.Generator().manual_seed(1234)
        ns = max(8, (1 << 26) // (t * v))
        lp = torch.log_softmax(torch.randn(ns, t, v, generator=g), dim=-1)
        tg = torch.randint(1, v, (ns, l), generator=g, dtype=torch.int32)
        il = torch.full((ns,), t, dtype=torch.int32)
        tl = torch.full((ns,), l, dtype=torch.int32)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lp_np, tg_np, il_np, tl_np = lp.numpy(), tg.numpy(), il.numpy(), tl.numpy()

    def run_oracle(k):
        if kind == "alpha":
            octc.ctc_alpha_nll(lp_np[:k], tg_np[:k], il_np[:k], tl_np[:k])
        else:
            octc.ctc_viterbi(lp_np[:k], tg_np[:k], il_np[:k], tl_np[:k])

    def run_torch(k):
        if kind == "alpha":
            torch.nn.functional.ctc_loss(lp[:k].transpose(0, 1), tg[:k].long(), il[:k].long(), tl[:k].long(),
                                         blank=0, reduction="none")
        else:
            import torchaudio.functional as AF
            for i in range(k):
                AF.forced_align(lp[i:i + 1, :int(il[i])], tg[i:i + 1, :int(tl[i])].long(), blank=0)

    out = {}
    n_avail = lp.shape[0]
    for name, fn in (("oracle_port", run_oracle), ("torch_cpu", run_torch)):
        try:
            probe = min(n_avail, max(cores, 16))
            fn(probe)  # warm-up (thread pools, page faults)
            t0 = time.perf_counter()
            fn(probe)
            per_window = (time.perf_counter() - t0) / probe
            total_steps = max(steps + warmup, 1)
            k = int(min(n_avail, max(probe, budget_s / total_steps / max(per_window, 1e-9))))
            for _ in range(warmup):
                fn(k)
            t0 = time.perf_counter()
            for _ in range(max(steps, 1)):
                fn(k)
            dt = (time.perf_counter() - t0) / max(steps, 1)
            hours = float(il_np[:k].astype(np.int64).sum()) * FRAME_SECONDS / 3600.0
            cells = float((il_np[:k].astype(np.int64) * (2 * tl_np[:k].astype(np.int64) + 1)).sum())
            out[name] = {"audio_h_per_s": hours / dt, "cells_per_s": cells / dt, "windows": k,
                         "seconds_per_step": dt}
        except Exception as exc:  # torchaudio may be absent on some box
            out[name] = {"error": repr(exc)}
    out["cores"] = cores
    out["oracle_threads"] = octc.num_threads()
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    legs = cpu_leg(args.workload, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    best_name = max((k for k in ("oracle_port", "torch_cpu") if "audio_h_per_s" in legs[k]),
                    key=lambda k: legs[k]["audio_h_per_s"])
    best = legs[best_name]
    kind = WORKLOADS[args.workload][0]
    line = {
        "impl": "reference", "metric": "aligned_audio_hours_per_s", "value": best["audio_h_per_s"],
        "unit": "audio-h/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": best["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "kernel": kind},
        "cells_per_s": best["cells_per_s"],
        "cpu_baseline": {
            "value": best["audio_h_per_s"], "unit": "audio-h/s", "cores": legs["cores"],
            "kind": "port" if best_name == "oracle_port" else "reference",
            "sample": f"{best['windows']} windows of the workload per step; faster of the OpenMP C "
                      f"restatement (oracle/) and the installed torch/torchaudio CPU comparator "
                      f"north_star names -- here: {best_name}",
            "oracle_port": legs["oracle_port"], "torch_cpu": legs["torch_cpu"]},
        "e2e": {"value": best["audio_h_per_s"], "unit": "audio-h/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import ipfa_b200 as ipfa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kind, n, t, l, v, ragged = WORKLOADS[args.workload]
    # rotate over several distinct input sets so no step finds its emissions in L2
    set_bytes = n * t * v * 4
    n_sets = max(2, min(8, int(np.ceil(3 * 126e6 / set_bytes)))) if set_bytes < 2e9 else 1
    sets = [make_inputs(args.workload, 1000 * rank + s, device=dev) for s in range(n_sets)]
    cells, hours, alg_bytes = work_units(args.workload, sets[0][2], sets[0][3])

    def step(i):
        lp, tg, il, tl = sets[i % n_sets]
        if kind == "alpha":
            return ipfa.ctc_alpha_nll(lp, tg, il, tl)
        return ipfa.ctc_forced_align(lp, tg, il, tl, tokens=False).paths

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ipfa.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step(i)
    ev[1].record()
    barrier()
    elapsed_ms = ev[0].elapsed_time(ev[1])
    launches = ipfa.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join()

    # per-launch duration of the dominant kernel (events around single launches)
    k_ms = []
    for i in range(min(args.steps, 50)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(i)
        b.record()
        b.synchronize()
        k_ms.append(a.elapsed_time(b))
    kernel_ms = float(np.median(k_ms))

    # end-to-end through the host-buffer C ABI: pinned host emissions, H2D inside the timed region
    lp, tg, il, tl = sets[0]
    h_sets = []
    for s in range(min(n_sets, 2)):
        lp_s, tg_s, il_s, tl_s = sets[s]
        h_sets.append((lp_s.cpu().pin_memory(), tg_s.cpu().pin_memory(), il_s.cpu().pin_memory(),
                       tl_s.cpu().pin_memory()))
    out_host = torch.empty(n, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        hl, ht, hi, htl = h_sets[i % len(h_sets)]
        if kind == "alpha":
            return ipfa.ctc_alpha_nll_host(hl.numpy(), ht.numpy(), hi.numpy(), htl.numpy(), out=out_host.numpy())
        return ipfa.ctc_forced_align_host(hl.numpy(), ht.numpy(), hi.numpy(), htl.numpy(), tokens=False)

    e2e_steps = max(3, min(args.steps, 20))
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d = int(set_bytes + tg.numel() * 4 + 8 * n)
    d2h = int(4 * n if kind == "alpha" else (8 * n * t + 8 * n))

    times = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(times[0]), float(times[1])

    if rank == 0:
        peak, peak_src = peaks()
        ms_per_step = elapsed_ms / args.steps
        value = world * hours / (ms_per_step * 1e-3)
        # one launch per step for the alpha workload: its average launch duration over the timed
        # region is ms_per_step (the GPU never idles: launches are issued ahead of execution)
        kernel_avg_ms = ms_per_step if world == 1 else float(np.mean(k_ms))
        achieved = alg_bytes / (kernel_avg_ms * 1e-3) / 1e9
        legs = cpu_leg(args.workload, budget_s=12.0)
        best_name = max((k for k in ("oracle_port", "torch_cpu") if "audio_h_per_s" in legs[k]),
                        key=lambda k: legs[k]["audio_h_per_s"])
        line = {
            "metric": "aligned_audio_hours_per_s", "value": value, "unit": "audio-h/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload], "kernel": kind,
                       "windows_per_gpu": n, "T": t, "L": l, "V": v,
                       "l2": f"{n_sets} rotating input sets of {set_bytes / 1e6:.0f} MB per GPU (> 126 MB L2)",
                       "sharding": "independent windows per rank, no collective on the data path"},
            "cells_per_s": world * cells / (ms_per_step * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_avg_ms,
                         "single_launch_median_ms": kernel_ms,
                         "note": "T-serial log-sum-exp recursion: MUFU/latency bound, see DESIGN.md"},
            "cpu_baseline": {"value": legs[best_name]["audio_h_per_s"], "unit": "audio-h/s",
                             "cores": legs["cores"],
                             "kind": "port" if best_name == "oracle_port" else "reference",
                             "sample": f"{legs[best_name]['windows']} windows of the same workload, "
                                       f"faster of oracle port / torch CPU: {best_name}",
                             "oracle_port": legs["oracle_port"], "torch_cpu": legs["torch_cpu"]},
            "e2e": {"value": world * hours / (e2e_ms * 1e-3), "unit": "audio-h/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "api": "ipfa_ctc_alpha_host" if kind == "alpha" else "ipfa_ctc_viterbi_host"},
            "gpu_launches": int(launches),
            "clocks": sampler.result(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
