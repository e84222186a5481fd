#!/usr/bin/env python
"""bench.py -- throughput of the alignment hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2v|c3|c4|seg|c5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic emissions.
Default workload = BASELINE.json configs[1] ("c2"): batched CTC-loss window
scoring, 1024 candidate windows x T=1000 frames x L=100 labels, V=32, fp32.
Weak scaling: every rank processes its own shard of independent windows (no
data-path collective -- SURVEY.md section 8(e)); after the timed region the
per-window results are gathered once over NCCL and checked.

value      whole-job aligned audio-hours/s (20 ms per frame, alignment_utils.py:87
           of the reference) with the emissions already resident in HBM
e2e        same metric through the host-buffer C-ABI call (ipfa_*_host): pinned
           host emissions -> H2D -> kernels -> D2H of the results, every step
roofline   algorithmic bytes of the step's kernels / their CUDA-event duration,
           against MEASURED_PEAKS.json's HBM copy bandwidth
cpu_baseline  the CPU restatement (oracle/) and, where it exists, the installed
           torch / torchaudio CPU comparator, on a bounded sample of the workload,
           using every host core

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


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload):
    """DRAM bytes per launch of the workload's dominant kernel from the committed `ncu --set full`
    capture (profiles/traffic_r01.json, written by tools/summarize_ncu.py); None when absent."""
    path = os.path.join(ROOT, "profiles", "traffic_r01.json")
    try:
        with open(path) as f:
            return json.load(f).get(workload, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


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
            time.sleep(0.002)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- workloads
class CtcWorkload:
    """kind 'alpha' (kernel 1) or 'viterbi' (kernel 2a) on the 2L+1 lattice."""

    def __init__(self, name, kind, n, t, l, v, ragged, text, vocab_major=False):
        self.name, self.kind, self.n, self.t, self.l, self.v, self.ragged, self.text = \
            name, kind, n, t, l, v, ragged, text
        self.vocab_major = vocab_major  # emissions stored [N, V, T] (viewed as [N, T, V] with stride_v = T)
        self.api = "ipfa_ctc_alpha_host" if kind == "alpha" else "ipfa_ctc_viterbi_host"
        self.set_bytes = n * t * v * 4
        self.shape = {"windows_per_gpu": n, "T": t, "L": l, "V": v}

    def make(self, seed, device=None, n=None):
        import torch
        n = n or self.n
        dev = device or "cpu"
        g = torch.Generator(device=dev).manual_seed(seed)
        lp = torch.log_softmax(torch.randn(n, self.t, self.v, generator=g, device=dev), dim=-1)
        if self.vocab_major:
            lp = lp.transpose(1, 2).contiguous().transpose(1, 2)  # same values, vocabulary-major storage
        tg = torch.randint(1, self.v, (n, self.l), generator=g, device=dev, dtype=torch.int32)
        if self.ragged:
            il = torch.randint(self.t // 2, self.t + 1, (n,), generator=g, device=dev, dtype=torch.int32)
            tl = torch.randint(self.l // 2, self.l + 1, (n,), generator=g, device=dev, dtype=torch.int32)
        else:
            il = torch.full((n,), self.t, dtype=torch.int32, device=dev)
            tl = torch.full((n,), self.l, dtype=torch.int32, device=dev)
        return lp, tg, il, tl

    def units(self, inputs):
        il = inputs[2].cpu().numpy().astype(np.int64)
        tl = inputs[3].cpu().numpy().astype(np.int64)
        cells = int((il * (2 * tl + 1)).sum())
        hours = float(il.sum()) * FRAME_SECONDS / 3600.0
        read = int((il * np.minimum(tl + 1, self.v) * 4 + tl * 4).sum())
        if self.kind == "alpha":
            write = 4 * len(il)
        else:  # backpointer planes, 1.5 bits per state and frame, written once and read once + paths + frame scores
            write = int((2 * ((il * (2 * tl + 1) * 3 + 15) // 16) + il * 8).sum())
        return cells, hours, read + write

    def step(self, ipfa, inputs):
        lp, tg, il, tl = inputs
        if self.kind == "alpha":
            return ipfa.ctc_alpha_nll(lp, tg, il, tl)
        return ipfa.ctc_forced_align(lp, tg, il, tl, tokens=False).total

    def to_host(self, inputs):
        import torch
        # (the host-buffer entry point takes [N, T, V] rows; a vocabulary-major set is laid out that way for it)
        host = [x.contiguous().cpu().pin_memory().numpy() for x in inputs]
        n, t = self.n, self.t
        if self.kind == "alpha":
            out = {"out": torch.empty(n, dtype=torch.float32).pin_memory().numpy()}
        else:
            out = {k: torch.empty(shape, dtype=dt).pin_memory().numpy() for k, shape, dt in
                   (("paths", (n, t), torch.int32), ("scores", (n, t), torch.float32),
                    ("total", (n,), torch.float32), ("status", (n,), torch.int32))}
        return host, out

    def e2e_step(self, ipfa, host, out):
        lp, tg, il, tl = host
        if self.kind == "alpha":
            return ipfa.ctc_alpha_nll_host(lp, tg, il, tl, out=out["out"])
        return ipfa.ctc_forced_align_host(lp, tg, il, tl, tokens=False, out=out)

    def traffic(self):
        h2d = int(self.set_bytes + self.n * self.l * 4 + 8 * self.n)
        d2h = int(4 * self.n if self.kind == "alpha" else (8 * self.n * self.t + 8 * self.n))
        return h2d, d2h

    # CPU legs: name -> callable(k windows); sample arrays prepared once
    def cpu_setup(self):
        import torch
        cap = max(8, min(self.n, (1 << 26) // (self.t * self.v)))
        lp, tg, il, tl = self.make(1234, n=cap)
        self._cpu = (lp, tg, il, tl, lp.numpy(), tg.numpy(), il.numpy(), tl.numpy())
        torch.set_num_threads(os.cpu_count() or 1)
        return cap

    def cpu_legs(self):
        import torch
        from oracle import ctc as octc
        lp, tg, il, tl, lp_np, tg_np, il_np, tl_np = self._cpu

        def run_oracle(k):
            if self.kind == "alpha":
                octc.ctc_alpha_nll(lp_np[:k], tg_np[:k], il_np[:k], tl_np[:k])
            else:
                octc.ctc_viterbi(lp_np[:k], tg_np[:k], il_np[:k], tl_np[:k])

        def run_torch(k):
            if self.kind == "alpha":
                torch.nn.functional.ctc_loss(lp[:k].transpose(0, 1), tg[:k].long(), il[:k].long(),
                                             tl[:k].long(), blank=0, reduction="none")
            else:
                import torchaudio.functional as AF
                for i in range(k):
                    AF.forced_align(lp[i:i + 1, :int(il[i])], tg[i:i + 1, :int(tl[i])].long(), blank=0)

        return {"oracle_port": run_oracle, "torch_cpu": run_torch}

    def cpu_units(self, k):
        il, tl = self._cpu[6][:k].astype(np.int64), self._cpu[7][:k].astype(np.int64)
        return float((il * (2 * tl + 1)).sum()), float(il.sum()) * FRAME_SECONDS / 3600.0


def _seg_cpu_worker(args):
    """One window through the reference's algorithm on the CPU (fill in C, backtrace and
    scoring interpreted, like ctc-segmentation itself)."""
    from oracle import ctcseg as oseg
    lp, gt, ub = args
    cfg = oseg.CtcSegmentationParameters(index_duration=FRAME_SECONDS)
    oseg.get_segments(cfg, lp, gt.reshape(-1, 1).astype(np.int64), ub.tolist(), [""] * (len(ub) - 1))
    return 0


class SegWorkload:
    """kind 'seg': one anchor-loop iteration for N files in flight (BASELINE configs[4] unit of
    work): CTC-segmentation fill + every-prefix backtrace/scoring + on-device selection on
    70 s windows (T = 3500 frames, the reference's max_window_size) with K utterances."""

    kind = "seg"
    api = "ipfa_ctcseg_host"

    def __init__(self, name, n, t, k, tokens, v, text, all_prefixes=True):
        self.name, self.n, self.t, self.k, self.tokens, self.v, self.text = name, n, t, k, tokens, v, text
        self.all_prefixes = all_prefixes  # False: only the full text is aligned (word level / search on speech)
        self.set_bytes = n * t * v * 4
        self.cols = 2 + k * (tokens + 1)
        self.shape = {"windows_per_gpu": n, "T": t, "K": k, "columns": self.cols, "V": v}

    def make(self, seed, device=None, n=None):
        import torch
        n = n or self.n
        dev = device or "cpu"
        g = torch.Generator(device=dev).manual_seed(seed)
        lp = torch.log_softmax(torch.randn(n, self.t, self.v, generator=g, device=dev), dim=-1)
        toks = torch.randint(1, self.v, (n, self.k, self.tokens), generator=g, device=dev, dtype=torch.int32)
        gt = torch.zeros((n, self.cols), dtype=torch.int32, device=dev)
        gt[:, 0] = -1
        ub = torch.zeros((n, self.k + 1), dtype=torch.int32, device=dev)
        for u in range(self.k):  # -1, (blank, tokens)*, blank
            b = 1 + u * (self.tokens + 1)
            ub[:, u] = b
            gt[:, b + 1:b + 1 + self.tokens] = toks[:, u]
        ub[:, self.k] = self.cols - 1
        il = torch.full((n,), self.t, dtype=torch.int32, device=dev)
        nc = torch.full((n,), self.cols, dtype=torch.int32, device=dev)
        nu = torch.full((n,), self.k, dtype=torch.int32, device=dev)
        tlen = torch.full((n, self.k), 60, dtype=torch.int32, device=dev)
        last = torch.zeros(n, dtype=torch.int32, device=dev)
        return lp, il, gt, nc, ub, nu, tlen, last

    def units(self, inputs):
        cells = int(self.n) * self.t * self.cols
        hours = self.n * self.t * FRAME_SECONDS / 3600.0
        read = self.n * self.t * min(self.cols, self.v) * 4 + self.n * self.cols * 4
        # 1-bit transition marks written once; the backtrace hops from switch to switch and reads only the
        # blocks it walks (profiles/r01_seg.txt: 191 MB of DRAM traffic per step), so its reads are not charged
        bp = (self.n * self.t * self.cols + 7) // 8
        out = self.n * self.k * self.k * 24 + self.n * self.k * 4
        return cells, hours, read + bp + out

    def step(self, ipfa, inputs):
        lp, il, gt, nc, ub, nu, tlen, last = inputs
        if not self.all_prefixes:
            return ipfa.ctcseg_align(lp, il, gt, nc, ub, nu, FRAME_SECONDS, flags=2, details=False).seg
        res = ipfa.ctcseg_align(lp, il, gt, nc, ub, nu, FRAME_SECONDS, flags=2 | 8, details=False)
        dec, anchor = ipfa.anchor_select(res.seg, nu, tlen, last)
        return dec

    def to_host(self, inputs):
        host = [x.cpu().pin_memory().numpy() for x in inputs[:6]]
        return host, {}

    def e2e_step(self, ipfa, host, out):
        lp, il, gt, nc, ub, nu = host
        return ipfa.ctcseg_align_host(lp, il, gt, nc, ub, nu, FRAME_SECONDS, flags=2 | (8 if self.all_prefixes else 0),
                                      details=False)

    def traffic(self):
        return int(self.set_bytes + self.n * (self.cols + self.k + 4) * 4), int(self.n * self.k * self.k * 24 +
                                                                              self.n * (self.k + 1) * 4)

    def cpu_setup(self):
        cap = min(self.n, 4 * (os.cpu_count() or 1))
        lp, il, gt, nc, ub, nu, _, _ = self.make(1234, n=cap)
        self._cpu = (lp.numpy(), gt.numpy(), ub.numpy())
        return cap

    def cpu_legs(self):
        import multiprocessing as mp
        lp, gt, ub = self._cpu
        cores = os.cpu_count() or 1
        pool = mp.get_context("fork").Pool(cores)
        self._pool = pool

        def run_oracle(k):
            pool.map(_seg_cpu_worker, [(lp[i], gt[i], ub[i]) for i in range(k)], chunksize=1)

        return {"oracle_port": run_oracle}

    def cpu_units(self, k):
        return float(k) * self.t * self.cols, k * self.t * FRAME_SECONDS / 3600.0


WORKLOADS = {
    "c2": CtcWorkload("c2", "alpha", 1024, 1000, 100, 32, False,
                      "BASELINE configs[1]: batched CTC-loss window scoring, 1024 windows x T=1000 x L=100, V=32, fp32"),
    "c2v": CtcWorkload("c2v", "viterbi", 1024, 1000, 100, 32, False,
                       "configs[1] shapes through the Viterbi + backtrace path"),
    "c3": CtcWorkload("c3", "viterbi", 65536, 500, 40, 32, True,
                      "BASELINE configs[2]: Viterbi forced align + backtrace, 65536 utterances, T<=500, L<=40, V=32"),
    "c4": CtcWorkload("c4", "alpha", 256, 3000, 400, 5000, False,
                      "BASELINE configs[3]: long-window large-vocab scoring, 256 windows x T=3000 x L=400, V=5000"),
    "c4v": CtcWorkload("c4v", "alpha", 256, 3000, 400, 5000, False,
                       "BASELINE configs[3] with vocabulary-major emissions ([N, V, T], stride_v = T): 256 windows x "
                       "T=3000 x L=400, V=5000 -- the layout in which the window scorer reads only its own columns",
                       vocab_major=True),
    "seg": SegWorkload("seg", 256, 3500, 6, 150, 32,
                       "BASELINE configs[4] unit: anchor-loop iteration for 256 files in flight, 70 s windows "
                       "(T=3500), 6 utterances x 150 chars, all prefixes + on-device selection, V=32"),
    "c3seg": SegWorkload("c3seg", 65536, 500, 6, 6, 32,
                         "BASELINE configs[2] through the reference's own aligner call: 65536 word-level windows "
                         "(word_level_alignment.py:69-103: [pre, sep, WORD, sep, post, sep]), T=500, 6 utterances x 6 "
                         "tokens (44 columns), CTC segmentation + backtrace + scoring of the full text, V=32",
                         all_prefixes=False),
}



# ----------------------------------------------------------------------------- configs[4]: anchor sweep
C5_TEXT = ("BASELINE configs[4]: full anchor-loop sweep over {h:g} h of synthetic emissions (files of 5-60 min, "
           "rows of 20-60 words, V=32), sharded by file (LPT) over the ranks, on-device window construction, "
           "all-prefix segmentation, selection and anchor update")


def _c5_specs(hours, world, rank):
    """Every rank derives the same corpus description and keeps its own LPT shard of the files."""
    import sweep_corpus
    from ipfa_b200 import sharding
    rng = np.random.default_rng(2024)
    minutes, left = [], hours * 60.0
    while left > 1e-9:
        m = min(float(rng.uniform(5.0, 60.0)), left)
        minutes.append(max(m, 0.5))
        left -= m
    shards = sharding.lpt_shards(minutes, world)
    mine = sorted(shards[rank])
    specs = [sweep_corpus.make_spec(f"f{i:04d}", minutes[i], 7000 + i, corrupt_frac=0.06, non_speech_every=9)
             for i in mine]
    return specs, minutes, mine


def _c5_cpu_worker(args):
    from oracle import sweep as osweep
    import importlib
    stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    spec, lp = args
    t0 = time.perf_counter()
    rows, status, stats = osweep.sweep_file(spec.file_id, spec.audio_path, lp, spec.n_samples, spec.rows,
                                            stub.CharTokenizer())
    return time.perf_counter() - t0, stats["cells"], len(rows)


def c5_cpu_leg(budget_s):
    """The CPU restatement (oracle/sweep.py: C table fill + interpreted backtrace / scoring / loop,
    one alignment per shrinking-transcript iteration) on a bounded sample: one short file per core."""
    import multiprocessing as mp
    import sweep_corpus
    cores = os.cpu_count() or 1
    minutes = max(0.5, min(6.0, budget_s / 6.0))   # ~4 s of CPU per audio minute and process
    specs = [sweep_corpus.make_spec(f"cpu{i:03d}", minutes, 9000 + i, corrupt_frac=0.06, non_speech_every=9)
             for i in range(cores)]
    jobs = [(s, sweep_corpus.emissions(s, "cpu", seed=i).numpy()) for i, s in enumerate(specs)]
    pool = mp.get_context("fork").Pool(cores)
    try:
        t0 = time.perf_counter()
        res = pool.map(_c5_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    finally:
        pool.terminate()
    hours = sum(s.n_samples for s in specs) / 16000 / 3600.0
    return {"cores": cores, "oracle_port": {"audio_h_per_s": hours / dt, "cells_per_s": sum(r[1] for r in res) / dt,
                                            "windows": len(specs), "seconds_per_step": dt,
                                            "files": len(specs), "minutes_per_file": minutes}}


def c5_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    legs = c5_cpu_leg(budget_s=60.0)
    best = legs["oracle_port"]
    sample = (f"{best['files']} synthetic files of {best['minutes_per_file']:g} min (one per core) through "
              "oracle/sweep.py: the reference's per-file anchor loop, C table fill + interpreted backtrace")
    line = {"impl": "reference", "metric": "aligned_audio_hours_per_s", "value": best["audio_h_per_s"],
            "unit": "audio-h/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": best["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C5_TEXT.format(h=args.hours), "kernel": "sweep"},
            "cells_per_s": best["cells_per_s"],
            "cpu_baseline": {"value": best["audio_h_per_s"], "unit": "audio-h/s", "cores": legs["cores"],
                             "kind": "port", "sample": sample},
            "e2e": {"value": best["audio_h_per_s"], "unit": "audio-h/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def c5_measure(hours_total, world, rank, dev, steps, warm, groups=32, use_graphs=True, capacity=None,
               with_e2e=True, mode=None):
    """BASELINE configs[4] on this job's ranks: ONE corpus of `hours_total` hours, its files sharded over
    the ranks (LPT by duration), every rank runs the whole anchor loop of its files.  A step = state
    reset + run until every file stops.  Collective calls (all ranks must enter).  Returns a dict of
    job-wide numbers (times are the max over ranks, counts the sums)."""
    import torch
    import torch.distributed as dist
    import sweep_corpus
    from ipfa_b200 import sweep as sw_mod
    stub = __import__("importlib").import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    specs, minutes, mine = _c5_specs(hours_total, world, rank)
    files = [sw_mod.SweepFile(s.file_id, s.audio_path, sweep_corpus.emissions(s, dev, seed=i), s.n_samples, s.rows)
             for i, s in zip(mine, specs)]
    files = sw_mod.sort_longest_first(files)
    corpus = sw_mod.SweepCorpus(files, stub.CharTokenizer())
    for f in files:
        f.lpz = None  # the corpus holds the only copy
    sweep = sw_mod.AnchorSweep(corpus, index_duration=FRAME_SECONDS, samples_to_frames_ratio=320.0,
                               groups=groups, use_graphs=use_graphs, capacity=capacity, mode=mode)
    hours_mine = sum(s.n_samples for s in specs) / 16000 / 3600.0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_sweep():
        sweep.reset()
        return sweep.run(steps_per_poll=16)

    for _ in range(warm):   # also settles the launch capacity (CAPACITY round trips happen here)
        status = one_sweep()
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    launches0 = sweep.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record()
    for _ in range(steps):
        status = one_sweep()
    ev[1].record()
    barrier()
    elapsed_ms = ev[0].elapsed_time(ev[1])
    launches = sweep.kernel_launches - launches0  # kernels executed, graph replays included
    sampler.stop_flag = True
    sampler.join()
    st = sweep.stats()

    e2e_ms, h2d_b, d2h_b = 0.0, 0.0, 0.0
    if with_e2e:
        # end to end: emissions start in pinned host memory; corpus upload + sweep + result rows back
        host_lp = corpus.lp.cpu().pin_memory()
        out_seg_h = torch.empty(sweep.out_seg.shape, dtype=torch.float64).pin_memory()
        out_info_h = torch.empty(sweep.out_info.shape, dtype=torch.int32).pin_memory()
        e2e_steps = max(2, min(steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            if sweep.mode == "resident":
                corpus.begin_upload(host_lp)   # file by file on a copy stream; the sweep starts on the first files
            else:
                corpus.lp.copy_(host_lp, non_blocking=True)
            one_sweep()
            out_seg_h.copy_(sweep.out_seg, non_blocking=True)
            out_info_h.copy_(sweep.out_info, non_blocking=True)
            torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
        h2d_b, d2h_b = float(host_lp.numel() * 4), float(out_seg_h.numel() * 8 + out_info_h.numel() * 4)

    vals = torch.tensor([elapsed_ms, e2e_ms, float(st["steps"]), float(len(files))], dtype=torch.float64, device=dev)
    sums = torch.tensor([hours_mine, float(st["cells"]), float(st["frames"]), float(st["windows"]),
                         float(len(files)), float((status == sw_mod.DONE).sum()), float(launches), h2d_b, d2h_b],
                        dtype=torch.float64, device=dev)
    gathered = None
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        # the job's only collective on results: per-utterance rows to rank 0 (not timed)
        from ipfa_b200 import sharding
        gathered = sharding.gather_objects([len(r) for r in sweep.file_rows()])
    out = dict(zip(("hours", "cells", "frames", "windows", "files", "files_done", "launches", "h2d", "d2h"),
                   (float(x) for x in sums)))
    out.update(elapsed_ms=float(vals[0]), e2e_ms=float(vals[1]), iterations_max=int(vals[2]),
               files_per_rank_max=int(vals[3]), steps=steps, warm=warm, capacity=sweep.capacity,
               iterations_rank0=st["steps"], clocks=sampler.result(), gathered=gathered, mode=sweep.mode,
               longest_file_minutes=float(max(minutes)))
    del sweep, corpus, files
    torch.cuda.empty_cache()
    return out


def c5_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps = min(args.steps, 10)
    warm = max(min(args.warmup, 3), 3)
    m = c5_measure(args.hours, world, rank, dev, steps, warm, groups=args.groups, use_graphs=not args.no_graphs,
                   capacity=[int(x) for x in args.capacity.split(',')] if args.capacity else None,
                   mode=args.sweep_mode)
    if rank == 0:
        peak, peak_src = peaks()
        ms_per_step = m["elapsed_ms"] / steps
        alg_bytes = m["frames"] * 32 * 4 + 2 * m["cells"] / 8  # emission panel rows + 1-bit backpointers written and read
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        legs = c5_cpu_leg(budget_s=20.0)
        best = legs["oracle_port"]
        line = {
            "metric": "aligned_audio_hours_per_s", "value": m["hours"] / (ms_per_step * 1e-3), "unit": "audio-h/s",
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": C5_TEXT.format(h=args.hours), "kernel": "sweep"},
            "run": {"l2": f"corpus emissions {m['h2d'] / 1e6:.0f} MB over all ranks, every window read once per sweep",
                    "sharding": "files sharded by duration (LPT), no collective on the data path",
                    "files": int(m["files"]), "files_done": int(m["files_done"]), "hours": m["hours"],
                    "iterations_rank0": m["iterations_rank0"], "iterations_max": m["iterations_max"],
                    "windows": int(m["windows"]), "capacity_T_C_K_rank0": m["capacity"],
                    "sweep_mode": m["mode"],
                    "groups_per_gpu": args.groups if m["mode"] == "lockstep" else None,
                    "cuda_graphs": (not args.no_graphs) if m["mode"] == "lockstep" else None, "V": 32},
            "cells_per_s": m["cells"] / (ms_per_step * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic("c5") if m["mode"] == "resident" else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_per_step,
                         "kernels_per_step": m["launches"] / max(steps, 1) / world,
                         "note": "a file's windows are a serial chain of T-serial fills, each on one SM (resident mode: "
                                 "one persistent launch, a CTA per file; lock step: three launches per iteration); the "
                                 "longest file's chain bounds the sweep: see DESIGN.md 5.6"},
            "cpu_baseline": {"value": best["audio_h_per_s"], "unit": "audio-h/s", "cores": legs["cores"],
                             "kind": "port",
                             "sample": f"{best['files']} files of {best['minutes_per_file']:g} min, one per core, "
                                       "through oracle/sweep.py (C table fill + interpreted backtrace/loop)",
                             "oracle_port": best},
            "e2e": {"value": m["hours"] / (m["e2e_ms"] * 1e-3), "unit": "audio-h/s", "h2d_bytes_per_step": int(m["h2d"]),
                    "d2h_bytes_per_step": int(m["d2h"]), "ms_per_step": m["e2e_ms"],
                    "api": "ipfa_sweep_resident_device" if m["mode"] == "resident" else "ipfa_sweep_step_device"},
            "gpu_launches": int(m["launches"]), "clocks": m["clocks"],
        }
        if world > 1:
            line["final_gather"] = {"files": int(sum(len(g) for g in m["gathered"])), "backend": "nccl"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def config_obj(wl):
    """The `config` of a line: the workload and its shape, the same object in both arms."""
    return dict({"workload": wl.text, "kernel": wl.kind}, **wl.shape)


# ----------------------------------------------------------------------------- CPU legs
def cpu_leg(wl, budget_s, steps=1, warmup=0):
    """Times the CPU path(s) on a bounded sample of the workload, all host cores."""
    cap = wl.cpu_setup()
    cores = os.cpu_count() or 1
    out = {"cores": cores}
    for name, fn in wl.cpu_legs().items():
        try:
            probe = min(cap, max(cores, 16))
            fn(probe)  # warm-up (thread pools, page faults)
            t0 = time.perf_counter()
            fn(probe)
            per_window = (time.perf_counter() - t0) / probe
            total_steps = max(steps + warmup, 1)
            k = int(min(cap, max(probe, budget_s / total_steps / max(per_window, 1e-9))))
            for _ in range(warmup):
                fn(k)
            t0 = time.perf_counter()
            for _ in range(max(steps, 1)):
                fn(k)
            dt = (time.perf_counter() - t0) / max(steps, 1)
            cells, hours = wl.cpu_units(k)
            out[name] = {"audio_h_per_s": hours / dt, "cells_per_s": cells / dt, "windows": k,
                         "seconds_per_step": dt}
        except Exception as exc:  # e.g. torchaudio absent on some box
            out[name] = {"error": repr(exc)}
    if getattr(wl, "_pool", None) is not None:
        wl._pool.terminate()
    try:
        from oracle import ctc as octc
        out["oracle_threads"] = octc.num_threads()
    except Exception:
        pass
    return out


def best_cpu(legs):
    names = [k for k in ("oracle_port", "torch_cpu") if k in legs and "audio_h_per_s" in legs[k]]
    best = max(names, key=lambda k: legs[k]["audio_h_per_s"])
    return best, legs[best]


def cpu_baseline_obj(legs, what):
    name, best = best_cpu(legs)
    obj = {"value": best["audio_h_per_s"], "unit": "audio-h/s", "cores": legs["cores"],
           "kind": "port" if name == "oracle_port" else "reference",
           "sample": f"{best['windows']} windows of the workload per step ({what}); fastest CPU leg: {name}"}
    for k in ("oracle_port", "torch_cpu"):
        if k in legs:
            obj[k] = legs[k]
    return obj


def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    wl = WORKLOADS[args.workload]
    legs = cpu_leg(wl, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    name, best = best_cpu(legs)
    what = ("oracle/ restatement of ctc-segmentation: C fill + interpreted backtrace/scoring, one process per core"
            if wl.kind == "seg" else
            "OpenMP C restatement (oracle/) vs installed torch/torchaudio CPU comparator north_star names")
    line = {
        "impl": "reference", "metric": "aligned_audio_hours_per_s", "value": best["audio_h_per_s"],
        "unit": "audio-h/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": best["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_obj(wl),
        "cells_per_s": best["cells_per_s"], "cpu_baseline": cpu_baseline_obj(legs, what),
        "e2e": {"value": best["audio_h_per_s"], "unit": "audio-h/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm
def _timed(fn, n, barrier, torch):
    """n calls of fn(i) between two CUDA events on the current stream; returns ms."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record()
    last = None
    for i in range(n):
        last = fn(i)
    ev[1].record()
    barrier()
    return ev[0].elapsed_time(ev[1]), last


def measure_resident(wl, ipfa, dev, rank, steps, warm, barrier, use_graphs, sustain_s=0.0):
    """Device-resident throughput of one workload on this rank: rotating input sets (> L2), optional CUDA
    graphs, K timed steps, then (sustain_s > 0) the same loop for at least that long."""
    import torch
    n_sets = max(3, min(8, int(np.ceil(3 * 126e6 / wl.set_bytes)))) if wl.set_bytes < 2e9 else 1
    sets = [wl.make(1000 * rank + s_, device=dev) for s_ in range(n_sets)]
    cells, hours, alg_bytes = wl.units(sets[0])
    graphs, graph_outs, launches_per_step = [], [], 0
    if use_graphs:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for s_ in sets:
                wl.step(ipfa, s_)
            side.synchronize()
            n0 = ipfa.launch_count()
            for s_ in sets:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    graph_outs.append(wl.step(ipfa, s_))
                graphs.append(g)
            launches_per_step = (ipfa.launch_count() - n0) // n_sets
        torch.cuda.synchronize()

    def do_step(i):
        if use_graphs:
            graphs[i % n_sets].replay()
            return graph_outs[i % n_sets]
        return wl.step(ipfa, sets[i % n_sets])

    for i in range(warm):
        do_step(i)
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    launches0 = ipfa.launch_count()
    elapsed_ms, last = _timed(do_step, steps, barrier, torch)
    launches = launches_per_step * steps if use_graphs else ipfa.launch_count() - launches0
    sustained = None
    if sustain_s > 0:
        n_sus = int(max(steps, min(2_000_000, sustain_s * 1e3 / max(elapsed_ms / steps, 1e-4))))
        sus_ms, _ = _timed(do_step, n_sus, barrier, torch)
        sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus}
    sampler.stop_flag = True
    sampler.join()
    # the dominant kernel alone, CUDA events on the launching stream around its launch (no graph)
    torch.cuda.synchronize()
    k = [0]

    def one_plain():
        k[0] += 1
        wl.step(ipfa, sets[k[0] % n_sets])
    kernel_ms = ipfa.ops.dominant_kernel_ms(one_plain, repeats=6)
    return {"sets": sets, "n_sets": n_sets, "cells": cells, "hours": hours, "alg_bytes": alg_bytes,
            "elapsed_ms": elapsed_ms, "steps": steps, "launches": launches, "last": last, "clocks": sampler.result(),
            "sustained": sustained, "kernel_ms": kernel_ms, "use_graphs": use_graphs}


def gpu_comparators(wl, sets, torch, budget_windows=64):
    """The kernels this image already has for the same jobs, on the same
    inputs: ATen's CUDA ctc_loss (whole batch, [T, N, C]) for the window scorer, torchaudio's CUDA
    forced_align (one launch per frame, batch size 1: a sample of the windows) for Viterbi."""
    lp, tg, il, tl = sets[0]
    out = {}
    try:
        if wl.kind == "alpha":
            n = lp.shape[0] if wl.set_bytes < 1e9 else min(lp.shape[0], budget_windows)
            x = lp[:n].transpose(0, 1).contiguous()
            tgt, a, b = tg[:n].long(), il[:n].long(), tl[:n].long()
            fn = lambda: torch.nn.functional.ctc_loss(x, tgt, a, b, blank=0, reduction="none")
            units = float(il[:n].sum()) * FRAME_SECONDS / 3600.0
            name = "torch.nn.functional.ctc_loss (CUDA, reduction='none')"
        else:
            import torchaudio.functional as AF
            n = min(lp.shape[0], budget_windows)
            items = [(lp[i:i + 1, :int(il[i])].contiguous(), tg[i:i + 1, :int(tl[i])].long()) for i in range(n)]

            def fn():
                for e, t_ in items:
                    AF.forced_align(e, t_, blank=0)
            units = float(il[:n].sum()) * FRAME_SECONDS / 3600.0
            name = "torchaudio.functional.forced_align (CUDA, one call per window)"
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 5 if wl.kind == "alpha" else 2
        ev[0].record()
        for _ in range(reps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        out = {"name": name, "windows": int(n), "ms": ms, "audio_h_per_s": units / (ms * 1e-3)}
    except Exception as exc:  # noqa: BLE001  (a comparator that does not run is reported, not fatal)
        out = {"error": repr(exc)[:200]}
    return out


def raw_copy_ceiling(nbytes, barrier, torch, seconds=0.25):
    """What the link gives this rank while every rank copies at once: one pinned buffer of the step's
    size, cudaMemcpyAsync host-to-device in a loop (the e2e step cannot be faster than this)."""
    host = torch.empty(int(nbytes), dtype=torch.uint8).pin_memory()
    devb = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        devb.copy_(host, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    n = 0
    while True:
        devb.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        n += 1
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return nbytes * n / dt / 1e9


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

    wl = WORKLOADS[args.workload]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # Window scoring is a memset and two or three short kernels per step: the step of every input set
    # is captured once in a CUDA graph and replayed (no host launch gaps); --no_graphs launches from the host.
    use_graphs = wl.kind == "alpha" and not args.no_graphs
    warm = max(args.warmup, 3)
    m = measure_resident(wl, ipfa, dev, rank, args.steps, warm, barrier, use_graphs, sustain_s=args.sustain)
    sets, n_sets, cells, hours, alg_bytes = m["sets"], m["n_sets"], m["cells"], m["hours"], m["alg_bytes"]
    elapsed_ms, launches, last = m["elapsed_ms"], m["launches"], m["last"]

    # the only collective of the job: one gather of the per-window results (not timed)
    gathered = None
    if world > 1:
        from ipfa_b200 import sharding
        res = last.float().reshape(last.shape[0], -1)
        full = sharding.gather_rows(res, list(range(rank * res.shape[0], (rank + 1) * res.shape[0])),
                                    world * res.shape[0])
        gathered = [int(full.shape[0]), bool(torch.isfinite(full).all())]

    # end-to-end through the host-buffer C ABI: pinned host buffers, H2D + D2H inside the timed region
    host_sets = [wl.to_host(sets[s_]) for s_ in range(min(n_sets, 2))]
    e2e_steps = max(3, min(args.steps, 20))
    for i in range(3):
        wl.e2e_step(ipfa, *host_sets[i % len(host_sets)])
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        wl.e2e_step(ipfa, *host_sets[i % len(host_sets)])
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    h2d, d2h = wl.traffic()
    raw_gbs = raw_copy_ceiling(h2d, barrier, torch)

    sus_ms = m["sustained"]["ms_per_step"] if m["sustained"] else 0.0
    times = torch.tensor([elapsed_ms, e2e_ms, sus_ms, -raw_gbs], dtype=torch.float64, device=dev)
    raw_sum = torch.tensor([raw_gbs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(raw_sum, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, sus_ms, raw_min = float(times[0]), float(times[1]), float(times[2]), -float(times[3])

    # BASELINE configs[4] on the same ranks (strong scaling of one corpus), and -- single GPU only -- short
    # runs of the other configs; every rank enters c5_measure (it is collective)
    extra = {}
    if args.extras and args.workload == "c2":
        del host_sets
        c5 = c5_measure(args.hours, world, rank, dev, steps=2, warm=2, with_e2e=False)
        ms5 = c5["elapsed_ms"] / c5["steps"]
        extra["c5"] = {"workload": C5_TEXT.format(h=args.hours), "scaling": "strong", "ms_per_sweep": ms5,
                       "audio_h_per_s": c5["hours"] / (ms5 * 1e-3), "files": int(c5["files"]),
                       "files_per_rank_max": c5["files_per_rank_max"], "iterations_longest_chain": c5["iterations_max"],
                       "longest_file_minutes": c5["longest_file_minutes"], "windows": int(c5["windows"]),
                       "cells_per_s": c5["cells"] / (ms5 * 1e-3), "kernels_per_sweep": c5["launches"] / c5["steps"] / world,
                       "sweep_mode": c5["mode"]}
        if world > 1:
            # the same sweep with the corpus growing with the job (100 h per GPU): a file's windows are a serial
            # chain, so ONE corpus stops scaling at its longest file; capacity is what more GPUs buy
            c5w = c5_measure(args.hours * world, world, rank, dev, steps=2, warm=2, with_e2e=False)
            msw = c5w["elapsed_ms"] / c5w["steps"]
            extra["c5_weak"] = {"workload": C5_TEXT.format(h=args.hours * world), "scaling": "weak",
                                "ms_per_sweep": msw, "audio_h_per_s": c5w["hours"] / (msw * 1e-3),
                                "files": int(c5w["files"]), "files_per_rank_max": c5w["files_per_rank_max"],
                                "iterations_longest_chain": c5w["iterations_max"], "windows": int(c5w["windows"]),
                                "sweep_mode": c5w["mode"]}
        if world == 1:
            for name in ("c2v", "c3", "c3seg", "c4", "c4v", "seg"):
                w2 = WORKLOADS[name]
                r2 = measure_resident(w2, ipfa, dev, rank, steps=5, warm=3, barrier=barrier, use_graphs=False)
                ms2 = r2["elapsed_ms"] / r2["steps"]
                extra[name] = {"workload": w2.text, "ms_per_step": ms2, "audio_h_per_s": r2["hours"] / (ms2 * 1e-3),
                               "cells_per_s": r2["cells"] / (ms2 * 1e-3), "kernel_ms": r2["kernel_ms"],
                               "roofline_frac": r2["alg_bytes"] / (ms2 * 1e-3) / 1e9 / peaks()[0],
                               "algorithmic_bytes_per_step": r2["alg_bytes"]}
                if w2.kind in ("alpha", "viterbi"):
                    extra[name]["gpu_comparator"] = gpu_comparators(w2, r2["sets"], torch)
                del r2
                torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = peaks()
        ms_per_step = elapsed_ms / args.steps
        kernel_ms = m["kernel_ms"] if m["kernel_ms"] else ms_per_step
        # roofline of the dominant kernel: its algorithmic bytes over its own event-timed duration
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        legs = cpu_leg(wl, budget_s=12.0)
        lin = wl.kind == "alpha" and wl.v <= 64 and wl.l + 1 <= 256 and not os.environ.get("IPFA_ALPHA_LOG")
        bound_note = {"alpha": ("T-serial recursion walked from both ends, states kept as scaled fp64 probabilities "
                                "(linear-domain instance): issue / latency bound at 3.5 resident warps per SM "
                                "sub-partition, not HBM bound -- DESIGN.md 5.1") if lin else
                               ("T-serial log-sum-exp recursion walked from both ends: MUFU bound (4 MUFU per state "
                                "pair and frame), not HBM bound -- DESIGN.md 5.1"),
                      "viterbi": "T-serial max-plus recursion + latency-bound backtrace: issue bound -- DESIGN.md 5.2",
                      "seg": "T-serial max-plus recursion, issue bound (13-15 instructions per cell) -- DESIGN.md 5.3"}[wl.kind]
        extra_roof = {}
        if wl.kind == "alpha" and wl.v <= 64:
            il = sets[0][2].cpu().numpy().astype(np.int64)
            tl = sets[0][3].cpu().numpy().astype(np.int64)
            clk = (m["clocks"].get("sm_mhz") or 1965.0) * 1e6
            if lin:
                # the pipes that bound the linear-domain instance: per frame and warp of
                # 32 lanes x P pairs, 3P + 1 FP64 warp-instructions (2 cycles each per SM sub-partition,
                # profiles/microbench_fp64_r01.txt) out of ~9P + 1 issued instructions
                pairs = int(tl.max()) + 1
                per_lane = 1 if pairs <= 32 else 2 if pairs <= 64 else 4 if pairs <= 128 else 8
                frames = float(il.sum())
                fp64_ms = frames * (3 * per_lane + 1) * 2 / (148 * 4) / clk * 1e3
                issue_ms = frames * (9 * per_lane + 1) / (148 * 4) / clk * 1e3
                extra_roof = {"fp64_pipe_floor_ms": fp64_ms, "issue_floor_ms": issue_ms,
                              "frac_of_issue_floor": issue_ms / kernel_ms}
            else:
                pair_warps = np.ceil((tl + 1) / 32.0)
                mufu_cycles = float((il * pair_warps).sum()) * 4 * 8
                floor_ms = mufu_cycles / (148 * 4) / clk * 1e3
                extra_roof = {"mufu_floor_ms": floor_ms, "frac_of_mufu_floor": floor_ms / kernel_ms}
        line = {
            "metric": "aligned_audio_hours_per_s", "value": world * hours / (ms_per_step * 1e-3),
            "unit": "audio-h/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if lin else "f32", "data": "synthetic",
            "config": config_obj(wl),
            "run": {"l2": f"{n_sets} rotating input sets of {wl.set_bytes / 1e6:.0f} MB per GPU (> 126 MB L2)",
                    "sharding": "independent windows per rank, no collective on the data path",
                    "cuda_graphs": use_graphs},
            "cells_per_s": world * cells / (ms_per_step * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(args.workload),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                         "kernel_ms_how": ("CUDA events on the launching stream around the step's longest kernel, "
                                           "mean of 6 launches outside the graph") if m["kernel_ms"] else "step time",
                         "step_ms": ms_per_step, "kernel_share_of_step": kernel_ms / ms_per_step,
                         "kernels_per_step": launches / max(args.steps, 1), "note": bound_note, **extra_roof},
            "cpu_baseline": cpu_baseline_obj(legs, "same shapes, same generator"),
            "e2e": {"value": world * hours / (e2e_ms * 1e-3), "unit": "audio-h/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "api": wl.api,
                    "raw_h2d_copy_gbs": {"sum_over_ranks": float(raw_sum[0]), "slowest_rank": raw_min,
                                         "how": "every rank copies one pinned buffer of h2d_bytes_per_step to its "
                                                "GPU in a loop at the same time (cudaMemcpyAsync); the e2e step "
                                                "moves the same bytes, so it cannot beat this"},
                    "link_share": (h2d / (e2e_ms * 1e-3) / 1e9) / max(raw_min, 1e-9)},
            "gpu_launches": int(launches),
            "clocks": m["clocks"],
        }
        if m["sustained"]:
            line["sustained"] = {"value": world * hours / (sus_ms * 1e-3), "unit": "audio-h/s", "ms_per_step": sus_ms,
                                 "steps": m["sustained"]["steps"], "seconds": m["sustained"]["seconds"],
                                 "note": "same loop run for >= --sustain seconds right after the K timed steps; "
                                         "`value` is the K-step (burst) figure, the clocks were sampled over both"}
        if wl.kind in ("alpha", "viterbi"):
            line["gpu_comparator"] = gpu_comparators(wl, sets, torch)
            if "audio_h_per_s" in line["gpu_comparator"]:
                line["gpu_comparator"]["ours_over_it"] = (hours / (ms_per_step * 1e-3)) / line["gpu_comparator"]["audio_h_per_s"]
        if extra:
            line["extra"] = extra
        if gathered is not None:
            line["final_gather"] = {"rows": gathered[0], "finite": gathered[1], "backend": "nccl"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--hours", type=float, default=100.0, help="c5: hours of audio in the corpus (all ranks)")
    ap.add_argument("--capacity", default="", help="c5: initial launch capacity T,C,K (default: from the rows)")
    ap.add_argument("--no_graphs", action="store_true", help="c5: launch every kernel from the host (no CUDA graph)")
    ap.add_argument("--sweep_mode", default="auto", choices=["auto", "resident", "lockstep"],
                    help="c5: one persistent launch with a CTA per file, or lock-step iterations")
    ap.add_argument("--groups", type=int, default=32, help="c5: independent file groups (streams) per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sustain", type=float, default=1.0, help="seconds of the sustained run after the K timed steps (0: none)")
    ap.add_argument("--no_extras", dest="extras", action="store_false",
                    help="c2 only: skip the extra.c5 / c2v / c3 / c4 / seg sub-results")
    args = ap.parse_args()
    if args.workload == "c5":
        return c5_reference_arm(args) if args.impl == "reference" else c5_arm(args)
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
