"""The anchor sweep's critical path: one 60-minute file alone (its windows are a serial chain)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ipfa_b200
from ipfa_b200 import sweep as sw, stub_asr
import sweep_corpus

for minutes in (60.0, 30.0):
    spec = sweep_corpus.make_spec("long", minutes, 7001, corrupt_frac=0.06, non_speech_every=9)
    lp = sweep_corpus.emissions(spec, "cuda", seed=1)
    f = sw.SweepFile(spec.file_id, spec.audio_path, lp, spec.n_samples, spec.rows)
    run = sw.AnchorSweep(sw.SweepCorpus([f], stub_asr.CharTokenizer()), index_duration=0.02, samples_to_frames_ratio=320.0)
    for _ in range(2):
        run.reset(); run.run(steps_per_poll=16)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    run.reset(); run.run(steps_per_poll=16)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = run.stats()
    print(f"{minutes:.0f}-minute file alone: {dt * 1e3:.1f} ms, {st['steps']} iterations, {st['windows']} windows, "
          f"{dt / max(st['windows'], 1) * 1e6:.0f} us per window, capacity {run.capacity}")
