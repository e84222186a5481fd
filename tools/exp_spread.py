import os, sys, time
import torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import ipfa_b200
from ipfa_b200 import sweep as sw, stub_asr
import sweep_corpus
spec = sweep_corpus.make_spec("long", 60.0, 7001, corrupt_frac=0.06, non_speech_every=9)
lp = sweep_corpus.emissions(spec, "cuda", seed=1)
for name, flags in (("no spread", 2), ("spread 2", 2 | 64), ("spread 4", 2 | 128)):
    f = sw.SweepFile(spec.file_id, spec.audio_path, lp, spec.n_samples, spec.rows)
    run = sw.AnchorSweep(sw.SweepCorpus([f], stub_asr.CharTokenizer()), index_duration=0.02, samples_to_frames_ratio=320.0, seg_flags=flags)
    for _ in range(2):
        run.reset(); run.run(steps_per_poll=16)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    run.reset(); run.run(steps_per_poll=16)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = run.stats()
    print(f"{name}: {dt * 1e3:.1f} ms, {dt / max(st['windows'], 1) * 1e6:.0f} us per window")
