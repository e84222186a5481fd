import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, ipfa_b200
from ipfa_b200 import sweep as sw, stub_asr
import sweep_corpus
spec = sweep_corpus.make_spec("long", 20.0, 7001, corrupt_frac=0.06, non_speech_every=9)
lp = sweep_corpus.emissions(spec, "cuda", seed=1)
f = sw.SweepFile(spec.file_id, spec.audio_path, lp, spec.n_samples, spec.rows)
run = sw.AnchorSweep(sw.SweepCorpus([f], stub_asr.CharTokenizer()), index_duration=0.02, samples_to_frames_ratio=320.0, seg_flags=int(os.environ.get("IPFA_EXP_FLAGS", "130")), use_graphs=False)
run.reset(); run.run(steps_per_poll=16)
torch.cuda.synchronize()
