#!/bin/bash
# per-kernel durations of a one-file sweep with the fill spread over 1 / 2 / 4 SMs
for f in 2 66 130; do
  IPFA_EXP_FLAGS=$f ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/spread_$f.csv python - <<'PY' > /dev/null 2>&1
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, ipfa_b200
from ipfa_b200 import sweep as sw, stub_asr
import sweep_corpus
spec = sweep_corpus.make_spec("long", 20.0, 7001, corrupt_frac=0.06, non_speech_every=9)
lp = sweep_corpus.emissions(spec, "cuda", seed=1)
f = sw.SweepFile(spec.file_id, spec.audio_path, lp, spec.n_samples, spec.rows)
run = sw.AnchorSweep(sw.SweepCorpus([f], stub_asr.CharTokenizer()), index_duration=0.02, samples_to_frames_ratio=320.0, seg_flags=int(os.environ["IPFA_EXP_FLAGS"]), use_graphs=False)
run.reset(); run.run(steps_per_poll=16)
torch.cuda.synchronize()
PY
  python - <<PY
import csv
agg={}
for row in csv.reader(open("gpurun_out/spread_$f.csv")):
    if len(row)>14 and row[12]=="gpu__time_duration.sum":
        name=row[4].split("(")[0].replace("void ","")[:60]
        a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=float(row[14])
print("flags $f")
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:5]:
    print("   %-62s n=%4d mean %8.1f us"%(k,v[0],v[1]/v[0]/1000))
PY
done
