#!/bin/bash
# Run on the GPU box (under gpurun): bench lines + ncu launch lists + ncu --set full captures.
# Every ncu run is preceded by the same command without ncu (exit code checked by &&).
mkdir -p gpurun_out
for w in c2 c2v c3 c4 seg; do
  (timeout 400 python bench.py --workload $w 2>&1 | tail -1) > gpurun_out/bench_$w.json
done
(IPFA_ALPHA_LOG=1 timeout 400 python bench.py --workload c2 2>&1 | tail -1) > gpurun_out/bench_c2_log.json
(timeout 600 python bench.py --workload c5 --steps 5 --warmup 2 2>&1 | tail -1) > gpurun_out/bench_c5.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c2_reference.json
timeout 300 python bench.py --impl reference --workload c5 2>&1 | tail -1 > gpurun_out/bench_c5_reference.json
for w in c2 c2v c3 seg; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/plain_$w.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ctc_|ctcseg|anchor|sweep" -c 60 --csv \
      --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/ncu_l_$w.log 2>&1
done
python bench.py --workload c5 --hours 10 --steps 1 --warmup 1 --groups 1 > gpurun_out/plain_c5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ctc_|ctcseg|anchor|sweep" -s 2000 -c 600 --csv \
    --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --hours 10 --steps 1 --warmup 1 --groups 1 > gpurun_out/ncu_l_c5.log 2>&1
python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/plain_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctc_alpha -s 4 -c 2 -f -o gpurun_out/prof_alpha \
    python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/ncu_a.log 2>&1
python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctc_alpha -s 2 -c 1 -f -o gpurun_out/prof_alpha_c4 \
    python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/ncu_c4.log 2>&1
python bench.py --workload c2v --steps 5 --warmup 3 > gpurun_out/plain_v.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:viterbi -s 6 -c 2 -f -o gpurun_out/prof_viterbi \
    python bench.py --workload c2v --steps 5 --warmup 3 > gpurun_out/ncu_v.log 2>&1
python bench.py --workload seg --steps 5 --warmup 3 > gpurun_out/plain_s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ctcseg|anchor" -s 9 -c 3 -f -o gpurun_out/prof_seg \
    python bench.py --workload seg --steps 5 --warmup 3 > gpurun_out/ncu_s.log 2>&1
[ -x tools/microbench_fp64 ] && tools/microbench_fp64 > gpurun_out/microbench_fp64.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/nvsmi.csv
ls -la gpurun_out | tail -30
