#!/bin/bash
# Run on the GPU box (under gpurun): bench lines + ncu launch list + ncu --set full capture for the
# window-scoring workload (BASELINE configs[1]) with the linear-domain instance (default) and with
# the log-domain instance alone (IPFA_ALPHA_LOG=1).  Every ncu run follows the same command without ncu.
mkdir -p gpurun_out
(timeout 400 python bench.py --workload c2 2>&1 | tail -1) > gpurun_out/bench_c2.json
(IPFA_ALPHA_LOG=1 timeout 400 python bench.py --workload c2 2>&1 | tail -1) > gpurun_out/bench_c2_log.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c2_reference.json
python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ctc_|length_bucket" -c 60 --csv \
    --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/ncu_l_c2.log 2>&1
python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/plain_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctc_alpha -s 4 -c 2 -f -o gpurun_out/prof_alpha \
    python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/ncu_a.log 2>&1
tools/microbench_fp64 > gpurun_out/microbench_fp64.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/nvsmi.csv
tail -c 600 gpurun_out/bench_c2.json
