#!/bin/bash
# Run on the GPU box (under gpurun): the anchor sweep's bench lines in both modes, the launch list of a
# file-resident sweep and one ncu --set full capture of the persistent kernel on the 100 h corpus.
# Every ncu run is preceded by the same command without ncu.
mkdir -p gpurun_out
(timeout 300 python bench.py --workload c5 --sweep_mode resident --steps 5 --warmup 3 2>&1 | tail -1) > gpurun_out/r02_bench_c5_resident.json
(timeout 300 python bench.py --workload c5 --sweep_mode lockstep --steps 5 --warmup 3 2>&1 | tail -1) > gpurun_out/r02_bench_c5_lockstep.json
bash tools/exp_resident.sh > gpurun_out/r02_exp_resident.txt 2>&1
python bench.py --workload c5 --sweep_mode resident --steps 1 --warmup 3 > gpurun_out/plain_c5r.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ctc_|ctcseg|anchor|sweep" -c 40 --csv \
    --log-file gpurun_out/r02_launches_c5_resident.csv python bench.py --workload c5 --sweep_mode resident --steps 1 --warmup 3 > gpurun_out/ncu_l_c5r.log 2>&1
python bench.py --workload c5 --sweep_mode resident --steps 1 --warmup 3 > gpurun_out/plain_c5r2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_resident -s 4 -c 1 -f -o gpurun_out/r02_prof_resident \
    python bench.py --workload c5 --sweep_mode resident --steps 1 --warmup 3 > gpurun_out/ncu_c5r.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r02_nvsmi_resident.csv
ls -la gpurun_out | grep "r02_.*resident\|r02_exp_res"
