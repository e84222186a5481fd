#!/bin/bash
# compute-sanitizer over a small-shape subset of the GPU tests (run under gpurun):
#   memcheck   out-of-bounds / misaligned accesses, incl. the bulk-copy (cp.async.bulk + mbarrier) pipe
#   racecheck  shared-memory hazards: the emission ring refilled after the in-place conversion
#              (ctc_alpha.cu, LIN overlap path), the exchange lines, the DSMEM ring (IPFA_SEG_CLUSTER=1)
#   synccheck  barrier / mbarrier misuse
# Summaries land in gpurun_out/sanitize_<tool>.txt; the last lines of each are what profiles/ keeps.
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SUBSET="tests/test_gpu_edges.py tests/test_gpu_ctc.py::test_alpha_golden tests/test_gpu_ctc.py::test_alpha_linear_instance tests/test_gpu_ctc.py::test_alpha_linear_instance_shapes tests/test_gpu_ctc.py::test_alpha_fp32_tier_is_exact_where_it_answers tests/test_gpu_ctc.py::test_viterbi_golden tests/test_gpu_ctcseg.py::test_flags_and_unpeaked tests/test_gpu_ctcseg.py::test_windowed_mode_equals_full_table_when_the_audio_fits tests/test_gpu_sweep.py::test_sweep_equals_cpu_oracle_on_a_synthetic_corpus"
for tool in memcheck racecheck synccheck; do
    echo "== $tool" > gpurun_out/sanitize_$tool.txt
    timeout ${SAN_TIMEOUT:-900} $SAN --tool $tool --print-limit 20 --error-exitcode 99 \
        python -m pytest $SUBSET -m gpu -q -x -p no:cacheprovider >> gpurun_out/sanitize_$tool.txt 2>&1
    echo "exit code $?" >> gpurun_out/sanitize_$tool.txt
done
# the cluster experiment of the segmentation fill (distributed shared memory ring)
echo "== racecheck, IPFA_SEG_CLUSTER=1" > gpurun_out/sanitize_cluster.txt
IPFA_SEG_CLUSTER=1 timeout ${SAN_TIMEOUT:-900} $SAN --tool racecheck --print-limit 20 --error-exitcode 99 \
    python -m pytest "tests/test_gpu_ctcseg.py::test_all_prefixes_vs_oracle" -m gpu -q -x -p no:cacheprovider >> gpurun_out/sanitize_cluster.txt 2>&1
echo "exit code $?" >> gpurun_out/sanitize_cluster.txt
for f in gpurun_out/sanitize_*.txt; do echo "---- $f"; tail -8 $f; done
