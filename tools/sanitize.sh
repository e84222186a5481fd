#!/bin/bash
# compute-sanitizer over a small-shape subset of the GPU tests (run under gpurun):
#   memcheck   out-of-bounds / misaligned accesses, incl. the bulk-copy (cp.async.bulk + mbarrier) pipe
#   racecheck  shared-memory hazards: the emission ring refilled after the in-place conversion
#              (ctc_alpha.cu, LIN overlap path), the exchange lines, the cluster hand-over of the segmentation fill (IPFA_SEG_SPREAD_2 / _4)
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
# the cluster variant of the segmentation fill (st.async hand-over through distributed shared memory)
echo "== racecheck, segmentation fill spread over a cluster" > gpurun_out/sanitize_cluster.txt
timeout ${SAN_TIMEOUT:-900} $SAN --tool racecheck --print-limit 20 --error-exitcode 99 \
    python -m pytest "tests/test_gpu_ctcseg.py::test_window_columns_spread_over_a_cluster" -m gpu -q -x -p no:cacheprovider >> gpurun_out/sanitize_cluster.txt 2>&1
echo "exit code $?" >> gpurun_out/sanitize_cluster.txt
for f in gpurun_out/sanitize_*.txt; do echo "---- $f"; tail -8 $f; done
