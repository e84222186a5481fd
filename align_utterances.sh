#!/bin/bash
# Utterance-level alignment driver -- same config zone and stage order as the reference's
# align_utterances.sh (:51-75, :113-152).  The n_process fan-out (:127-137) is one process
# per GPU: files are sharded deterministically inside the script, no claim files.

# config zone
alignment_name="sample"                 # alignment name, comment to use timestamp instead
tsv_path=data/sample/tsv/sample.tsv     # source file with metadata
vad_segments_filtered_filepath=""       # <name>_vad_segments_filtered.tsv (VAD stage is out of scope here)
merge_files=true
generate_stm_results=true
n_process=1                             # = number of GPUs of this box to use

threshold=-2.0
short_utterance_len=30
max_words_sequence=100
max_window_size=70.0
window_to_stop=500.0
min_text_to_audio_prop=0.8
max_text_to_audio_prop_exec=10

asr_hub="stub"                          # e.g. "Voyager1/asr-wav2vec2-commonvoice-es" when speechbrain is installed
asr_savedir="data/asr/"

if [ ! -z ${alignment_name+set} ]; then wip_dir="data/wip_"$alignment_name; else wip_dir="data/wip_"$(date +%s); fi
results_dir=$wip_dir"/results"; logs_dir=$wip_dir"/logs"
mkdir -p $results_dir $logs_dir
find $results_dir -type f -empty -print -delete

echo "Starting alignment..."
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n_process --master-addr 127.0.0.1 --master-port 29517 \
    src/iterative_utterance_alignment.py --tsv $tsv_path --vad_segments_tsv $vad_segments_filtered_filepath \
    --dst $results_dir --asr_hub $asr_hub --asr_savedir $asr_savedir --threshold $threshold \
    --logs_path $logs_dir --short_utterance_len $short_utterance_len --max_words_sequence $max_words_sequence \
    --max_window_size $max_window_size --window_to_stop $window_to_stop --min_text_to_audio_prop $min_text_to_audio_prop \
    --max_text_to_audio_prop_exec $max_text_to_audio_prop_exec > $logs_dir"/global.log"

if $merge_files; then python -u src/merge_aligned_files.py --global_tsv $tsv_path --src $results_dir; fi
if $generate_stm_results; then
    stm_dir=$results_dir/stm; mkdir -p $stm_dir
    python -u src/tsv_to_stm.py --src_path $results_dir --dst_path $stm_dir
fi
