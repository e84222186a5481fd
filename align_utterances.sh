#!/bin/bash
#
# Utterance-level iterative pseudo-forced alignment -- same config zone, WIP layout and stage order
# as the reference's align_utterances.sh (:51-75 config, :77-107 directories, :127-152 stages).
#
# What differs: the n_process background copies of the entry point that claimed files through empty
# result TSVs (:127-137) become one process per GPU (torchrun); the files are sharded inside the entry
# point, deterministically and balanced by length.  The VAD stage (:109-123, a SpeechBrain VAD model)
# is outside this repo: point vad_segments_filtered_filepath at its output, or keep
# generate_vad_segments=true where the reference's src/preprocess scripts are available.

#########################################################
###################### DEFINITIONS ######################
#########################################################

# config zone
alignment_name="sample"                 # alignment name, comment to use timestamp instead
tsv_path=data/sample/tsv/sample.tsv     # source file with metadata
merge_files=true                        # merge aligned files in a single tsv
generate_vad_segments=false             # true: run the reference's VAD scripts (src/preprocess, not shipped here)
generate_stm_results=true               # generate stm files from tsv results
n_process=1                             # number of GPUs of this box to use (one process per GPU)
resident_emissions=false                # true: encode every file once, anchor loops of all files on the GPU

# VAD configuration
max_non_speech_segments=20.0            # vad segments to filter

# Alignment parameters
threshold=-2.0                          # anchors threshold
short_utterance_len=30                  # minimum sequence of chars to select anchors
max_words_sequence=100                  # measured from CommonVoice
max_window_size=70.0                    # seconds
window_to_stop=500.0                    # seconds, window to stop execution
min_text_to_audio_prop=0.8              # min text to audio proportion
max_text_to_audio_prop_exec=10          # number of consecutive exceptions to stop

# trained ASR: a SpeechBrain EncoderASR source; "stub" = random-init emitter (no meaning, smoke runs only)
asr_hub="Voyager1/asr-wav2vec2-commonvoice-es"
asr_savedir="data/asr/"

#########################################################
####################### ALIGNMENT #######################
#########################################################

tsv_filename=$(basename $tsv_path)

if [ ! -z ${alignment_name+set} ]; then
    wip_dir="data/wip_"$alignment_name
    echo "Alignment name defined, WIP folder is: "$wip_dir
else
    wip_dir="data/wip_"$(date +%s)
    echo "Alignment name not defined, WIP folder is: "$wip_dir
fi

vad_dir=$wip_dir"/vad"
results_dir=$wip_dir"/results"
logs_dir=$wip_dir"/logs"
mkdir -p $wip_dir $vad_dir $results_dir $logs_dir

echo "Removing previous empty files from: "$results_dir
find $results_dir -type f -empty -print -delete

vad_segments_tsv=${tsv_filename/.tsv/_vad_segments.tsv}
vad_segments_filtered_filepath=${vad_segments_filtered_filepath:-$vad_dir"/"${vad_segments_tsv/.tsv/_filtered.tsv}}
if $generate_vad_segments; then
    echo "Generating VAD segments: "$tsv_path
    python -u src/preprocess/get_vad_segments_speechbrain.py --src $tsv_path --dst $vad_dir
    echo "Filtering VAD segments..."
    python -u src/preprocess/filter_non_speech_segments.py --src $vad_dir"/"$vad_segments_tsv --dst $vad_dir \
        --length $max_non_speech_segments
fi

echo "Starting alignment..."
resident_flag=""; if $resident_emissions; then resident_flag="--resident_emissions"; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n_process --master-addr 127.0.0.1 --master-port 29517 \
    src/iterative_utterance_alignment.py --tsv $tsv_path --vad_segments_tsv $vad_segments_filtered_filepath \
    --dst $results_dir --asr_hub $asr_hub --asr_savedir $asr_savedir --threshold $threshold \
    --logs_path $logs_dir --short_utterance_len $short_utterance_len --max_words_sequence $max_words_sequence \
    --max_window_size $max_window_size --window_to_stop $window_to_stop --min_text_to_audio_prop $min_text_to_audio_prop \
    --max_text_to_audio_prop_exec $max_text_to_audio_prop_exec $resident_flag > $logs_dir"/global.log"

if $merge_files; then
    echo "Merging aligned files from: "$results_dir
    python -u src/postprocess/merge_aligned_files.py --global_tsv $tsv_path --src $results_dir
fi

if $generate_stm_results; then
    echo "Generating stm files from: "$results_dir
    stm_dir=$results_dir/stm
    mkdir -p $stm_dir
    python -u src/scripts/tsv_to_stm.py --src_path $results_dir --dst_path $stm_dir
fi
