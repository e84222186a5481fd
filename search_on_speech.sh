#!/bin/bash
#
# Search on speech -- same config zone, WIP layout and stage as the reference's search_on_speech.sh
# (:43-56 config, :63-79 directories, :81-85 stage): align one text against every segment of a TSV
# and keep the score.

#########################################################
###################### DEFINITIONS ######################
#########################################################

# config zone
alignment_name="sample_sos"             # alignment name, comment to use timestamp instead
tsv_path=data/wip_sample/results/sample_aligned.tsv  # source file with metadata
speech_to_search="conmigo"              # text that will be searched in all segments

# alignment corrections: better apply this after
collar=0.0                              # collar to alignment in seconds
offset_time=0.0                         # alignment shift to right in seconds
left_offset=0.0                         # start shift in seconds
right_offset=0.0                        # end shift in seconds

# trained ASR: a SpeechBrain EncoderASR source; "stub" = random-init emitter (no meaning, smoke runs only)
asr_hub="Voyager1/asr-wav2vec2-commonvoice-es"
asr_savedir="data/asr/"

#########################################################
####################### ALIGNMENT #######################
#########################################################

if [ ! -z ${alignment_name+set} ]; then
    wip_dir="data/wip_"$alignment_name
    echo "Alignment name defined, WIP folder is: "$wip_dir
else
    wip_dir="data/wip_"$(date +%s)
    echo "Alignment name not defined, WIP folder is: "$wip_dir
fi

results_dir=$wip_dir"/results"
logs_dir=$wip_dir"/logs"
mkdir -p $wip_dir $results_dir $logs_dir

echo "Starting search on speech..."
python -u src/search_on_speech.py --tsv_path $tsv_path \
    --dst_path $results_dir --asr_hub $asr_hub --asr_savedir $asr_savedir \
    --logs_path $logs_dir --text="$speech_to_search" --collar $collar --offset_time $offset_time \
    --left_offset $left_offset --right_offset $right_offset
