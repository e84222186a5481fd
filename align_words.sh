#!/bin/bash
#
# Word-level alignment -- same config zone, WIP layout and two stages as the reference's
# align_words.sh (:43-57 config, :64-84 directories, :86-96 stages): search the wanted words in the
# transcriptions, then align them.  Input: a TSV with the columns listed in that file's header
# (Sample_ID, Sample_Path, Channel, Audio_Length, Start, End, Transcription, Speaker_ID, Database).

#########################################################
###################### DEFINITIONS ######################
#########################################################

# config zone
config_file=config/words.json           # json config file: contains an array with the wanted words
alignment_name="sample_words"           # alignment name, comment to use timestamp instead
tsv_path=data/wip_sample/results/sample_aligned.tsv  # source file with metadata
text_column="Transcription"             # column name in tsv that contains the utterance text reference

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

tsv_filename=$(basename $tsv_path)
filtered_tsv_dir=$results_dir"/"${tsv_filename/.tsv/_filtered.tsv}

echo "Searching words in source data..."
python -u src/search_words.py --tsv_path $tsv_path --dst $results_dir \
    --config_file $config_file --text_column $text_column

echo "Starting word-level alignment..."
python -u src/word_level_alignment.py --tsv_path $filtered_tsv_dir \
    --dst_path $results_dir --asr_hub $asr_hub --asr_savedir $asr_savedir \
    --logs_path $logs_dir --use_time_info --collar $collar --offset_time $offset_time \
    --left_offset $left_offset --right_offset $right_offset
