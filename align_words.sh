#!/bin/bash
# Word-level alignment driver (reference: align_words.sh:43-57, :91-96).
tsv_path=data/sample/tsv/sample.tsv
config_file=config/words.json
dst=data/wip_words; logs_dir=$dst/logs
asr_hub="stub"; asr_savedir="data/asr/"
mkdir -p $dst $logs_dir
python -u src/search_words.py --tsv_path $tsv_path --dst $dst --config_file $config_file
name=$(basename $tsv_path .tsv)
python -u src/word_level_alignment.py --tsv_path $dst/${name}_filtered.tsv --logs_path $logs_dir \
    --asr_hub $asr_hub --asr_savedir $asr_savedir --use_time_info
