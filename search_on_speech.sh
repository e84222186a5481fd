#!/bin/bash
# Search-on-speech driver (reference: search_on_speech.sh:43-56).
tsv_path=data/sample/tsv/sample.tsv
text="hola"
dst=data/wip_sos; logs_dir=$dst/logs
asr_hub="stub"; asr_savedir="data/asr/"
mkdir -p $dst $logs_dir
python -u src/search_on_speech.py --tsv_path $tsv_path --dst_path $dst --logs_path $logs_dir --text "$text" \
    --asr_hub $asr_hub --asr_savedir $asr_savedir
