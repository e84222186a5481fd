"""Concatenate per-file TSVs into <name>_aligned.tsv (reference: src/postprocess/merge_aligned_files.py:24-25)."""
import argparse
import os

import pandas as pd


def main(args):
    if not (os.path.isfile(args.global_tsv) and os.path.isdir(args.src)):
        print('Source file or source directory does not exist')
        return
    name = args.global_tsv.split('/')[-1].replace('.tsv', '')
    parts = []
    for audio_path in pd.read_csv(args.global_tsv, header=0, sep='\t')['Sample_Path'].unique():
        f = os.path.join(args.src, audio_path.split('/')[-1].replace('.wav', '.tsv'))
        if os.path.isfile(f) and os.path.getsize(f) > 0:
            parts.append(pd.read_csv(f, header=0, sep='\t'))
    if parts:
        pd.concat(parts, ignore_index=True).to_csv(os.path.join(args.src, name + '_aligned.tsv'), index=None, sep='\t')


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description="Script merge aligned files")
    parser.add_argument("--global_tsv", default="")
    parser.add_argument("--src", default="")
    main(parser.parse_args())
