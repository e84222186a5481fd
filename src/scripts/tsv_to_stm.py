"""TSV -> STM lines "<file> <channel> <speaker> <start> <end> <,,> <text>" (reference: src/scripts/tsv_to_stm.py:32-39)."""
import argparse
import os

import pandas as pd


def main(args):
    if not (os.path.isdir(args.src_path) and os.path.isdir(args.dst_path)):
        print("Non valid arguments to convert tsv files to stm.")
        return
    for file in sorted(os.listdir(args.src_path)):
        if not file.endswith('.tsv') or file.startswith('.'):
            continue
        src = pd.read_csv(os.path.join(args.src_path, file), header=0, sep='\t')
        name = file.replace('.tsv', '')
        with open(os.path.join(args.dst_path, name + '.stm'), 'w') as stm:
            for _, row in src.iterrows():
                stm.write("{0} {1} {2} {3} {4} <,,> {5}\n".format(
                    name, row['Channel'], row['Speaker_ID'], round(row['Start'], 3), round(row['End'], 3),
                    str(row['Transcription']).lower()))


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description="Script to generate stm files from tsv")
    parser.add_argument('--src_path', default="")
    parser.add_argument("--dst_path", default="")
    main(parser.parse_args())
