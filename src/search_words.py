"""Filter a TSV for rows containing the wanted words -- drop-in for
/root/reference/src/search_words.py (pure text; restated for pandas 3: no DataFrame.append)."""
import argparse
import json
import os

import pandas as pd

from _common import words


def main(args):
    wanted_words = json.load(open(args.config_file, 'r'))["words"]
    df = pd.read_csv(os.path.join(args.tsv_path), header=0, sep='\t')
    df['Normalized_Transcription'] = df[args.text_column].apply(lambda x: words.normalize_transcript(x).upper())
    if wanted_words != ["*"]:
        frames, wanted = [], []
        for word in wanted_words:
            hit = df[df['Normalized_Transcription'].str.contains(word.upper(), regex=False)]
            frames.append(hit)
            wanted += [word.upper()] * len(hit.index)
        filtered = pd.concat(frames) if frames else df.iloc[:0].copy()
        filtered['Wanted_Text'] = wanted
    else:
        df = df[~df["Normalized_Transcription"].str.contains("UNKNOWN")]
        df = df[~df["Normalized_Transcription"].str.contains("UNTRANSCRIBED")]
        filtered = df.copy()
        filtered['Wanted_Text'] = filtered['Normalized_Transcription']
    print('Found following occurrences: \n' + str(filtered['Wanted_Text'].value_counts()))
    filtered = filtered.drop_duplicates(keep='first')
    tsv_name = args.tsv_path.split('/')[-1].replace('.tsv', '')
    filtered.to_csv(os.path.join(args.dst, tsv_name + '_filtered.tsv'), sep='\t', index=None)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description="Script to search wanted words in a tsv file")
    parser.add_argument("--tsv_path", default="")
    parser.add_argument("--dst", default="")
    parser.add_argument("--config_file", default="config/words.json")
    parser.add_argument("--text_column", default="Transcription")
    main(parser.parse_args())
