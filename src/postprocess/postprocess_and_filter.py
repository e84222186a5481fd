"""Score-threshold filter with offsets and collar -- drop-in for
/root/reference/src/postprocess/postprocess_and_filter.py:54-77 (same CLI, same output name)."""
import argparse
import os
import wave

import pandas as pd


def _audio_seconds(path):
    """torchaudio.info(path).num_frames / sample_rate (:22-23) from the WAV header."""
    with wave.open(path, 'rb') as w:
        return w.getnframes() / w.getframerate()


def main(args):
    if not os.path.isfile(args.tsv):
        print('tsv file does not exist ({0})'.format(args.tsv))
        return
    name = args.tsv.split('/')[-1].replace('.tsv', '')
    df = pd.read_csv(args.tsv, header=0, sep='\t')
    keep = df['Segment_Score'] >= args.score if args.comp == 'gt' else df['Segment_Score'] < args.score
    out = df[keep].copy()
    out['Start'] = out['Start'] + args.offset_time + args.left_offset
    out['End'] = out['End'] + args.offset_time + args.right_offset
    if args.collar > 0.0:
        delta = args.collar / 2
        out['Start'] = out['Start'].apply(lambda x: x - delta if (x - delta) > 0.0 else 0.0)
        lengths = {}
        ends = []
        for _, row in out.iterrows():
            path = row['Sample_Path']
            if path not in lengths:
                lengths[path] = _audio_seconds(path)
            end = float(row['End'])
            ends.append(lengths[path] if (end + delta) > lengths[path] else end + delta)
        out = out.drop('End', axis=1)
        out['End'] = ends
    out['Audio_Length'] = out['End'] - out['Start']
    print('Total audio length {0} seconds'.format(out['Audio_Length'].sum()))
    out.to_csv(os.path.join(os.path.dirname(args.tsv), name + '_' + args.comp + '_' + str(args.score) + '_filtered.tsv'),
               index=None, sep='\t')


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Script that post-processes aligned files")
    parser.add_argument("--tsv", default="")
    parser.add_argument('--score', type=float, default=-1.0)
    parser.add_argument('--comp', type=str, default="gt", choices=['gt', 'lt'])
    parser.add_argument('--offset_time', type=float, default=0.0)
    parser.add_argument("--left_offset", type=float, default=0.0)
    parser.add_argument("--right_offset", type=float, default=0.0)
    parser.add_argument('--collar', type=float, default=0.0)
    main(parser.parse_args())
