"""Search on speech -- drop-in for /root/reference/src/search_on_speech.py (same CLI :132-154;
writes <dst_path>/<tsv>_sos.tsv like :126-127).  One target text against every segment, batched."""
import argparse
import os

import pandas as pd

from _common import CTCSegmentation, hostglue, load_asr, words


def main(args):
    log_name = args.tsv_path.split('/')[-1].replace('.tsv', '')
    logger = hostglue.alignment_logger(args.logs_path, f"{log_name}")
    logger.debug('Starting word search in file: ' + str(args.tsv_path))
    if args.text == '':  # :25-28
        raise Exception("Sorry, empty text cannot be searched on speech.")
    wanted_text = words.normalize_transcript(args.text).upper()
    asr_model = load_asr(args.asr_hub, args.asr_savedir)
    aligner = CTCSegmentation(asr_model, kaldi_style_text=False, time_stamps="fixed")
    df = pd.read_csv(args.tsv_path, header=0, sep='\t')
    out = words.search_on_speech(aligner, asr_model, df, wanted_text, offset_time=args.offset_time,
                                 left_offset=args.left_offset, right_offset=args.right_offset, logger=logger)
    tsv_name = args.tsv_path.split('/')[-1].replace('.tsv', '')
    out.to_csv(os.path.join(args.dst_path, tsv_name + '_sos.tsv'), sep='\t', index=None)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description="Script to search words in speech")
    parser.add_argument("--asr_hub", help="ASR source path", default="")
    parser.add_argument("--asr_savedir", help="ASR save dir to store a symbolic link", default="")
    parser.add_argument("--tsv_path", help="metadata with audio segments", default="")
    parser.add_argument("--dst_path", help="path to place results", default="")
    parser.add_argument('--offset_time', type=float, default=0.0, help='temporal shift in seconds of alignment')
    parser.add_argument("--left_offset", type=float, default=0.0, help='left offset in seconds')
    parser.add_argument("--right_offset", type=float, default=0.0, help='right offset in seconds')
    parser.add_argument('--collar', type=float, default=0.0, help='collar to apply to alignment in seconds')
    parser.add_argument("--logs_path", help="path to place logs", default="")
    parser.add_argument("--text", help="text that we want to search on speech", default="")
    main(parser.parse_args())
