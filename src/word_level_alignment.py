"""Word-level alignment -- drop-in for /root/reference/src/word_level_alignment.py
(same CLI; writes <tsv>_words.tsv).  Rows are aligned in batches on the GPU."""
import argparse

import pandas as pd

from _common import CTCSegmentation, hostglue, load_asr, words


def main(args):
    log_name = args.tsv_path.split('/')[-1].replace('.tsv', '')
    logger = hostglue.alignment_logger(args.logs_path, f"{log_name}")
    logger.debug('Starting word alignment for file: ' + str(args.tsv_path))
    asr_model = load_asr(args.asr_hub, args.asr_savedir)
    aligner = CTCSegmentation(asr_model, kaldi_style_text=False, time_stamps="fixed")
    df = pd.read_csv(args.tsv_path, header=0, sep='\t')
    out = words.align_words(aligner, asr_model, df, time_info=args.time_info, offset_time=args.offset_time,
                            left_offset=args.left_offset, right_offset=args.right_offset, logger=logger)
    out.to_csv(args.tsv_path.replace('_filtered.tsv', '_words.tsv'), sep='\t', index=None)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description="Word-level alignment")
    parser.add_argument("--tsv_path", default="")
    parser.add_argument("--logs_path", default="")
    parser.add_argument("--asr_hub", default="")
    parser.add_argument("--asr_savedir", default="")
    parser.add_argument("--offset_time", type=float, default=0.0)
    parser.add_argument("--left_offset", type=float, default=0.0)
    parser.add_argument("--right_offset", type=float, default=0.0)
    parser.add_argument("--use_time_info", dest="time_info", action="store_true")
    main(parser.parse_args())
