"""Word-level alignment -- drop-in for /root/reference/src/word_level_alignment.py
(same CLI :146-168; writes <tsv>_words.tsv next to the input like :141).  Rows are aligned in
batches on the GPU instead of one ``get_segments`` call per row."""
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
    parser = argparse.ArgumentParser(description="Script to generate word-level segmentation")
    parser.add_argument('--use_time_info', dest='time_info', action='store_true', help='use source temporal information')
    parser.add_argument("--asr_hub", help="ASR source path", default="")
    parser.add_argument("--asr_savedir", help="ASR save dir to store a symbolic link", default="")
    parser.add_argument("--tsv_path", help="metadata with filtered audio", default="")
    # accepted and unused, as in the reference (:157, :163): align_words.sh:96 passes --dst_path
    parser.add_argument("--dst_path", help="path to place results", default="")
    parser.add_argument('--offset_time', type=float, default=0.0, help='temporal shift in seconds of alignment')
    parser.add_argument("--left_offset", type=float, default=0.0, help='left offset in seconds')
    parser.add_argument("--right_offset", type=float, default=0.0, help='right offset in seconds')
    parser.add_argument('--collar', type=float, default=0.0, help='collar to apply to alignment in seconds')
    parser.add_argument("--logs_path", help="path to place logs", default="")
    main(parser.parse_args())
