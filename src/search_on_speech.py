"""Search on speech -- drop-in for /root/reference/src/search_on_speech.py
(same CLI; writes <tsv>_sos.tsv).  One target text against every segment, batched."""
import argparse
import os

import pandas as pd

from _common import CTCSegmentation, hostglue, load_asr, words


def main(args):
    log_name = args.tsv_path.split('/')[-1].replace('.tsv', '')
    logger = hostglue.alignment_logger(args.logs_path, f"{log_name}")
    asr_model = load_asr(args.asr_hub, args.asr_savedir)
    aligner = CTCSegmentation(asr_model, kaldi_style_text=False, time_stamps="fixed")
    df = pd.read_csv(args.tsv_path, header=0, sep='\t')
    out = words.search_on_speech(aligner, asr_model, df, args.text.upper(), offset_time=args.offset_time,
                                 left_offset=args.left_offset, right_offset=args.right_offset, logger=logger)
    tsv_name = args.tsv_path.split('/')[-1].replace('.tsv', '')
    out.to_csv(os.path.join(args.dst_path, tsv_name + '_sos.tsv'), sep='\t', index=None)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description="Search on speech")
    parser.add_argument("--tsv_path", default="")
    parser.add_argument("--dst_path", default="")
    parser.add_argument("--logs_path", default="")
    parser.add_argument("--text", default="")
    parser.add_argument("--asr_hub", default="")
    parser.add_argument("--asr_savedir", default="")
    parser.add_argument("--offset_time", type=float, default=0.0)
    parser.add_argument("--left_offset", type=float, default=0.0)
    parser.add_argument("--right_offset", type=float, default=0.0)
    main(parser.parse_args())
