"""Iterative pseudo-forced alignment, utterance level -- drop-in for
/root/reference/src/iterative_utterance_alignment.py (same CLI, same per-file TSV).

Differences: the numerics run on the B200 path (one table fill per window serves every
candidate iteration; the accept/shrink/revert decision is taken on the device), and the
``n_process`` OS processes that claimed files through empty TSVs become one process per
GPU (``torchrun``) with deterministic length-balanced file shards."""
import argparse
import os

import pandas as pd

from _common import CTCSegmentation, anchor, hostglue, load_asr, rank_world, sharding


def main(args):
    import torch
    rank, world = rank_world()
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    asr_model = load_asr(args.asr_hub, args.asr_savedir)
    aligner = CTCSegmentation(asr_model, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    samples_to_frames_ratio = aligner.estimate_samples_to_frames_ratio()
    aligner.samples_to_frames_ratio = samples_to_frames_ratio

    df_path = os.path.normpath(args.tsv)
    vad_path = os.path.normpath(args.vad_segments_tsv)
    if not (os.path.isfile(df_path) and os.path.isfile(vad_path)):
        print('{0} or {1} file does not exists, please create it.'.format(df_path, vad_path))
        return
    df = pd.read_csv(df_path, header=0, sep='\t')
    vad_df = pd.read_csv(vad_path, header=0, sep='\t')
    audio_paths = list(df['Sample_Path'].unique())
    # shard by file, heaviest first (cost ~ audio length x text length)
    costs = [float(df[df['Sample_Path'] == p]['End'].max()) * max(1, int(df[df['Sample_Path'] == p]
             ['Transcription'].astype(str).str.len().sum())) for p in audio_paths]
    mine = sharding.lpt_shards(costs, world)[rank]
    os.makedirs(args.dst, exist_ok=True)

    todo = []
    for i in mine:
        audio_path = audio_paths[i]
        tsv_result_file = os.path.join(args.dst, audio_path.split('/')[-1].replace('.wav', '.tsv'))
        if not (os.path.isfile(tsv_result_file) and os.path.getsize(tsv_result_file) > 0):
            todo.append(i)
    resident = {}
    if args.resident_emissions and todo:
        # every file encoded once, the anchor loops of all files in lock step on the device;
        # files that reach the reference's time re-spreading branch (:119-146) are redone below
        from _common import sweep as sweep_mod
        jobs = [(audio_paths[i], df[df['Sample_Path'] == audio_paths[i]].reset_index(drop=True),
                 vad_df[vad_df['Sample_Path'] == audio_paths[i]].reset_index(drop=True)) for i in todo]
        rows_per_file, status = sweep_mod.align_files_resident(
            asr_model, aligner, jobs, samples_to_frames_ratio, threshold=args.threshold,
            short_utterance_len=args.short_utterance_len, max_words_sequence=args.max_words_sequence,
            max_window_size=args.max_window_size, window_to_stop=args.window_to_stop,
            min_text_to_audio_prop=args.min_text_to_audio_prop,
            max_text_to_audio_prop_exec=args.max_text_to_audio_prop_exec)
        for i, rows, st in zip(todo, rows_per_file, status):
            if st in (sweep_mod.DONE, sweep_mod.STOP_WINDOW, sweep_mod.STOP_EXCEPTIONS):
                resident[i] = rows

    for i in mine:
        audio_path = audio_paths[i]
        tsv_result_file = os.path.join(args.dst, audio_path.split('/')[-1].replace('.wav', '.tsv'))
        if os.path.isfile(tsv_result_file) and os.path.getsize(tsv_result_file) > 0:
            print('File ' + tsv_result_file + ' already exist, skipping the alignment generation.')
            continue
        file_df = df[df['Sample_Path'] == audio_path].reset_index(drop=True)
        vad_file_df = vad_df[vad_df['Sample_Path'] == audio_path].reset_index(drop=True)
        rows = resident[i] if i in resident else anchor.get_file_iterative_segmentation(
            asr_model, aligner, audio_path, file_df, vad_file_df, samples_to_frames_ratio,
            logs_path=args.logs_path, threshold=args.threshold, short_utterance_len=args.short_utterance_len,
            max_words_sequence=args.max_words_sequence, min_words_sequence=args.min_words_sequence,
            max_window_size=args.max_window_size, window_to_stop=args.window_to_stop,
            min_text_to_audio_prop=args.min_text_to_audio_prop,
            max_text_to_audio_prop_exec=args.max_text_to_audio_prop_exec)
        out = pd.DataFrame(rows, columns=anchor.RESULT_COLUMNS)
        out = hostglue.remove_artefacts(out, args.short_utterance_len)
        out.to_csv(tsv_result_file, sep='\t', index=None)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description="Iterative pseudo-forced alignment algorithm")
    parser.add_argument("--tsv", help="tsv file with comming from a stm file", default="")
    parser.add_argument("--vad_segments_tsv", help="tsv file of filtered speech segments", default="")
    parser.add_argument("--dst", help="path to place results", default="")
    parser.add_argument("--logs_path", help="path to place logs", default="")
    parser.add_argument("--asr_hub", help="ASR source path", default="")
    parser.add_argument("--asr_savedir", help="ASR save dir to store a symbolic link", default="")
    parser.add_argument('--threshold', type=float, default=-2.0, help='alignment threshold')
    parser.add_argument('--short_utterance_len', type=int, default=30)
    parser.add_argument('--max_words_sequence', type=int, default=24)
    parser.add_argument('--min_words_sequence', type=int, default=None)
    parser.add_argument('--max_window_size', type=float, default=70.0)
    parser.add_argument('--window_to_stop', type=float, default=500.0)
    parser.add_argument('--min_text_to_audio_prop', type=float, default=0.8)
    parser.add_argument('--max_text_to_audio_prop_exec', type=int, default=10)
    parser.add_argument('--resident_emissions', action='store_true',
                        help='encode every file once and run the anchor loops of all files on the GPU '
                             '(windows are frame slices of the file-level emissions)')
    main(parser.parse_args())
