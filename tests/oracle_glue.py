"""Glue between the product's host mirror and the CPU oracle (test infrastructure): the
candidate-iteration loop of /root/reference/src/iterative_utterance_alignment.py:203-385 run the
way the reference runs it -- one CPU alignment per shrinking-transcript iteration
(``oracle.anchor.anchor_window`` over ``oracle.ctcseg.get_segments``) -- with the signature of
``anchor.align_window`` so it can stand in for the CUDA window function."""
import numpy as np


def oracle_cfg(task):
    from oracle import ctcseg as oseg
    return oseg.CtcSegmentationParameters(index_duration=task.config.index_duration,
                                          score_min_mean_over_L=task.config.score_min_mean_over_L,
                                          char_list=task.config.char_list)


def oracle_fields(aligner, text, lpz, name, n_samples):
    """``str(task).split`` fields of one CPU alignment (raises AssertionError like the reference)."""
    from oracle import ctcseg as oseg
    lp = lpz.cpu().numpy() if hasattr(lpz, "cpu") else np.asarray(lpz)
    task = aligner.prepare_segmentation_task(text, lpz, name, n_samples)
    res = oseg.get_segments(oracle_cfg(task), lp, task.ground_truth_mat, task.utt_begin_indices, task.text)
    return [s.split(" ", 5) for s in oseg.task_str(name, task.text, res["segments"]).strip().split("\n")]


def oracle_window_fn(aligner, transcript, lpz, name, n_samples, clip_start, is_last_segment, new_segment_start,
                     threshold, short_utterance_len, file_id, audio_path, channel, speaker_id, database,
                     logger=None):
    from oracle.anchor import anchor_window
    rows, nss, disc, n_iter = anchor_window(
        transcript, lambda tr: oracle_fields(aligner, tr, lpz, name, n_samples), clip_start, is_last_segment,
        new_segment_start, [], threshold, short_utterance_len, file_id, audio_path,
        {"Channel": channel, "Speaker_ID": speaker_id, "Database": database})
    return rows, nss, disc, n_iter
