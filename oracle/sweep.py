"""oracle/sweep.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's per-file anchor loop,
/root/reference/src/iterative_utterance_alignment.py:67-402, over files whose emissions
were computed once (BASELINE.json configs[4]): the clip the
reference would cut from the audio and re-encode (:149-201) is the frame slice
``lpz[int(clip_start*sr) // frame_shift : ... + audio_samples // frame_shift]`` of the
file's emissions.  The candidate loop (:203-385) is ``oracle.anchor.anchor_window`` -- one
CPU alignment per shrinking-transcript iteration, the way the reference runs it -- and the
alignment is ``oracle.ctcseg.get_segments``.

The reference's ``fix_text_to_time_proportion`` branch (:119-146, pandas + VAD table) is
host policy outside this restatement: the file stops with status ``needs_recalc`` and the
rows found so far, exactly where the CUDA sweep hands the file back to the host.
"""
import numpy as np

from . import anchor as oanchor
from . import ctcseg as oseg


def _text_to_audio(audio_length, text_length, sample_rate):
    # alignment_utils.py:84-106
    return text_length * 0.08 * 3 * sample_rate / audio_length


def _find_a_valid_text_to_audio_proportion(audio_length, transcript, samples_to_frames_ratio):
    # alignment_utils.py:174-196
    original = transcript
    max_chars = int(audio_length / samples_to_frames_ratio)
    dropped = []
    for _ in range(1, len(transcript) + 1):
        if len(" ".join(transcript)) < max_chars:
            return transcript, dropped
        dropped.append(transcript[-1])
        transcript = transcript[:-1]
    return original, []


def sweep_file(file_id, audio_path, lpz, n_samples, rows, tokenizer, index_duration=0.02,
               samples_to_frames_ratio=320.0, frame_shift=320, sample_rate=16000, threshold=-2.0,
               short_utterance_len=30, max_window_size=70.0, window_to_stop=500.0,
               min_text_to_audio_prop=0.8, max_text_to_audio_prop_exec=10, scoring_length=30, blank=0,
               recalc_fn=None):
    """Returns ``(file_alignments, status, stats)``; ``rows`` as in the product's ``SweepFile.rows``.
    ``recalc_fn(row_index, rows, clip_start) -> rows`` stands for ``fix_text_to_time_proportion``
    (:127-146); without it the file stops with status ``needs_recalc`` at that point."""
    lpz = np.asarray(lpz, dtype=np.float32)
    cfg = oseg.CtcSegmentationParameters(index_duration=index_duration, score_min_mean_over_L=scoring_length,
                                         blank=blank)
    unk = tokenizer.unk_id() if hasattr(tokenizer, "unk_id") else -1
    new_segment_start = None
    discarded_transcripts = []
    file_alignments = []
    exceptions_counter = 0
    text_to_audio_proportion = 0.0
    next_row_is_non_speech = False
    following_row = None
    last_clip = None
    n_rows = len(rows)
    status = "done"
    stats = {"windows": 0, "alignments": 0, "cells": 0, "frames": 0}

    for row_index in range(n_rows):
        row = rows[row_index]
        if row["Type"] == "Non-Speech":  # :73-77
            new_segment_start = float(row["End"])
            continue
        is_last_segment = (row_index + 1) == n_rows
        clip_start = new_segment_start if new_segment_start is not None else float(row["Start"])
        clip_end = float(row["End"])
        clip_length = clip_end - clip_start
        transcript = list(row["utterances"])
        if discarded_transcripts:  # :94-96
            transcript = discarded_transcripts[::-1] + transcript
            discarded_transcripts = []
        text_length = len(" ".join(transcript))
        try:  # :101-109
            text_to_audio_proportion = _text_to_audio(int(clip_length * sample_rate), text_length, sample_rate)
        except ZeroDivisionError:
            pass
        if not is_last_segment:
            following_row = rows[row_index + 1]
            next_row_is_non_speech = following_row["Type"] == "Non-Speech"
        speech_ending = (text_to_audio_proportion > 10.0 and next_row_is_non_speech and following_row is not None
                         and abs(float(following_row["Start"]) - clip_start) > 5.0)
        recalculate_time_references = clip_length >= max_window_size or speech_ending  # :119-123
        if clip_length >= window_to_stop:  # :125
            status = "window_to_stop"
            break
        if recalculate_time_references:
            if recalc_fn is None:
                status = "needs_recalc"
                break
            rows = recalc_fn(row_index, rows, clip_start)  # :127-146
            row = rows[row_index]
            clip_end = float(row["End"])
            clip_length = clip_end - clip_start

        # :149-160 torchaudio.load(frame_offset, num_frames) of torchaudio==0.11 clamps to the file and
        # refuses num_frames other than -1 or > 0; the reference then keeps the previous clip (:157-159)
        offset, num_frames = int(clip_start * sample_rate), int(clip_length * sample_rate)
        if offset >= 0 and (num_frames > 0 or num_frames == -1):
            offset = min(offset, n_samples)
            audio_length = n_samples - offset if num_frames == -1 else min(num_frames, n_samples - offset)
            last_clip = (offset, audio_length)
        elif last_clip is None:
            raise NameError("name 'audio_length' is not defined")  # :162 of the reference
        else:
            offset, audio_length = last_clip
        if audio_length > 0:
            text_to_audio_proportion = _text_to_audio(audio_length, text_length, sample_rate)
        if not is_last_segment:  # :167-192
            if not audio_length > 0:
                discarded_transcripts = transcript[::-1]
                continue
            elif text_to_audio_proportion < min_text_to_audio_prop:
                discarded_transcripts = transcript[::-1]
                new_segment_start = clip_start
                continue
            if text_to_audio_proportion > 10.0 and next_row_is_non_speech and \
                    abs(float(following_row["Start"]) - clip_start) > 5.0:
                transcript, discarded_transcripts = _find_a_valid_text_to_audio_proportion(
                    audio_length, transcript, samples_to_frames_ratio)

        f0 = offset // frame_shift
        n_frames = min(audio_length // frame_shift, max(0, lpz.shape[0] - f0))
        window = lpz[f0:f0 + n_frames]
        try:
            def align_fn(tr):
                token_list = []
                for utt in tr:
                    ids = np.asarray(tokenizer.encode_as_ids(utt))
                    token_list.append(ids[ids != unk] if ids.size else ids)
                gt, ub = oseg.prepare_token_list(cfg, token_list)
                res = oseg.get_segments(cfg, window, gt, ub, tr)
                stats["alignments"] += 1
                if stats["_first"]:
                    stats["_first"] = False
                    stats["cells"] += window.shape[0] * len(gt)
                    stats["frames"] += window.shape[0]
                return [s.split(" ", 5) for s in
                        oseg.task_str(row.get("Sample_ID", file_id), tr, res["segments"]).strip().split("\n")]

            stats["_first"] = True
            got, new_segment_start, discarded_transcripts, _ = oanchor.anchor_window(
                transcript, align_fn, clip_start, is_last_segment, new_segment_start, discarded_transcripts,
                threshold, short_utterance_len, file_id, audio_path,
                {"Channel": row.get("Channel"), "Speaker_ID": row.get("Speaker_ID"),
                 "Database": row.get("Database")})
            file_alignments += got
            stats["windows"] += 1
            exceptions_counter = 0
        except AssertionError:  # :390-402
            discarded_transcripts += transcript[::-1]
            exceptions_counter += 1
            if exceptions_counter >= max_text_to_audio_prop_exec:
                status = "exceptions_limit"
                break
            continue
    stats.pop("_first", None)
    return file_alignments, status, stats
