"""The iterative anchor loop on top of the CUDA path.

Restates /root/reference/src/iterative_utterance_alignment.py:14-404
(``get_file_iterative_segmentation``) with one structural change: the reference's
candidate-iteration loop (:203-385) re-aligns the SAME emissions against a shrinking
utterance list, once per iteration; here ONE ``ipfa_ctcseg_device`` call aligns every
prefix of the list (one table fill), ``ipfa_anchor_select_device`` evaluates the
accept / shrink / revert state machine (:221-379) on the device, and only the decision
(16 bytes) plus the accepted prefix's segments return to the host.  Results are the
rows the reference would have appended to ``file_alignments``.
"""
import numpy as np
import torch

from . import hostglue as hg
from . import ops
from .ctc_segmentation import CTCSegmentation

SEL_ACCEPT_CURRENT, SEL_KEEP_PREVIOUS, SEL_DISCARD_ALL, SEL_LAST_SEGMENT = 0, 1, 2, 3


def _fmt2(x):
    return float(f"{x:.2f}")


def _fmt4(x):
    return float(f"{x:3.4f}")


def rows_from_decision(decision, seg_k, texts, clip_start, new_segment_start, threshold,
                       short_utterance_len, file_id, audio_path, channel, speaker_id, database):
    """Turn one window's device decision into (rows, new_segment_start, discarded).

    ``seg_k``: float64 [k, 3] segments of the accepted prefix (None when k == 0);
    ``texts``: the window's utterance list.  Values go through the same
    ``.2f`` / ``.4f`` text round trip as ``str(task)`` -> ``float()`` (:219-230)."""
    k, _, _, anchor_u = (int(v) for v in decision)
    rows = []
    for u in range(k):
        text = texts[u]
        start, end, score = _fmt2(seg_k[u, 0]), _fmt2(seg_k[u, 1]), _fmt4(seg_k[u, 2])
        abs_start, abs_end = clip_start + start, clip_start + end
        if len(text) < short_utterance_len:  # :241 short utterances are never anchors
            score += 2 * threshold
        segment_id = "_".join([file_id, str(abs_start), str(abs_end)])
        rows.append([segment_id, audio_path, channel, end - start, abs_start, abs_end, score, text,
                     speaker_id, database])
    if anchor_u == -1:
        new_segment_start = clip_start
    elif anchor_u >= 0:
        new_segment_start = clip_start + _fmt2(seg_k[anchor_u, 1])
    discarded = list(reversed(texts[k:]))  # dropped last-first, like the appends at :273,:282,:350,:373
    return rows, new_segment_start, discarded


def align_windows(tasks, is_last, threshold=-2.0, short_utterance_len=30):
    """Batched core: all prefixes of every task + on-device selection.

    Returns ``(decision int32 [N,4] (host), seg_accepted list of float64 [k,3] or None, status [N])``.
    Only the decisions, the status words and the accepted segments cross PCIe."""
    res, n_utts = CTCSegmentation.prefix_segments(tasks)
    kmax = res.seg.shape[1]
    text_len = np.zeros((len(tasks), kmax), np.int32)
    for i, t in enumerate(tasks):
        text_len[i, :len(t.text)] = [len(x) for x in t.text]
    decision, _anchor = ops.anchor_select(res.seg, n_utts, text_len, np.asarray(is_last, np.int32),
                                          threshold=threshold, short_len=short_utterance_len)
    decision = decision.cpu().numpy()
    status = res.status.cpu().numpy()
    # gather the accepted prefix of every window in one indexed read
    idx_w = torch.arange(len(tasks), device=res.seg.device)
    idx_k = torch.as_tensor(np.maximum(decision[:, 0] - 1, 0), device=res.seg.device, dtype=torch.long)
    picked = res.seg[idx_w, idx_k].cpu().numpy()  # [N, kmax, 3]
    segs = [picked[i, :decision[i, 0]] if decision[i, 0] > 0 else None for i in range(len(tasks))]
    return decision, segs, status


def align_window(aligner, transcript, lpz, name, n_samples, clip_start, is_last_segment,
                 new_segment_start, threshold, short_utterance_len, file_id, audio_path, channel,
                 speaker_id, database, logger=None):
    """Lines :203-385 for one window.  Raises ``AssertionError`` when the audio is shorter
    than the text (caught by the caller like the reference does at :390)."""
    task = aligner.prepare_segmentation_task(transcript, lpz, name, n_samples)
    if len(task.ground_truth_mat) > lpz.shape[0]:
        raise AssertionError("Audio is shorter than text!")
    decision, segs, status = align_windows([task], [is_last_segment], threshold, short_utterance_len)
    if status[0] & 4:
        raise AssertionError("Audio is shorter than text!")
    rows, nss, discarded = rows_from_decision(decision[0], segs[0], task.text, clip_start,
                                              new_segment_start, threshold, short_utterance_len,
                                              file_id, audio_path, channel, speaker_id, database)
    if logger is not None:
        for r in rows:
            logger.debug('{0} | {1} | {2} | {3}'.format(round(r[4], 3), round(r[5], 3), round(r[6], 3), r[7]))
        logger.debug('Not included transcripts: ' + str(discarded))
    return rows, nss, discarded, int(decision[0][1])


def get_file_iterative_segmentation(asr_model, aligner, audio_path, file_df, vad_file_df,
                                    samples_to_frames_ratio, logs_path, threshold=-2.0,
                                    short_utterance_len=30, max_words_sequence=24, min_words_sequence=None,
                                    max_window_size=70.0, window_to_stop=500.0, min_text_to_audio_prop=0.8,
                                    max_text_to_audio_prop_exec=10, window_fn=None, console_log=False):
    """Anchor loop over one audio file (:14-404).  ``window_fn`` replaces :195-385 (default:
    :func:`align_window`, the CUDA path; the parity tests pass the CPU oracle here)."""
    window_fn = window_fn or align_window
    log_name = audio_path.split('/')[-1].replace('.wav', '')
    logger = hg.alignment_logger(logs_path, f"{log_name}", console=console_log)
    logger.debug('Starting iterative alignment for file: ' + str(audio_path))

    new_segment_start = None
    discarded_transcripts = []
    n_segments = len(file_df.index)
    list_of_splits = []
    file_alignments = []

    info = hg.audio_info(audio_path)
    real_audio_length = info.num_frames / info.sample_rate
    logger.debug('Audio length: ' + str(round(real_audio_length, 2)))
    logger.debug('Labels length: ' + str(round(float(file_df.iloc[n_segments - 1]['End']), 2)))

    file_df = hg.fix_time_reference(file_df, vad_file_df, real_audio_length, n_segments)
    n_rows = len(file_df.index)
    exceptions_counter = 0
    file_id = audio_path.split('/')[-1].replace('.wav', '')
    text_to_audio_proportion = 0.0
    next_row_is_non_speech = False
    following_row = None
    audio_normalized, audio_length, sr = None, 0, info.sample_rate

    for row_index in range(n_rows):
        row = file_df.iloc[row_index]
        if row['Type'] == 'Non-Speech':  # :73-77
            new_segment_start = float(row['End'])
            continue

        is_last_segment = (row_index + 1) == n_rows
        clip_start = new_segment_start if new_segment_start is not None else float(row['Start'])
        clip_end = float(row['End'])
        clip_length = clip_end - clip_start
        database = row['Database']
        transcript = hg.prepare_text(str(row['Transcription']).upper(), max_words_sequence=max_words_sequence)
        if isinstance(transcript, str):
            transcript = [transcript]
        list_of_splits.append(len(transcript))
        if discarded_transcripts:  # :94-96
            transcript = discarded_transcripts[::-1] + transcript
            discarded_transcripts = []
        text_length = hg.count_text_length(transcript)
        try:  # :101-109 (a clip shorter than one sample divides by zero; the value of the previous row stays)
            text_to_audio_proportion = hg.get_text_to_audio_proportion(
                int(clip_length * info.sample_rate), text_length, info.sample_rate)
        except ZeroDivisionError:
            pass
        if not is_last_segment:
            following_row = file_df.iloc[row_index + 1]
            next_row_is_non_speech = following_row['Type'] == 'Non-Speech'

        speech_ending = (text_to_audio_proportion > 10.0 and next_row_is_non_speech and following_row is not None
                         and abs(float(following_row['Start']) - clip_start) > 5.0)
        recalculate_time_references = clip_length >= max_window_size or speech_ending  # :119-123
        if clip_length >= window_to_stop:  # :125
            break
        if recalculate_time_references:
            logger.debug('Recalculating time references, using last anchor as beginning...')
            file_df = hg.fix_text_to_time_proportion(
                file_df, vad_file_df, real_audio_length - clip_start,
                hg.get_n_aligned_rows(list_of_splits, len(file_alignments)), n_segments, clip_start, logger)
            row = file_df.iloc[row_index]
            clip_end = float(row['End'])
            clip_length = clip_end - clip_start

        try:  # :147-159
            audio, sr = hg.audio_load(audio_path, frame_offset=int(clip_start * info.sample_rate),
                                      num_frames=int(clip_length * info.sample_rate), channels_first=False)
            audio_normalized = asr_model.audio_normalizer(audio, sr)
            audio_length = audio.shape[0]
        except RuntimeError:
            # A clip that ends before it starts (Non-Speech rows land one row early when a file has
            # several VAD gaps, alignment_utils.py:160-169) makes torchaudio.load refuse its num_frames;
            # the reference prints and goes on with the audio of the previous clip of the file.
            if audio_normalized is None:
                raise NameError("name 'audio_normalized' is not defined")  # what the reference dies of

        if audio_length > 0:
            text_to_audio_proportion = hg.get_text_to_audio_proportion(audio_length, text_length, sr)
        logger.debug('Text to audio proportion: ' + str(text_to_audio_proportion))

        if not is_last_segment:  # :167-192
            if not audio_length > 0:
                discarded_transcripts = transcript[::-1]
                continue
            elif text_to_audio_proportion < min_text_to_audio_prop:
                logger.debug('Low quantity of text compared to audio, reading more audio and text...')
                discarded_transcripts = transcript[::-1]
                new_segment_start = clip_start
                continue
            if text_to_audio_proportion > 10.0 and next_row_is_non_speech and \
                    abs(float(following_row['Start']) - clip_start) > 5.0:
                transcript, discarded_transcripts = hg.find_a_valid_text_to_audio_proportion(
                    audio_length, transcript, samples_to_frames_ratio)

        try:
            lpz = aligner.get_lpz(audio_normalized)  # :201
            rows, new_segment_start, newly_discarded, _ = window_fn(
                aligner, transcript, lpz, row['Sample_ID'], audio_normalized.shape[0], clip_start,
                is_last_segment, new_segment_start, threshold, short_utterance_len, file_id, audio_path,
                row['Channel'], row['Speaker_ID'], database, logger)
            file_alignments += rows
            discarded_transcripts = discarded_transcripts + newly_discarded
            exceptions_counter = 0
        except AssertionError as e:  # :390-402
            logger.debug(e)
            discarded_transcripts += transcript[::-1]
            exceptions_counter += 1
            if exceptions_counter >= max_text_to_audio_prop_exec:
                break
            continue
    for h in list(logger.handlers):
        h.close()
        logger.removeHandler(h)
    return file_alignments


RESULT_COLUMNS = ['Sample_ID', 'Sample_Path', 'Channel', 'Audio_Length', 'Start', 'End', 'Segment_Score',
                  'Transcription', 'Speaker_ID', 'Database']
