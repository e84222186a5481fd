"""oracle/anchor.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement of the candidate-iteration loop of the reference's anchor search,
/root/reference/src/iterative_utterance_alignment.py:195-385 (SURVEY.md section 8(a)
row A10): the same emissions are re-aligned against a shrinking list of
utterances, and a threshold state machine decides to accept, drop the last
utterance and repeat, or revert to the previous iteration.

``align_fn(transcript)`` stands for lines 208-219 (prepare task, get_segments,
``str(task)`` split into 6 fields) and returns ``list_of_segments``: one
``[utt_name, name, start, end, score, text]`` list of strings per utterance.
"""


def anchor_window(transcript, align_fn, clip_start, is_last_segment, new_segment_start,
                  discarded_transcripts, threshold=-2.0, short_utterance_len=30,
                  file_id="file", audio_path="file.wav", row_meta=None, log=None):
    """Runs lines 195-385 for one window.

    Returns ``(rows, new_segment_start, discarded_transcripts, n_iterations)``
    where ``rows`` are the ``file_alignments`` entries this window contributed
    (lines 258-260 after all splices)."""
    row_meta = row_meta or {"Channel": 1, "Speaker_ID": "spk", "Database": "db"}
    log = log or (lambda *_: None)
    transcript = list(transcript)
    discarded_transcripts = list(discarded_transcripts)
    file_alignments = []

    bad_alignment = True
    repeat_alignment = True
    previous_segmentation = []
    previous_absolute_end = 0.0
    n_iterations = 0
    segment_score = None
    absolute_end = None

    while bad_alignment or repeat_alignment:  # :203
        n_iterations += 1
        list_of_segments = align_fn(transcript)  # :208-219

        for segment in list_of_segments:  # :221
            if len(segment) != 6:
                continue
            segment_transcript = segment[-1]
            segment_start = float(segment[2])
            segment_end = float(segment[3])
            segment_score = float(segment[4])
            absolute_start = clip_start + segment_start
            absolute_end = clip_start + segment_end
            segment_length = segment_end - segment_start
            segment_id = "_".join([file_id, str(absolute_start), str(absolute_end)])
            if len(segment_transcript) < short_utterance_len:  # :241
                segment_score += 2 * threshold
            if segment_score < threshold:  # :245
                bad_alignment = True
            else:
                bad_alignment = False
                new_segment_start = absolute_end
            file_alignments.append([
                segment_id, audio_path, row_meta["Channel"], segment_length, absolute_start,
                absolute_end, segment_score, segment_transcript, row_meta["Speaker_ID"],
                row_meta["Database"]])

        K = len(list_of_segments)
        if is_last_segment:  # :263
            bad_alignment = False
            repeat_alignment = False
        else:
            if bad_alignment and repeat_alignment and not previous_segmentation:  # :269
                if not transcript[:-1]:  # :272
                    discarded_transcripts.append(transcript[-1])
                    bad_alignment = False
                    repeat_alignment = False
                    file_alignments = file_alignments[:len(file_alignments) - K]
                    new_segment_start = clip_start
                else:  # :281
                    discarded_transcripts.append(transcript[-1])
                    transcript = transcript[:-1]
                    file_alignments = file_alignments[:len(file_alignments) - K]
                    log("repeat-low-score")
            elif (not bad_alignment or previous_segmentation) and repeat_alignment:  # :289
                if previous_segmentation:
                    previous_score = float(previous_segmentation[-2][4])  # :293
                    if len(previous_segmentation[-2][-1]) < short_utterance_len:
                        previous_score += 2 * threshold
                    if segment_score > -1.0 and not (previous_score == segment_score):  # :298
                        new_segment_start = absolute_end
                        bad_alignment = False
                        repeat_alignment = False
                        file_alignments = (file_alignments[:len(file_alignments) - (2 * K + 1)]
                                           + file_alignments[len(file_alignments) - K:])
                    elif previous_score >= segment_score:  # :306
                        repeat_alignment = False
                        bad_alignment = False
                        new_segment_start = previous_absolute_end
                        discarded_transcripts = discarded_transcripts[:-1]
                        file_alignments = file_alignments[:len(file_alignments) - K]
                    else:  # :316
                        if not transcript[:-1]:  # :319
                            if not bad_alignment:
                                new_segment_start = absolute_end
                                file_alignments = (
                                    file_alignments[:len(file_alignments) - (2 * K + 1)]
                                    + file_alignments[len(file_alignments) - K:])
                            else:
                                discarded_transcripts.append(transcript[-1])
                                file_alignments = file_alignments[:len(file_alignments) - (2 * K + 1)]
                                new_segment_start = clip_start
                            bad_alignment = False
                            repeat_alignment = False
                        else:  # :336
                            if bad_alignment:  # :340
                                repeat_alignment = False
                                bad_alignment = False
                                new_segment_start = previous_absolute_end
                                discarded_transcripts = discarded_transcripts[:-1]
                                file_alignments = file_alignments[:len(file_alignments) - K]
                            else:  # :348
                                previous_segmentation = list_of_segments
                                previous_absolute_end = absolute_end
                                discarded_transcripts.append(transcript[-1])
                                transcript = transcript[:-1]
                                file_alignments = (
                                    file_alignments[:len(file_alignments) - (2 * K + 1)]
                                    + file_alignments[len(file_alignments) - K:])
                else:  # :357 first repetition
                    if segment_score > -1.0:  # :360
                        new_segment_start = absolute_end
                        bad_alignment = False
                        repeat_alignment = False
                    elif not transcript[:-1]:  # :367
                        new_segment_start = absolute_end
                        bad_alignment = False
                        repeat_alignment = False
                    else:  # :372
                        discarded_transcripts.append(transcript[-1])
                        transcript = transcript[:-1]
                        previous_segmentation = list_of_segments
                        previous_absolute_end = absolute_end

    return file_alignments, new_segment_start, discarded_transcripts, n_iterations
