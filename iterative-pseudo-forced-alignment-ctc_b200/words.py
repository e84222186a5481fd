"""Word-level alignment and search-on-speech on the CUDA path.

Restates the row loops of /root/reference/src/word_level_alignment.py:35-141 and
/root/reference/src/search_on_speech.py:45-127.  Every TSV row is an independent
utterance, so instead of one ``get_segments`` call per row the rows are aligned
in batches (``CTCSegmentation.get_segments_batch`` -> one fill + one backtrace
launch per batch).  Outputs keep the reference's columns and value formatting.
"""
import string
import re

import pandas as pd

from . import hostglue as hg
from .ctc_segmentation import CTCSegmentation

WORD_COLUMNS = ['Sample_ID', 'Sample_Path', 'Audio_Length', 'Start', 'End', 'Segment_Score', 'Transcription',
                'Speaker_ID', 'Word', 'Database']
SOS_COLUMNS = ['Sample_ID', 'Sample_Path', 'Audio_Length', 'Start', 'End', 'Segment_Score', 'Speaker_ID', 'Word',
               'Database']


def _spanish_number_words():
    """``num2words(n, lang='es')`` when the package is installed (text_utils.py:4, :39-46)."""
    try:
        from num2words import num2words
    except ImportError:
        return None
    return lambda n: num2words(n, lang='es')


def normalize_transcript(transcript, number_to_words=None):
    """text_utils.py:49-78.  Numbers are spelled out through ``number_to_words`` (int -> str),
    by default num2words' Spanish like the reference; where that package is missing (this image)
    a transcript that contains digits is an error rather than a silently different ground truth."""
    out = re.sub(r"<font color=\"#[0-9a-fA-F]{6}\">", "", transcript)
    out = re.sub(r"</font>", "", out).replace('\n', ' ')
    out = out.translate(str.maketrans('', '', string.punctuation)).lower()
    out = out.replace('!', '').replace('¡', '').replace('?', '').replace('¿', '')
    out = out.replace('   ', ' ').replace('  ', ' ')
    numbers = [int(s) for s in out.split() if s.isdigit()]
    if numbers:
        number_to_words = number_to_words or _spanish_number_words()
        if number_to_words is None:
            raise ImportError("the transcript contains numbers and num2words is not installed "
                              "(pass number_to_words=...)")
        for number in numbers:
            out = out.replace(str(number), number_to_words(number))
    return out


def word_sentence(sentence, wanted_text, normalize=True):
    """[pre, '·', WORD, '·', post, '·'] with empties removed (word_level_alignment.py:69-84)."""
    base = normalize_transcript(sentence).upper() if normalize else sentence
    parts = base.split(wanted_text)
    parts.insert(1, wanted_text)
    parts = [p.strip() for p in parts if p != ""]
    out = []
    for i in range(1, 2 * len(parts)):
        out.append("·" if (i + 1) % 2 else parts[int(i / 2)])
    out.append("·")
    return out


def _parse(task):
    return [seg.split(" ", 5) for seg in str(task).strip().split("\n")]


def _run_batches(aligner, asr_model, jobs, batch_size):
    """jobs: list of dicts with audio_path/clip_start/clip_length/text/name.  Yields (job, fields|AssertionError)."""
    for b in range(0, len(jobs), batch_size):
        chunk = jobs[b:b + batch_size]
        tasks, kept = [], []
        for job in chunk:
            info = hg.audio_info(job['audio_path'])
            audio, sr = hg.audio_load(job['audio_path'], frame_offset=int(job['clip_start'] * info.sample_rate),
                                      num_frames=int(job['clip_length'] * info.sample_rate), channels_first=False)
            audio_n = asr_model.audio_normalizer(audio, sr)
            lpz = aligner.get_lpz(audio_n)
            task = aligner.prepare_segmentation_task(job['text'], lpz, job['name'], audio_n.shape[0])
            if len(task.ground_truth_mat) > lpz.shape[0]:
                yield job, AssertionError("Audio is shorter than text!")
                continue
            tasks.append(task)
            kept.append(job)
        for job, task, res in zip(kept, tasks, CTCSegmentation.get_segments_batch(tasks)):
            if isinstance(res, AssertionError):
                yield job, res
                continue
            task.set(**res)
            yield job, _parse(task)


def align_words(aligner, asr_model, df, time_info=True, offset_time=0.0, left_offset=0.0, right_offset=0.0,
                batch_size=256, logger=None):
    """DataFrame of <tsv>_filtered.tsv rows -> DataFrame of *_words.tsv rows."""
    jobs = []
    for _, row in df.iterrows():
        clip_start = float(row['Start']) if time_info else 0.0
        clip_length = float(row['End']) - clip_start if time_info else float(row['Audio_Length'])
        jobs.append({'audio_path': row['Sample_Path'], 'clip_start': clip_start, 'clip_length': clip_length,
                     'text': word_sentence(row['Normalized_Transcription'], row['Wanted_Text']),
                     'name': row['Sample_ID'], 'row': row})
    out = []
    for job, fields in _run_batches(aligner, asr_model, jobs, batch_size):
        row = job['row']
        wanted_text = row['Wanted_Text']
        if isinstance(fields, AssertionError):
            if logger:
                logger.debug(fields)
            continue
        audio_name = row['Sample_Path'].split('/')[-1]
        ext = audio_name.split('.')[-1]
        for seg in fields:
            if len(seg) != 6 or seg[-1] != wanted_text:
                continue
            start = float(seg[2]) + offset_time + left_offset
            end = float(seg[3]) + offset_time + right_offset
            abs_start, abs_end = job['clip_start'] + start, job['clip_start'] + end
            sample_id = "_".join([audio_name.replace(ext, ''), str(abs_start), str(abs_end)])
            out.append([sample_id, row['Sample_Path'], end - start, abs_start, abs_end, float(seg[4]),
                        row['Normalized_Transcription'], row['Speaker_ID'], wanted_text.lower(), row['Database']])
    return pd.DataFrame(out, columns=WORD_COLUMNS)


def search_on_speech(aligner, asr_model, df, wanted_text, offset_time=0.0, left_offset=0.0, right_offset=0.0,
                     batch_size=256, logger=None):
    """One target against every segment (search_on_speech.py:45-127) -> *_sos.tsv rows."""
    sentence = "·" + wanted_text.strip() + "·"
    jobs = []
    for _, row in df.iterrows():
        clip_start = float(row['Start'])
        jobs.append({'audio_path': row['Sample_Path'], 'clip_start': clip_start,
                     'clip_length': float(row['End']) - clip_start, 'text': sentence, 'name': row['Sample_ID'],
                     'row': row})
    out = []
    for job, fields in _run_batches(aligner, asr_model, jobs, batch_size):
        row = job['row']
        if isinstance(fields, AssertionError):
            if logger:
                logger.debug(fields)
            continue
        audio_name = row['Sample_Path'].split('/')[-1]
        ext = audio_name.split('.')[-1]
        for seg in fields:
            if len(seg) != 6 or seg[-1] != sentence:
                continue
            start = float(seg[2]) + offset_time + left_offset
            end = float(seg[3]) + offset_time + right_offset
            abs_start, abs_end = job['clip_start'] + start, job['clip_start'] + end
            sample_id = "_".join([audio_name.replace(ext, ''), str(abs_start), str(abs_end)])
            out.append([sample_id, row['Sample_Path'], end - start, abs_start, abs_end, float(seg[4]),
                        row['Speaker_ID'], wanted_text.lower(), row['Database']])
    return pd.DataFrame(out, columns=SOS_COLUMNS)
