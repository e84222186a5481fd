"""Host-side bookkeeping around the alignment hot path, restated for this image.

Behavioural restatement of /root/reference/src/utils/alignment_utils.py and
src/utils/text_utils.py:81-95.  The reference code
itself cannot run here: pandas 3 has no ``DataFrame.append`` and rejects
``float(one_row_Series)``; torchaudio 2.11 has no ``torchaudio.info`` and its
``load`` needs the absent torchcodec -- so frames are
built with ``pd.concat`` / ``.iloc[i]['col']`` and audio comes from stdlib ``wave``.
Nothing here touches the GPU.
"""
import logging
import os
import sys
import wave

import numpy as np
import pandas as pd
import torch

ROW_COLUMNS = ['Sample_ID', 'Sample_Path', 'Audio_Length', 'Start', 'End', 'Transcription', 'Speaker_ID',
               'Database', 'Channel', 'Text_Length', 'Type']


def alignment_logger(logs_path, logger_name, level=logging.DEBUG, console=True):
    """Per-file logger to stdout and <logs_path>/<name>.log (alignment_utils.py:10-32)."""
    logger = logging.getLogger(logger_name)
    fmt = logging.Formatter("%(asctime)s [%(name)s] %(message)s")
    logger.setLevel(level)
    logger.propagate = False
    for h in list(logger.handlers):
        logger.removeHandler(h)
    if console:
        ch = logging.StreamHandler(sys.stdout)
        ch.setLevel(level)
        ch.setFormatter(fmt)
        logger.addHandler(ch)
    if logs_path:
        os.makedirs(logs_path, exist_ok=True)
        fh = logging.FileHandler(os.path.join(logs_path, logger_name + '.log'), mode='w')
        fh.setLevel(level)
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    return logger


# ----------------------------------------------------------------------------- text
def split_long_transcript(transcript, max_words_sequence=24):
    """Chunks of at most ``max_words_sequence`` words (text_utils.py:81-95)."""
    words = transcript.split(' ')
    chunks = []
    for i in range(int(len(words) / max_words_sequence) + 1):
        chunk = ' '.join(words[i * max_words_sequence:(i + 1) * max_words_sequence]).strip()
        if chunk:
            chunks.append(chunk)
    return chunks


def prepare_text(transcript, max_words_sequence=None, min_words_sequence=None):
    """str -> list of utterances (alignment_utils.py:35-68)."""
    if max_words_sequence or min_words_sequence:
        if max_words_sequence:
            if len(transcript.split(' ')) > max_words_sequence:
                transcript = split_long_transcript(transcript, max_words_sequence=max_words_sequence)
            else:
                transcript = [transcript]
        if min_words_sequence:
            raise Exception("Min word sequence not implemented")
    return transcript


def get_n_aligned_rows(list_of_splits, n_aligned_splits):
    """alignment_utils.py:71-77."""
    index = 0
    for i in range(1, len(list_of_splits)):
        if sum(list_of_splits[:-i]) <= n_aligned_splits:
            index = i
            break
    return len(list_of_splits[:-index])


def count_text_length(transcript):
    return len(" ".join(transcript))


def get_text_to_audio_proportion(audio_length, text_length, sample_rate):
    """> 1: more text than audio (80 ms per character, x3 margin; alignment_utils.py:84-106)."""
    maximum_text_duration_samples = text_length * 0.08 * 3 * sample_rate
    return maximum_text_duration_samples / audio_length


def find_a_valid_text_to_audio_proportion(audio_length, transcript, samples_to_frames_ratio):
    """Drop trailing utterances until the text fits the frames (alignment_utils.py:174-196)."""
    original = transcript
    max_chars = int(audio_length / samples_to_frames_ratio)
    dropped = []
    for _ in range(1, len(transcript) + 1):
        if count_text_length(transcript) < max_chars:
            return transcript, dropped
        dropped.append(transcript[-1])
        transcript = transcript[:-1]
    return original, []


# ----------------------------------------------------------------------------- time references
def insert_row(idx, df, values):
    """Insert one row at position idx keeping the fixed column order (alignment_utils.py:109-115)."""
    if isinstance(values, pd.Series):
        values = [values[c] for c in ROW_COLUMNS]
    new = pd.DataFrame([list(values)], columns=ROW_COLUMNS)
    top, bottom = df.iloc[:idx], df.iloc[idx:]
    frames = [f for f in (top, new, bottom) if len(f.index)]
    out = pd.concat(frames, ignore_index=True, sort=False)
    return out.reset_index(drop=True)


def _spread_text_over_speech(file_df, vad_df, first, n_segments, total_text, speech_length, acc,
                             real_audio_length):
    """Shared loop of fix_time_reference / fix_text_to_time_proportion: every row gets audio
    in proportion to its characters, jumping over non-speech gaps."""
    vad_i = 0
    gap_after = []
    for index in range(first, n_segments):
        share = file_df.loc[index, 'Text_Length'] / total_text * speech_length
        file_df.loc[index, 'Start'] = acc
        file_df.loc[index, 'End'] = acc + share
        acc += share
        speech_end = float(vad_df.iloc[vad_i]['End'])
        if acc >= speech_end:
            file_df.loc[index, 'End'] = speech_end
            if vad_i + 1 < len(vad_df.index):
                acc = float(vad_df.iloc[vad_i + 1]['Start'])
                vad_i += 1
                gap_after.append(index)
        if index + 1 == n_segments and acc < real_audio_length:
            file_df.loc[index, 'End'] = real_audio_length
    return file_df, gap_after


def _insert_non_speech(file_df, vad_df, gap_after, legacy_order):
    sample = file_df.iloc[0]
    path, channel, database = str(sample['Sample_Path']), int(sample['Channel']), str(sample['Database'])
    for i, idx in enumerate(gap_after):
        end_i = float(vad_df.iloc[i]['End'])
        start_next = float(vad_df.iloc[i + 1]['Start'])
        if legacy_order:
            # alignment_utils.py:164-167 pairs these values with ROW_COLUMNS positionally
            values = ['Non-speech-' + str(i), path, start_next - end_i, end_i, start_next, "Non-Speech",
                      "Non-Speech", database, channel, 0, 'Non-Speech']
        else:
            # alignment_utils.py:261-265 lists channel third; positional pairing is kept
            values = ['Non-speech-' + str(i), path, channel, start_next - end_i, end_i, start_next,
                      "Non-Speech", "Non-Speech", database, 0, 'Non-Speech']
        file_df = insert_row(idx + 1, file_df, values)
    return file_df


def fix_time_reference(file_df, vad_file_df, real_audio_length, n_segments):
    """Initial proportional time references + Non-Speech rows (alignment_utils.py:118-171)."""
    file_df = file_df.copy()
    for col in ('Start', 'End'):
        file_df[col] = file_df[col].astype(float)
    file_df['Text_Length'] = file_df['Transcription'].astype(str).apply(len)
    total_text = file_df['Text_Length'].sum()
    speech_length = vad_file_df['Segment_Length'].sum()
    file_df, gaps = _spread_text_over_speech(file_df, vad_file_df, 0, n_segments, total_text, speech_length,
                                             0.0, real_audio_length)
    file_df['Type'] = 'Speech'
    file_df = file_df[[c for c in ROW_COLUMNS if c in file_df.columns]]
    return _insert_non_speech(file_df, vad_file_df, gaps, legacy_order=True)


def fix_text_to_time_proportion(file_df, vad_file_df, real_audio_length, n_aligned, n_segments,
                                last_anchor_time, logger):
    """Re-spread the not-yet-aligned text from the last anchor on (alignment_utils.py:199-274)."""
    total_text = file_df.iloc[n_aligned:]['Text_Length'].sum()
    vad = vad_file_df[vad_file_df['End'] > last_anchor_time].reset_index(drop=True).copy()
    vad.loc[0, 'Start'] = last_anchor_time
    vad['Segment_Length'] = vad['End'] - vad['Start']
    speech_length = vad['Segment_Length'].sum()
    logger.debug('Remaining audio: {0} | Remaining text: {1} | Remaining speech length {2}'.format(
        real_audio_length, total_text, speech_length))
    logger.debug('Aligned index: {0}, Total segments: {1}'.format(n_aligned, n_segments))
    old_non_speech = file_df[file_df['Type'] == 'Non-Speech']
    file_df = file_df[file_df['Type'] == 'Speech'].reset_index(drop=True)
    file_df, gaps = _spread_text_over_speech(file_df, vad, n_aligned, n_segments, total_text, speech_length,
                                             last_anchor_time, real_audio_length)
    file_df['Type'] = 'Speech'
    file_df = _insert_non_speech(file_df, vad, gaps, legacy_order=False)
    if len(old_non_speech.index) > len(gaps):
        rows = old_non_speech.iterrows()
        for _ in range(len(old_non_speech.index) - len(gaps)):
            index, row = next(rows)
            file_df = insert_row(index, file_df, row)
    return file_df


def remove_artefacts(df, length):
    """Give short utterances their +4.0 back (alignment_utils.py:277-281; = -2*threshold at the default)."""
    short = df['Transcription'].apply(len) < length
    df.loc[short, 'Segment_Score'] = df.loc[short, 'Segment_Score'] + 4.0
    return df


# ----------------------------------------------------------------------------- audio
class AudioInfo:
    def __init__(self, num_frames, sample_rate, num_channels):
        self.num_frames, self.sample_rate, self.num_channels = num_frames, sample_rate, num_channels


def audio_info(path):
    """Stand-in for torchaudio.info (absent in torchaudio 2.11): PCM WAV header through stdlib wave."""
    with wave.open(path, 'rb') as w:
        return AudioInfo(w.getnframes(), w.getframerate(), w.getnchannels())


def audio_load(path, frame_offset=0, num_frames=-1, channels_first=False):
    """Stand-in for torchaudio.load(..., frame_offset, num_frames, channels_first=False) of the
    reference's torchaudio==0.11.0 (sox_io backend): float32 in [-1, 1).  Like that backend it raises
    ``RuntimeError`` for ``frame_offset < 0`` and for ``num_frames`` other than -1 or a positive count;
    the anchor loop depends on that (see ``anchor.get_file_iterative_segmentation``)."""
    if frame_offset < 0:
        raise RuntimeError("Invalid argument: frame_offset must be non-negative.")
    if num_frames is None:
        num_frames = -1
    if not (num_frames == -1 or num_frames > 0):
        raise RuntimeError("Invalid argument: num_frames must be -1 or greater than 0.")
    with wave.open(path, 'rb') as w:
        sr, ch, width, total = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        frame_offset = min(int(frame_offset), total)
        n = total - frame_offset if num_frames == -1 else min(int(num_frames), total - frame_offset)
        w.setpos(frame_offset)
        raw = w.readframes(n)
    if width == 2:
        data = np.frombuffer(raw, dtype='<i2').astype(np.float32) / 32768.0
    elif width == 4:
        data = np.frombuffer(raw, dtype='<i4').astype(np.float32) / 2147483648.0
    elif width == 1:
        data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported sample width {width}")
    audio = torch.from_numpy(data.reshape(-1, ch).copy())
    return (audio.t().contiguous() if channels_first else audio), sr


def write_wav(path, samples, sample_rate=16000):
    """Mono 16-bit PCM writer (synthetic fixtures for tests / config 1)."""
    pcm = np.clip(np.asarray(samples, dtype=np.float64), -1.0, 1.0 - 1.0 / 32768)
    pcm = (pcm * 32768.0).astype('<i2')
    with wave.open(path, 'wb') as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())
