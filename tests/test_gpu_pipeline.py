"""End-to-end drop-in parity on the GPU: the anchor loop over a synthetic file
(BASELINE configs[0]) and the word-level / search-on-speech row loops, CUDA path vs the
same host glue driven by the CPU oracle (oracle/ctcseg + oracle/anchor)."""
import importlib

import numpy as np
import pandas as pd
import pytest

from scripted_asr import ScriptedASR, make_schedule, ramp_audio

pytestmark = pytest.mark.gpu

PKG = "iterative-pseudo-forced-alignment-ctc_b200"
hg = importlib.import_module(PKG + ".hostglue")
anchor = importlib.import_module(PKG + ".anchor")
words = importlib.import_module(PKG + ".words")
cs = importlib.import_module(PKG + ".ctc_segmentation")

WORDS = ("hola que tal estamos aqui para probar el alineamiento forzado iterativo con anclas "
         "sobre un audio largo y un texto que no siempre coincide con lo que se dice").split()


def _utterances(rng, n):
    return [" ".join(rng.choice(WORDS, size=int(rng.integers(6, 14)))) for _ in range(n)]


def _oracle_cfg(task):
    from oracle import ctcseg as oseg
    return oseg.CtcSegmentationParameters(index_duration=task.config.index_duration,
                                          score_min_mean_over_L=task.config.score_min_mean_over_L,
                                          char_list=task.config.char_list)


def oracle_window_fn(aligner, transcript, lpz, name, n_samples, clip_start, is_last_segment, new_segment_start,
                     threshold, short_utterance_len, file_id, audio_path, channel, speaker_id, database, logger=None):
    """Lines :203-385 the way the reference runs them: one CPU alignment per iteration."""
    from oracle import ctcseg as oseg
    from oracle.anchor import anchor_window
    lp = lpz.cpu().numpy()

    def align_fn(tr):
        task = aligner.prepare_segmentation_task(tr, lpz, name, n_samples)
        res = oseg.get_segments(_oracle_cfg(task), lp, task.ground_truth_mat, task.utt_begin_indices, task.text)
        return [s.split(" ", 5) for s in oseg.task_str(name, task.text, res["segments"]).strip().split("\n")]

    rows, nss, disc, n_iter = anchor_window(
        transcript, align_fn, clip_start, is_last_segment, new_segment_start, [], threshold, short_utterance_len,
        file_id, audio_path, {"Channel": channel, "Speaker_ID": speaker_id, "Database": database})
    return rows, nss, disc, n_iter


def _make_file(d, n_utts, seed=11):
    rng = np.random.default_rng(seed)
    utts = _utterances(rng, n_utts)
    tok = importlib.import_module(PKG + ".stub_asr").CharTokenizer()
    frames, spans = make_schedule([u.upper() for u in utts], tok, rng)
    total = (len(frames) + 40) * 320
    wav = str(d / "talk.wav")
    hg.write_wav(wav, ramp_audio(total))
    dur = total / 16000
    # the TSV's own times are deliberately useless (the loop re-derives them, :53)
    df = pd.DataFrame({'Sample_ID': [f"talk_{i}" for i in range(n_utts)], 'Sample_Path': [wav] * n_utts,
                       'Channel': [1] * n_utts, 'Audio_Length': [dur / n_utts] * n_utts,
                       'Start': [0.0] * n_utts, 'End': [dur] * n_utts, 'Transcription': utts,
                       'Speaker_ID': ['spk1'] * n_utts, 'Database': ['synthetic'] * n_utts})
    vad = pd.DataFrame({'Sample_Path': [wav], 'Start': [0.0], 'End': [dur], 'Segment_Length': [dur]})
    # 16-bit PCM quantises the ramp: the emitter rounds to the nearest sample, exact up to ~2^15 samples/step
    return wav, df, vad, frames, spans, total, utts, str(d)


@pytest.fixture(scope="module")
def synthetic_file(tmp_path_factory):
    return _make_file(tmp_path_factory.mktemp("file"), 14)


def test_five_minute_file_config0(tmp_path):
    """BASELINE configs[0]: iterative_utterance_alignment on one synthetic ~5-min 16 kHz file."""
    wav, df, vad, frames, spans, total, utts, d = _make_file(tmp_path, 80, seed=5)
    assert total / 16000 > 280
    asr = ScriptedASR(frames, total, device="cuda", corrupt=((2000, 2300), (9000, 9100)))
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    aligner.samples_to_frames_ratio = 320.0
    kw = dict(threshold=-2.0, short_utterance_len=30, max_words_sequence=24, max_window_size=70.0)
    got = anchor.get_file_iterative_segmentation(asr, aligner, wav, df.copy(), vad.copy(), 320.0, d, **kw)
    ref = anchor.get_file_iterative_segmentation(asr, aligner, wav, df.copy(), vad.copy(), 320.0, d,
                                                 window_fn=oracle_window_fn, **kw)
    assert len(ref) >= 60 and got == ref


@pytest.mark.parametrize("corrupt", [(), ((300, 420),)])
def test_anchor_loop_file_matches_oracle(synthetic_file, corrupt):
    wav, df, vad, frames, spans, total, utts, d = synthetic_file
    asr = ScriptedASR(frames, total, device="cuda", corrupt=corrupt)
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    aligner.samples_to_frames_ratio = 320.0
    kw = dict(threshold=-2.0, short_utterance_len=30, max_words_sequence=8, max_window_size=70.0)
    got = anchor.get_file_iterative_segmentation(asr, aligner, wav, df.copy(), vad.copy(), 320.0, d, **kw)
    ref = anchor.get_file_iterative_segmentation(asr, aligner, wav, df.copy(), vad.copy(), 320.0, d,
                                                 window_fn=oracle_window_fn, **kw)
    assert len(ref) > 0
    assert got == ref  # identical rows: ids, times (0.01 s), scores (1e-4), texts
    out = hg.remove_artefacts(pd.DataFrame(got, columns=anchor.RESULT_COLUMNS), 30)
    assert list(out.columns) == anchor.RESULT_COLUMNS
    if not corrupt:
        # clean emissions: every utterance chunk is accepted and sits on its scheduled frames
        assert (out['Segment_Score'] > -2.0).mean() > 0.7
        assert out['End'].is_monotonic_increasing


def test_align_window_batch_equals_single(synthetic_file):
    """N windows in one launch == N single-window calls (ragged T, K)."""
    wav, df, vad, frames, spans, total, utts, d = synthetic_file
    asr = ScriptedASR(frames, total, device="cuda")
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    aligner.samples_to_frames_ratio = 320.0
    tasks, flags = [], []
    for i, (a, b) in enumerate([(0, 4), (2, 5), (5, 6), (6, 11)]):
        s0, s1 = spans[a][0] * 320, min(total, (spans[b - 1][1] + 10) * 320)
        audio, sr = hg.audio_load(wav, s0, s1 - s0)
        lpz = aligner.get_lpz(asr.audio_normalizer(audio, sr))
        tasks.append(aligner.prepare_segmentation_task([u.upper() for u in utts[a:b]], lpz, f"w{i}", s1 - s0))
        flags.append(i == 3)
    dec, segs, status = anchor.align_windows(tasks, flags)
    for i, t in enumerate(tasks):
        d1, s1_, st1 = anchor.align_windows([t], [flags[i]])
        assert np.array_equal(dec[i], d1[0]) and status[i] == st1[0]
        assert (segs[i] is None) == (s1_[0] is None)
        if segs[i] is not None:
            assert np.array_equal(segs[i], s1_[0])


def test_word_level_and_search_on_speech_match_oracle(synthetic_file):
    from oracle import ctcseg as oseg
    wav, df, vad, frames, spans, total, utts, d = synthetic_file
    asr = ScriptedASR(frames, total, device="cuda")
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed")
    aligner.samples_to_frames_ratio = 320.0
    rows = []
    for i, (a, b) in enumerate(spans[:8]):
        start, end = max(0, a - 10) * 0.02, (b + 10) * 0.02
        word = utts[i].split()[len(utts[i].split()) // 2].upper()
        rows.append({'Sample_ID': f"talk_{i}", 'Sample_Path': wav, 'Audio_Length': end - start, 'Start': start,
                     'End': end, 'Normalized_Transcription': utts[i].upper(), 'Wanted_Text': word,
                     'Speaker_ID': 'spk1', 'Database': 'synthetic'})
    wdf = pd.DataFrame(rows)
    got = words.align_words(aligner, asr, wdf, time_info=True, batch_size=3)

    def oracle_fields(text, row):
        info = hg.audio_info(wav)
        cs0, cl = float(row['Start']), float(row['End']) - float(row['Start'])
        audio, sr = hg.audio_load(wav, int(cs0 * info.sample_rate), int(cl * info.sample_rate))
        lpz = aligner.get_lpz(asr.audio_normalizer(audio, sr))
        task = aligner.prepare_segmentation_task(text, lpz, row['Sample_ID'], audio.shape[0])
        res = oseg.get_segments(_oracle_cfg(task), lpz.cpu().numpy(), task.ground_truth_mat,
                                task.utt_begin_indices, task.text)
        return [s.split(" ", 5) for s in oseg.task_str(row['Sample_ID'], task.text, res["segments"]).strip().split("\n")]

    k = 0
    for _, row in wdf.iterrows():
        for seg in oracle_fields(words.word_sentence(row['Normalized_Transcription'], row['Wanted_Text']), row):
            if seg[-1] == row['Wanted_Text']:
                g = got.iloc[k]
                assert g['Start'] == float(row['Start']) + float(seg[2]) and g['End'] == float(row['Start']) + float(seg[3])
                assert g['Segment_Score'] == float(seg[4]) and g['Word'] == row['Wanted_Text'].lower()
                k += 1
    assert k == len(got) and k >= 8
    # the wanted word is found close to where the schedule put it, with a confident score
    assert (got['Segment_Score'] > -5.0).all() and (got['Segment_Score'] > -1.5).mean() >= 0.5

    target = utts[2].split()[1].upper()
    sos = words.search_on_speech(aligner, asr, wdf, target, batch_size=4)
    assert len(sos) == len(wdf)
    for (_, row), (_, g) in zip(wdf.iterrows(), sos.iterrows()):
        seg = [s for s in oracle_fields("·" + target + "·", row) if s[-1] == "·" + target + "·"][0]
        assert g['Start'] == float(row['Start']) + float(seg[2]) and g['Segment_Score'] == float(seg[4])
