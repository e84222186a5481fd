"""CPU tests of the host glue restated from /root/reference/src/utils/alignment_utils.py,
text_utils.py, search_words.py, tsv_to_stm.py, merge_aligned_files.py."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hg = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.hostglue")
words = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.words")
anchor = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.anchor")


def test_prepare_text_and_split():
    text = " ".join(f"w{i}" for i in range(50))
    parts = hg.prepare_text(text, max_words_sequence=24)
    assert [len(p.split(' ')) for p in parts] == [24, 24, 2]
    assert hg.prepare_text("A B C", max_words_sequence=24) == ["A B C"]
    assert hg.count_text_length(["AB", "CDE"]) == 6
    with pytest.raises(Exception):
        hg.prepare_text("A", min_words_sequence=3)


def test_text_to_audio_proportion_and_trim():
    # 100 chars * 80 ms * 3 = 24 s of "maximum text duration" against 12 s of audio
    assert hg.get_text_to_audio_proportion(12 * 16000, 100, 16000) == pytest.approx(2.0)
    kept, dropped = hg.find_a_valid_text_to_audio_proportion(320 * 30, ["A" * 20, "B" * 20, "C" * 5], 320.0)
    assert kept == ["A" * 20] and dropped == ["C" * 5, "B" * 20]
    kept, dropped = hg.find_a_valid_text_to_audio_proportion(320 * 5, ["A" * 20], 320.0)
    assert kept == ["A" * 20] and dropped == []
    assert hg.get_n_aligned_rows([2, 1, 3, 1], 3) == 2


def _file_df(n=6):
    return pd.DataFrame({
        'Sample_ID': [f"s{i}" for i in range(n)], 'Sample_Path': ['a/b/file.wav'] * n, 'Channel': [1] * n,
        'Audio_Length': [1.0] * n, 'Start': [float(i) for i in range(n)], 'End': [float(i + 1) for i in range(n)],
        'Transcription': ["x" * (10 * (i + 1)) for i in range(n)], 'Speaker_ID': ['spk'] * n,
        'Database': ['db'] * n})


def test_fix_time_reference_spreads_text_and_inserts_non_speech():
    vad = pd.DataFrame({'Sample_Path': ['a/b/file.wav'] * 2, 'Start': [0.0, 60.0], 'End': [50.0, 100.0],
                        'Segment_Length': [50.0, 40.0]})
    out = hg.fix_time_reference(_file_df(), vad, 100.0, 6)
    speech = out[out['Type'] == 'Speech'].reset_index(drop=True)
    non = out[out['Type'] == 'Non-Speech']
    assert len(speech) == 6 and len(non) == 1
    # proportional to characters: 10..60 of 210 chars over 90 s of speech
    assert speech.loc[0, 'Start'] == 0.0 and speech.loc[0, 'End'] == pytest.approx(10 / 210 * 90)
    assert float(non.iloc[0]['Start']) == 50.0 and float(non.iloc[0]['End']) == 60.0
    assert speech.iloc[-1]['End'] == 100.0
    # rows stay ordered in time
    ends = out['End'].astype(float).values
    assert np.all(np.diff(ends) >= -1e-9)


def test_insert_row_and_remove_artefacts():
    df = hg.fix_time_reference(_file_df(3), pd.DataFrame({'Start': [0.0], 'End': [9.0], 'Segment_Length': [9.0]}),
                               9.0, 3)
    row = ['n0', 'p', 1.0, 2.0, 3.0, 'Non-Speech', 'Non-Speech', 'db', 1, 0, 'Non-Speech']
    out = hg.insert_row(1, df, row)
    assert len(out) == 4 and out.loc[1, 'Sample_ID'] == 'n0' and out.loc[2, 'Sample_ID'] == 's1'
    res = pd.DataFrame({'Transcription': ["short", "x" * 40], 'Segment_Score': [-4.5, -0.5]})
    res = hg.remove_artefacts(res, 30)
    assert res['Segment_Score'].tolist() == [-0.5, -0.5]


def test_wav_round_trip(tmp_path):
    x = np.sin(np.arange(16000) / 20.0) * 0.5
    p = str(tmp_path / "t.wav")
    hg.write_wav(p, x)
    info = hg.audio_info(p)
    assert (info.num_frames, info.sample_rate, info.num_channels) == (16000, 16000, 1)
    audio, sr = hg.audio_load(p, frame_offset=1000, num_frames=500, channels_first=False)
    assert audio.shape == (500, 1) and sr == 16000
    np.testing.assert_allclose(audio[:, 0].numpy(), x[1000:1500], atol=1 / 32768 + 1e-7)


def test_word_sentence_layout():
    s = words.word_sentence("hola, buenos días a todos", "BUENOS")
    assert s == ["HOLA", "·", "BUENOS", "·", "DÍAS A TODOS", "·"]
    assert words.word_sentence("buenos días", "BUENOS") == ["BUENOS", "·", "DÍAS", "·"]
    assert words.normalize_transcript('<font color="#ffffff">¡Hola,   mundo!</font>') == "hola mundo"


def test_rows_from_decision_formats_like_the_reference():
    seg = np.array([[0.014, 1.236, -0.123456], [1.236, 2.5, -1.9]])
    rows, nss, disc = anchor.rows_from_decision((2, 1, 0, 1), seg, ["x" * 40, "short", "dropped"], 10.0, None,
                                                -2.0, 30, "file", "p/file.wav", 1, "spk", "db")
    assert rows[0][:2] == ["file_10.01_11.24", "p/file.wav"]
    assert rows[0][4:7] == [10.01, 11.24, -0.1235]
    assert rows[1][6] == pytest.approx(-1.9 - 4.0)  # short utterance penalty, :241
    assert nss == 12.5 and disc == ["dropped"]
    rows, nss, disc = anchor.rows_from_decision((0, 2, 2, -1), None, ["a", "b"], 3.0, 7.0, -2.0, 30, "f", "p", 1,
                                                "s", "d")
    assert rows == [] and nss == 3.0 and disc == ["b", "a"]


def test_search_words_and_stm_cli(tmp_path):
    tsv = tmp_path / "set.tsv"
    pd.DataFrame({'Sample_ID': ['a', 'b', 'c'], 'Sample_Path': ['x.wav'] * 3, 'Channel': [1] * 3,
                  'Start': [0.0, 1.0, 2.0], 'End': [1.0, 2.0, 3.123456], 'Speaker_ID': ['s'] * 3,
                  'Transcription': ['Hola mundo', 'adios', 'hola otra vez'], 'Database': ['d'] * 3}
                 ).to_csv(tsv, sep='\t', index=None)
    cfg = tmp_path / "words.json"
    cfg.write_text(json.dumps({"words": ["hola"]}))
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, os.path.join(ROOT, "src", "search_words.py"), "--tsv_path", str(tsv), "--dst",
                    str(tmp_path), "--config_file", str(cfg)], check=True, env=env, capture_output=True)
    out = pd.read_csv(tmp_path / "set_filtered.tsv", sep='\t')
    assert out['Sample_ID'].tolist() == ['a', 'c'] and out['Wanted_Text'].tolist() == ['HOLA', 'HOLA']
    stm_dir = tmp_path / "stm"
    stm_dir.mkdir()
    subprocess.run([sys.executable, os.path.join(ROOT, "src", "scripts", "tsv_to_stm.py"), "--src_path", str(tmp_path),
                    "--dst_path", str(stm_dir)], check=True, env=env, capture_output=True)
    lines = (stm_dir / "set.stm").read_text().splitlines()
    assert lines[2] == "set 1 s 2.0 3.123 <,,> hola otra vez"
