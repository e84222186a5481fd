"""oracle/sweep.py (the CPU restatement of the per-file anchor loop over resident emissions)
against the host mirror of the reference loop driven by the same CPU oracle -- two independent
restatements of /root/reference/src/iterative_utterance_alignment.py:67-402 must produce the
same rows.  CPU only."""
import importlib

import numpy as np
import pandas as pd
import pytest
import torch

import sweep_corpus
from scripted_asr import ScriptedASR

PKG = "iterative-pseudo-forced-alignment-ctc_b200"
hg = importlib.import_module(PKG + ".hostglue")
anchor = importlib.import_module(PKG + ".anchor")
cs = importlib.import_module(PKG + ".ctc_segmentation")
stub = importlib.import_module(PKG + ".stub_asr")


class ExactASR(ScriptedASR):
    @torch.no_grad()
    def encode_batch(self, wavs, wav_lens=None):
        x = wavs[0].double()
        n = x.shape[0] // self.stride
        if n == 0:
            return torch.zeros(1, 0, self.logits.shape[1], device=self.device)
        f0 = int(x[0].item()) // self.stride
        idx = torch.arange(f0, f0 + n).clamp(max=self.logits.shape[0] - 1)
        return self.logits[idx].unsqueeze(0).to(self.device)


def _oracle_window_fn(aligner, transcript, lpz, name, n_samples, clip_start, is_last_segment, new_segment_start,
                      threshold, short_utterance_len, file_id, audio_path, channel, speaker_id, database,
                      logger=None):
    from oracle import ctcseg as oseg
    from oracle.anchor import anchor_window
    lp = np.asarray(lpz)

    def align_fn(tr):
        task = aligner.prepare_segmentation_task(tr, lpz, name, n_samples)
        cfg = oseg.CtcSegmentationParameters(index_duration=task.config.index_duration,
                                             score_min_mean_over_L=task.config.score_min_mean_over_L)
        res = oseg.get_segments(cfg, lp, task.ground_truth_mat, task.utt_begin_indices, task.text)
        return [s.split(" ", 5) for s in oseg.task_str(name, task.text, res["segments"]).strip().split("\n")]

    rows, nss, disc, n_iter = anchor_window(
        transcript, align_fn, clip_start, is_last_segment, new_segment_start, [], threshold, short_utterance_len,
        file_id, audio_path, {"Channel": channel, "Speaker_ID": speaker_id, "Database": database})
    return rows, nss, disc, n_iter


@pytest.mark.parametrize("seed,minutes,ns", [(3, 2.0, 0), (4, 3.0, 3)])
def test_oracle_sweep_equals_host_mirror(monkeypatch, tmp_path, seed, minutes, ns):
    from oracle import sweep as osweep
    spec = sweep_corpus.make_spec("talk", minutes, seed, corrupt_frac=0.15, non_speech_every=ns)
    total = spec.n_samples
    asr = ExactASR(spec.frame_tokens, total, device="cpu", seed=seed)
    wav = spec.audio_path

    monkeypatch.setattr(hg, "audio_info", lambda path: hg.AudioInfo(total, 16000, 1))

    def load(path, frame_offset=0, num_frames=-1, channels_first=False):
        frame_offset = max(0, min(int(frame_offset), total))
        n = total - frame_offset if num_frames is None or num_frames < 0 else \
            max(0, min(int(num_frames), total - frame_offset))
        return torch.arange(frame_offset, frame_offset + n, dtype=torch.float64).reshape(-1, 1), 16000
    monkeypatch.setattr(hg, "audio_load", load)

    # the TSV the reference would read: one row per transcript segment; a VAD table whose gaps
    # become Non-Speech rows in fix_time_reference
    n = len(spec.rows)
    dur = total / 16000
    df = pd.DataFrame({'Sample_ID': [r["Sample_ID"] for r in spec.rows], 'Sample_Path': [wav] * n,
                       'Channel': [1] * n, 'Audio_Length': [dur / n] * n, 'Start': [0.0] * n, 'End': [dur] * n,
                       'Transcription': [" ".join(r["utterances"]) for r in spec.rows],
                       'Speaker_ID': ["spk_talk"] * n, 'Database': ['synthetic'] * n})
    if ns:
        cut = dur * 0.45
        vad = pd.DataFrame({'Sample_Path': [wav] * 2, 'Start': [0.0, cut + 4.0], 'End': [cut, dur],
                            'Segment_Length': [cut, dur - cut - 4.0]})
    else:
        vad = pd.DataFrame({'Sample_Path': [wav], 'Start': [0.0], 'End': [dur], 'Segment_Length': [dur]})
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", scoring_length=30,
                                 keep_lpz_on_device=False)
    aligner.samples_to_frames_ratio = 320.0
    kw = dict(threshold=-2.0, short_utterance_len=30, max_words_sequence=24, max_window_size=70.0)
    ref = anchor.get_file_iterative_segmentation(asr, aligner, wav, df.copy(), vad.copy(), 320.0, str(tmp_path),
                                                 window_fn=_oracle_window_fn, **kw)

    sweep = importlib.import_module(PKG + ".sweep")
    fixed = hg.fix_time_reference(df, vad, dur, n)
    rows = sweep.rows_from_dataframe(fixed, 24)
    lpz = torch.log_softmax(asr.logits[: total // 320].float(), dim=-1).numpy()
    got, status, stats = osweep.sweep_file("talk", wav, lpz, total, rows, stub.CharTokenizer())
    assert stats["windows"] >= 3
    if status == "needs_recalc":
        assert len(got) > 0 and got == ref[:len(got)]
    else:
        assert status == "done" and got == ref
    if ns:
        assert any(r["Type"] == "Non-Speech" for r in rows)
