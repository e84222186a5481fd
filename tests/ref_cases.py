"""Deterministic end-to-end cases for the reference's own host loops (test infrastructure).

One definition of every case serves three users:
* ``tests/golden/make_ref_golden.py`` -- runs /root/reference's UNMODIFIED loops (through
  ``tests/ref_shim.py``, aligner = ``oracle.sb_aligner``) on these cases in the build container
  and commits what they wrote under ``tests/golden/ref/``;
* the CPU tests -- check that the committed fixtures are what the reference writes today (when
  the reference tree is present) and that the repo's host mirror / oracle restatements agree;
* the ``-m gpu`` tests -- run the product's entry points (CUDA path) on the same cases and compare
  the files byte for byte with the fixtures (the GPU box has no reference tree).

Everything a case needs is rebuilt from a seed: the WAV (samples that spell their own position, so the
emitter can tell where a clip starts), the TSVs, and the acoustic-model double whose emissions are quantised to a
2^-10 grid so they are bit-identical on every machine.
"""
import hashlib
import importlib
import os
import types
import wave

import numpy as np
import pandas as pd
import torch

stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")

WORDS = ("hola que tal estamos aqui para probar el alineamiento forzado iterativo con anclas sobre un audio "
         "largo y un texto que no siempre coincide con lo que se dice en la grabacion del pleno de hoy").split()
STRIDE = 320
SR = 16000


def position_pcm(total):
    """16-bit samples from which ANY clip can tell its absolute start: sample pair q = p // 2 holds the
    low 15 bits of q at the even position (>= 0) and -(q >> 15) - 1 at the odd one (< 0)."""
    q = np.arange((total + 1) // 2, dtype=np.int64)
    pcm = np.empty(2 * q.size, np.int64)
    pcm[0::2] = q & 0x7fff
    pcm[1::2] = -(q >> 15) - 1
    return pcm[:total].astype("<i2")


def clip_position(x0, x1):
    """Absolute sample index of a clip from its first two samples (floats, int16 / 32768)."""
    a, b = int(round(float(x0) * 32768.0)), int(round(float(x1) * 32768.0))
    if a >= 0:
        return 2 * (a + ((-b - 1) << 15))
    return 2 * (((-a - 1) << 15) + ((b - 1) & 0x7fff)) + 1


def write_wav(path, pcm):
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(SR)
        w.writeframes(np.asarray(pcm, dtype="<i2").tobytes())


class GridASR:
    """Acoustic-model double with the EncoderASR surface the reference uses
    (audio_normalizer, encode_batch, hparams.log_softmax / sample_rate, tokenizer, device).

    The file's log-probabilities are fixed at construction: seeded logits peaked on a
    character schedule, normalised in fp64 and rounded to multiples of 2^-10, so the fp32 values
    do not depend on the machine's exp/log.  ``encode_batch`` returns the rows of the clip it is
    handed (the samples spell their own position, ``position_pcm``); ``log_softmax`` is the identity.
    """

    def __init__(self, frame_tokens, total_samples, device="cpu", peak=7.0, noise=1.0, seed=0, corrupt=()):
        self.device = torch.device(device)
        self.tokenizer = stub.CharTokenizer()
        self.total = int(total_samples)
        v = self.tokenizer.vocab_size()
        rng = np.random.default_rng(seed)
        n_frames = self.total // STRIDE + 1
        logits = np.round(rng.standard_normal((n_frames, v)) * noise * 256.0) / 256.0
        ft = np.zeros(n_frames, np.int64)
        m = min(len(frame_tokens), n_frames)
        ft[:m] = np.asarray(frame_tokens[:m], np.int64)
        logits[np.arange(n_frames), ft] += peak
        for a, b in corrupt:  # frames where the audio does not say what the text says
            logits[a:b] = np.round(rng.standard_normal((b - a, v)) * 3.0 * 256.0) / 256.0
        lse = np.log(np.exp(logits).sum(-1, keepdims=True))
        lp = np.round((logits - lse) * 1024.0) / 1024.0
        self.lp_host = torch.from_numpy(lp.astype(np.float32))
        self.lp = self.lp_host.to(self.device)
        self.hparams = types.SimpleNamespace(sample_rate=SR, log_softmax=lambda x: x)

    def digest(self):
        return hashlib.sha256(self.lp_host.numpy().tobytes()).hexdigest()

    def audio_normalizer(self, audio, sample_rate):
        return audio.mean(dim=1) if audio.dim() == 2 else audio

    @torch.no_grad()
    def encode_batch(self, wavs, wav_lens=None):
        x = wavs[0]
        n = x.shape[0] // STRIDE
        if n == 0:
            return torch.zeros(1, 0, self.lp.shape[1], device=self.device)
        # (estimate_samples_to_frames_ratio hands in noise: any rows do, only their count matters)
        f0 = max(0, clip_position(x[0], x[1])) // STRIDE
        idx = torch.arange(f0, f0 + n, device=self.device).clamp(max=self.lp.shape[0] - 1)
        return self.lp[idx].unsqueeze(0)


def _utterance(rng, lo=6, hi=14):
    return " ".join(rng.choice(WORDS, size=int(rng.integers(lo, hi))))


def _schedule(segments, tok, rng, frames_per_char=(2, 5), gap_frames=(8, 30), lead=20):
    """segments: list of (list of utterance strings, silence frames after the segment).
    Returns (frame_tokens, per-utterance frame spans, per-segment (first frame, last frame))."""
    frames = [0] * lead
    spans, seg_spans = [], []
    for utts, silence in segments:
        seg_start = len(frames)
        for utt in utts:
            start = len(frames)
            for t in tok.encode_as_ids(utt):
                frames += [t] * int(rng.integers(*frames_per_char))
                if rng.random() < 0.3:
                    frames += [0]
            spans.append((start, len(frames)))
            frames += [0] * int(rng.integers(*gap_frames))
        seg_spans.append((seg_start, len(frames)))
        frames += [0] * silence
    return np.array(frames, dtype=np.int64), spans, seg_spans


class AnchorCase:
    """One audio file + its TSV rows + VAD table + loop parameters."""

    def __init__(self, name, seed, rows_per_segment, silences, corrupt=(), loop=None, words_per_row=(6, 14),
                 frames_per_char=(2, 5), gap_frames=(8, 30), text_only_rows=0, peak=7.0):
        self.name, self.seed, self.peak = name, seed, peak
        self.loop = dict(threshold=-2.0, short_utterance_len=30, max_words_sequence=8, min_words_sequence=None,
                         max_window_size=70.0, window_to_stop=500.0, min_text_to_audio_prop=0.8,
                         max_text_to_audio_prop_exec=10)
        self.loop.update(loop or {})
        rng = np.random.default_rng(seed)
        tok = stub.CharTokenizer()
        segments, texts = [], []
        for n_rows, silence in zip(rows_per_segment, silences):
            utts = [_utterance(rng, *words_per_row) for _ in range(n_rows)]
            texts += utts
            segments.append(([u.upper() for u in utts], silence))
        self.frame_tokens, self.spans, seg_spans = _schedule(segments, tok, rng, frames_per_char=frames_per_char,
                                                             gap_frames=gap_frames)
        # rows whose text was never spoken (the transcript runs ahead of the audio)
        texts += [_utterance(rng, 30, 40) for _ in range(text_only_rows)]
        self.total = (len(self.frame_tokens) + 40) * STRIDE
        self.corrupt = tuple(corrupt)
        self.texts = texts
        self.wav_rel = f"audio/{name}.wav"
        dur = self.total / SR
        n = len(texts)
        # the TSV's own times are dummies: the loop re-derives them (iterative_utterance_alignment.py:53)
        self.df = pd.DataFrame({
            "Sample_ID": [f"{name}_{i}" for i in range(n)], "Sample_Path": [self.wav_rel] * n,
            "Channel": [1] * n, "Audio_Length": [dur / n] * n, "Start": [0.0] * n, "End": [dur] * n,
            "Segment_Score": [0.0] * n, "Transcription": texts, "Speaker_ID": ["spk1"] * n,
            "Database": ["synthetic"] * n})
        vad = []
        for i, (a, b) in enumerate(seg_spans):
            s = 0.0 if i == 0 else a * 0.02
            e = dur if i == len(seg_spans) - 1 else (b + 5) * 0.02
            vad.append({"Sample_Path": self.wav_rel, "Start": s, "End": e, "Segment_Length": e - s})
        self.vad = pd.DataFrame(vad)

    def materialise(self, root):
        """Write the WAV and the TSVs under ``root`` (paths inside the TSVs stay relative to it)."""
        os.makedirs(os.path.join(root, "audio"), exist_ok=True)
        os.makedirs(os.path.join(root, "tsv"), exist_ok=True)
        write_wav(os.path.join(root, self.wav_rel), position_pcm(self.total))
        self.tsv_rel = f"tsv/{self.name}.tsv"
        self.vad_rel = f"tsv/{self.name}_vad_segments_filtered.tsv"
        self.df.to_csv(os.path.join(root, self.tsv_rel), sep="\t", index=None)
        self.vad.to_csv(os.path.join(root, self.vad_rel), sep="\t", index=None)
        return self

    def asr(self, device="cpu"):
        return GridASR(self.frame_tokens, self.total, device=device, seed=self.seed, corrupt=self.corrupt,
                       peak=self.peak)


def anchor_cases():
    """name -> AnchorCase.  Together they walk every branch of the reference loop the product
    restates (``manifest.json`` records which log lines each case produced)."""
    A = AnchorCase
    return {
        # every utterance is where the text says: accept / shrink-and-keep paths
        "clean": A("clean", 11, [14], [0]),
        # stretches of audio that do not match the text: bad alignments, reverts, discards
        "corrupt": A("corrupt", 12, [16], [0], corrupt=((300, 420), (900, 1010))),
        # three VAD speech segments: Non-Speech rows (:73-77)
        "nonspeech": A("nonspeech", 13, [6, 5, 6], [260, 340, 0], corrupt=((500, 560),)),
        # small max_window_size: fix_text_to_time_proportion re-spreads the remaining rows (:119-146)
        "recalc": A("recalc", 14, [9, 8], [300, 0], corrupt=((350, 700),), loop=dict(max_window_size=8.0)),
        # slow speech: "Low quantity of text compared to audio" (:175-181), windows grow, re-spreading
        "sparse": A("sparse", 17, [10], [0], frames_per_char=(14, 18)),
        # soft emissions, scores between the threshold and -1: the iterate-to-improve branches (:289-377)
        "mushy": A("mushy", 31, [8], [0], peak=3.6, words_per_row=(16, 40)),
        "mushy2": A("mushy2", 24, [12], [0], peak=3.75, words_per_row=(4, 20)),
        # much text, little audio, a speech segment about to end: the trimming branch (:185-192) and
        # "Audio is shorter than text!" (:390-402) below the limit
        "ending": A("ending", 19, [4, 4], [300, 0], frames_per_char=(1, 2), gap_frames=(2, 5),
                    words_per_row=(40, 50), text_only_rows=2),
        # transcript far longer than the audio: the exception limit stops the file (:397-399)
        "dense": A("dense", 15, [4], [0], text_only_rows=8, frames_per_char=(1, 2),
                   loop=dict(max_text_to_audio_prop_exec=3)),
        # window_to_stop reached: lost in alignment (:125-126)
        "lost": A("lost", 16, [12], [0], corrupt=((150, 1400),), loop=dict(window_to_stop=18.0, max_window_size=400.0)),
    }


LOG_MARKS = {
    "non_speech_row": "Skipping non-speech segment",
    "recalc": "Recalculating time references",
    "low_text": "Low quantity of text compared to audio",
    "segment_ending": "As the speech segment is finishing",
    "shrink_bad": "Misalignment detected. Repeating alignment due low score.",
    "first_good": "Good results. Starting iteration...",
    "improved": "Results have improved. Continuing iteration...",
    "not_improved": "Not improved results keeping previous alignment",
    "improved_but_bad": "Improved results but score is under the treshold",
    "very_nice": "Not repeating because last alignment is very nice",
    "single_stored": "Storing this last alignment",
    "single_discarded": "Reading more audio to better align",
    "shorter_than_text": "is shorter than text",
    "exception_limit": "Number of reached the limit",
}
# plus "load_refused": clips torchaudio.load refused (counted from the print at :158, not a log line)


def count_marks(log_text):
    return {k: log_text.count(v) for k, v in LOG_MARKS.items()}


# ----------------------------------------------------------------------------- word level / search on speech
class WordsCase:
    """Rows of an utterance-level TSV (clips of the ``clean`` file) for word_level_alignment.py
    (through search_words.py) and search_on_speech.py."""

    def __init__(self, anchor_case, n_rows=8, wanted=("alineamiento", "audio", "texto")):
        self.base = anchor_case
        self.wanted = list(wanted)
        rows = []
        for i, (a, b) in enumerate(anchor_case.spans[:n_rows]):
            start, end = max(0, a - 10) * 0.02, (b + 10) * 0.02
            rows.append({"Sample_ID": f"{anchor_case.name}_{i}", "Sample_Path": anchor_case.wav_rel, "Channel": 1,
                         "Audio_Length": end - start, "Start": start, "End": end, "Segment_Score": -0.5,
                         "Transcription": anchor_case.texts[i].capitalize() + ".", "Speaker_ID": "spk1",
                         "Database": "synthetic"})
        # a clip far too short for its text: "Audio is shorter than text!" (word_level_alignment.py:130)
        rows.append(dict(rows[0], Sample_ID=f"{anchor_case.name}_short", End=rows[0]["Start"] + 0.3,
                         Audio_Length=0.3))
        self.df = pd.DataFrame(rows)
        self.search_text = "el alineamiento"

    def materialise(self, root):
        os.makedirs(os.path.join(root, "words"), exist_ok=True)
        self.tsv_rel = "words/utterances.tsv"
        self.df.to_csv(os.path.join(root, self.tsv_rel), sep="\t", index=None)
        self.config_rel = "words/words.json"
        import json
        with open(os.path.join(root, self.config_rel), "w") as f:
            json.dump({"words": self.wanted}, f)
        return self


def asr_factory(savedir, device):
    """``--asr_hub py:ref_cases:asr_factory --asr_savedir <case name>``: the acoustic-model double
    of an anchor case for the product's entry points (``src/_common.load_asr``)."""
    return anchor_cases()[savedir].asr(device)
