"""Synthetic corpora for the anchor sweep (BASELINE configs[4]): files of a given duration with
TSV-like rows (20-60 words each, split into <= 24-word utterances like ``prepare_text``), a
character schedule, text-proportional row times (what ``fix_time_reference`` produces for one
VAD segment) and emissions peaked on the scheduled characters; a fraction of the utterances is
'not what was said' (random frames) so the loop takes its shrink / discard paths.

Shared by tests/ and bench.py (workload c5)."""
import importlib

import numpy as np

stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")

WORDS = ("hola que tal estamos aqui para probar el alineamiento forzado iterativo con anclas sobre un audio "
         "largo y un texto que no siempre coincide con lo que se dice en la grabacion del pleno de hoy").upper().split()


class Spec:
    """Host-side description of one synthetic file (no emissions yet)."""

    def __init__(self, file_id, frame_tokens, n_samples, rows):
        self.file_id, self.frame_tokens, self.n_samples, self.rows = file_id, frame_tokens, n_samples, rows
        self.audio_path = f"/synthetic/{file_id}.wav"


def make_spec(file_id, minutes, seed, corrupt_frac=0.06, non_speech_every=0, max_words_sequence=24,
              frames_per_char=(2, 5), gap_frames=(8, 30)):
    rng = np.random.default_rng(seed)
    tok = stub.CharTokenizer()
    word_ids = [np.asarray(tok.encode_as_ids(w), dtype=np.int8) for w in WORDS]
    space = np.asarray(tok.encode_as_ids("a b")[1:2], dtype=np.int8)
    target = int(minutes * 60 * 50)
    pieces = [np.zeros(20, np.int8)]
    n_frames = 20
    rows = []
    while n_frames < target:
        n_words = int(rng.integers(20, 61))
        picks = rng.integers(0, len(WORDS), size=n_words)
        utterances, row_chars = [], 0
        for a in range(0, n_words, max_words_sequence):
            chunk = picks[a:a + max_words_sequence]
            text = " ".join(WORDS[i] for i in chunk)
            ids = []
            for j, i in enumerate(chunk):
                if j:
                    ids.append(space)
                ids.append(word_ids[i])
            ids = np.concatenate(ids)
            dur = rng.integers(frames_per_char[0], frames_per_char[1], size=ids.size)
            frames = np.repeat(ids, dur)
            if rng.random() < corrupt_frac:  # the audio says something else here
                frames = rng.integers(0, tok.vocab_size(), size=frames.size).astype(np.int8)
            gap = int(rng.integers(*gap_frames))
            pieces += [frames, np.zeros(gap, np.int8)]
            n_frames += frames.size + gap
            utterances.append(text)
            row_chars += len(text)
        rows.append({"Type": "Speech", "utterances": utterances, "chars": len(" ".join(utterances)),
                     "Channel": 1, "Speaker_ID": "spk_" + file_id, "Database": "synthetic",
                     "Sample_ID": f"{file_id}_{len(rows)}"})
        if non_speech_every and len(rows) % non_speech_every == 0:
            sil = int(rng.integers(150, 400))
            pieces.append(np.zeros(sil, np.int8))
            n_frames += sil
    pieces.append(np.zeros(40, np.int8))
    frame_tokens = np.concatenate(pieces)
    n_samples = int(frame_tokens.size) * 320
    # text-proportional times over the whole file (fix_time_reference with one VAD segment)
    dur = n_samples / 16000
    total = sum(r["chars"] for r in rows)
    acc = 0.0
    for r in rows:
        share = r["chars"] / total * dur
        r["Start"], r["End"] = acc, acc + share
        acc += share
    rows[-1]["End"] = dur
    return Spec(file_id, frame_tokens, n_samples, rows)


def emissions(spec, device, seed=0, peak=7.0, noise=1.0, vocab=None):
    """fp32 log-softmax emissions [T, V] of the file on ``device`` (T = n_samples // 320)."""
    import torch
    v = vocab or stub.CharTokenizer().vocab_size()
    g = torch.Generator(device=device).manual_seed(seed)
    ft = torch.as_tensor(spec.frame_tokens.astype(np.int64), device=device)
    logits = torch.randn(ft.shape[0], v, generator=g, device=device) * noise
    logits[torch.arange(ft.shape[0], device=device), ft] += peak
    return torch.log_softmax(logits, dim=-1)


def corpus_specs(total_hours, seed=0, min_minutes=5.0, max_minutes=60.0, **kw):
    """Files of 5-60 minutes adding up to ``total_hours``."""
    rng = np.random.default_rng(seed)
    specs, left, i = [], total_hours * 60.0, 0
    while left > 0:
        m = float(rng.uniform(min_minutes, max_minutes))
        m = min(m, max(left, min(min_minutes, left)))
        specs.append(make_spec(f"f{i:04d}", m, seed * 100003 + i, **kw))
        left -= m
        i += 1
    return specs
