"""Test double for the acoustic model: emissions that follow a known character schedule.

The audio of a synthetic file is a ramp (sample value = absolute position / total), so
the emitter can recover the absolute time of any window it is handed and emit
log-probabilities peaked on the character scheduled there.  This makes the anchor loop
take its accept / shrink paths on synthetic data (a random-init model only produces
rejections)."""
import types

import numpy as np
import torch

from cases import ctc_case  # noqa: F401  (keeps tests/ on sys.path semantics identical)

import importlib

stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")


def make_schedule(utterances, tokenizer, rng, frames_per_char=(2, 5), gap_frames=(8, 30), lead=20):
    """Returns (frame_tokens int array, utterance frame spans)."""
    frames = [0] * lead
    spans = []
    for utt in utterances:
        start = len(frames)
        for tok in tokenizer.encode_as_ids(utt):
            frames += [tok] * int(rng.integers(*frames_per_char))
            if rng.random() < 0.3:
                frames += [0]
        spans.append((start, len(frames)))
        frames += [0] * int(rng.integers(*gap_frames))
    return np.array(frames, dtype=np.int64), spans


class ScriptedASR:
    def __init__(self, frame_tokens, total_samples, device="cpu", stride=320, sample_rate=16000, peak=7.0,
                 noise=1.0, seed=0, corrupt=()):
        self.device = torch.device(device)
        self.tokenizer = stub.CharTokenizer()
        self.stride, self.total = stride, total_samples
        v = self.tokenizer.vocab_size()
        g = torch.Generator().manual_seed(seed)
        n_frames = total_samples // stride + 1
        logits = torch.randn(n_frames, v, generator=g) * noise
        ft = torch.zeros(n_frames, dtype=torch.long)
        ft[:min(len(frame_tokens), n_frames)] = torch.as_tensor(frame_tokens[:n_frames])
        logits[torch.arange(n_frames), ft] += peak
        for a, b in corrupt:  # frames where the audio does not match the text
            logits[a:b] = torch.randn(b - a, v, generator=g) * 3.0
        self.logits = logits
        self.hparams = types.SimpleNamespace(sample_rate=sample_rate,
                                             log_softmax=lambda x: torch.log_softmax(x, dim=-1))

    def audio_normalizer(self, audio, sample_rate):
        return audio.mean(dim=1) if audio.dim() == 2 else audio

    @torch.no_grad()
    def encode_batch(self, wavs, wav_lens=None):
        x = wavs[0].double()
        n = x.shape[0] // self.stride
        if n == 0:
            return torch.zeros(1, 0, self.logits.shape[1], device=self.device)
        first = int(round(float(x[0]) * self.total))  # absolute sample index of the window start
        f0 = first // self.stride
        idx = torch.arange(f0, f0 + n).clamp(max=self.logits.shape[0] - 1)
        return self.logits[idx].unsqueeze(0).to(self.device)


def ramp_audio(total_samples):
    return (np.arange(total_samples, dtype=np.float64) / total_samples)
