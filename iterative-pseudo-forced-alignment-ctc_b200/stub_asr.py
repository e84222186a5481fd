"""Deterministic random-init character-CTC acoustic model with the attribute
surface the reference uses from ``speechbrain.pretrained.EncoderASR``
(/root/reference/src/iterative_utterance_alignment.py:415-420, src/test/test_asr.py:35-39).

SpeechBrain and its wav2vec2 checkpoints are out of scope (BASELINE.json north_star)
and not installed in this image; BASELINE config 1 names a "random-init char-CTC
EncoderASR (V~32)".  This stand-in produces emissions of the right shape
(one frame per 320 samples at 16 kHz, V = 32, blank = 0) from a fixed-seed
conv + linear stack; it is NOT an ASR model.
"""
import types

import torch

PIECES = ["<blank>", "<unk>", "▁"] + list("abcdefghijklmnopqrstuvwxyz") + ["ñ", "'", "·"]


class CharTokenizer:
    """SentencePiece-like surface: encode_as_ids / encode_as_pieces / id_to_piece / vocab_size / unk_id."""

    def __init__(self, pieces=PIECES):
        self.pieces = list(pieces)
        self.index = {p: i for i, p in enumerate(self.pieces)}

    def vocab_size(self):
        return len(self.pieces)

    def unk_id(self):
        return 1

    def id_to_piece(self, i):
        return self.pieces[i]

    def encode_as_pieces(self, text):
        out = []
        for ch in text.strip().lower():
            out.append("▁" if ch.isspace() else (ch if ch in self.index else "<unk>"))
        return out

    def encode_as_ids(self, text):
        return [self.index[p] for p in self.encode_as_pieces(text)]


class StubEncoderASR:
    def __init__(self, device="cpu", seed=1234, sample_rate=16000, hidden=64, stride=320, kernel=400):
        self.device = torch.device(device)
        self.tokenizer = CharTokenizer()
        g = torch.Generator().manual_seed(seed)
        v = self.tokenizer.vocab_size()
        self.stride, self.kernel = stride, kernel
        self.w1 = (torch.randn(hidden, 1, kernel, generator=g) / kernel ** 0.5).to(self.device)
        self.w2 = (torch.randn(v, hidden, generator=g) * 2.0).to(self.device)
        self.b2 = torch.zeros(v)
        self.b2[0] = 1.5  # blank-heavy like a CTC model
        self.b2 = self.b2.to(self.device)
        self.hparams = types.SimpleNamespace(sample_rate=sample_rate,
                                             log_softmax=lambda x: torch.log_softmax(x, dim=-1))

    def audio_normalizer(self, audio, sample_rate):
        """[samples, channels] -> mono [samples] (speechbrain AudioNormalizer surface)."""
        if audio.dim() == 2:
            audio = audio.mean(dim=1)
        return audio

    @torch.no_grad()
    def encode_batch(self, wavs, wav_lens=None):
        x = wavs.to(self.device).float().unsqueeze(1)  # [B, 1, S]
        pad = max(self.kernel - self.stride, 0)
        x = torch.nn.functional.pad(x, (0, pad))
        h = torch.tanh(torch.nn.functional.conv1d(x * 30.0, self.w1, stride=self.stride))  # [B, H, T]
        return h.transpose(1, 2) @ self.w2.t() + self.b2  # [B, T, V] logits
