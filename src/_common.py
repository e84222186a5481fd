"""Shared bootstrap of the drop-in entry points: repo root on sys.path, acoustic model loader.

The reference loads ``speechbrain.pretrained.EncoderASR.from_hparams`` (out of scope,
not installed here).  ``load_asr`` returns that model when SpeechBrain is importable and
otherwise the deterministic stub emitter, so the alignment path runs end to end on
synthetic audio (BASELINE.json configs[0])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import ipfa_b200  # noqa: E402  (fails loudly when libipfa_b200.so is missing)
from ipfa_b200 import anchor, hostglue, sharding, sweep, words  # noqa: E402,F401
from ipfa_b200.ctc_segmentation import CTCSegmentation  # noqa: E402,F401


def load_asr(asr_hub, asr_savedir, device=None):
    import torch
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    if asr_hub and asr_hub != "stub":
        try:
            from speechbrain.pretrained import EncoderASR
            return EncoderASR.from_hparams(source=asr_hub, savedir=asr_savedir, run_opts={"device": device})
        except ImportError:
            print("speechbrain is not installed: using the random-init stub emitter (stub_asr.py)")
    from ipfa_b200.stub_asr import StubEncoderASR
    return StubEncoderASR(device=device)


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
