"""Shared bootstrap of the drop-in entry points: repo root on sys.path, acoustic model loader.

The reference loads ``speechbrain.pretrained.EncoderASR.from_hparams``
(/root/reference/src/iterative_utterance_alignment.py:415); that model is out of scope here
(BASELINE.json north_star) and SpeechBrain is not part of this image.  ``load_asr`` keeps the
reference's behaviour -- it FAILS when the model cannot be loaded -- and adds two explicit
opt-ins for machines without SpeechBrain:

* ``--asr_hub stub``                  the deterministic random-init char-CTC emitter of
                                      BASELINE.json configs[0] (``stub_asr.py``; not an ASR model);
* ``--asr_hub py:<module>:<factory>`` any object with the EncoderASR surface the aligner uses,
                                      built by ``factory(savedir, device)``.
"""
import importlib
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
    if not asr_hub:
        raise ValueError("--asr_hub is empty: name a SpeechBrain EncoderASR source, 'stub', or 'py:<module>:<factory>'")
    if asr_hub == "stub":
        from ipfa_b200.stub_asr import StubEncoderASR
        print("asr_hub=stub: random-init character emitter, the alignments carry no meaning", file=sys.stderr)
        return StubEncoderASR(device=device)
    if asr_hub.startswith("py:"):
        _, module, factory = asr_hub.split(":", 2)
        return getattr(importlib.import_module(module), factory)(asr_savedir, device)
    try:
        from speechbrain.pretrained import EncoderASR
    except ImportError as e:
        raise ImportError(f"--asr_hub {asr_hub!r} needs speechbrain (speechbrain==0.5.11 in the reference's "
                          "requirements.txt:87), which is not installed; there is no fallback model") from e
    return EncoderASR.from_hparams(source=asr_hub, savedir=asr_savedir, run_opts={"device": device})


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
