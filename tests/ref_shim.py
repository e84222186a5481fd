"""Run the reference's OWN host code as the oracle (test infrastructure).

``/root/reference`` is a pure-Python repo whose host loops are on disk but do not import in
this image: speechbrain / num2words are absent, torchaudio 2.11 lost ``info`` and its ``load``
needs torchcodec, pandas 3 lost ``DataFrame.append`` and ``float(one_row_Series)``
(SURVEY.md section 0.5).  This module installs the smallest set of shims that lets those files
run UNMODIFIED, and hands them an aligner of the caller's choice in place of
``speechbrain.alignment.ctc_segmentation.CTCSegmentation``:

* ``oracle.sb_aligner.CTCSegmentation`` (CPU restatement) -> the reference loop is the oracle
  for rows A10, A11, (f)1, (f)3 of SURVEY.md section 8;
* the product's CUDA-backed mirror -> the reference loop drives the product's aligner.

Only available where ``/root/reference`` exists (this container).  The GPU box has no
reference tree: there the tests compare against the fixtures under ``tests/golden/ref_*``
that ``tests/golden/make_ref_golden.py`` wrote with this module.

Nothing here is imported by the product.
"""
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import types
import wave

import numpy as np
import pandas as pd
import torch

REFERENCE_ROOT = os.environ.get("IPFA_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")


def available():
    return os.path.isfile(os.path.join(REFERENCE_SRC, "iterative_utterance_alignment.py"))


# ----------------------------------------------------------------------------- torchaudio
class _Info:
    def __init__(self, num_frames, sample_rate, num_channels):
        self.num_frames, self.sample_rate, self.num_channels = num_frames, sample_rate, num_channels


def _ta_info(path):
    """torchaudio.info(path) for PCM WAV through stdlib ``wave`` (torchaudio==0.11 API)."""
    with wave.open(path, "rb") as w:
        return _Info(w.getnframes(), w.getframerate(), w.getnchannels())


def _ta_load(path, frame_offset=0, num_frames=-1, normalize=True, channels_first=True, format=None):
    """torchaudio.load of torchaudio==0.11.0 (/root/reference/requirements.txt:94, sox_io backend) for
    16-bit PCM WAV: float32 in [-1, 1), [channels, time] unless channels_first=False.  Like that
    backend it REJECTS ``frame_offset < 0`` and ``num_frames`` other than -1 or a positive count
    ("Invalid argument: num_frames must be -1 or greater than 0.") -- the reference reaches that with
    a clip whose end lies before its start and then carries on with the previous clip's audio
    (iterative_utterance_alignment.py:147-159, bare ``except``)."""
    if frame_offset < 0:
        raise RuntimeError("Invalid argument: frame_offset must be non-negative.")
    if not (num_frames == -1 or num_frames > 0):
        raise RuntimeError("Invalid argument: num_frames must be -1 or greater than 0.")
    with wave.open(path, "rb") as w:
        sr, ch, width, total = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        assert width == 2, "shimmed torchaudio.load reads 16-bit PCM only"
        frame_offset = min(int(frame_offset), total)
        n = total - frame_offset if num_frames == -1 else min(int(num_frames), total - frame_offset)
        w.setpos(frame_offset)
        raw = w.readframes(n)
    data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    audio = torch.from_numpy(data.reshape(-1, ch).copy())
    return (audio.t().contiguous() if channels_first else audio), sr


# ----------------------------------------------------------------------------- pandas 1.x behaviour
def _df_append(self, other, ignore_index=False, verify_integrity=False, sort=False):
    """DataFrame.append of pandas < 2 (alignment_utils.py:114, search_words.py:29)."""
    if isinstance(other, pd.Series):
        if other.name is None and not ignore_index:
            raise TypeError("Can only append a Series if ignore_index=True or if the Series has a name")
        other = pd.DataFrame([other.values], columns=list(other.index),
                             index=None if ignore_index else [other.name])
    if len(self.index) == 0 and len(self.columns) == 0:
        out = other.copy()
    elif len(other.index) == 0:
        out = self.copy()
    else:
        out = pd.concat([self, other], ignore_index=ignore_index, sort=sort)
    return out


def _series_float(self):
    """float(one_element_Series) of pandas < 3 (iterative_utterance_alignment.py:47 ...)."""
    if len(self) != 1:
        raise TypeError("cannot convert the series to <class 'float'>")
    return float(self.iloc[0])


@contextlib.contextmanager
def legacy_pandas():
    had_append = hasattr(pd.DataFrame, "append")
    old_float = pd.Series.__dict__.get("__float__")
    if not had_append:
        pd.DataFrame.append = _df_append
    pd.Series.__float__ = _series_float
    try:
        yield
    finally:
        if not had_append:
            del pd.DataFrame.append
        if old_float is None:
            del pd.Series.__float__
        else:
            pd.Series.__float__ = old_float


# ----------------------------------------------------------------------------- module stubs
def _num2words(number, lang="es", **kw):
    raise NotImplementedError("num2words is not installed; the synthetic transcripts carry no digits")


class _NoModel:
    """speechbrain.pretrained.EncoderASR placeholder; tests pass their own acoustic model."""
    _instance = None

    @classmethod
    def from_hparams(cls, source=None, savedir=None, **kw):
        if cls._instance is None:
            raise RuntimeError("ref_shim: set EncoderASR._instance to the test's acoustic model first")
        return cls._instance


_STUBS = {}


def _install_stubs(aligner_cls):
    sb = types.ModuleType("speechbrain")
    pre = types.ModuleType("speechbrain.pretrained")
    pre.EncoderASR = _NoModel
    ali = types.ModuleType("speechbrain.alignment")
    seg = types.ModuleType("speechbrain.alignment.ctc_segmentation")
    seg.CTCSegmentation = aligner_cls
    sb.pretrained, sb.alignment, ali.ctc_segmentation = pre, ali, seg
    n2w = types.ModuleType("num2words")
    n2w.num2words = _num2words
    mods = {"speechbrain": sb, "speechbrain.pretrained": pre, "speechbrain.alignment": ali,
            "speechbrain.alignment.ctc_segmentation": seg, "num2words": n2w}
    for k, m in mods.items():
        _STUBS[k] = sys.modules.get(k)
        sys.modules[k] = m
    return pre


def _remove_stubs():
    for k, old in _STUBS.items():
        if old is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = old
    _STUBS.clear()


@contextlib.contextmanager
def reference_modules(aligner_cls, asr_model=None, quiet=True):
    """Context in which the reference's files are imported and run unmodified.

    Yields a namespace with ``iua`` (iterative_utterance_alignment), ``wla``
    (word_level_alignment), ``sos`` (search_on_speech), ``search_words`` and
    ``alignment_utils`` / ``text_utils`` -- fresh module objects every time, bound to
    ``aligner_cls``."""
    assert available(), "no reference tree at " + REFERENCE_ROOT
    import torchaudio
    saved_ta = {k: getattr(torchaudio, k, None) for k in ("info", "load")}
    saved_path = list(sys.path)
    saved_utils = {k: v for k, v in sys.modules.items() if k == "utils" or k.startswith("utils.")}
    for k in saved_utils:
        del sys.modules[k]
    pre = _install_stubs(aligner_cls)
    pre.EncoderASR._instance = asr_model
    torchaudio.info, torchaudio.load = _ta_info, _ta_load
    sys.path.insert(0, REFERENCE_SRC)
    loaded = []
    sink = io.StringIO()
    try:
        with legacy_pandas(), (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            ns = types.SimpleNamespace()

            def load(name, rel):
                spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REFERENCE_SRC, rel))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                loaded.append("ref_" + name)
                return mod

            ns.iua = load("iua", "iterative_utterance_alignment.py")
            ns.wla = load("wla", "word_level_alignment.py")
            ns.sos = load("sos", "search_on_speech.py")
            ns.search_words = load("search_words", "search_words.py")
            ns.alignment_utils = importlib.import_module("utils.alignment_utils")
            ns.text_utils = importlib.import_module("utils.text_utils")
            ns.stdout = sink
            yield ns
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved_utils)
        for k, v in saved_ta.items():
            if v is None:
                if hasattr(torchaudio, k):
                    delattr(torchaudio, k)
            else:
                setattr(torchaudio, k, v)
        pre.EncoderASR._instance = None
        _remove_stubs()


def close_logger(name):
    """The reference never closes its per-file handlers (alignment_utils.py:10-32); tests do."""
    import logging
    lg = logging.getLogger(name)
    for h in list(lg.handlers):
        h.close()
        lg.removeHandler(h)


def run_script(rel_path, argv, cwd=None):
    """Run a reference script that needs no shim beyond pandas (merge_aligned_files.py,
    tsv_to_stm.py, postprocess_and_filter.py) as ``__main__`` in a subprocess."""
    import subprocess
    boot = ("import sys, runpy; sys.path.insert(0, %r); sys.path.insert(0, %r); import ref_shim; "
            "import torchaudio; torchaudio.info = ref_shim._ta_info; torchaudio.load = ref_shim._ta_load\n"
            "sys.argv = [%r] + %r\n"
            "with ref_shim.legacy_pandas():\n"
            "    runpy.run_path(%r, run_name='__main__')\n") % (
                os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.join(REFERENCE_SRC, rel_path)),
                rel_path, list(argv), os.path.join(REFERENCE_SRC, rel_path))
    return subprocess.run([sys.executable, "-c", boot], cwd=cwd, capture_output=True, text=True, check=True)
