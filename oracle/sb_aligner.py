"""oracle/sb_aligner.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU stand-in for ``speechbrain.alignment.ctc_segmentation.CTCSegmentation`` (speechbrain==0.5.11,
/root/reference/requirements.txt:87; not vendored, not installed: PARITY UNPINNED for the wrapper
itself) with the surface the reference's entry points use:

    CTCSegmentation(asr_model, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    .estimate_samples_to_frames_ratio()  .get_lpz(speech)
    .prepare_segmentation_task(text, lpz, name, speech_len)  .get_segments(task)
    task.set(**segments)  str(task)
        (/root/reference/src/iterative_utterance_alignment.py:418-420,201-219,
         /root/reference/src/word_level_alignment.py:26,89-103, src/search_on_speech.py:36,74-88)

``tests/ref_shim.py`` installs this class under the speechbrain module name so that the
reference's own loops run unmodified on the CPU -- the loops are then the oracle for the rows
of SURVEY.md section 8 that live in the reference tree (A10, A11, (f)1, (f)3).  The numerics
under it are ``oracle/ctcseg.py`` + ``ctcseg_oracle.c``.
"""
import numpy as np
import torch

from . import ctcseg as oseg


class CTCSegmentationTask:
    text = None
    ground_truth_mat = None
    utt_begin_indices = None
    timings = None
    char_probs = None
    state_list = None
    segments = None
    config = None
    done = False
    name = "utt"
    utt_ids = None
    lpz = None

    def __init__(self, **kwargs):
        self.set(**kwargs)

    def set(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __str__(self):
        return oseg.task_str(self.name, self.text, self.segments, self.utt_ids)


class CTCSegmentation:
    def __init__(self, asr_model, kaldi_style_text=True, text_converter="tokenize", time_stamps="auto",
                 scoring_length=None, min_window_size=None, max_window_size=None, gratis_blank=None,
                 set_blank=None, fs=None):
        self.asr_model = asr_model
        self.kaldi_style_text = kaldi_style_text
        self.time_stamps = time_stamps
        assert text_converter == "tokenize"
        self.fs = float(fs if fs is not None else asr_model.hparams.sample_rate)
        self.samples_to_frames_ratio = None
        tok = asr_model.tokenizer
        self.config = oseg.CtcSegmentationParameters(
            char_list=[tok.id_to_piece(i) for i in range(tok.vocab_size())])
        if scoring_length is not None:
            self.config.score_min_mean_over_L = scoring_length
        if min_window_size is not None:
            self.config.min_window_size = min_window_size
        if max_window_size is not None:
            self.config.max_window_size = max_window_size
        if gratis_blank is not None:
            self.config.blank_transition_cost_zero = gratis_blank
        if set_blank is not None:
            self.config.blank = set_blank

    def estimate_samples_to_frames_ratio(self, speech_len=215040):
        lpz = self.get_lpz(torch.rand(speech_len))
        return speech_len / lpz.shape[0]

    @torch.no_grad()
    def get_lpz(self, speech):
        """encode_batch -> log_softmax -> squeeze -> host NumPy."""
        if isinstance(speech, np.ndarray):
            speech = torch.tensor(speech)
        dev = self.asr_model.device
        enc = self.asr_model.encode_batch(speech.unsqueeze(0).to(dev), torch.tensor([1.0]).to(dev))
        return self.asr_model.hparams.log_softmax(enc).detach().squeeze(0).cpu().numpy()

    def _split_text(self, text):
        if isinstance(text, str):
            text = text.splitlines()
        text = [u for u in text if len(u)]
        if not self.kaldi_style_text:
            return None, text
        pairs = [u.split(" ", 1) for u in text]
        pairs = [p for p in pairs if len(p) == 2]
        return [p[0] for p in pairs], [p[1] for p in pairs]

    def prepare_segmentation_task(self, text, lpz, name=None, speech_len=None):
        if self.time_stamps == "fixed":
            if self.samples_to_frames_ratio is None:
                self.samples_to_frames_ratio = self.estimate_samples_to_frames_ratio()
            self.config.index_duration = self.samples_to_frames_ratio / self.fs
        else:
            self.config.index_duration = speech_len / lpz.shape[0] / self.fs
        utt_ids, text = self._split_text(text)
        tok = self.asr_model.tokenizer
        unk = tok.unk_id()
        token_list = []
        for utt in text:
            ids = np.array(tok.encode_as_ids(utt))
            token_list.append(ids[ids != unk] if ids.size else ids)
        gt, ub = oseg.prepare_token_list(self.config, token_list)
        return CTCSegmentationTask(config=self.config, name=name, text=text, ground_truth_mat=gt,
                                   utt_begin_indices=ub, utt_ids=utt_ids, lpz=lpz)

    @staticmethod
    def get_segments(task):
        res = oseg.get_segments(task.config, np.asarray(task.lpz), task.ground_truth_mat,
                                task.utt_begin_indices, task.text)
        res.update(name=task.name, done=True)
        return res
