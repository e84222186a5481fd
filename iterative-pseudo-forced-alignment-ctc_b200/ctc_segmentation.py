"""Host-side mirror of the reference's aligner interface, backed by the CUDA path.

The reference drives ``speechbrain.alignment.ctc_segmentation.CTCSegmentation``
(speechbrain==0.5.11, /root/reference/requirements.txt:87; not vendored) with
four calls -- ``get_lpz -> prepare_segmentation_task -> get_segments -> str(task)``:
  /root/reference/src/iterative_utterance_alignment.py:418-420, 201-219
  /root/reference/src/word_level_alignment.py:26, 89-103
  /root/reference/src/search_on_speech.py:36, 74-88
This module re-provides that surface (same names, argument meaning and error
behaviour) so the entry points change only their import line.  What differs:
``get_lpz`` may leave the emissions on the GPU (a CUDA tensor flows into
``prepare_segmentation_task`` unchanged), ``get_segments`` runs
``ipfa_ctcseg_device`` instead of the Cython fill + Python backtrace, and
``get_segments_batch`` / ``prefix_segments`` expose the batched and all-prefix
forms the anchor loop uses.  There is no CPU fallback.
"""
import logging

import numpy as np
import torch

from . import ops

logger = logging.getLogger(__name__)


class CtcSegmentationParameters:
    """Same attributes and defaults as ``ctc_segmentation.CtcSegmentationParameters``
    (ctc-segmentation==1.7.1)."""

    max_prob = -10000000000.0
    skip_prob = -10000000000.0
    min_window_size = 8000
    max_window_size = 100000
    index_duration = 0.025
    score_min_mean_over_L = 30
    space = "·"
    blank = 0
    replace_spaces_with_blanks = False
    blank_transition_cost_zero = False
    preamble_transition_cost_zero = True
    backtrack_from_max_t = False
    self_transition = "ε"
    start_of_ground_truth = "#"
    excluded_characters = ".,»«•❍·"
    tokenized_meta_symbol = "▁"
    char_list = None
    # scoring-window index rounding: "floor" or "round" (see oracle/ctcseg.py SEG_INDEX_ROUNDING)
    seg_index_rounding = "floor"
    # windowed table mode (T > min_window_size), the two spots of the package that could not be
    # re-checked here: largest window step "int+1" | "ceil"; cur_offset carry "ascending" | "shift"
    window_step_rule = "int+1"
    offset_cascade = "ascending"

    def __init__(self, **kwargs):
        self.set(**kwargs)

    def set(self, **kwargs):
        for key in kwargs:
            setattr(self, key, kwargs[key])

    @property
    def index_duration_in_seconds(self):
        return self.index_duration

    @property
    def flags(self):
        f = int(self.blank_transition_cost_zero) + 2 * int(self.preamble_transition_cost_zero)
        if self.seg_index_rounding == "round":
            f |= ops.SEG_ROUND_NEAREST
        if self.window_step_rule == "ceil":
            f |= ops.SEG_WINDOW_STEP_CEIL
        if self.offset_cascade == "shift":
            f |= ops.SEG_OFFSET_SHIFT
        return f

    def __str__(self):
        return str({k: getattr(self, k) for k in dir(self) if not k.startswith("_") and k != "set"})


def prepare_token_list(config, text):
    """``[-1] + (blank + tokens)* + blank`` with one blank between utterances
    (ctc_segmentation.prepare_token_list)."""
    ground_truth = [-1]
    utt_begin_indices = []
    for utt in text:
        if not ground_truth[-1] == config.blank:
            ground_truth += [config.blank]
        utt_begin_indices.append(len(ground_truth) - 1)
        ground_truth += np.asarray(utt).tolist()
    if not ground_truth[-1] == config.blank:
        ground_truth += [config.blank]
    utt_begin_indices.append(len(ground_truth) - 1)
    ground_truth_mat = np.array(ground_truth, dtype=np.int64).reshape(-1, 1)
    return ground_truth_mat, utt_begin_indices


def prepare_text(config, text, char_list=None):
    """The ``classic`` text converter (ctc_segmentation.prepare_text):
    the ground truth is the character string ``#·utt0·utt1·...·``; row i of ``ground_truth_mat``
    holds, in column s, the token that spells the last s+1 characters ending at i (or -1), so a
    multi-character token is one transition that skips s columns."""
    if char_list is not None:
        config.char_list = char_list
    blank = config.char_list[config.blank]
    ground_truth = config.start_of_ground_truth
    utt_begin_indices = []
    for utt in text:
        if not ground_truth.endswith(config.space):
            ground_truth += config.space
        utt_begin_indices.append(len(ground_truth) - 1)
        for char in utt:
            if char.isspace() and config.replace_spaces_with_blanks:
                if not ground_truth.endswith(config.space):
                    ground_truth += config.space
            elif char in config.char_list and char not in config.excluded_characters:
                ground_truth += char
    if not ground_truth.endswith(config.space):
        ground_truth += config.space
    utt_begin_indices.append(len(ground_truth) - 1)
    max_char_len = max(len(c) for c in config.char_list)
    index = {c: i for i, c in reversed(list(enumerate(config.char_list)))}  # list.index: first match
    ground_truth_mat = np.ones([len(ground_truth), max_char_len], np.int64) * -1
    for i in range(len(ground_truth)):
        for s in range(max_char_len):
            if i - s < 0:
                continue
            span = ground_truth[i - s:i + 1].replace(config.space, blank)
            if span in index:
                ground_truth_mat[i, s] = index[span]
    return ground_truth_mat, utt_begin_indices


class CTCSegmentationTask:
    """Task object for CTC segmentation (speechbrain CTCSegmentationTask)."""

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
    print_confidence_score = True
    print_utterance_text = True

    def __init__(self, **kwargs):
        self.set(**kwargs)

    def set(self, **kwargs):
        for key in kwargs:
            setattr(self, key, kwargs[key])

    def __str__(self):
        """One line per utterance: ``<utt_id> <name> <start:.2f> <end:.2f> <score:3.4f> <text>``
        (parsed with ``split(" ", 5)`` at iterative_utterance_alignment.py:218-219)."""
        output = ""
        num_utts = len(self.segments)
        if self.utt_ids is None:
            utt_names = [f"{self.name}_{i:04}" for i in range(num_utts)]
        else:
            utt_names = self.utt_ids
        for i, boundary in enumerate(self.segments):
            utt_entry = f"{utt_names[i]} {self.name} {boundary[0]:.2f} {boundary[1]:.2f}"
            output += utt_entry
            if self.print_confidence_score:
                output += f" {boundary[2]:3.4f}"
            if self.print_utterance_text:
                output += f" {self.text[i]}"
            output += "\n"
        return output


def _as_device_lpz(lpz, device=None):
    if isinstance(lpz, np.ndarray):
        lpz = torch.from_numpy(np.ascontiguousarray(lpz, dtype=np.float32))
    if not lpz.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("the CTC-segmentation path needs a CUDA device (no CPU fallback)")
        lpz = lpz.to(device or "cuda", non_blocking=True)
    return lpz.float()


class CTCSegmentation:
    """Align text to audio using CTC segmentation (CUDA-backed drop-in).

    ``asr_model`` must expose what the reference's EncoderASR does:
    ``encode_batch``, ``hparams.log_softmax``, ``hparams.sample_rate``,
    ``tokenizer`` (``encode_as_ids``, ``id_to_piece``, ``vocab_size``, ``unk_id``)
    and ``device``.
    """

    fs = 16000
    kaldi_style_text = True
    samples_to_frames_ratio = None
    time_stamps = "auto"
    choices_time_stamps = ["auto", "fixed"]
    text_converter = "tokenize"
    choices_text_converter = ["tokenize", "classic"]
    warned_about_misconfiguration = False

    def __init__(self, asr_model, kaldi_style_text=True, text_converter="tokenize", time_stamps="auto",
                 keep_lpz_on_device=True, **ctc_segmentation_args):
        if not hasattr(asr_model, "encode_batch") or not hasattr(asr_model, "tokenizer"):
            raise AttributeError("The given asr_model has no CTC module!")
        self.config = CtcSegmentationParameters()
        self.asr_model = asr_model
        self._encode = asr_model.encode_batch
        if hasattr(asr_model, "mods") and hasattr(getattr(asr_model.mods, "decoder", None), "ctc_forward_step"):
            self._ctc = asr_model.mods.decoder.ctc_forward_step
        else:
            self._ctc = asr_model.hparams.log_softmax
        self._tokenizer = asr_model.tokenizer
        self.keep_lpz_on_device = keep_lpz_on_device
        self.set_config(fs=asr_model.hparams.sample_rate, time_stamps=time_stamps,
                        kaldi_style_text=kaldi_style_text, text_converter=text_converter,
                        **ctc_segmentation_args)
        char_list = [asr_model.tokenizer.id_to_piece(i) for i in range(asr_model.tokenizer.vocab_size())]
        self.config.char_list = char_list
        max_char_len = max(len(c) for c in char_list)
        if len(char_list) > 500 and max_char_len >= 8:
            logger.warning("The dictionary has %d tokens with a max length of %d.", len(char_list), max_char_len)

    def set_config(self, time_stamps=None, fs=None, samples_to_frames_ratio=None, set_blank=None,
                   replace_spaces_with_blanks=None, kaldi_style_text=None, text_converter=None,
                   gratis_blank=None, min_window_size=None, max_window_size=None, scoring_length=None):
        if time_stamps is not None:
            if time_stamps not in self.choices_time_stamps:
                raise NotImplementedError(f"Parameter ´time_stamps´ has to be one of {self.choices_time_stamps}")
            self.time_stamps = time_stamps
        if fs is not None:
            self.fs = float(fs)
        if samples_to_frames_ratio is not None:
            self.samples_to_frames_ratio = float(samples_to_frames_ratio)
        if set_blank is not None:
            self.config.blank = set_blank
        if replace_spaces_with_blanks is not None:
            self.config.replace_spaces_with_blanks = replace_spaces_with_blanks
        if kaldi_style_text is not None:
            self.kaldi_style_text = kaldi_style_text
        if text_converter is not None:
            if text_converter not in self.choices_text_converter:
                raise NotImplementedError(
                    f"Parameter ´text_converter´ has to be one of {self.choices_text_converter}")
            self.text_converter = text_converter
        if min_window_size is not None:
            self.config.min_window_size = min_window_size
        if max_window_size is not None:
            self.config.max_window_size = max_window_size
        if gratis_blank is not None:
            self.config.blank_transition_cost_zero = gratis_blank
        if (self.config.blank_transition_cost_zero and self.config.replace_spaces_with_blanks
                and not self.warned_about_misconfiguration):
            logger.error("Blanks are inserted between words, and also the transition cost of blank is zero.")
            self.warned_about_misconfiguration = True
        if scoring_length is not None:
            self.config.score_min_mean_over_L = scoring_length

    def get_timing_config(self, speech_len=None, lpz_len=None):
        timing_cfg = {"index_duration": self.config.index_duration}
        if self.time_stamps == "fixed":
            if self.samples_to_frames_ratio is None:
                self.samples_to_frames_ratio = self.estimate_samples_to_frames_ratio()
            index_duration = self.samples_to_frames_ratio / self.fs
        else:
            assert self.time_stamps == "auto"
            samples_to_frames_ratio = speech_len / lpz_len
            index_duration = samples_to_frames_ratio / self.fs
        timing_cfg["index_duration"] = index_duration
        return timing_cfg

    def estimate_samples_to_frames_ratio(self, speech_len=215040):
        random_input = torch.rand(speech_len)
        lpz = self.get_lpz(random_input)
        return speech_len / lpz.shape[0]

    @torch.no_grad()
    def get_lpz(self, speech):
        """Log CTC posteriors [T, V].  Stays on the GPU when the model is there
        (the reference's ``.cpu().numpy()`` hop is the D2H/H2D round trip this
        path removes)."""
        if isinstance(speech, np.ndarray):
            speech = torch.tensor(speech)
        speech = speech.unsqueeze(0).to(self.asr_model.device)
        wav_lens = torch.tensor([1.0]).to(self.asr_model.device)
        enc = self._encode(speech, wav_lens)
        lpz = self._ctc(enc).detach().squeeze(0)
        if self.keep_lpz_on_device and lpz.is_cuda:
            return lpz.float().contiguous()
        return lpz.cpu().numpy()

    def _split_text(self, text):
        if isinstance(text, str):
            text = text.splitlines()
        text = list(filter(len, text))
        if self.kaldi_style_text:
            utt_ids_and_text = [utt.split(" ", 1) for utt in text]
            utt_ids_and_text = list(filter(lambda ui: len(ui) == 2, utt_ids_and_text))
            utt_ids = [utt[0] for utt in utt_ids_and_text]
            text = [utt[1] for utt in utt_ids_and_text]
        else:
            utt_ids = None
        return utt_ids, text

    def prepare_segmentation_task(self, text, lpz, name=None, speech_len=None):
        config = self.config
        lpz_len = lpz.shape[0]
        timing_cfg = self.get_timing_config(speech_len, lpz_len)
        config.set(**timing_cfg)
        utt_ids, text = self._split_text(text)
        if self.text_converter == "tokenize":
            token_list = [np.array(self._tokenizer.encode_as_ids(utt)) for utt in text]
            unk = self._tokenizer.unk_id() if hasattr(self._tokenizer, "unk_id") else -1
            token_list = [utt[utt != unk] if utt.size else utt for utt in token_list]
            ground_truth_mat, utt_begin_indices = prepare_token_list(config, token_list)
        else:
            assert self.text_converter == "classic"
            text_pieces = ["".join(self._tokenizer.encode_as_pieces(utt)) for utt in text]
            text_pieces = [utt.replace("<unk>", "") for utt in text_pieces]
            ground_truth_mat, utt_begin_indices = prepare_text(config, text_pieces)
        return CTCSegmentationTask(config=config, name=name, text=text, ground_truth_mat=ground_truth_mat,
                                   utt_begin_indices=utt_begin_indices, utt_ids=utt_ids, timings=None,
                                   char_probs=None, state_list=None, segments=None, done=False, lpz=lpz)

    # ------------------------------------------------------------------ alignment
    @staticmethod
    def _pack(tasks):
        n = len(tasks)
        cfg = tasks[0].config
        t_max = max(int(t.lpz.shape[0]) for t in tasks)
        c_max = max(len(t.ground_truth_mat) for t in tasks)
        k_max = max(len(t.utt_begin_indices) - 1 for t in tasks)
        g_cols = max(np.asarray(t.ground_truth_mat).reshape(len(t.ground_truth_mat), -1).shape[1] for t in tasks)
        gt = np.full((n, c_max) if g_cols == 1 else (n, c_max, g_cols), -1, np.int32)
        ub = np.zeros((n, k_max + 1), np.int32)
        n_cols = np.zeros(n, np.int32)
        n_utts = np.zeros(n, np.int32)
        in_len = np.zeros(n, np.int32)
        for i, t in enumerate(tasks):
            g = np.asarray(t.ground_truth_mat).reshape(len(t.ground_truth_mat), -1)
            if g_cols == 1:
                gt[i, :len(g)] = g[:, 0]
            else:
                gt[i, :len(g), :g.shape[1]] = g
            n_cols[i] = len(g)
            k = len(t.utt_begin_indices) - 1
            ub[i, :k + 1] = t.utt_begin_indices
            ub[i, k + 1:] = t.utt_begin_indices[-1]
            n_utts[i] = k
            in_len[i] = t.lpz.shape[0]
        if n == 1:
            lp = _as_device_lpz(tasks[0].lpz)[None]
        else:
            dev_lpz = [_as_device_lpz(t.lpz) for t in tasks]
            v = dev_lpz[0].shape[1]
            lp = torch.zeros((n, t_max, v), dtype=torch.float32, device=dev_lpz[0].device)
            for i, x in enumerate(dev_lpz):
                lp[i, :x.shape[0]] = x
        return cfg, lp, in_len, gt, n_cols, ub, n_utts

    @staticmethod
    def _run(tasks, all_prefixes, details):
        cfg, lp, in_len, gt, n_cols, ub, n_utts = CTCSegmentation._pack(tasks)
        flags = cfg.flags | (ops.SEG_ALL_PREFIXES if all_prefixes else 0)
        kw = dict(blank=cfg.blank, score_len=cfg.score_min_mean_over_L, flags=flags, details=details)
        if lp.shape[1] <= cfg.min_window_size:
            # (a multi-column ground truth runs the general kernels with window = T)
            res = ops.ctcseg_align(lp, in_len, gt, n_cols, ub, n_utts, cfg.index_duration_in_seconds, **kw)
            return cfg, res, n_utts
        # ctc-segmentation's windowed table mode: audio longer than min_window_size frames.  When the
        # backtrace leaves the window the reference catches IndexError, doubles the window and starts
        # over until max_window_size ("Check data for large repetitions or noise").
        window = int(cfg.min_window_size)
        while True:
            res = ops.ctcseg_align(lp, in_len, gt, n_cols, ub, n_utts, cfg.index_duration_in_seconds,
                                   window=window, **kw)
            if not bool((res.status & ops.WIN_WINDOW_TOO_SMALL).any()):
                return cfg, res, n_utts
            window *= 2
            if window >= cfg.max_window_size:
                raise IndexError("Maximum window size reached. Check data for large repetitions or noise.")

    @staticmethod
    def _result(task, cfg, res, i, k):
        """Reference-typed result of window i, prefix k (1-based length)."""
        n_cols_k = task.utt_begin_indices[k] + 1
        t_len = int(task.lpz.shape[0])
        timing = res.timing[i, k - 1, :n_cols_k].cpu().numpy()
        timings = np.where(timing < 0, 0.0, timing.astype(np.float64) * cfg.index_duration_in_seconds)
        char_probs = res.char_prob[i, k - 1, :t_len].cpu().numpy().astype(np.float64)
        state = res.state[i, k - 1, :t_len].cpu().numpy()
        gt = np.asarray(task.ground_truth_mat).reshape(len(task.ground_truth_mat), -1)
        state_list = [""] * t_len
        for t in np.nonzero(state != -2)[0]:
            s = int(state[t])
            if s == -1:
                state_list[t] = cfg.self_transition
            else:
                tok = int(gt[s & 0xffffff, s >> 24])  # column | (candidate index << 24)
                state_list[t] = cfg.char_list[tok] if cfg.char_list is not None else tok
        seg = res.seg[i, k - 1, :k].cpu().numpy()
        segments = [(seg[u, 0], seg[u, 1], seg[u, 2]) for u in range(k)]
        return {"name": task.name, "timings": timings, "char_probs": char_probs, "state_list": state_list,
                "segments": segments, "done": True}

    @staticmethod
    def get_segments(task):
        """Obtain segments for one task (rows A4-A7).  Raises ``AssertionError`` when
        the text is longer than the audio, like the reference (caught at
        iterative_utterance_alignment.py:390, word_level_alignment.py:130)."""
        assert type(task) == CTCSegmentationTask
        assert task.config is not None
        if len(task.ground_truth_mat) > task.lpz.shape[0] and task.config.skip_prob <= task.config.max_prob:
            raise AssertionError("Audio is shorter than text!")
        cfg, res, n_utts = CTCSegmentation._run([task], all_prefixes=False, details=True)
        return CTCSegmentation._result(task, cfg, res, 0, int(n_utts[0]))

    @staticmethod
    def get_segments_batch(tasks):
        """Many independent tasks in one launch (word-level alignment / search on
        speech: /root/reference/src/word_level_alignment.py:35, search_on_speech.py:45).
        Returns a list with a result dict, or an ``AssertionError`` instance, per task."""
        if not tasks:
            return []
        cfg, res, n_utts = CTCSegmentation._run(tasks, all_prefixes=False, details=True)
        status = res.status.cpu().numpy()
        out = []
        for i, task in enumerate(tasks):
            if status[i] & 4:
                out.append(AssertionError("Audio is shorter than text!"))
            else:
                out.append(CTCSegmentation._result(task, cfg, res, i, int(n_utts[i])))
        return out

    @staticmethod
    def prefix_segments(tasks):
        """All utterance-prefixes of every task from ONE table fill (the shrinking
        transcript iterations of iterative_utterance_alignment.py:203-385).
        Returns the device-side :class:`ops.SegAlignment` and ``n_utts``."""
        cfg, res, n_utts = CTCSegmentation._run(tasks, all_prefixes=True, details=False)
        return res, n_utts

    def __call__(self, speech, text, name=None):
        if name is None:
            name = "utt"
        utt_ids, text = self._split_text(text)
        lpz = self.get_lpz(speech)
        task = self.prepare_segmentation_task(text, lpz, name, speech.shape[0])
        segments = self.get_segments(task)
        task.set(**segments)
        assert task.done
        return task
