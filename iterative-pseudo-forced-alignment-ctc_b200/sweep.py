"""Corpus-resident anchor sweep: the iterative anchor loop of many files in lock step.

Drives ``ipfa_sweep_step_device`` (``csrc/anchor_sweep.cu``), which restates the per-row
control flow of /root/reference/src/iterative_utterance_alignment.py:67-402 on the device
for files whose emissions were computed once and stay in
HBM.  One iteration aligns one window of every active file -- window construction, the
all-prefix CTC segmentation, the accept / shrink / revert decision and the anchor update
-- with no host round trip; the host polls the per-file status words every few iterations.

The result rows are the ``file_alignments`` entries the reference appends (:258-260),
``RESULT_COLUMNS`` order, ready for ``to_csv(sep='\\t')``.
"""
import ctypes

import numpy as np
import torch

from . import hostglue as hg
from ._lib import check, lib

ACTIVE, DONE, STOP_WINDOW, STOP_EXCEPTIONS, NEEDS_RECALC, CAPACITY, NO_AUDIO = range(7)
STATUS_NAMES = ["active", "done", "window_to_stop", "exceptions_limit", "needs_recalc", "capacity", "no_audio"]

DEFAULT_MODE = "auto"   # AnchorSweep(mode=None): "auto" | "resident" | "lockstep"

_vp, _i32, _i64, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double


class _Corpus(ctypes.Structure):
    _fields_ = [("lp", _vp), ("stride_t", _i64), ("V", _i32), ("blank", _i32), ("n_files", _i32),
                ("reserved", _i32), ("file_frame0", _vp), ("file_frames", _vp), ("file_samples", _vp),
                ("row_first", _vp), ("row_type", _vp), ("row_start", _vp), ("row_end", _vp),
                ("row_utt_end", _vp), ("utt_first", _vp), ("utt_col", _vp), ("utt_chars", _vp),
                ("file_tok0", _vp), ("tokens", _vp), ("file_ready", _vp)]


class _State(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("row", "utt", "anchor", "prop", "next_ns", "follow_start", "exc", "status",
                                   "need", "recalc_row", "n_windows", "cells", "frames", "clip")]


class _Params(ctypes.Structure):
    _fields_ = [("threshold", _f64), ("max_window_size", _f64), ("window_to_stop", _f64),
                ("min_text_to_audio_prop", _f64), ("samples_to_frames_ratio", _f64), ("index_duration", _f64),
                ("short_len", _i32), ("max_exceptions", _i32), ("sample_rate", _i32), ("frame_shift", _i32),
                ("score_len", _i32), ("seg_flags", _i32)]


class SweepFile:
    """One audio file of the corpus.

    ``lpz``: fp32 [T, V] log-posteriors of the WHOLE file (torch tensor, any device);
    ``n_samples``: audio samples (``torchaudio.info(...).num_frames``);
    ``rows``: the file's TSV rows after ``fix_time_reference`` (:52-57), each a dict with
    ``Type`` ('Speech' / 'Non-Speech'), ``Start``, ``End``, ``utterances`` (list of str, the
    output of ``prepare_text`` at :89) and the metadata the result rows carry
    (``Channel``, ``Speaker_ID``, ``Database``)."""

    def __init__(self, file_id, audio_path, lpz, n_samples, rows):
        self.file_id, self.audio_path, self.lpz, self.n_samples, self.rows = \
            file_id, audio_path, lpz, int(n_samples), rows


def rows_from_dataframe(file_df, max_words_sequence=24):
    """TSV rows (after ``fix_time_reference``) -> ``SweepFile.rows`` (:79-89)."""
    rows = []
    for _, row in file_df.iterrows():
        if row['Type'] == 'Non-Speech':
            rows.append({"Type": 'Non-Speech', "Start": float(row['Start']), "End": float(row['End']),
                         "utterances": []})
            continue
        transcript = hg.prepare_text(str(row['Transcription']).upper(), max_words_sequence=max_words_sequence)
        if isinstance(transcript, str):
            transcript = [transcript]
        rows.append({"Type": row['Type'], "Start": float(row['Start']), "End": float(row['End']),
                     "utterances": list(transcript), "Channel": row['Channel'],
                     "Speaker_ID": row['Speaker_ID'], "Database": row['Database']})
    return rows


def pack_files(files, tokenizer, blank=0):
    """Flat host arrays of ``ipfa_sweep_corpus`` (include/ipfa_b200.h) + per-file utterance texts.

    File f's token stream is ``blank, tokens(u0), blank, tokens(u1), ..., blank``: the
    ground-truth column ``prepare_token_list`` builds for utterances [a, b) of the file is
    ``[-1] + stream[utt_col[a] : utt_col[b] + 1]``."""
    unk = tokenizer.unk_id() if hasattr(tokenizer, "unk_id") else -1
    a = {k: [] for k in ("file_frame0", "file_frames", "file_samples", "row_type", "row_start", "row_end",
                         "row_utt_end", "utt_col", "utt_chars", "file_tok0", "tokens")}
    a["row_first"], a["utt_first"] = [0], [0]
    all_texts = []
    frame0 = 0
    for f in files:
        a["file_frame0"].append(frame0)
        a["file_frames"].append(int(f.lpz.shape[0]))
        frame0 += int(f.lpz.shape[0])
        a["file_samples"].append(f.n_samples)
        a["file_tok0"].append(len(a["tokens"]))
        texts = []
        col = 0
        for row in f.rows:
            speech = row["Type"] != 'Non-Speech'
            a["row_type"].append(0 if speech else 1)
            a["row_start"].append(float(row["Start"]))
            a["row_end"].append(float(row["End"]))
            if speech:
                for utt in row["utterances"]:
                    ids = np.asarray(tokenizer.encode_as_ids(utt), dtype=np.int64)
                    ids = ids[ids != unk] if ids.size else ids
                    if ids.size == 0 or ids[-1] == blank or ids[0] == blank:
                        # prepare_token_list would merge the neighbouring blanks; such windows are
                        # not slices of one token stream
                        raise ValueError(f"utterance without tokens in {f.file_id!r}: {utt!r}")
                    a["utt_col"].append(col)
                    a["utt_chars"].append(len(utt))
                    a["tokens"].append(blank)
                    a["tokens"].extend(int(i) for i in ids)
                    col += 1 + ids.size
                    texts.append(utt)
            a["row_utt_end"].append(len(texts))
        a["utt_col"].append(col)   # closing blank
        a["utt_chars"].append(0)
        a["tokens"].append(blank)
        a["row_first"].append(len(a["row_type"]))
        a["utt_first"].append(len(a["utt_col"]))
        all_texts.append(texts)
    dtypes = dict(file_frame0=np.int64, file_samples=np.int64, file_tok0=np.int64, row_start=np.float64,
                  row_end=np.float64)
    return {k: np.asarray(v, dtype=dtypes.get(k, np.int32)) for k, v in a.items()}, all_texts


def sort_longest_first(files):
    """Files ordered by emission frames, longest first: the anchor loop of a file is a serial chain,
    so the longest files set the duration of the sweep and go into the smallest lock-step groups."""
    return sorted(files, key=lambda f: -int(f.lpz.shape[0]))


class SweepCorpus:
    """Files packed into the device arrays of ``ipfa_sweep_corpus``; emissions resident in HBM."""

    def __init__(self, files, tokenizer, blank=0, device=None):
        if not files:
            raise ValueError("empty corpus")
        self.files = files
        device = torch.device(device) if device is not None else files[0].lpz.device
        if device.type != "cuda":
            raise ValueError("the anchor sweep runs on a CUDA device (no CPU fallback)")
        self.device = device
        self.blank = blank
        self.host, self.texts = pack_files(files, tokenizer, blank)
        self.lp = torch.cat([f.lpz.to(device=device, dtype=torch.float32) for f in files], dim=0).contiguous()
        self.V = int(self.lp.shape[1])
        self.n_slots = len(self.host["utt_col"])
        self.arrays = {k: torch.as_tensor(v, device=device) for k, v in self.host.items()}
        self._ready, self._copy_stream = None, None

    def struct(self):
        c = _Corpus()
        c.lp = self.lp.data_ptr()
        c.stride_t = self.lp.stride(0)
        c.V, c.blank, c.n_files = self.V, self.blank, len(self.files)
        for k, t in self.arrays.items():
            setattr(c, k, t.data_ptr())
        c.file_ready = self._ready.data_ptr() if self._ready is not None else None
        return c

    # ------------------------------------------------------------------ uploads overlapping the sweep
    def begin_upload(self, host_lp):
        """Start copying the emissions of all files from pinned host memory (``host_lp``: fp32
        [total frames, V], the files in corpus order) on a copy stream, file by file, each followed by its
        word of ``ipfa_sweep_corpus.file_ready``: a file-resident sweep launched right after starts on the
        first (longest) files while the others are still in flight.  The lock-step path waits for the whole
        upload (``finish_upload``).  Asynchronous."""
        dev = self.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._ones = torch.ones(len(self.files), dtype=torch.int32).pin_memory()
            self._ready_buf = torch.zeros(len(self.files), dtype=torch.int32, device=dev)
        cur = torch.cuda.current_stream(dev)
        self._ready = self._ready_buf
        self._ready.zero_()                      # on the caller's stream, before the sweep's launch ...
        zeroed = torch.cuda.Event()
        zeroed.record(cur)
        self._copy_stream.wait_event(zeroed)     # ... and before the first flag
        frame0, frames = self.host["file_frame0"], self.host["file_frames"]
        with torch.cuda.stream(self._copy_stream):
            for f in range(len(self.files)):
                a, b = int(frame0[f]), int(frame0[f]) + int(frames[f])
                self.lp[a:b].copy_(host_lp[a:b], non_blocking=True)
                self._ready[f:f + 1].copy_(self._ones[f:f + 1], non_blocking=True)

    def finish_upload(self):
        """Wait (on the current stream) for an upload started by :meth:`begin_upload`."""
        if self._copy_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._copy_stream)

    def total_frames(self):
        return int(self.lp.shape[0])


class AnchorSweep:
    """Runs the anchor loop of every file of a :class:`SweepCorpus` on its device.

    ``mode``: ``"resident"`` -- one persistent launch, a CTA per file, every file runs its own chain of
    windows on one SM and the SMs take files from a queue (``ipfa_sweep_resident_device``; list the files
    longest first); ``"lockstep"`` -- one window of every active file per iteration, three launches per
    iteration (``ipfa_sweep_step_device``); ``"auto"`` (default): resident whenever the library supports
    the corpus (dense emissions with ``V % 4 == 0``, default table flags), lock step otherwise.  Both
    produce the same rows, status words and counters.

    Lock-step mode only: ``groups``: independent lock-step groups of files, each on its own stream with its own launch
    capacity; ``use_graphs``: the launches of ``steps_per_poll`` iterations of all groups are
    captured once into a CUDA graph and replayed (the sweep is a chain of short dependent kernels,
    so with many groups the host's launch rate would otherwise bound it)."""

    def __init__(self, corpus, index_duration, samples_to_frames_ratio, frame_shift=None, sample_rate=16000,
                 threshold=-2.0, short_utterance_len=30, max_window_size=70.0, window_to_stop=500.0,
                 min_text_to_audio_prop=0.8, max_text_to_audio_prop_exec=10, scoring_length=30, seg_flags=2,
                 capacity=None, groups=1, use_graphs=True, mode=None):
        mode = DEFAULT_MODE if mode is None else mode
        if mode not in ("auto", "resident", "lockstep"):
            raise ValueError(f"mode {mode!r}")
        self.corpus = corpus
        self.groups = int(groups)
        self.use_graphs = bool(use_graphs)
        p = _Params()
        p.threshold, p.max_window_size, p.window_to_stop = threshold, max_window_size, window_to_stop
        p.min_text_to_audio_prop = min_text_to_audio_prop
        p.samples_to_frames_ratio = float(samples_to_frames_ratio)
        p.index_duration = float(index_duration)
        p.short_len, p.max_exceptions, p.sample_rate = short_utterance_len, max_text_to_audio_prop_exec, sample_rate
        p.frame_shift = int(frame_shift if frame_shift is not None else round(samples_to_frames_ratio))
        p.score_len, p.seg_flags = scoring_length, seg_flags
        self.params = p
        self.state = None
        self.reset()
        self.mode = mode
        if mode != "lockstep":
            # the resident kernel has one capacity for all files; "auto" asks the library
            probe = list(capacity) if capacity else self._initial_capacity(0, len(corpus.files))
            supported = self._resident_bytes(probe) > 0
            if mode == "resident" and not supported:
                raise ValueError("the file-resident sweep does not cover this corpus (see ipfa_b200.h); "
                                 "use mode='lockstep'")
            self.mode = "resident" if supported else "lockstep"
        if self.mode == "resident":
            self.groups = 1
        self.ranges = self._group_ranges()
        first = list(capacity) if capacity else None
        self.capacities = [list(first) if first else self._initial_capacity(lo, hi) for lo, hi in self.ranges]
        self._ws = [None] * len(self.ranges)
        self._ws_cap = [None] * len(self.ranges)
        dev = corpus.device
        self._streams = [torch.cuda.Stream(dev) for _ in self.ranges] if len(self.ranges) > 1 else None
        self._graph, self._graph_key, self._graph_launches = None, None, 0
        self.kernel_launches = 0

    @property
    def capacity(self):
        """Largest (Tmax, Cmax, Kmax) over the groups."""
        return [max(c[i] for c in self.capacities) for i in range(3)]

    def reset(self):
        """Initial loop state of every file (:36-43) and empty outputs (in place: graphs keep pointers)."""
        dev, n, nan = self.corpus.device, len(self.corpus.files), float("nan")
        if self.state is None:
            i32 = lambda *shape: torch.zeros(shape, dtype=torch.int32, device=dev)
            f64 = lambda: torch.zeros(n, dtype=torch.float64, device=dev)
            self.state = dict(row=i32(n), utt=i32(n), anchor=f64(), prop=f64(), next_ns=i32(n), follow_start=f64(),
                              exc=i32(n), status=i32(n), need=i32(n, 3), recalc_row=i32(n), n_windows=i32(n),
                              cells=torch.zeros(n, dtype=torch.int64, device=dev),
                              frames=torch.zeros(n, dtype=torch.int64, device=dev),
                              clip=torch.zeros((n, 2), dtype=torch.int64, device=dev))
            self.out_seg = torch.zeros((self.corpus.n_slots, 4), dtype=torch.float64, device=dev)
            self.out_info = torch.zeros((self.corpus.n_slots, 2), dtype=torch.int32, device=dev)
        for k, t in self.state.items():
            t.fill_(nan if k in ("anchor", "follow_start") else (-1 if k in ("recalc_row", "clip") else 0))
        self.out_seg.zero_()
        self.out_info.fill_(-1)
        self.steps = 0

    def _initial_capacity(self, lo, hi):
        """(Tmax, Cmax, Kmax) that holds every single row's own window of files [lo, hi)."""
        c, p = self.corpus, self.params
        h = c.host
        t_max, c_max, k_max = 64, 16, 2
        for f in range(lo, hi):
            u_prev = 0
            for r in range(h["row_first"][f], h["row_first"][f + 1]):
                if h["row_type"][r] == 1:
                    continue
                u1 = int(h["row_utt_end"][r])
                s0 = int(h["utt_first"][f])
                c_max = max(c_max, int(h["utt_col"][s0 + u1] - h["utt_col"][s0 + u_prev]) + 2)
                k_max = max(k_max, u1 - u_prev)
                t_max = max(t_max, int((h["row_end"][r] - h["row_start"][r]) * p.sample_rate) // p.frame_shift + 1)
                u_prev = u1
        # head-room for pending utterances and anchors that lag behind the row start
        return [min(8000, int(t_max * 1.5) + 8), int(c_max * 2) + 8, int(k_max * 2) + 2]

    # ------------------------------------------------------------------ groups
    # Lock step makes every iteration as long as its longest window, and one launch is sized for
    # the widest window in flight.  Files are therefore cut into `groups` contiguous ranges that
    # iterate independently on their own streams with their own capacity (a group is the same
    # corpus / state with the per-file pointers shifted), so one file's long window only holds back
    # its own group and the groups' kernels overlap on the GPU.
    def _group_ranges(self):
        """Contiguous file ranges of growing size (1 : 2 : 3 : ...).  With the files ordered longest
        first (:func:`sort_longest_first`) the files that set the sweep's critical path share their
        lock step with few others."""
        n = len(self.corpus.files)
        g = max(1, min(self.groups, n))
        total = g * (g + 1) // 2
        edges = [round(n * (i * (i + 1) // 2) / total) for i in range(g + 1)]
        return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]

    _PER_FILE_CORPUS = ("file_frame0", "file_frames", "file_samples", "row_first", "utt_first", "file_tok0")

    def _structs(self, lo, hi):
        c = self.corpus
        cs = c.struct()
        cs.n_files = hi - lo
        for k in self._PER_FILE_CORPUS:
            t = c.arrays[k]
            setattr(cs, k, t.data_ptr() + lo * t.element_size())
        ss = _State()
        for k, t in self.state.items():
            setattr(ss, k, t.data_ptr() + lo * t.stride(0) * t.element_size())
        return cs, ss

    def _workspace(self, g):
        cap = tuple(self.capacities[g])
        if self._ws_cap[g] != cap:
            lo, hi = self.ranges[g]
            nbytes = lib().ipfa_sweep_workspace_bytes(hi - lo, cap[0], cap[1], cap[2], self.corpus.V)
            self._ws[g] = None
            self._ws[g] = torch.empty(nbytes, dtype=torch.uint8, device=self.corpus.device)
            self._ws_cap[g] = cap
        return self._ws[g]

    def _issue(self, n_steps):
        """Queue ``n_steps`` iterations of every group (fork from / join into the current stream)."""
        c = self.corpus
        cur = torch.cuda.current_stream(c.device)
        multi = self._streams is not None
        if multi:
            start = torch.cuda.Event()
            start.record(cur)
        for g, (lo, hi) in enumerate(self.ranges):
            st = self._streams[g] if multi else cur
            if multi:
                st.wait_event(start)
            ws, cap = self._workspace(g), self.capacities[g]
            cs, ss = self._structs(lo, hi)
            rc = lib().ipfa_sweep_step_device(
                ctypes.byref(cs), ctypes.byref(self.params), ctypes.byref(ss), self.out_seg.data_ptr(),
                self.out_info.data_ptr(), self.steps, n_steps, cap[0], cap[1], cap[2], ws.data_ptr(), ws.numel(),
                st.cuda_stream)
            check(rc, "ipfa_sweep_step_device")
        if multi:
            for st in self._streams:
                done = torch.cuda.Event()
                done.record(st)
                cur.wait_event(done)

    def step(self, n_steps=1):
        """``n_steps`` iterations of every group; asynchronous on the current stream."""
        c = self.corpus
        for g in range(len(self.ranges)):
            self._workspace(g)
        with torch.cuda.device(c.device):
            if not self.use_graphs:
                before = lib().ipfa_launch_count()
                self._issue(n_steps)
                self.kernel_launches += lib().ipfa_launch_count() - before
            else:
                key = (n_steps, tuple(tuple(cap) for cap in self.capacities), tuple(w.data_ptr() for w in self._ws))
                if self._graph_key != key:
                    self._graph = None
                    torch.cuda.synchronize(c.device)
                    graph = torch.cuda.CUDAGraph()
                    before = lib().ipfa_launch_count()
                    with torch.cuda.graph(graph):
                        self._issue(n_steps)
                    self._graph_launches = lib().ipfa_launch_count() - before
                    self._graph, self._graph_key = graph, key
                self._graph.replay()
                self.kernel_launches += self._graph_launches
        self.steps += n_steps

    # ------------------------------------------------------------------ file-resident mode
    def _resident_bytes(self, cap):
        cs, _ = self._structs(0, len(self.corpus.files))
        return int(lib().ipfa_sweep_resident_workspace_bytes(ctypes.byref(cs), ctypes.byref(self.params),
                                                             int(cap[0]), int(cap[1]), int(cap[2])))

    def run_resident(self):
        """One persistent launch: every ACTIVE file runs until it leaves ACTIVE.  Asynchronous on the
        current stream."""
        c = self.corpus
        cap = self.capacities[0]
        nbytes = self._resident_bytes(cap)
        if nbytes <= 0:   # a window outgrew the resident kernel's range: the lock-step path takes over
            self.mode = "lockstep"
            return False
        if self._ws_cap[0] != ("resident",) + tuple(cap):
            self._ws[0] = None
            self._ws[0] = torch.empty(nbytes, dtype=torch.uint8, device=c.device)
            self._ws_cap[0] = ("resident",) + tuple(cap)
        ws = self._ws[0]
        cs, ss = self._structs(0, len(c.files))
        with torch.cuda.device(c.device):
            before = lib().ipfa_launch_count()
            rc = lib().ipfa_sweep_resident_device(
                ctypes.byref(cs), ctypes.byref(self.params), ctypes.byref(ss), self.out_seg.data_ptr(),
                self.out_info.data_ptr(), cap[0], cap[1], cap[2], ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(c.device).cuda_stream)
            check(rc, "ipfa_sweep_resident_device")
            self.kernel_launches += lib().ipfa_launch_count() - before
        return True

    def run(self, steps_per_poll=8, max_steps=100000, recalc_fn=None):
        """Iterate until no file is active.  ``recalc_fn(sweep, file_index)`` may re-spread the rows
        of a file that stopped with NEEDS_RECALC (the reference's ``fix_text_to_time_proportion``,
        :127-146) and return True to resume it.  Returns the per-file status array (host)."""
        while self.steps < max_steps:
            if self.mode == "resident":
                if not self.run_resident():
                    continue
                self.steps = int(self.state["n_windows"].max().item())  # windows of the longest chain
            else:
                self.corpus.finish_upload()   # only the file-resident kernel honours the per-file ready words
                self.step(steps_per_poll)
            status = self.state["status"].cpu().numpy()
            if (status == CAPACITY).any():
                need = self.state["need"].cpu().numpy()
                grew, too_long = False, False
                for g, (lo, hi) in enumerate(self.ranges):
                    mask = status[lo:hi] == CAPACITY
                    if not mask.any():
                        continue
                    want = need[lo:hi][mask].max(axis=0)
                    cap = self.capacities[g]
                    for i in range(3):
                        if want[i] > cap[i]:
                            cap[i] = int(want[i] * 1.25) + 4
                            grew = True
                    cap[0] = min(cap[0], 8000)
                    too_long |= bool(want[0] > 8000)
                if too_long or not grew:
                    # windowed table mode (T > 8000 frames) is not on the device path
                    break
                self.state["status"][self.state["status"] == CAPACITY] = ACTIVE
                continue
            if recalc_fn is not None and (status == NEEDS_RECALC).any():
                resumed = False
                for f in np.nonzero(status == NEEDS_RECALC)[0]:
                    resumed |= bool(recalc_fn(self, int(f)))
                if resumed:
                    continue
            if not (status == ACTIVE).any():
                break
        return self.state["status"].cpu().numpy()

    # ------------------------------------------------------------------ host policy hand-off
    def set_row_times(self, f, starts, ends):
        """New Start / End (seconds) of file f's rows, e.g. after ``fix_text_to_time_proportion``."""
        r0, r1 = int(self.corpus.host["row_first"][f]), int(self.corpus.host["row_first"][f + 1])
        assert len(starts) == len(ends) == r1 - r0
        dev = self.corpus.device
        self.corpus.arrays["row_start"][r0:r1] = torch.as_tensor(np.array(starts, np.float64), device=dev)
        self.corpus.arrays["row_end"][r0:r1] = torch.as_tensor(np.array(ends, np.float64), device=dev)
        self.corpus.host["row_start"][r0:r1] = starts
        self.corpus.host["row_end"][r0:r1] = ends

    def resume_after_recalc(self, f):
        """File f stopped with NEEDS_RECALC and the host has re-spread its rows (:127-146): continue
        with the alignment of the same row."""
        self.state["recalc_row"][f] = self.state["row"][f]
        self.state["status"][f] = ACTIVE

    # ------------------------------------------------------------------ results
    def file_rows(self):
        """Per file: the rows the reference would have appended to ``file_alignments`` (:258-260)."""
        c = self.corpus
        seg = self.out_seg.cpu().numpy()
        info = self.out_info.cpu().numpy()
        out = []
        for f, sf in enumerate(c.files):
            s0 = int(c.host["utt_first"][f])
            rows = []
            for u, text in enumerate(c.texts[f]):
                if info[s0 + u, 0] < 0:
                    continue
                clip_start, start, end, score = (float(x) for x in seg[s0 + u])
                meta = sf.rows[int(info[s0 + u, 1])]
                abs_start, abs_end = clip_start + start, clip_start + end
                segment_id = "_".join([sf.file_id, str(abs_start), str(abs_end)])
                rows.append([segment_id, sf.audio_path, meta.get("Channel"), end - start, abs_start, abs_end,
                             score, text, meta.get("Speaker_ID"), meta.get("Database")])
            out.append(rows)
        return out

    def stats(self):
        return {"steps": self.steps, "windows": int(self.state["n_windows"].sum().item()),
                "cells": int(self.state["cells"].sum().item()),
                "frames": int(self.state["frames"].sum().item()),
                "status": {STATUS_NAMES[k]: int(v) for k, v in
                           zip(*np.unique(self.state["status"].cpu().numpy(), return_counts=True))}}


class _Quiet:
    def debug(self, *a, **k):
        pass


def dataframe_recalc(frames, vads, real_lengths, logger=None):
    """``recalc_fn`` for :meth:`AnchorSweep.run` that applies the reference's
    ``fix_text_to_time_proportion`` (:127-146) on the host when a file stops with NEEDS_RECALC.

    ``frames[f]``: the file's TSV rows after ``fix_time_reference`` (a DataFrame, updated in
    place here like ``file_df`` in the reference); ``vads[f]``: its VAD table; ``real_lengths[f]``:
    audio seconds.  The device keeps the row STRUCTURE it was given, so the file is resumed only
    when the re-spreading leaves the number and types of rows unchanged (one VAD segment, the
    common case); otherwise it stays NEEDS_RECALC for the per-file loop."""
    logger = logger or _Quiet()

    def recalc(sweep, f):
        h = sweep.corpus.host
        r0 = int(h["row_first"][f])
        row = int(sweep.state["row"][f])
        anchor = float(sweep.state["anchor"][f])
        df = frames[f]
        clip_start = anchor if anchor == anchor else float(df.iloc[row]['Start'])
        speech = [i for i in range(row + 1) if df.iloc[i]['Type'] != 'Non-Speech']
        ends = h["row_utt_end"][r0:r0 + row + 1]
        list_of_splits = [int(ends[i] - (ends[i - 1] if i else 0)) for i in speech]  # :90
        n_aligned_utts = int(sweep.state["utt"][f])                                  # len(file_alignments)
        n_segments = int((df['Type'] != 'Non-Speech').sum())
        new = hg.fix_text_to_time_proportion(df, vads[f], real_lengths[f] - clip_start,
                                             hg.get_n_aligned_rows(list_of_splits, n_aligned_utts),
                                             n_segments, clip_start, logger)
        if len(new.index) != len(df.index) or list(new['Type']) != list(df['Type']):
            return False
        frames[f] = new
        sweep.set_row_times(f, new['Start'].astype(float).to_numpy(), new['End'].astype(float).to_numpy())
        sweep.resume_after_recalc(f)
        return True
    return recalc


def align_files_resident(asr_model, aligner, jobs, samples_to_frames_ratio, threshold=-2.0, short_utterance_len=30,
                         max_words_sequence=24, max_window_size=70.0, window_to_stop=500.0,
                         min_text_to_audio_prop=0.8, max_text_to_audio_prop_exec=10, groups=None, mode=None):
    """Anchor loop of several files with file-level emissions resident on the GPU.

    ``jobs``: list of ``(audio_path, file_df, vad_file_df)`` like the arguments of
    ``get_file_iterative_segmentation``.  Every file is encoded ONCE (``aligner.get_lpz`` on the
    whole normalised audio) instead of once per window (:149-201).  Returns
    ``(rows_per_file, status_per_file)``; files whose status is not DONE stopped where the
    reference's host-side policy takes over (see ``ipfa_b200.h``) and carry the rows found so far."""
    files, frames, vads, lengths = [], [], [], []
    for audio_path, file_df, vad_file_df in jobs:
        info = hg.audio_info(audio_path)
        audio, sr = hg.audio_load(audio_path, channels_first=False)
        lpz = aligner.get_lpz(asr_model.audio_normalizer(audio, sr))
        if not torch.is_tensor(lpz):
            lpz = torch.as_tensor(lpz)
        real_length = info.num_frames / info.sample_rate
        fixed = hg.fix_time_reference(file_df, vad_file_df, real_length, len(file_df.index))
        file_id = audio_path.split('/')[-1].replace('.wav', '')
        files.append(SweepFile(file_id, audio_path, lpz.cuda(), info.num_frames,
                               rows_from_dataframe(fixed, max_words_sequence)))
        frames.append(fixed)
        vads.append(vad_file_df)
        lengths.append(real_length)
    order = sorted(range(len(files)), key=lambda i: -int(files[i].lpz.shape[0]))  # longest first
    files = [files[i] for i in order]
    frames, vads, lengths = ([x[i] for i in order] for x in (frames, vads, lengths))
    corpus = SweepCorpus(files, asr_model.tokenizer, blank=aligner.config.blank)
    if groups is None:  # measured on 201 files: 32 groups of ~6 files (bench.py --workload c5)
        groups = min(32, max(1, len(files) // 6))
    fs = int(asr_model.hparams.sample_rate)
    sweep = AnchorSweep(corpus, index_duration=samples_to_frames_ratio / fs,
                        samples_to_frames_ratio=samples_to_frames_ratio, sample_rate=fs, threshold=threshold,
                        short_utterance_len=short_utterance_len, max_window_size=max_window_size,
                        window_to_stop=window_to_stop, min_text_to_audio_prop=min_text_to_audio_prop,
                        max_text_to_audio_prop_exec=max_text_to_audio_prop_exec,
                        scoring_length=aligner.config.score_min_mean_over_L, seg_flags=aligner.config.flags,
                        groups=groups, mode=mode)
    status = sweep.run(recalc_fn=dataframe_recalc(frames, vads, lengths))
    rows = sweep.file_rows()
    inverse = {src: pos for pos, src in enumerate(order)}  # back to the order of `jobs`
    return [rows[inverse[i]] for i in range(len(order))], np.asarray([status[inverse[i]] for i in range(len(order))])
