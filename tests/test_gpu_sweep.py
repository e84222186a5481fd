"""Device-resident anchor sweep (BASELINE configs[4]) vs the per-file host loop driven by the
CPU oracle: identical ``file_alignments`` rows for every file of a small synthetic corpus.

The acoustic model is a test double whose emissions are a pure function of the absolute
frame index, so "re-encode the clip" (the reference, :149-201) and "slice the file's
emissions" (the sweep) see the same log-probabilities."""
import importlib

import numpy as np
import pandas as pd
import pytest
import torch

from scripted_asr import ScriptedASR, make_schedule
from test_gpu_pipeline import _utterances, oracle_window_fn

pytestmark = pytest.mark.gpu

PKG = "iterative-pseudo-forced-alignment-ctc_b200"
hg = importlib.import_module(PKG + ".hostglue")
anchor = importlib.import_module(PKG + ".anchor")
cs = importlib.import_module(PKG + ".ctc_segmentation")
sweep = importlib.import_module(PKG + ".sweep")
stub = importlib.import_module(PKG + ".stub_asr")


@pytest.fixture(autouse=True, params=["resident", "lockstep"])
def sweep_mode(request, monkeypatch):
    """Every test of this module runs on both device paths: the file-resident persistent kernel and the
    lock-step iterations.  (The synthetic corpora have V = 32 dense emissions, which both cover.)"""
    monkeypatch.setattr(sweep, "DEFAULT_MODE", request.param)
    return request.param


class ExactASR(ScriptedASR):
    """The synthetic 'audio' carries its absolute sample index exactly (float64)."""

    @torch.no_grad()
    def encode_batch(self, wavs, wav_lens=None):
        x = wavs[0].double()
        n = x.shape[0] // self.stride
        if n == 0:
            return torch.zeros(1, 0, self.logits.shape[1], device=self.device)
        f0 = int(x[0].item()) // self.stride
        idx = torch.arange(f0, f0 + n).clamp(max=self.logits.shape[0] - 1)
        return self.logits[idx].unsqueeze(0).to(self.device)


def _corpus_file(name, n_utts, seed, corrupt=(), vad_gap=None):
    rng = np.random.default_rng(seed)
    utts = _utterances(rng, n_utts)
    tok = stub.CharTokenizer()
    frames, spans = make_schedule([u.upper() for u in utts], tok, rng)
    total = (len(frames) + 40) * 320
    dur = total / 16000
    wav = f"/synthetic/{name}.wav"
    df = pd.DataFrame({'Sample_ID': [f"{name}_{i}" for i in range(n_utts)], 'Sample_Path': [wav] * n_utts,
                       'Channel': [1] * n_utts, 'Audio_Length': [dur / n_utts] * n_utts,
                       'Start': [0.0] * n_utts, 'End': [dur] * n_utts, 'Transcription': utts,
                       'Speaker_ID': ['spk' + name] * n_utts, 'Database': ['synthetic'] * n_utts})
    if vad_gap is None:
        vad = pd.DataFrame({'Sample_Path': [wav], 'Start': [0.0], 'End': [dur], 'Segment_Length': [dur]})
    else:
        a, b = vad_gap
        vad = pd.DataFrame({'Sample_Path': [wav] * 2, 'Start': [0.0, b], 'End': [a, dur],
                            'Segment_Length': [a, dur - b]})
    asr = ExactASR(frames, total, device="cuda", corrupt=corrupt, seed=seed)
    return dict(name=name, wav=wav, df=df, vad=vad, total=total, asr=asr)


@pytest.fixture()
def exact_audio(monkeypatch):
    """audio_info / audio_load stand-ins: sample i of every synthetic file has the value i."""
    totals = {}

    def info(path):
        return hg.AudioInfo(totals[path], 16000, 1)

    def load(path, frame_offset=0, num_frames=-1, channels_first=False):
        total = totals[path]
        frame_offset = max(0, min(int(frame_offset), total))
        n = total - frame_offset if num_frames is None or num_frames < 0 else \
            max(0, min(int(num_frames), total - frame_offset))
        audio = torch.arange(frame_offset, frame_offset + n, dtype=torch.float64).reshape(-1, 1)
        return (audio.t().contiguous() if channels_first else audio), 16000

    monkeypatch.setattr(hg, "audio_info", info)
    monkeypatch.setattr(hg, "audio_load", load)
    return totals


KW = dict(threshold=-2.0, short_utterance_len=30, max_words_sequence=8, max_window_size=70.0)


def _reference_rows(spec, tmp, window_fn):
    aligner = cs.CTCSegmentation(spec["asr"], kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    aligner.samples_to_frames_ratio = 320.0
    return anchor.get_file_iterative_segmentation(spec["asr"], aligner, spec["wav"], spec["df"].copy(),
                                                  spec["vad"].copy(), 320.0, str(tmp), window_fn=window_fn, **KW)


def _sweep_files(specs):
    files = []
    for s in specs:
        n_segments = len(s["df"].index)
        fixed = hg.fix_time_reference(s["df"], s["vad"], s["total"] / 16000, n_segments)
        lpz = torch.log_softmax(s["asr"].logits[: s["total"] // 320].float(), dim=-1).cuda()
        files.append(sweep.SweepFile(s["name"], s["wav"], lpz, s["total"],
                                     sweep.rows_from_dataframe(fixed, KW["max_words_sequence"])))
    return files


def test_sweep_rows_equal_the_per_file_loop(exact_audio, tmp_path):
    specs = [_corpus_file("a", 14, 11),
             _corpus_file("b", 22, 12, corrupt=((300, 420), (900, 960))),
             _corpus_file("c", 9, 13, corrupt=((150, 260),)),
             _corpus_file("d", 18, 14, vad_gap=(20.0, 24.0)),
             _corpus_file("e", 3, 15)]
    for s in specs:
        exact_audio[s["wav"]] = s["total"]
    ref = [_reference_rows(s, tmp_path, oracle_window_fn) for s in specs]
    assert sum(len(r) for r in ref) > 40

    corpus = sweep.SweepCorpus(_sweep_files(specs), stub.CharTokenizer())
    sw = sweep.AnchorSweep(corpus, index_duration=320.0 / 16000, samples_to_frames_ratio=320.0,
                           threshold=KW["threshold"], short_utterance_len=KW["short_utterance_len"],
                           max_window_size=KW["max_window_size"])
    status = sw.run(steps_per_poll=4)
    got = sw.file_rows()
    st = sw.stats()
    assert st["windows"] > 0 and st["cells"] > 0
    for f, s in enumerate(specs):
        if status[f] == sweep.NEEDS_RECALC:
            # the host-policy hand-off: rows up to the hand-off point must still agree
            assert got[f] == ref[f][:len(got[f])], s["name"]
        else:
            assert status[f] == sweep.DONE, (s["name"], status[f])
            assert got[f] == ref[f], s["name"]
    assert (status == sweep.DONE).sum() >= 3


def test_sweep_is_independent_of_batching_and_capacity(exact_audio, tmp_path):
    """One file alone == the same file inside a corpus; a tiny initial capacity only costs
    CAPACITY round trips."""
    specs = [_corpus_file("a", 12, 21), _corpus_file("b", 16, 22, corrupt=((200, 300),))]
    files = _sweep_files(specs)
    tok = stub.CharTokenizer()
    kw = dict(index_duration=0.02, samples_to_frames_ratio=320.0)
    both = sweep.AnchorSweep(sweep.SweepCorpus(files, tok), groups=2, **kw)
    both.run()
    rows_both = both.file_rows()
    for i in range(2):
        solo = sweep.AnchorSweep(sweep.SweepCorpus([files[i]], tok), capacity=(64, 16, 1), **kw)
        status = solo.run(steps_per_poll=1)
        assert solo.file_rows()[0] == rows_both[i]
        assert status[0] == both.state["status"].cpu().numpy()[i]


def test_sweep_equals_cpu_oracle_on_a_synthetic_corpus():
    """A small configs[4]-style corpus (files of different lengths, non-speech rows, utterances
    whose audio says something else): rows, final status and work counters of every file equal
    the CPU restatement oracle/sweep.py, which aligns once per shrinking-transcript iteration."""
    import sweep_corpus
    from oracle import sweep as osweep
    tok = stub.CharTokenizer()
    specs = [sweep_corpus.make_spec(f"f{i}", m, 100 + i, corrupt_frac=c, non_speech_every=ns)
             for i, (m, c, ns) in enumerate([(1.0, 0.1, 0), (4.0, 0.2, 3), (2.5, 0.0, 0), (6.0, 0.12, 5),
                                             (0.6, 0.3, 0), (3.0, 0.5, 2)])]
    lps = [sweep_corpus.emissions(s, "cuda", seed=7 + i) for i, s in enumerate(specs)]
    files = [sweep.SweepFile(s.file_id, s.audio_path, lp, s.n_samples, s.rows) for s, lp in zip(specs, lps)]
    sw = sweep.AnchorSweep(sweep.SweepCorpus(files, tok), index_duration=0.02, samples_to_frames_ratio=320.0,
                           groups=3)
    status = sw.run(steps_per_poll=8)
    got = sw.file_rows()
    windows = sw.state["n_windows"].cpu().numpy()
    cells = sw.state["cells"].cpu().numpy()
    n_rows = 0
    for f, (s, lp) in enumerate(zip(specs, lps)):
        ref, ref_status, stats = osweep.sweep_file(s.file_id, s.audio_path, lp.cpu().numpy(), s.n_samples,
                                                   s.rows, tok)
        assert sweep.STATUS_NAMES[status[f]] == ref_status, s.file_id
        assert got[f] == ref, s.file_id
        assert windows[f] == stats["windows"] and cells[f] == stats["cells"], s.file_id
        n_rows += len(ref)
    assert n_rows > 100


def test_entry_point_with_resident_emissions(tmp_path):
    """The drop-in CLI in --resident_emissions mode: every file encoded once, all anchor loops on
    the device, same per-file TSV layout as the reference (:410, :469-473)."""
    import os
    import subprocess
    import sys
    rng = np.random.default_rng(0)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rows, vads = [], []
    for name, secs in (("a", 40), ("b", 25)):
        wav = str(tmp_path / f"{name}.wav")
        hg.write_wav(wav, rng.normal(0, 0.1, secs * 16000))
        utts = _utterances(rng, 5)
        for i, u in enumerate(utts):
            rows.append([f"{name}_{i}", wav, 1, secs / 5, 0.0, float(secs), u, "spk", "db"])
        vads.append([wav, 0.0, float(secs), float(secs)])
    tsv, vad = str(tmp_path / "in.tsv"), str(tmp_path / "vad.tsv")
    pd.DataFrame(rows, columns=['Sample_ID', 'Sample_Path', 'Channel', 'Audio_Length', 'Start', 'End',
                                'Transcription', 'Speaker_ID', 'Database']).to_csv(tsv, sep='\t', index=None)
    pd.DataFrame(vads, columns=['Sample_Path', 'Start', 'End', 'Segment_Length']).to_csv(vad, sep='\t', index=None)
    dst = str(tmp_path / "out")
    cmd = [sys.executable, os.path.join(root, "src", "iterative_utterance_alignment.py"), "--tsv", tsv,
           "--vad_segments_tsv", vad, "--dst", dst, "--logs_path", str(tmp_path), "--asr_hub", "stub",
           "--max_words_sequence", "8", "--resident_emissions"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    for name in ("a", "b"):
        out = pd.read_csv(os.path.join(dst, f"{name}.tsv"), sep='\t')
        assert list(out.columns) == anchor.RESULT_COLUMNS
        assert len(out.index) >= 1  # the last row of a file is always kept (:263)
        assert (out['End'] >= out['Start']).all()


@pytest.mark.parametrize("case", ["needs_recalc", "exceptions_limit", "window_to_stop", "low_text"])
def test_sweep_terminal_states_equal_the_oracle(case):
    """The branches around the alignment call (:119-125, :167-192, :390-402): hand-off to host policy,
    consecutive text-longer-than-audio errors, the window_to_stop break, 'read more audio and text'."""
    import copy
    import sweep_corpus
    from oracle import sweep as osweep
    tok = stub.CharTokenizer()
    spec = sweep_corpus.make_spec("t", 3.0, 77, corrupt_frac=0.0)
    rows = copy.deepcopy(spec.rows)
    kw = {}
    if case == "needs_recalc":
        kw = dict(max_window_size=12.0)          # the second or third clip reaches 12 s
    elif case == "exceptions_limit":
        for i, r in enumerate(rows):             # half-second clips for ~150 characters of text
            r["Start"], r["End"] = 0.5 * i, 0.5 * (i + 1)
        kw = dict(max_text_to_audio_prop_exec=3)
    elif case == "window_to_stop":
        kw = dict(max_window_size=1000.0, window_to_stop=15.0)
    elif case == "low_text":
        kw = dict(min_text_to_audio_prop=4.5)    # most clips have "too little text": merged with the next row
    lp = sweep_corpus.emissions(spec, "cuda", seed=3)
    f = sweep.SweepFile(spec.file_id, spec.audio_path, lp, spec.n_samples, rows)
    sw = sweep.AnchorSweep(sweep.SweepCorpus([f], tok), index_duration=0.02, samples_to_frames_ratio=320.0, **kw)
    status = sw.run(steps_per_poll=2)
    ref_rows, ref_status, stats = osweep.sweep_file(spec.file_id, spec.audio_path, lp.cpu().numpy(), spec.n_samples,
                                                    rows, tok, **kw)
    assert sweep.STATUS_NAMES[status[0]] == ref_status
    if case != "low_text":
        assert ref_status == case
    assert sw.file_rows()[0] == ref_rows
    assert int(sw.state["n_windows"][0]) == stats["windows"]


def test_sweep_resumes_after_host_recalc_and_trims_text_before_silence():
    """:119-146 + :185-192: a speech segment that ends (next row is Non-Speech) with far more text than
    audio first goes to the host's time re-spreading (here: the identity), then aligns only as many
    utterances as fit (find_a_valid_text_to_audio_proportion)."""
    import copy
    import sweep_corpus
    from oracle import sweep as osweep
    tok = stub.CharTokenizer()
    spec = sweep_corpus.make_spec("t", 2.0, 91, corrupt_frac=0.0)
    rows = copy.deepcopy(spec.rows)
    first = rows[0]
    assert len(first["utterances"]) >= 2
    n_chars = len(" ".join(first["utterances"]))
    clip = n_chars * 0.019                                # text_to_audio_proportion = 0.24 / 0.019 = 12.6
    assert clip > 5.0 and int(clip * 50) <= n_chars       # more characters than frames: text must be trimmed
    first["Start"], first["End"] = 0.0, clip
    silence = {"Type": "Non-Speech", "Start": clip, "End": clip + 1.0, "utterances": []}
    rows = [first, silence] + rows[1:]
    rows[2]["Start"] = clip + 1.0
    lp = sweep_corpus.emissions(spec, "cuda", seed=4)
    f = sweep.SweepFile(spec.file_id, spec.audio_path, lp, spec.n_samples, rows)
    sw = sweep.AnchorSweep(sweep.SweepCorpus([f], tok), index_duration=0.02, samples_to_frames_ratio=320.0)
    calls = []

    def recalc(s, fi):
        calls.append(int(s.state["row"][fi]))
        s.resume_after_recalc(fi)
        return True

    status = sw.run(steps_per_poll=1, recalc_fn=recalc)
    ref_rows, ref_status, stats = osweep.sweep_file(spec.file_id, spec.audio_path, lp.cpu().numpy(), spec.n_samples,
                                                    rows, tok, recalc_fn=lambda i, r, c: r)
    assert calls and calls[0] == 0
    assert sweep.STATUS_NAMES[status[0]] == ref_status == "done"
    assert sw.file_rows()[0] == ref_rows and len(ref_rows) > 3
    assert int(sw.state["n_windows"][0]) == stats["windows"]


def test_sweep_with_host_time_respreading_equals_the_per_file_loop(exact_audio, tmp_path):
    """max_window_size small enough that the reference's `fix_text_to_time_proportion` (:119-146)
    fires several times per file: the sweep hands the file to the host callback, takes the new
    row times and resumes -- same rows as the per-file loop driven by the CPU oracle."""
    specs = [_corpus_file("a", 14, 31), _corpus_file("b", 20, 32, corrupt=((300, 420),)),
             _corpus_file("c", 9, 33)]
    for s in specs:
        exact_audio[s["wav"]] = s["total"]
    kw = dict(KW, max_window_size=5.5)
    ref = []
    for s in specs:
        aligner = cs.CTCSegmentation(s["asr"], kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
        aligner.samples_to_frames_ratio = 320.0
        ref.append(anchor.get_file_iterative_segmentation(s["asr"], aligner, s["wav"], s["df"].copy(),
                                                          s["vad"].copy(), 320.0, str(tmp_path),
                                                          window_fn=oracle_window_fn, **kw))
    frames, vads, lengths, files = [], [], [], []
    for s in specs:
        fixed = hg.fix_time_reference(s["df"], s["vad"], s["total"] / 16000, len(s["df"].index))
        lpz = torch.log_softmax(s["asr"].logits[: s["total"] // 320].float(), dim=-1).cuda()
        files.append(sweep.SweepFile(s["name"], s["wav"], lpz, s["total"],
                                     sweep.rows_from_dataframe(fixed, kw["max_words_sequence"])))
        frames.append(fixed)
        vads.append(s["vad"])
        lengths.append(s["total"] / 16000)
    sw = sweep.AnchorSweep(sweep.SweepCorpus(files, stub.CharTokenizer()), index_duration=0.02,
                           samples_to_frames_ratio=320.0, max_window_size=kw["max_window_size"])
    n_calls = []
    inner = sweep.dataframe_recalc(frames, vads, lengths)

    def counted(s_, f_):
        n_calls.append(f_)
        return inner(s_, f_)

    status = sw.run(steps_per_poll=2, recalc_fn=counted)
    got = sw.file_rows()
    assert len(n_calls) >= 2                       # the host policy really ran
    for f, s in enumerate(specs):
        assert status[f] == sweep.DONE, (s["name"], sweep.STATUS_NAMES[status[f]])
        assert got[f] == ref[f], s["name"]


# ---------------------------------------------------------------------------------------------------
# Instances of the file-resident kernel that the corpora above do not reach on their own
def _synthetic_files(n_files=3, minutes=1.0, pad_v=0, seed=40):
    import sweep_corpus
    specs = [sweep_corpus.make_spec(f"r{i}", minutes, seed + i, corrupt_frac=0.1, non_speech_every=5)
             for i in range(n_files)]
    lps = [sweep_corpus.emissions(sp, "cuda", seed=seed + 9 * i) for i, sp in enumerate(specs)]
    if pad_v:  # extra vocabulary entries nobody can prefer
        lps = [torch.cat([lp, torch.full((lp.shape[0], pad_v), -40.0, device=lp.device)], dim=1).contiguous()
               for lp in lps]
    return [sweep.SweepFile(sp.file_id, sp.audio_path, lp, sp.n_samples, sp.rows) for sp, lp in zip(specs, lps)]


def _run(files, mode, **kw):
    run = sweep.AnchorSweep(sweep.SweepCorpus(files, stub.CharTokenizer()), index_duration=0.02,
                            samples_to_frames_ratio=320.0, mode=mode, **kw)
    status = run.run()
    stats = {k: v for k, v in run.stats().items() if k != "steps"}
    return run, (status.tolist(), run.file_rows(), stats)


@pytest.mark.parametrize("kc", [1, 2, 4])
def test_resident_columns_per_thread_instances(kc, sweep_mode):
    """Every columns-per-thread instance of the fill / walk (chosen per window by default) on the same
    windows: identical rows, status words and counters."""
    if sweep_mode != "resident":
        pytest.skip("resident kernel only")
    ipfa = importlib.import_module(PKG)
    files = _synthetic_files()
    _, want = _run(files, "lockstep")
    with ipfa.tuning(IPFA_SWEEP_KC=str(kc)):
        run, got = _run(files, "resident")
    assert run.mode == "resident" and got == want and sum(len(r) for r in got[1]) > 20


def test_resident_runtime_pitch_and_fallback(sweep_mode):
    """V = 36: dense rows of another width than 32 (the run-time pitch instance); V = 34: rows that 16-byte
    pieces cannot move -> `auto` runs the lock-step path, `resident` refuses."""
    if sweep_mode != "resident":
        pytest.skip("resident kernel only")
    files36 = _synthetic_files(pad_v=4)
    run, got = _run(files36, "auto")
    assert run.mode == "resident" and run.corpus.V == 36
    assert got == _run(files36, "lockstep")[1]
    files34 = _synthetic_files(pad_v=2)
    run34, got34 = _run(files34, "auto")
    assert run34.mode == "lockstep" and got34[1] == got[1]  # the padding changes no alignment
    with pytest.raises(ValueError):
        _run(files34, "resident")


def test_resident_capacity_growth_and_tiny_launch(sweep_mode):
    """A launch capacity far too small: the files stop with CAPACITY, the host grows it and relaunches; the
    state carries over (same rows as one launch with room)."""
    if sweep_mode != "resident":
        pytest.skip("resident kernel only")
    files = _synthetic_files(n_files=2, minutes=0.7)
    _, want = _run(files, "resident")
    run, got = _run(files, "resident", capacity=(64, 16, 1))
    assert got == want and run.capacity[1] > 16


def test_resident_sweep_overlapping_the_upload(sweep_mode):
    """Emissions arriving file by file from pinned host memory while the kernel is already running
    (``ipfa_sweep_corpus.file_ready``): same rows as with everything resident, twice in a row (the ready
    words are re-armed), and the lock-step path simply waits for the upload."""
    files = _synthetic_files(n_files=6, minutes=0.6)
    _, want = _run(files, sweep_mode)
    corpus = sweep.SweepCorpus(files, stub.CharTokenizer())
    host_lp = corpus.lp.cpu().pin_memory()
    run = sweep.AnchorSweep(corpus, index_duration=0.02, samples_to_frames_ratio=320.0, mode=sweep_mode)
    for _ in range(2):
        corpus.lp.fill_(float("nan"))           # whatever is aligned must have come through the upload
        run.reset()
        corpus.begin_upload(host_lp)
        status = run.run()
        torch.cuda.synchronize()
        stats = {k: v for k, v in run.stats().items() if k != "steps"}
        assert (status.tolist(), run.file_rows(), stats) == want


@pytest.mark.parametrize("pad_v", [0, 4])
def test_resident_two_files_per_sm(pad_v, sweep_mode):
    """Corpora of many short files run two files per SM (the 64-register instances of the kernel, fixed and
    run-time panel pitch); forced here on a small corpus: identical rows, status words and counters."""
    if sweep_mode != "resident":
        pytest.skip("resident kernel only")
    ipfa = importlib.import_module(PKG)
    files = _synthetic_files(n_files=5, minutes=0.8, pad_v=pad_v)
    _, want = _run(files, "lockstep")
    with ipfa.tuning(IPFA_SWEEP_CTAS="2"):
        run, got = _run(files, "resident")
    assert run.mode == "resident" and got == want
