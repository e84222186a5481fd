"""The product's entry points (CUDA path) against what the reference's own code wrote.

``tests/golden/ref/`` is the output of /root/reference's UNMODIFIED entry points on the cases of
``tests/ref_cases.py`` (see ``tests/test_reference_loops.py``, which regenerates it wherever the
reference tree exists).  Here the repo's drop-in entry points -- ``src/iterative_utterance_alignment.py``,
``src/search_words.py`` + ``src/word_level_alignment.py``, ``src/search_on_speech.py`` -- run the same
cases on the GPU, through the same CLI the reference's shell drivers use, and every file they write
must be byte-identical: ids, times, scores, texts, column order, float formatting.
"""
import json
import os
import runpy
import sys

import pytest

import ref_cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref")
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))
CASES = list(ref_cases.anchor_cases())


def _read(path):
    with open(path, "rb") as f:
        return f.read()


def _run_entry(script, argv, monkeypatch):
    """``python src/<script> <argv>`` inside this process (so the library the driver sees loaded is
    the one under test)."""
    monkeypatch.setattr(sys, "argv", [script] + [str(a) for a in argv])
    src = os.path.join(ROOT, "src")
    monkeypatch.syspath_prepend(src)
    for m in ("_common",):
        sys.modules.pop(m, None)
    runpy.run_path(os.path.join(src, script), run_name="__main__")


def _loop_argv(case):
    argv = []
    for k, v in case.loop.items():
        if v is not None:
            argv += ["--" + k, v]
    return argv


@pytest.mark.parametrize("resident", [False, True], ids=["per_file_loop", "resident_emissions"])
@pytest.mark.parametrize("name", CASES)
def test_utterance_alignment_writes_the_reference_tsv(name, resident, tmp_path, monkeypatch):
    case = ref_cases.anchor_cases()[name].materialise(str(tmp_path))
    assert case.asr().digest() == MANIFEST["anchor"][name]["emissions_sha256"]
    monkeypatch.chdir(tmp_path)
    os.makedirs("results")
    argv = ["--tsv", case.tsv_rel, "--vad_segments_tsv", case.vad_rel, "--dst", "results", "--logs_path", "logs",
            "--asr_hub", "py:ref_cases:asr_factory", "--asr_savedir", name] + _loop_argv(case)
    if resident:
        argv.append("--resident_emissions")
    _run_entry("iterative_utterance_alignment.py", argv, monkeypatch)
    assert _read(os.path.join("results", name + ".tsv")) == _read(os.path.join(GOLDEN, "results", name + ".tsv"))


def test_word_level_alignment_and_search_on_speech_write_the_reference_tsvs(tmp_path, monkeypatch):
    root = str(tmp_path)
    base = ref_cases.anchor_cases()["clean"].materialise(root)
    wc = ref_cases.WordsCase(base).materialise(root)
    monkeypatch.chdir(tmp_path)
    os.makedirs("logs")
    _run_entry("search_words.py", ["--tsv_path", wc.tsv_rel, "--dst", "words", "--config_file", wc.config_rel,
                                   "--text_column", "Transcription"], monkeypatch)
    common = ["--asr_hub", "py:ref_cases:asr_factory", "--asr_savedir", "clean", "--dst_path", "words",
              "--logs_path", "logs", "--left_offset", "-0.02", "--right_offset", "0.03"]
    # align_words.sh:96 of the reference spells the flag "--tsv" (argparse prefix match)
    _run_entry("word_level_alignment.py", ["--tsv", "words/utterances_filtered.tsv", "--use_time_info"] + common,
               monkeypatch)
    _run_entry("search_on_speech.py", ["--tsv", wc.tsv_rel, "--text=" + wc.search_text] + common, monkeypatch)
    for f in ("utterances_filtered.tsv", "utterances_words.tsv", "utterances_sos.tsv"):
        assert _read(os.path.join("words", f)) == _read(os.path.join(GOLDEN, "words", f)), f
    assert MANIFEST["words"]["utterances_words.tsv"] >= 8
