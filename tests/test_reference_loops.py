"""The reference's own host code as the oracle (CPU).

``tests/golden/ref/`` holds what /root/reference's UNMODIFIED ``iterative_utterance_alignment.main``
(:407-478 over :14-404 and utils/alignment_utils.py), ``word_level_alignment.main``,
``search_on_speech.main``, ``search_words.main`` and the post-step scripts wrote for the cases of
``tests/ref_cases.py`` (generator: ``tests/golden/make_ref_golden.py``; shims: ``tests/ref_shim.py``).

* where the reference tree is present (the build container) the fixtures are REGENERATED and must
  be byte-identical -- the golden files are the reference's output, not a recollection of it;
* everywhere, the repo's restatements are held against them on the CPU: the host mirror of the row
  loop (``anchor.py`` + ``hostglue.py``) driven by the CPU oracle, ``oracle/anchor.py``, and the
  drop-in writers under ``src/``.
The ``-m gpu`` half (the CUDA path writing the same files) is ``tests/test_gpu_reference_loops.py``.
"""
import filecmp
import importlib
import json
import os
import shutil
import subprocess
import sys

import pandas as pd
import pytest

import ref_cases
import ref_shim
from oracle_glue import oracle_window_fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref")
PKG = "iterative-pseudo-forced-alignment-ctc_b200"
hg = importlib.import_module(PKG + ".hostglue")
anchor = importlib.import_module(PKG + ".anchor")
cs = importlib.import_module(PKG + ".ctc_segmentation")

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ref_golden  # noqa: E402

MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))
CASES = list(ref_cases.anchor_cases())


def _read(path):
    with open(path, "rb") as f:
        return f.read()


@pytest.mark.skipif(not ref_shim.available(), reason="no reference tree on this machine")
def test_fixtures_are_what_the_reference_writes(tmp_path):
    """Runs the reference end to end (all cases, post-steps, word level, search on speech)."""
    manifest = make_ref_golden.run_reference(str(tmp_path))
    assert manifest == MANIFEST
    files = make_ref_golden.golden_files(str(tmp_path))
    assert sorted(files) == sorted(os.path.relpath(os.path.join(d, f), GOLDEN)
                                   for d, _, fs in os.walk(GOLDEN) for f in fs if f != "manifest.json")
    for rel in files:
        assert _read(os.path.join(str(tmp_path), rel)) == _read(os.path.join(GOLDEN, rel)), rel


def test_cases_walk_every_branch_of_the_reference_loop():
    total = {}
    for c in MANIFEST["anchor"].values():
        for k, v in c["branches"].items():
            total[k] = total.get(k, 0) + v
    assert set(total) == set(ref_cases.LOG_MARKS) | {"load_refused"} and all(v > 0 for v in total.values()), total


@pytest.mark.parametrize("name", CASES)
def test_emissions_are_reproducible(name):
    case = ref_cases.anchor_cases()[name]
    assert case.asr().digest() == MANIFEST["anchor"][name]["emissions_sha256"]


@pytest.mark.parametrize("name", CASES)
def test_host_mirror_with_cpu_oracle_writes_the_reference_tsv(name, tmp_path, monkeypatch):
    """anchor.get_file_iterative_segmentation (host mirror of :14-404) + hostglue
    (alignment_utils.py) + oracle/anchor.py (:203-385), all on the CPU == the reference's file."""
    case = ref_cases.anchor_cases()[name].materialise(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    asr = case.asr()
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    ratio = aligner.estimate_samples_to_frames_ratio()
    df = pd.read_csv(case.tsv_rel, header=0, sep="\t")
    vad = pd.read_csv(case.vad_rel, header=0, sep="\t")
    rows = anchor.get_file_iterative_segmentation(asr, aligner, case.wav_rel, df, vad, ratio, str(tmp_path / "logs"),
                                                  window_fn=oracle_window_fn, **case.loop)
    out = hg.remove_artefacts(pd.DataFrame(rows, columns=anchor.RESULT_COLUMNS), case.loop["short_utterance_len"])
    out.to_csv("got.tsv", sep="\t", index=None)
    assert _read("got.tsv") == _read(os.path.join(GOLDEN, "results", name + ".tsv"))


def _run_src(script, *argv, cwd):
    return subprocess.run([sys.executable, os.path.join(ROOT, "src", script), *argv], cwd=cwd, check=True,
                          capture_output=True, text=True)


def test_post_step_writers_match_the_reference_scripts(tmp_path):
    """src/merge_aligned_files.py, tsv_to_stm.py, postprocess_and_filter.py on the reference's
    per-file TSVs == what the reference's own scripts wrote from them (merge_aligned_files.py:7-30,
    tsv_to_stm.py:32-39, postprocess_and_filter.py:54-77)."""
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "results", "stm"))
    cases = ref_cases.anchor_cases()
    for name, case in cases.items():
        case.materialise(root)  # the collar step reads the WAV headers
        shutil.copy(os.path.join(GOLDEN, "results", name + ".tsv"), os.path.join(root, "results"))
    pd.concat([c.df for c in cases.values()], ignore_index=True).to_csv(os.path.join(root, "tsv", "all.tsv"), sep="\t",
                                                                        index=None)
    _run_src("postprocess/merge_aligned_files.py", "--global_tsv", "tsv/all.tsv", "--src", "results", cwd=root)
    _run_src("scripts/tsv_to_stm.py", "--src_path", "results", "--dst_path", "results/stm", cwd=root)
    _run_src("postprocess/postprocess_and_filter.py", "--tsv", "results/all_aligned.tsv", "--score", "-1.0", "--comp", "gt",
             "--collar", "0.2", "--left_offset", "-0.05", "--right_offset", "0.05", cwd=root)
    for rel in ["results/all_aligned.tsv", "results/all_aligned_gt_-1.0_filtered.tsv"] + \
            ["results/stm/" + f for f in sorted(os.listdir(os.path.join(GOLDEN, "results", "stm")))]:
        assert _read(os.path.join(root, rel)) == _read(os.path.join(GOLDEN, rel)), rel
    assert sorted(os.listdir(os.path.join(root, "results", "stm"))) == \
        sorted(os.listdir(os.path.join(GOLDEN, "results", "stm")))


def test_search_words_matches_the_reference_script(tmp_path):
    root = str(tmp_path)
    wc = ref_cases.WordsCase(ref_cases.anchor_cases()["clean"]).materialise(root)
    _run_src("search_words.py", "--tsv_path", wc.tsv_rel, "--dst", "words", "--config_file", wc.config_rel, cwd=root)
    assert _read(os.path.join(root, "words", "utterances_filtered.tsv")) == \
        _read(os.path.join(GOLDEN, "words", "utterances_filtered.tsv"))
