"""Writes tests/golden/ref/: what the reference's OWN, UNMODIFIED host code produces on the
cases of tests/ref_cases.py.

    python tests/golden/make_ref_golden.py          (needs /root/reference; CPU only)

For every anchor case the reference's ``iterative_utterance_alignment.main`` runs end to end
(/root/reference/src/iterative_utterance_alignment.py:407-478 -> :14-404, with
utils/alignment_utils.py under it) through ``tests/ref_shim.py``; the aligner it is handed in
place of speechbrain's is the CPU oracle (``oracle/sb_aligner.py``).  Then the reference's
post-steps run as subprocesses on those results (postprocess/merge_aligned_files.py,
scripts/tsv_to_stm.py, postprocess/postprocess_and_filter.py), and search_words.py,
word_level_alignment.py and search_on_speech.py run on the word-level case.

``manifest.json`` records, per case, the sha256 of the emissions (so a consumer can tell a
non-reproducible input from a wrong result), the loop parameters and how often each branch of
the reference loop logged its message.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (ROOT, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_cases  # noqa: E402
import ref_shim  # noqa: E402


def run_reference(root, aligner_cls=None, device="cpu"):
    """Run every case under ``root`` with the reference's code; returns the manifest dict.
    Files land in ``root/results`` (per-file TSVs, merged TSV, stm/, filtered TSV) and
    ``root/words`` (filtered / words / sos TSVs)."""
    if aligner_cls is None:
        from oracle.sb_aligner import CTCSegmentation as aligner_cls
    manifest = {"anchor": {}, "words": {}}
    cwd = os.getcwd()
    os.chdir(root)
    try:
        os.makedirs("results", exist_ok=True)
        os.makedirs("logs", exist_ok=True)
        cases = ref_cases.anchor_cases()
        all_df = []
        for name, case in cases.items():
            case.materialise(root)
            all_df.append(case.df)
            asr = case.asr(device)
            with ref_shim.reference_modules(aligner_cls, asr_model=asr) as ns:
                args = argparse.Namespace(tsv=case.tsv_rel, vad_segments_tsv=case.vad_rel, dst="results",
                                          logs_path="logs", asr_hub="synthetic", asr_savedir="unused", **case.loop)
                ns.iua.main(args)
                refused = ns.stdout.getvalue().count("Start frame:")  # the print of :158
            ref_shim.close_logger(name)
            log = open(os.path.join("logs", name + ".log")).read()
            n_rows = sum(1 for _ in open(os.path.join("results", name + ".tsv"))) - 1
            manifest["anchor"][name] = {"emissions_sha256": asr.digest(), "loop": case.loop, "rows": n_rows,
                                        "branches": dict(ref_cases.count_marks(log), load_refused=refused)}
        import pandas as pd
        pd.concat(all_df, ignore_index=True).to_csv("tsv/all.tsv", sep="\t", index=None)
        # post-steps, the reference's own scripts as __main__
        ref_shim.run_script("postprocess/merge_aligned_files.py", ["--global_tsv", "tsv/all.tsv", "--src", "results"],
                            cwd=root)
        os.makedirs("results/stm", exist_ok=True)
        ref_shim.run_script("scripts/tsv_to_stm.py", ["--src_path", "results", "--dst_path", "results/stm"], cwd=root)
        ref_shim.run_script("postprocess/postprocess_and_filter.py",
                            ["--tsv", "results/all_aligned.tsv", "--score", "-1.0", "--comp", "gt", "--collar", "0.2",
                             "--left_offset", "-0.05", "--right_offset", "0.05"], cwd=root)
        # word level / search on speech on clips of the clean file
        wc = ref_cases.WordsCase(cases["clean"]).materialise(root)
        asr = cases["clean"].asr(device)
        with ref_shim.reference_modules(aligner_cls, asr_model=asr) as ns:
            ns.search_words.main(argparse.Namespace(tsv_path=wc.tsv_rel, dst="words", config_file=wc.config_rel,
                                                    text_column="Transcription"))
            common = dict(asr_hub="synthetic", asr_savedir="unused", dst_path="words", offset_time=0.0,
                          left_offset=-0.02, right_offset=0.03, collar=0.0, logs_path="logs")
            ns.wla.main(argparse.Namespace(time_info=True, tsv_path="words/utterances_filtered.tsv", **common))
            ref_shim.close_logger("utterances_filtered")
            ns.sos.main(argparse.Namespace(tsv_path=wc.tsv_rel, text=wc.search_text, **common))
            ref_shim.close_logger("utterances")
        for f in ("utterances_filtered.tsv", "utterances_words.tsv", "utterances_sos.tsv"):
            manifest["words"][f] = sum(1 for _ in open(os.path.join("words", f))) - 1
    finally:
        os.chdir(cwd)
    return manifest


GOLDEN_FILES = (["results/%s.tsv" % n for n in ref_cases.anchor_cases()]
                + ["results/all_aligned.tsv", "results/all_aligned_gt_-1.0_filtered.tsv"]
                + ["words/utterances_filtered.tsv", "words/utterances_words.tsv", "words/utterances_sos.tsv"])


def golden_files(root):
    out = list(GOLDEN_FILES)
    stm = os.path.join(root, "results", "stm")
    if os.path.isdir(stm):
        out += ["results/stm/" + f for f in sorted(os.listdir(stm))]
    return out


def main():
    assert ref_shim.available(), "the reference tree is needed to (re)generate the fixtures"
    dst = os.path.join(HERE, "ref")
    with tempfile.TemporaryDirectory() as root:
        manifest = run_reference(root)
        shutil.rmtree(dst, ignore_errors=True)
        for rel in golden_files(root):
            os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
            shutil.copy(os.path.join(root, rel), os.path.join(dst, rel))
        with open(os.path.join(dst, "manifest.json"), "w") as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
