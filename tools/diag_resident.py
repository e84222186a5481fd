"""Diagnostic: run one ref_cases anchor case through the resident-emissions sweep on the GPU and print
the first rows that differ from tests/golden/ref (usage: python tools/diag_resident.py mushy)."""
import importlib
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pandas as pd  # noqa: E402

import ref_cases  # noqa: E402

PKG = "iterative-pseudo-forced-alignment-ctc_b200"
hg = importlib.import_module(PKG + ".hostglue")
anchor = importlib.import_module(PKG + ".anchor")
sweep = importlib.import_module(PKG + ".sweep")
cs = importlib.import_module(PKG + ".ctc_segmentation")

for name in sys.argv[1:]:
    case = ref_cases.anchor_cases()[name]
    root = tempfile.mkdtemp()
    case.materialise(root)
    os.chdir(root)
    asr = case.asr("cuda")
    aligner = cs.CTCSegmentation(asr, kaldi_style_text=False, time_stamps="fixed", scoring_length=30)
    ratio = aligner.estimate_samples_to_frames_ratio()
    df = pd.read_csv(case.tsv_rel, sep="\t")
    vad = pd.read_csv(case.vad_rel, sep="\t")
    kw = {k: v for k, v in case.loop.items() if k != "min_words_sequence"}
    rows, status = sweep.align_files_resident(asr, aligner, [(case.wav_rel, df, vad)], ratio, **kw)
    out = hg.remove_artefacts(pd.DataFrame(rows[0], columns=anchor.RESULT_COLUMNS), 30)
    out.to_csv("got.tsv", sep="\t", index=None)
    g = open("got.tsv").read().split("\n")
    r = open(os.path.join(ROOT, "tests/golden/ref/results", name + ".tsv")).read().split("\n")
    print(name, "status", status, "rows", len(g), len(r))
    n = 0
    for i, (a, b) in enumerate(zip(g, r)):
        if a != b:
            print(i, "\n got", a, "\n ref", b)
            n += 1
            if n >= 3:
                break
