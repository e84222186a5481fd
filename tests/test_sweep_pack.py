"""CPU checks of the anchor-sweep corpus packing: a window's ground-truth column, built on the
device as a slice of the file's token stream, must equal what ``prepare_token_list`` builds
for the same utterances."""
import importlib

import numpy as np
import torch

PKG = "iterative-pseudo-forced-alignment-ctc_b200"
sweep = importlib.import_module(PKG + ".sweep")
cs = importlib.import_module(PKG + ".ctc_segmentation")
stub = importlib.import_module(PKG + ".stub_asr")


def _files(rng, n_files=3):
    words = "uno dos tres cuatro cinco seis siete ocho nueve diez".split()
    files = []
    for f in range(n_files):
        rows, t = [], 0.0
        for r in range(int(rng.integers(3, 7))):
            if r == 2 and f == 1:
                rows.append({"Type": "Non-Speech", "Start": t, "End": t + 3.0, "utterances": []})
                t += 3.0
            utts = [" ".join(rng.choice(words, size=int(rng.integers(1, 6)))).upper()
                    for _ in range(int(rng.integers(1, 4)))]
            rows.append({"Type": "Speech", "Start": t, "End": t + 10.0, "utterances": utts, "Channel": 1,
                         "Speaker_ID": "s", "Database": "d"})
            t += 10.0
        files.append(sweep.SweepFile(f"file{f}", f"/x/file{f}.wav", torch.zeros(int(t * 50), 8), int(t * 16000),
                                     rows))
    return files


def test_window_ground_truth_is_a_slice_of_the_file_token_stream():
    rng = np.random.default_rng(3)
    tok = stub.CharTokenizer()
    files = _files(rng)
    host, texts = sweep.pack_files(files, tok, blank=0)
    cfg = cs.CtcSegmentationParameters()
    for f, sf in enumerate(files):
        s0 = host["utt_first"][f]
        n_utt = host["utt_first"][f + 1] - s0 - 1
        assert n_utt == len(texts[f]) == sum(len(r["utterances"]) for r in sf.rows)
        stream = host["tokens"][host["file_tok0"][f]:]
        for _ in range(20):
            a = int(rng.integers(0, n_utt))
            b = int(rng.integers(a + 1, n_utt + 1))
            token_list = [np.array(tok.encode_as_ids(u)) for u in texts[f][a:b]]
            gt, ub = cs.prepare_token_list(cfg, token_list)
            col_a, col_b = host["utt_col"][s0 + a], host["utt_col"][s0 + b]
            mine = np.concatenate([[-1], stream[col_a:col_b + 1]])
            assert np.array_equal(mine, gt.reshape(-1))
            assert [1 + host["utt_col"][s0 + a + k] - col_a for k in range(b - a + 1)] == list(ub)
            assert [host["utt_chars"][s0 + a + k] for k in range(b - a)] == [len(u) for u in texts[f][a:b]]
        # rows: utterance ranges are cumulative, non-speech rows own none
        r0, r1 = host["row_first"][f], host["row_first"][f + 1]
        ends = host["row_utt_end"][r0:r1]
        assert np.all(np.diff(ends) >= 0) and ends[-1] == n_utt
        assert [int(t) for t in host["row_type"][r0:r1]] == [int(r["Type"] == "Non-Speech") for r in sf.rows]
    assert host["file_frame0"].tolist() == np.cumsum([0] + [int(f.lpz.shape[0]) for f in files])[:-1].tolist()


def test_sweep_refuses_cpu_tensors():
    import pytest
    tok = stub.CharTokenizer()
    files = _files(np.random.default_rng(1), 1)
    if not torch.cuda.is_available():
        with pytest.raises(ValueError, match="CUDA"):
            sweep.SweepCorpus(files, tok)


def test_struct_layout_matches_header():
    """ctypes mirrors of the ABI structs: field order / count as declared in include/ipfa_b200.h."""
    import re
    lib_mod = importlib.import_module(PKG + "._lib")
    text = open(lib_mod.HEADER).read()
    for name, cls in (("ipfa_sweep_corpus", sweep._Corpus), ("ipfa_sweep_state", sweep._State),
                      ("ipfa_sweep_params", sweep._Params)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1].lstrip("*")
            fields.append(first)
            fields += [n.strip().lstrip("*") for n in names[1:]]
        assert fields == [f[0] for f in cls._fields_], name
