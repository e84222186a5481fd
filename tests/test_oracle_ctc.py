"""Pins oracle/ctc_oracle.c against the committed outputs of the installed
torch ctc_loss / torchaudio forced_align (tests/golden/ctc_golden.npz)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES
from oracle import ctc as octc


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_alpha_matches_torch_golden(golden, name):
    lp = golden[f"{name}/lp"]
    nll = octc.ctc_alpha_nll(lp, golden[f"{name}/targets"], golden[f"{name}/in_len"],
                             golden[f"{name}/tgt_len"])
    ref = golden[f"{name}/nll"]
    assert np.array_equal(np.isinf(nll), np.isinf(ref))
    fin = np.isfinite(ref)
    # tolerance from BASELINE.json north_star: loss within 1e-4 relative in fp32
    np.testing.assert_allclose(nll[fin], ref[fin], rtol=1e-4, atol=1e-5)


def test_alpha_empty_target(golden):
    lp = golden["empty/lp"]
    n = lp.shape[0]
    nll = octc.ctc_alpha_nll(lp, np.zeros((n, 0), np.int32), golden["empty/in_len"],
                             np.zeros(n, np.int32))
    np.testing.assert_allclose(nll, golden["empty/nll"], rtol=1e-5)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_viterbi_matches_torchaudio_golden(golden, name):
    lp = golden[f"{name}/lp"]
    paths, scores, status = octc.ctc_viterbi(lp, golden[f"{name}/targets"],
                                             golden[f"{name}/in_len"], golden[f"{name}/tgt_len"])
    assert np.array_equal(status, golden[f"{name}/fa_status"])
    ok = status == 0
    in_len = golden[f"{name}/in_len"]
    for i in np.nonzero(ok)[0]:
        t = in_len[i]
        # bit-exact paths (these inputs are tie-free)
        assert np.array_equal(paths[i, :t], golden[f"{name}/paths"][i, :t]), (name, i)
        assert np.array_equal(scores[i, :t], golden[f"{name}/scores"][i, :t])


def test_viterbi_tie_rules(golden):
    """Exact ties: strict-greater comparisons, ties fall to stay."""
    for i in range(int(golden["ties/count"])):
        lp = golden[f"ties/{i}/lp"][None]
        tg = golden[f"ties/{i}/targets"][None]
        paths, scores, status = octc.ctc_viterbi(lp, tg, [lp.shape[1]], [tg.shape[1]])
        assert status[0] == 0
        assert np.array_equal(paths[0], golden[f"ties/{i}/paths"]), i
        assert np.array_equal(scores[0], golden[f"ties/{i}/scores"])


def test_merge_tokens():
    path = np.array([0, 1, 1, 0, 2, 2, 2, 0, 2])
    sc = np.arange(9, dtype=np.float32)
    spans = octc.merge_tokens(path, sc)
    assert [(s[0], s[1], s[2]) for s in spans] == [(1, 1, 3), (2, 4, 7), (2, 8, 9)]
    assert spans[1][3] == pytest.approx(5.0)
