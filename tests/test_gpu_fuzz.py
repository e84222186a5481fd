"""A bounded, fixed-seed slice of tools/fuzz_parity.py inside ``-m gpu``: random shapes through every
kernel family -- window scorer (default instance, forced linear-domain shapes, log-domain alone,
sharpened emissions through the redo list, length buckets), Viterbi + backtrace, CTC segmentation
(all prefixes), windowed table mode with window doubling, small corpora through the anchor sweep
(with and without CUDA graphs) -- each case against the CPU oracle."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

CASES = {"alpha": 150, "viterbi": 150, "seg": 120, "windowed": 80, "sweep": 12}


@pytest.mark.parametrize("kind", list(CASES))
def test_random_cases_against_the_oracle(kind):
    import fuzz_parity
    n, bad, counts = fuzz_parity.run_fuzz(seed=20261018 + len(kind), budget_s=60.0, cases_per_kind={kind: CASES[kind]},
                                          only=[kind])
    assert not bad, bad
    # the time budget is a guard, not the plan: most of the slice must have run
    assert n >= CASES[kind] // 3, (n, counts)
