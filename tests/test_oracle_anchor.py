"""Behavioural tests of the anchor-loop state machine restatement (oracle/anchor.py,
/root/reference/src/iterative_utterance_alignment.py:203-379)."""
from oracle.anchor import anchor_window

LONG = "x" * 40   # >= short_utterance_len: no penalty


def seg(i, start, end, score, text=LONG):
    return [f"u_{i:04}", "u", f"{start:.2f}", f"{end:.2f}", f"{score:3.4f}", text]


def table_align(results):
    """results[k] = list of k segments for the k-utterance prefix."""
    calls = []

    def fn(transcript):
        calls.append(len(transcript))
        return results[len(transcript)]
    fn.calls = calls
    return fn


def test_very_good_first_alignment_is_accepted_at_once():
    res = {2: [seg(0, 0, 1, -0.3), seg(1, 1, 2, -0.5)]}
    fn = table_align(res)
    rows, nss, disc, n = anchor_window([LONG, LONG], fn, 10.0, False, None, [])
    assert n == 1 and len(rows) == 2 and nss == 12.0 and disc == []


def test_bad_then_good_prefix():
    res = {3: [seg(0, 0, 1, -0.3), seg(1, 1, 2, -0.4), seg(2, 2, 3, -5.0)],
           2: [seg(0, 0, 1, -0.3), seg(1, 1, 2.5, -0.4)]}
    fn = table_align(res)
    rows, nss, disc, n = anchor_window(["a" + LONG, "b" + LONG, "c" + LONG], fn, 0.0, False, None, [])
    assert fn.calls == [3, 2]
    assert [r[5] for r in rows] == [1.0, 2.5]
    assert nss == 2.5 and disc == ["c" + LONG]


def test_single_bad_utterance_is_discarded_and_anchor_rewinds():
    res = {1: [seg(0, 0, 1, -7.0)]}
    rows, nss, disc, n = anchor_window([LONG], table_align(res), 4.0, False, 99.0, [])
    assert rows == [] and nss == 4.0 and disc == [LONG]


def test_mediocre_score_iterates_and_keeps_previous_when_not_improved():
    # first alignment ok (>= threshold) but not "very good" (> -1): iterate
    res = {3: [seg(0, 0, 1, -1.2), seg(1, 1, 2, -1.5), seg(2, 2, 3, -1.6)],
           2: [seg(0, 0, 1, -1.2), seg(1, 1, 2, -1.7)]}
    fn = table_align(res)
    rows, nss, disc, n = anchor_window(["a" + LONG, "b" + LONG, "c" + LONG], fn, 0.0, False, None, [])
    # previous[-2] score -1.5 >= current -1.7 -> keep the 3-utterance alignment
    assert fn.calls == [3, 2] and len(rows) == 3 and nss == 3.0 and disc == []


def test_mediocre_improves_until_very_good():
    res = {3: [seg(0, 0, 1, -1.2), seg(1, 1, 2, -1.5), seg(2, 2, 3, -1.6)],
           2: [seg(0, 0, 1, -1.2), seg(1, 1, 2.2, -0.6)]}
    fn = table_align(res)
    rows, nss, disc, n = anchor_window(["a" + LONG, "b" + LONG, "c" + LONG], fn, 0.0, False, None, [])
    assert len(rows) == 2 and nss == 2.2 and disc == ["c" + LONG]


def test_last_segment_keeps_everything():
    res = {2: [seg(0, 0, 1, -9.0), seg(1, 1, 2, -9.0)]}
    rows, nss, disc, n = anchor_window([LONG, LONG], table_align(res), 0.0, True, 7.0, [])
    assert n == 1 and len(rows) == 2 and nss == 7.0


def test_short_utterance_penalty():
    res = {1: [seg(0, 0, 1, -0.1, text="short")]}
    rows, nss, disc, n = anchor_window(["short"], table_align(res), 0.0, False, None, [])
    # -0.1 + 2*(-2.0) = -4.1 < threshold -> bad, single utterance -> discarded
    assert rows == [] and disc == ["short"]
