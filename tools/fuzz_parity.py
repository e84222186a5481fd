"""Randomised parity sweep on the GPU against the CPU oracle, every kernel family:

    python tools/fuzz_parity.py [seconds] [alpha,viterbi,seg,windowed,sweep]     (under gpurun)

``tests/test_gpu_fuzz.py`` runs a bounded, fixed-seed slice of it (``run_fuzz``) inside ``-m gpu``."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ipfa_b200 as ipfa
from cases import ctc_case, seg_case
from oracle import ctc as octc
from oracle import ctcseg as oseg
from test_gpu_ctcseg import _pack

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
KINDS = ["alpha", "viterbi", "seg", "windowed", "sweep"]
WEIGHTS = [0.3, 0.3, 0.2, 0.15, 0.05]


def one_case(kind, rng, counts):
    """One random case of ``kind``; returns (ok, description)."""
    seed = int(rng.integers(1 << 30))
    try:
        if kind in ("alpha", "viterbi"):
            v = int(rng.choice([3, 8, 29, 32, 33, 64, 100, 700]))
            l = int(rng.integers(0, 300)) if rng.random() < 0.8 else int(rng.integers(300, 1200))
            t = int(rng.integers(max(1, l), 2 * l + 60))
            n = int(rng.choice([1, 2, 5, 17, 64, 300])) if l < 200 else int(rng.choice([1, 3, 9]))
            if rng.random() < 0.15 and l <= 48:
                n, t = int(rng.integers(4096, 5000)), int(rng.integers(max(8, l), 64))   # length buckets
            lp, tg, il, tl = ctc_case(seed, n, t, max(l, 1), v, ragged=bool(rng.random() < 0.7),
                                      repeats=bool(rng.random() < 0.5), peaked=bool(rng.random() < 0.5))
            if l == 0:
                tl[:] = 0
            desc = f"{kind} n={n} t={t} l={l} v={v}"
            if kind == "alpha":
                ref = octc.ctc_alpha_nll(lp, tg, il, tl)
                # the window scorer's instances: default choice, a forced linear-domain shape, log-domain alone
                mode = rng.choice(["default", "shape", "log"], p=[0.5, 0.35, 0.15])
                switches = {}
                if mode == "shape":
                    switches["IPFA_ALPHA_LIN_SHAPE"] = str(rng.choice(["1,1", "2,1", "4,1", "8,1", "1,2", "2,2", "4,2"]))
                    desc += " lin_shape=" + switches["IPFA_ALPHA_LIN_SHAPE"]
                elif mode == "log":
                    switches["IPFA_ALPHA_LOG"] = "1"
                    desc += " log"
                if rng.random() < 0.2:   # sharper emissions: some windows go through the redo list
                    sc = float(rng.choice([4.0, 20.0, 60.0]))
                    lp = torch.from_numpy(lp * sc).log_softmax(-1).numpy()
                    ref = octc.ctc_alpha_nll(lp, tg, il, tl)
                    desc += f" sharp x{sc}"
                with ipfa.tuning(**switches):
                    got = ipfa.ctc_alpha_nll(dev(lp), dev(tg), dev(il), dev(tl)).cpu().numpy()
                counts["alpha_redo"] = counts.get("alpha_redo", 0) + ipfa.ctc_alpha_redo_count(n)
                fin = np.isfinite(ref)
                ok = np.array_equal(np.isfinite(got), fin) and np.allclose(got[fin], ref[fin], rtol=1e-4, atol=1e-4)
            else:
                paths, scores, status = octc.ctc_viterbi(lp, tg, il, tl)
                res = ipfa.ctc_forced_align(dev(lp), dev(tg), dev(il), dev(tl))
                gp, gs, gst = res.paths.cpu().numpy(), res.scores.cpu().numpy(), res.status.cpu().numpy()
                ok = np.array_equal(gst & 1, status)
                for i in range(n):
                    if not status[i]:
                        ok &= np.array_equal(gp[i, :il[i]], paths[i, :il[i]]) and np.array_equal(gs[i, :il[i]], scores[i, :il[i]])
        elif kind == "sweep":
            import importlib
            import sweep_corpus
            from ipfa_b200 import sweep as sw
            from oracle import sweep as osweep
            stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
            n_files = int(rng.integers(1, 4))
            specs = [sweep_corpus.make_spec(f"z{i}", float(rng.uniform(0.3, 1.2)), seed + i,
                                            corrupt_frac=float(rng.uniform(0, 0.5)),
                                            non_speech_every=int(rng.choice([0, 2, 4]))) for i in range(n_files)]
            lps = [sweep_corpus.emissions(sp, "cuda", seed=seed + 7 * i, peak=float(rng.uniform(3, 8)))
                   for i, sp in enumerate(specs)]
            kw = dict(max_window_size=float(rng.choice([12.0, 25.0, 70.0])),
                      min_text_to_audio_prop=float(rng.choice([0.8, 3.0])))
            files = [sw.SweepFile(sp.file_id, sp.audio_path, lp_, sp.n_samples, sp.rows) for sp, lp_ in zip(specs, lps)]
            run = sw.AnchorSweep(sw.SweepCorpus(files, stub.CharTokenizer()), index_duration=0.02,
                                 samples_to_frames_ratio=320.0, groups=int(rng.integers(1, 3)),
                                 use_graphs=bool(rng.random() < 0.5),
                                 mode=str(rng.choice(["resident", "lockstep"])), **kw)
            status = run.run(steps_per_poll=int(rng.choice([1, 4, 16])))
            got = run.file_rows()
            desc = f"sweep files={n_files} mode={run.mode} {kw}"
            ok = True
            for f, (sp, lp_) in enumerate(zip(specs, lps)):
                ref_rows, ref_status, _ = osweep.sweep_file(sp.file_id, sp.audio_path, lp_.cpu().numpy(), sp.n_samples,
                                                            sp.rows, stub.CharTokenizer(), **kw)
                ok &= got[f] == ref_rows and sw.STATUS_NAMES[status[f]] == ref_status
        else:
            v = int(rng.choice([8, 32, 40, 300]))
            k_utts = int(rng.integers(1, 9))
            lo = int(rng.integers(1, 30)); hi = lo + int(rng.integers(0, 60))
            t = int(rng.integers(k_utts * hi + k_utts + 8, k_utts * hi * 4 + 200))
            n = int(rng.choice([1, 2, 4]))
            window = None
            if kind == "windowed":
                window = int(rng.integers(max(64, t // 4), t + 50))
            lp, in_len, utts = seg_case(seed, n, t, v, k_utts, lo, hi, peaked=bool(rng.random() < 0.8),
                                        ragged=(kind == "seg"))
            cfg = oseg.CtcSegmentationParameters(index_duration=0.02, score_min_mean_over_L=int(rng.choice([5, 30])))
            gt, ubs, n_cols, n_utts = _pack(cfg, utts)
            desc = f"{kind} n={n} t={t} v={v} k={k_utts} tok={lo}-{hi} window={window}"
            ok = True
            for i in range(n):
                if n_cols[i] > in_len[i]:
                    continue
                k = int(n_utts[i])
                w = window
                while True:
                    res = ipfa.ctcseg_align(dev(lp[i:i + 1]), in_len[i:i + 1], gt[i:i + 1], n_cols[i:i + 1],
                                            ubs[i:i + 1], n_utts[i:i + 1], 0.02, score_len=cfg.score_min_mean_over_L,
                                            flags=2, window=w)
                    if w is None or not int(res.status[0]) & 8:
                        break
                    w *= 2
                if window is not None:
                    cfg.min_window_size = window
                g, ub = oseg.prepare_token_list(cfg, utts[i])
                try:
                    timings, char_probs, state_list = oseg.ctc_segmentation(cfg, lp[i, :in_len[i]], g)
                except IndexError:
                    continue
                segs = oseg.determine_utterance_segments(cfg, ub, char_probs, timings, [""] * k)
                timing = res.timing[0, k - 1, :len(g)].cpu().numpy()
                ok &= np.array_equal(np.where(timing < 0, 0.0, timing * 0.02), timings)
                ok &= np.array_equal(res.char_prob[0, k - 1, :in_len[i]].cpu().numpy().astype(np.float64), char_probs)
                seg = res.seg[0, k - 1, :k].cpu().numpy()
                ok &= all(seg[u, 0] == segs[u][0] and seg[u, 1] == segs[u][1] and
                          np.isclose(seg[u, 2], segs[u][2], rtol=1e-12, atol=0) for u in range(k))
    except Exception as exc:  # noqa: BLE001
        ok, desc = False, f"{kind}: {exc!r}"
    return ok, f"{desc} seed={seed}"


def run_fuzz(seed, budget_s=None, cases_per_kind=None, only=None):
    """Random cases until the time budget or the per-kind case count is used up.
    Returns (number of cases, list of mismatch descriptions, counts per kind)."""
    rng = np.random.default_rng(seed)
    kinds = [k for k in KINDS if not only or k in only]
    counts, bad, n_cases = {}, [], 0
    t_end = time.time() + budget_s if budget_s else None
    while True:
        if t_end and time.time() >= t_end:
            break
        open_kinds = [k for k in kinds if cases_per_kind is None or counts.get(k, 0) < cases_per_kind[k]]
        if not open_kinds:
            break
        w = np.array([WEIGHTS[KINDS.index(k)] for k in open_kinds])
        kind = str(rng.choice(open_kinds, p=w / w.sum()))
        counts[kind] = counts.get(kind, 0) + 1
        ok, desc = one_case(kind, rng, counts)
        n_cases += 1
        if not ok:
            bad.append(desc)
            print("MISMATCH", desc, flush=True)
    return n_cases, bad, counts


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    n, bad, counts = run_fuzz(int(time.time()) % 100000, budget_s=budget, only=only)
    print(f"{n} random cases {counts}, {len(bad)} mismatches")
