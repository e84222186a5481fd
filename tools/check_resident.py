"""File-resident sweep against the lock-step sweep on a synthetic corpus (under gpurun):

    python tools/check_resident.py [minutes per file] [files]

prints both run times and whether rows, status words and counters are identical.  A watchdog ends the
process if the persistent kernel does not come back (it has no host round trips to time out on)."""
import importlib, os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sweep_corpus
from ipfa_b200 import sweep as sw
stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 1
specs = [sweep_corpus.make_spec(f"z{i}", minutes, 11 + i) for i in range(n_files)]
lps = [sweep_corpus.emissions(sp, "cuda", seed=3 + i) for i, sp in enumerate(specs)]
files = [sw.SweepFile(sp.file_id, sp.audio_path, lp_, sp.n_samples, sp.rows) for sp, lp_ in zip(specs, lps)]


def watchdog():
    time.sleep(20)
    print("WATCHDOG: the sweep did not come back within 20 s", flush=True)
    os._exit(3)


threading.Thread(target=watchdog, daemon=True).start()
out = {}
for mode in ("resident", "lockstep"):
    run = sw.AnchorSweep(sw.SweepCorpus(files, stub.CharTokenizer()), index_duration=0.02,
                         samples_to_frames_ratio=320.0, mode=mode, groups=min(32, max(1, n_files // 6)))
    run.run()
    torch.cuda.synchronize()
    t0 = time.time()
    run.reset()
    t1 = time.time()
    if mode == "resident" and os.environ.get("PROFILE_HOST"):
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        status = run.run()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
    else:
        status = run.run()
    t2 = time.time()
    torch.cuda.synchronize()
    t_run = time.time() - t0
    print(f"  host: reset {1e3 * (t1 - t0):.2f} ms, run {1e3 * (t2 - t1):.2f} ms", flush=True)
    if mode == "resident":   # the launch alone, CUDA events
        for _ in range(2):
            run.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            run.run_resident()
            e1.record()
            torch.cuda.synchronize()
            print(f"  resident launch alone: {e0.elapsed_time(e1):.3f} ms", flush=True)
    out[mode] = (status.tolist(), run.file_rows(), {k: v for k, v in run.stats().items() if k != "steps"})
    print(f"{mode:9s} {1e3 * t_run:8.2f} ms  capacity {run.capacity}  {run.stats()}", flush=True)
same = out["resident"] == out["lockstep"]
print("identical rows / status / counters:", same)
import ctypes
ctypes.CDLL(None).fflush(None)  # device printf (IPFA_SWEEP_PHASES=1) sits in the C stdio buffer
os._exit(0 if same else 1)
