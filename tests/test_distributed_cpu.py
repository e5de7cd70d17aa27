"""world_size-2 `gloo` test of the multi-rank exchange logic (robopoker_b200/distributed.py) on CPU.

The compute object is the oracle (tests may use it); what is under test is the host orchestration: tree-id sharding
by rank, the all-gather of blocked partials, the rank-ordered fold — ranks must end bit-identical to each other and
to a single-process emulation of the same two shards."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, epochs, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import binding as oracle
    from robopoker_b200.distributed import ShardedSolver

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s = oracle.OracleSolver("leduc", "LinearRegret", "LinearWeight", "ExternalSampling", batch=batch, seed=21)
    ShardedSolver(s, dist).step(epochs)
    np.save(os.path.join(out_dir, f"rows{rank}.npy"), s.profile_rows())
    np.save(os.path.join(out_dir, f"counters{rank}.npy"), np.array(list(s.counters().values()), dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_rank_gloo_matches_single_process_emulation(tmp_path, oracle, world):
    import torch.multiprocessing as mp

    batch, epochs = (96, 25) if world < 8 else (32, 12)  # world 8 = the driver's largest scaling point
    mp.spawn(_worker, args=(world, _free_port(), batch, epochs, str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / f"rows{r}.npy") for r in range(world)]
    assert all(rows[0].tobytes() == r.tobytes() for r in rows[1:])
    # single-process emulation: one handle per rank, concatenated partials, same fold
    r = [oracle.OracleSolver("leduc", "LinearRegret", "LinearWeight", "ExternalSampling", batch=batch, seed=21) for _ in range(world)]
    for k, s in enumerate(r):
        s.set_fold(1, k, world)
    for _ in range(epochs):
        words = np.concatenate([s.sample() for s in r])
        for s in r:
            s.fold_gathered(words, world)
    assert r[0].profile_rows().tobytes() == rows[0].tobytes()
    # shards are disjoint and complete: per-rank counters add up to the whole epoch's work
    c = [np.load(tmp_path / f"counters{k}.npy") for k in range(world)]
    assert sum(int(x[0]) for x in c) == sum(s.counters()["nodes"] for s in r)
    assert rows[0]["visits"].sum() > 0


def test_batched_fold_equals_ordered_at_batch_one(oracle):
    a = oracle.OracleSolver("leduc", batch=1, seed=3).step(3000)
    b = oracle.OracleSolver("leduc", batch=1, seed=3)
    b.set_fold(1)
    b.step(3000)
    assert a.profile_rows().tobytes() == b.profile_rows().tobytes()


def test_batched_fold_converges(oracle):
    s = oracle.OracleSolver("leduc", batch=256, seed=1)
    s.set_fold(1)
    s.step(512)
    assert s.exploitability() < 0.080  # the reference's Leduc threshold (crates/leduc/src/solver.rs:121-123)


def _kmeans_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from lloyd_data import turn_histograms

    from oracle import binding as oracle
    from robopoker_b200.distributed import sharded_kmeans_step

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pts = turn_histograms(3000, seed=5)
    full = oracle.OracleKmeans(pts, 12, threads=2)
    chosen = full.init_centroids(4)                      # seeding on the full set (replicated)
    lo, hi = 3000 * rank // world, 3000 * (rank + 1) // world
    shard = oracle.OracleKmeans(pts[lo:hi], 12, threads=2)
    shard.set_centroids_from_counts(pts[chosen].astype(np.uint64))
    shard.init_bounds()
    drifts = []
    for _ in range(5):
        d, sizes, re = sharded_kmeans_step(shard, dist)
        drifts.append(d)
    np.save(os.path.join(out_dir, f"kassign{rank}.npy"), shard.bounds()[0])
    np.save(os.path.join(out_dir, f"kcent{rank}.npy"), shard.future()[0])
    np.save(os.path.join(out_dir, f"kdrift{rank}.npy"), np.stack(drifts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_kmeans_allreduce_matches_single_process(tmp_path, oracle):
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from lloyd_data import turn_histograms

    world = 2
    mp.spawn(_kmeans_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    pts = turn_histograms(3000, seed=5)
    one = oracle.OracleKmeans(pts, 12, threads=4)
    one.init_centroids(4)
    one.init_bounds()
    drifts = np.stack([one.step()[0] for _ in range(5)])
    assign = np.concatenate([np.load(tmp_path / f"kassign{r}.npy") for r in range(world)])
    assert np.array_equal(assign, one.bounds()[0])                                   # sharding changes nothing
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"kcent{r}.npy"), one.future()[0])  # centroids identical on every rank
        assert np.array_equal(np.load(tmp_path / f"kdrift{r}.npy").view(np.uint32), drifts.view(np.uint32))


def _nlhe_worker(rank, world, port, batch, epochs, out_dir, mode="replicated"):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import binding as oracle
    from robopoker_b200.distributed import ShardedNlhe

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s = oracle.OracleNlhe(seed=13, batch=batch)
    ShardedNlhe(s, dist, mode=mode).step(epochs)
    np.save(os.path.join(out_dir, f"nlhe{rank}.npy"), s.export())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_nlhe_records_exchange(tmp_path, oracle):
    """Ragged all-gather of update records + ordered fold: both ranks end identical to ONE process running 2*batch trees."""
    import torch.multiprocessing as mp

    world, batch, epochs = 2, 24, 3
    mp.spawn(_nlhe_worker, args=(world, _free_port(), batch, epochs, str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / f"nlhe{r}.npy") for r in range(world)]
    assert rows[0].tobytes() == rows[1].tobytes()
    whole = oracle.OracleNlhe(seed=13, batch=world * batch)
    whole.step(epochs)
    assert whole.export().tobytes() == rows[0].tobytes() and len(rows[0]) > 1000


@pytest.mark.parametrize("world", [2, 4])
def test_two_rank_gloo_nlhe_owner_sharded_fold(tmp_path, oracle, world):
    """Owner-sharded exchange (all-to-all of records to the infoset's owner, fold, all-gather of the touched rows): every
    replica ends identical to ONE process running world*batch trees."""
    import torch.multiprocessing as mp

    batch, epochs = 24, 3
    mp.spawn(_nlhe_worker, args=(world, _free_port(), batch, epochs, str(tmp_path), "owner"), nprocs=world, join=True)
    rows = [np.load(tmp_path / f"nlhe{r}.npy") for r in range(world)]
    assert all(rows[0].tobytes() == r.tobytes() for r in rows[1:])
    whole = oracle.OracleNlhe(seed=13, batch=world * batch)
    whole.step(epochs)
    assert whole.export().tobytes() == rows[0].tobytes() and len(rows[0]) > 1000
