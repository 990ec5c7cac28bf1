"""Host-side logic of the multi-GPU paths, exercised with world_size=2 gloo on CPU.

The orchestration (row ranges, feature ownership, tiled reduce-scatter, block sampling) is the
product's distributions_b200.sharding; the per-tile compute is injected -- here the oracle stands in
for the CUDA kernels, which is exactly the role the oracle is allowed (checker, in tests only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from distributions_b200 import sharding, synth  # noqa: E402


def test_row_shard_partition():
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [sharding.row_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_feature_shard_partition():
    for F in (1, 5, 256):
        for world in (1, 2, 4, 8):
            owned = [sharding.feature_shard(F, r, world) for r in range(world)]
            assert sorted(sum(owned, [])) == list(range(F))


def _worker(rank, world, port, n, G, F, tile_rows, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    from oracle.pyoracle import Oracle
    o = Oracle()
    cc = synth.crosscat(901, G, n, n_gp=F // 2, n_bb=F - F // 2)
    feats = cc["features"]
    prior = o.py_prior(synth.PY_ALPHA, synth.PY_D, cc["sizes"])
    mine = sharding.feature_shard(len(feats), rank, world)

    def score_partial(lo, hi, out):
        sub = [dict(feats[f], values=feats[f]["values"][lo:hi]) for f in mine]
        part = cases.oracle_scores(o, sub, n=hi - lo, prior=prior if rank == 0 else None) if sub else \
            np.zeros((hi - lo, G), np.float32)
        out.copy_(torch.from_numpy(part))

    def sample_block(scores, u, out):
        out.copy_(torch.from_numpy(o.sample_rows(scores.numpy().copy(), u.numpy())))

    u = torch.from_numpy(cc["u"])
    assigns, rows = sharding.feature_sharded_score_sample(score_partial, sample_block, n, G, u, torch.device("cpu"),
                                                          tile_rows=tile_rows)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), rows=np.array(rows),
             assign=np.concatenate([a.numpy() for a in assigns]) if assigns else np.zeros(0, np.int32))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,tile_rows", [(1000, 256), (777, 100), (64, 4096)])
def test_feature_sharded_matches_single_process(tmp_path, n, tile_rows):
    import cases
    from oracle.pyoracle import Oracle
    world, G, F = 2, 19, 6
    port = 29500 + (os.getpid() + n) % 2000
    mp.spawn(_worker, args=(world, port, n, G, F, tile_rows, str(tmp_path)), nprocs=world, join=True)
    o = Oracle()
    cc = synth.crosscat(901, G, n, n_gp=F // 2, n_bb=F - F // 2)
    prior = o.py_prior(synth.PY_ALPHA, synth.PY_D, cc["sizes"])
    full = cases.oracle_scores(o, cc["features"], prior=prior)
    want = o.sample_rows(full.copy(), cc["u"])
    got = np.full(n, -1, np.int32)
    seen = np.zeros(n, bool)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        off = 0
        for lo, hi in d["rows"]:
            got[lo:hi] = d["assign"][off:off + hi - lo]
            assert not seen[lo:hi].any()  # every row is owned by exactly one rank
            seen[lo:hi] = True
            off += hi - lo
    assert seen.all()
    # summing features in a different association order moves scores by rounding only: indices agree
    # except near-ties
    ok = cases.explained_mismatch(full.astype(np.float64), cc["u"], got, want, 2e-5)
    assert ok.all()
    assert np.mean(got == want) > 0.99


# ---- row-sharded update row (batched add_value): accumulate -> all-reduce -> merge -----------------
def _update_worker(rank, world, port, n, G, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.gp(515, G, n)
    assign = np.random.default_rng(3).integers(0, G, n)
    lo, hi = sharding.row_shard(n, rank, world)
    stats = {"count": w["count"].astype(np.int64), "sum": w["sum"].astype(np.int64)}  # this rank's replica

    def accumulate(xchg):  # stand-in for dist_b200_rows_accumulate: [F=1][4][G] float64
        x = xchg.numpy()
        x[:] = 0
        x[0, 0] = np.bincount(assign[lo:hi], minlength=G)
        x[0, 1] = np.bincount(assign[lo:hi], weights=w["values"][lo:hi].astype(np.float64), minlength=G)

    def merge(xchg, sign):  # stand-in for dist_b200_rows_merge
        x = xchg.numpy()
        stats["count"] += sign * x[0, 0].astype(np.int64)
        stats["sum"] += sign * x[0, 1].astype(np.int64)

    xchg = torch.zeros((1, 4, G), dtype=torch.float64)
    sharding.row_sharded_update(accumulate, merge, xchg, +1)
    np.savez(os.path.join(out_dir, "upd%d.npz" % rank), count=stats["count"], sum=stats["sum"])
    dist.destroy_process_group()


def test_row_sharded_update_replicas_agree(tmp_path):
    world, n, G = 2, 3001, 17
    port = 29800 + os.getpid() % 1000
    mp.spawn(_update_worker, args=(world, port, n, G, str(tmp_path)), nprocs=world, join=True)
    w = synth.gp(515, G, n)
    assign = np.random.default_rng(3).integers(0, G, n)
    count = w["count"].astype(np.int64) + np.bincount(assign, minlength=G)
    total = w["sum"].astype(np.int64) + np.bincount(assign, weights=w["values"].astype(np.float64), minlength=G).astype(np.int64)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "upd%d.npz" % r))
        assert np.array_equal(d["count"], count) and np.array_equal(d["sum"], total)


def test_recommended_row_shards():
    from distributions_b200 import sharding
    assert [sharding.recommended_row_shards(w) for w in (1, 2, 4, 8)] == [1, 1, 2, 4]
