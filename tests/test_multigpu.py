"""Two-GPU tests of the feature-shard paths (NCCL reduce-scatter and the fused NVLink peer push).
Needs >= 2 CUDA devices: skipped on single-GPU boxes (run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, G, F, out_dir, row_shards):
    import torch.distributed as dist
    from distributions_b200 import capi, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cc = synth.crosscat(911, G, n, n_gp=F // 2, n_bb=F - F // 2)
    ids = {"gp": capi.GP, "bb": capi.BB}
    ctx = capi.Context(rank)
    u = torch.from_numpy(cc["u"]).to(dev)
    prior = torch.empty(G, device=dev)
    ctx.prior_pitman_yor(synth.PY_ALPHA, synth.PY_D, cc["sizes"], prior)

    def load(features, r0, r1):
        feats = [ctx.feature(ids[cc["features"][f]["model"]]).update_all(cc["features"][f]) for f in features]
        cols = [torch.from_numpy(np.ascontiguousarray(cc["features"][f]["values"][r0:r1],
                                                      dtype=capi.COLUMN_DTYPE[ids[cc["features"][f]["model"]]])).to(dev) for f in features]
        return feats, cols

    # (1) fused peer push: pure feature shards (row_shards = 1) or the feature x row hybrid
    shards = sharding.PeerFeatureShards(ctx, n, G, row_shards=row_shards)
    r0, r1 = shards.rows()
    feats, cols = load(shards.features(F), r0, r1)
    lo, hi = shards.owned()
    a_push = torch.full((hi - lo,), -1, device=dev, dtype=torch.int32)
    for _ in range(3):  # three steps: the two slot buffers alternate and are reused, flags advance by epoch
        a_push.fill_(-1)
        shards.step(feats, cols, prior, u[lo:hi], a_push)
    torch.cuda.synchronize()

    # (2) NCCL reduce-scatter orchestration (pure feature shards over the whole world)
    feats, cols = load(sharding.feature_shard(F, rank, world), 0, n)

    def score_partial(l, h, out):
        ctx.score_batch(feats, [c[l:h] for c in cols], h - l, prior if rank == 0 else None, out)

    def sample_block(scores, ub, out):
        ctx.sample_from_scores(scores, scores.shape[0], G, ub, out)

    assigns, rows = sharding.feature_sharded_score_sample(score_partial, sample_block, n, G, u, dev, tile_rows=1024,
                                                          comm_stream=torch.cuda.Stream(device=dev))
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, a_push=a_push.cpu().numpy(), rows=np.array(rows),
             a_rs=np.concatenate([a.cpu().numpy() for a in assigns]))
    shards.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,row_shards", [(2, 1), (4, 1), (4, 2), (8, 1), (8, 4)])
def test_feature_shards_multi_gpu(tmp_path, oracle, world, row_shards):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d CUDA devices" % world)
    import torch.multiprocessing as mp
    import cases
    from distributions_b200 import synth
    n, G, F = 5003, 40, 2 * world + 2
    port = 29600 + (os.getpid() + 17 * world + row_shards) % 1000
    mp.spawn(_worker, args=(world, port, n, G, F, str(tmp_path), row_shards), nprocs=world, join=True)
    cc = synth.crosscat(911, G, n, n_gp=F // 2, n_bb=F - F // 2)
    prior = oracle.py_prior(synth.PY_ALPHA, synth.PY_D, cc["sizes"])
    full = cases.oracle_scores(oracle, cc["features"], prior=prior)
    want = oracle.sample_rows(full.copy(), cc["u"])
    got_push = np.full(n, -1, np.int32)
    got_rs = np.full(n, -1, np.int32)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        got_push[int(d["lo"]):int(d["hi"])] = d["a_push"]
        off = 0
        for lo, hi in d["rows"]:
            got_rs[lo:hi] = d["a_rs"][off:off + hi - lo]
            off += hi - lo
    # fp32 sums of F feature terms in another association than the oracle's (per-rank partials, then the slots / NCCL's
    # reduction order): each addition may differ by half an ulp of |score| <= 64, i.e. 4e-6 relative to the row's total
    eps = 4e-6 * F
    for label, got in (("push", got_push), ("reduce-scatter", got_rs)):
        assert got.min() >= 0, label
        ok = cases.explained_mismatch(full.astype(np.float64), cc["u"], got, want, eps)
        assert ok.all(), "%s: %d of %d mismatching rows are not near-ties at eps %.1e (rows %s)" % (
            label, int((~ok).sum()), int((got != want).sum()), eps, np.nonzero(~ok)[0][:8])
        assert np.mean(got == want) > 0.995, label


# ---- row shards: the update row (batched add_value) with its all-reduce -----------------------------
def _update_case(capi, synth, G, n):
    """pooled models and the dd / dpd count tables in one exchange buffer"""
    ws = [synth.nich(701, G, n), synth.gp(702, G, n), synth.dd(705, G, n, dim=8), synth.bb(703, G, n), synth.bnb(704, G, n, r=2),
          synth.dpd(706, G, n, V=61)]
    ids = [capi.NICH, capi.GP, capi.DD, capi.BB, capi.BNB, capi.DPD]
    stat_bytes = (12 * G, 8 * G, 4 * G * 8, 8 * G, 8 * G, 4 * G * 61)
    cache_rows = (4, 3, 8, 2, 3, 62)
    return ws, ids, stat_bytes, cache_rows


def _update_worker(rank, world, port, n, G, out_dir):
    import torch.distributed as dist
    from distributions_b200 import capi, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ws, ids, stat_bytes, cache_rows = _update_case(capi, synth, G, n)
    assign = np.random.default_rng(12).integers(0, G, n).astype(np.int32)
    lo, hi = sharding.row_shard(n, rank, world)
    ctx = capi.Context(rank)
    feats = [ctx.feature(i).update_all(w) for i, w in zip(ids, ws)]
    cols = [torch.from_numpy(np.ascontiguousarray(w["values"][lo:hi], dtype=capi.COLUMN_DTYPE[i])).to(dev) for i, w in zip(ids, ws)]
    a_dev = torch.from_numpy(assign[lo:hi].copy()).to(dev)
    xchg = torch.zeros(ctx.rows_xchg_doubles(feats), dtype=torch.float64, device=dev)
    sharding.row_sharded_update(lambda x: ctx.rows_accumulate(feats, cols, a_dev, hi - lo, x),
                                lambda x, sign: ctx.rows_merge(feats, x, sign), xchg, +1)
    torch.cuda.synchronize()
    out = {}
    for k, (f, nb, rows) in enumerate(zip(feats, stat_bytes, cache_rows)):
        out["stats%d" % k] = f.download_stats(nb)
        out["caches%d" % k] = f.download_caches(rows)
    np.savez(os.path.join(out_dir, "upd%d.npz" % rank), **out)
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_update_two_gpus(tmp_path):
    """two ranks, each with half of the rows: after accumulate -> all-reduce -> merge both replicas hold the
    statistics a single GPU gets from all rows"""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import torch.multiprocessing as mp
    from distributions_b200 import capi, synth
    world, n, G = 2, 6001, 33
    port = 29700 + os.getpid() % 1000
    mp.spawn(_update_worker, args=(world, port, n, G, str(tmp_path)), nprocs=world, join=True)
    ws, ids, stat_bytes, cache_rows = _update_case(capi, synth, G, n)
    assign = np.random.default_rng(12).integers(0, G, n).astype(np.int32)
    ctx = capi.Context(0)
    feats = [ctx.feature(i).update_all(w) for i, w in zip(ids, ws)]
    cols = [torch.from_numpy(np.ascontiguousarray(w["values"], dtype=capi.COLUMN_DTYPE[i])).cuda() for i, w in zip(ids, ws)]
    ctx.add_rows_batch(feats, cols, torch.from_numpy(assign).cuda(), n)
    d = [np.load(os.path.join(str(tmp_path), "upd%d.npz" % r)) for r in range(world)]
    for k, (f, nb, rows) in enumerate(zip(feats, stat_bytes, cache_rows)):
        single = f.download_stats(nb)
        assert np.array_equal(d[0]["stats%d" % k], d[1]["stats%d" % k])          # replicas bit-identical
        assert np.array_equal(d[0]["caches%d" % k], d[1]["caches%d" % k])
        if k == 0:  # nich: counts exact, float statistics up to the summation order of the double accumulators
            assert np.array_equal(d[0]["stats0"][:4 * G], single[:4 * G])
            np.testing.assert_allclose(d[0]["stats0"][4 * G:].view(np.float32), single[4 * G:].view(np.float32), rtol=1e-6, atol=1e-6)
        else:
            assert np.array_equal(d[0]["stats%d" % k], single)
    ctx.close()
