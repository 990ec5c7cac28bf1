"""Batched Group::add_value on the device (the first 'next' row of SURVEY.md §8f): the statistics and
caches after dist_b200_feature_add_rows must equal the reference's one-value-at-a-time add_value
(restated in the oracle / simple integer sums) followed by update_all; counts exactly, nich's float
statistics within the pairwise-vs-sequential rounding (1e-5 relative)."""
import numpy as np
import pytest

import cases
from distributions_b200 import synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from distributions_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _expected_after_add(oracle, w, assign):
    """apply Group::add_value for every (value, group) pair sequentially, the reference's way"""
    w2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
    m, vals = w["model"], w["values"]
    G = w["sizes"].size
    if m == "nich":
        for g in range(G):
            xs = vals[assign == g]
            if len(xs):
                c, mean, ctv = oracle.nich_group_update(+1, int(w["count"][g]), float(w["mean"][g]), float(w["ctv"][g]), xs)
                w2["count"][g], w2["mean"][g], w2["ctv"][g] = c, mean, ctv
    elif m == "gp":
        w2["count"] = w["count"] + np.bincount(assign, minlength=G).astype(np.uint32)
        w2["sum"] = w["sum"] + np.bincount(assign, weights=vals.astype(np.float64), minlength=G).astype(np.uint32)
    elif m == "bb":
        w2["heads"] = w["heads"] + np.bincount(assign[vals != 0], minlength=G).astype(np.int32)
        w2["tails"] = w["tails"] + np.bincount(assign[vals == 0], minlength=G).astype(np.int32)
    elif m == "dd":
        np.add.at(w2["counts"], (assign, vals), 1)
    elif m == "dpd":
        known = vals != 0xFFFFFFFF
        np.add.at(w2["counts"], (assign[known], vals[known].astype(np.int64)), 1)
    return w2


@pytest.mark.parametrize("skew", [0.0, 0.85])
@pytest.mark.parametrize("name,G,n", [("nich", 37, 5003), ("nich", 3000, 20000), ("gp", 21, 5001), ("bb", 13, 5002),
                                      ("dd", 29, 5000), ("dpd", 19, 4999)])
def test_add_rows_matches_sequential_add_value(ctx, oracle, name, G, n, skew):
    from distributions_b200 import capi
    ids = {"dd": capi.DD, "dpd": capi.DPD, "bb": capi.BB, "gp": capi.GP, "nich": capi.NICH}
    kw = dict(dim=16) if name == "dd" else (dict(V=100, other_frac=0.05) if name == "dpd" else {})
    w = getattr(synth, name)(4000 + G, G, n, **kw)
    rng = np.random.default_rng(G)
    assign = rng.integers(0, G, n).astype(np.int32)
    # a CRP-like skew: most rows in two groups, so that many lanes of a warp share a group
    assign = np.where(rng.random(n) < skew, np.where(rng.random(n) < 0.7, 3, 5), assign).astype(np.int32)
    assign[::17] = -1  # rows left unassigned are skipped
    f = ctx.feature(ids[name]).update_all(w)
    col = dev(w["values"].astype(capi.COLUMN_DTYPE[ids[name]]))
    f.add_rows(col, dev(assign), n)
    keep = assign >= 0
    want = _expected_after_add(oracle, dict(w, values=w["values"][keep]), assign[keep])
    # statistics
    if name == "nich":
        raw = f.download_stats(12 * G)
        cnt = raw[:4 * G].view(np.int32)
        mean = raw[4 * G:8 * G].view(np.float32)
        ctv = raw[8 * G:].view(np.float32)
        assert np.array_equal(cnt, want["count"])
        np.testing.assert_allclose(mean, want["mean"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(ctv, want["ctv"], rtol=2e-4, atol=1e-3)
    elif name == "gp":
        raw = f.download_stats(8 * G).view(np.uint32)
        assert np.array_equal(raw[:G], want["count"]) and np.array_equal(raw[G:], want["sum"])
    elif name == "bb":
        raw = f.download_stats(8 * G).view(np.int32)
        assert np.array_equal(raw[:G], want["heads"]) and np.array_equal(raw[G:], want["tails"])
    elif name == "dd":
        raw = f.download_stats(4 * G * 16).view(np.int32).reshape(G, 16)
        assert np.array_equal(raw, want["counts"])
    else:
        raw = f.download_stats(4 * G * 100).view(np.int32).reshape(G, 100)
        assert np.array_equal(raw, want["counts"])
    # caches == update_all on the reference-side statistics
    rows = {"nich": 4, "gp": 3, "bb": 2, "dd": 16, "dpd": 101}[name]
    got = f.download_caches(rows)
    exp = cases.oracle_caches(oracle, want)
    if name in ("dd", "dpd"):
        exp = exp[:-1] - exp[-1][None, :]
        assert np.array_equal(got, exp)
    elif name == "nich":
        np.testing.assert_allclose(got, exp, rtol=3e-4, atol=3e-4)
    else:
        np.testing.assert_allclose(got, exp, rtol=2e-6, atol=2e-6)


def test_device_side_sweep(ctx, oracle):
    """score -> sample -> add_value -> prior update -> score again, all on the device; the second pass
    must equal the oracle's scores for the statistics the first pass produced."""
    from distributions_b200 import capi
    G, n = 50, 4000
    w = synth.nich(77, G, n)
    f = ctx.feature(capi.NICH).update_all(w)
    col, u = dev(w["values"]), dev(w["u"])
    sizes = dev(w["sizes"].astype(np.int32))
    prior = torch.empty(G, device="cuda")
    ctx.prior_pitman_yor_dev(synth.PY_ALPHA, synth.PY_D, G, sizes, prior)
    assign = torch.empty(n, device="cuda", dtype=torch.int32)
    ctx.score_sample_batch([f], [col], n, prior, u, assign)
    f.add_rows(col, assign, n)
    ctx.count_assignments(assign, n, G, sizes, accumulate=True)
    ctx.prior_pitman_yor_dev(synth.PY_ALPHA, synth.PY_D, G, sizes, prior)
    scores = torch.empty((n, G), device="cuda")
    assign2 = torch.empty(n, device="cuda", dtype=torch.int32)
    ctx.score_sample_batch([f], [col], n, prior, u, assign2, scores)
    torch.cuda.synchronize()
    a1 = assign.cpu().numpy()
    want = _expected_after_add(oracle, w, a1)
    sizes2 = w["sizes"] + np.bincount(a1, minlength=G).astype(np.int32)
    assert np.array_equal(sizes.cpu().numpy(), sizes2)
    prior2 = oracle.py_prior(synth.PY_ALPHA, synth.PY_D, sizes2)
    np.testing.assert_allclose(prior.cpu().numpy(), prior2, atol=2e-6)
    exp = cases.oracle_scores(oracle, [want], prior=prior2)
    got = scores.cpu().numpy()
    coeff = np.abs(oracle.nich_caches(want["shared"], want["count"], want["mean"], want["ctv"])[1])[None, :]
    # statistics agree to ~1e-5 (pairwise vs sequential); a perturbed argument can cross a fast_log step
    assert np.all(np.abs(got - exp) <= 2e-4 * (1 + np.abs(exp)) + 6.2e-5 * coeff)
    assert np.median(np.abs(got - exp)) < 1e-4


def test_add_rows_batch_mixed_kind(ctx, oracle):
    """all features of a kind in one call (more than one 128-feature launch, mixed models)"""
    from distributions_b200 import capi
    G, n = 24, 3000
    ids = {"bb": capi.BB, "gp": capi.GP, "nich": capi.NICH, "dd": capi.DD}
    names = ["gp", "bb", "nich"] * 50 + ["dd"]
    ws = [getattr(synth, m)(900 + i, G, n, **(dict(dim=8) if m == "dd" else {})) for i, m in enumerate(names)]
    rng = np.random.default_rng(5)
    assign = rng.integers(0, G, n).astype(np.int32)
    feats = [ctx.feature(ids[m]).update_all(w) for m, w in zip(names, ws)]
    cols = [dev(w["values"].astype(capi.COLUMN_DTYPE[ids[m]])) for m, w in zip(names, ws)]
    ctx.add_rows_batch(feats, cols, dev(assign), n)
    for m, w, f in zip(names, ws, feats):
        want = _expected_after_add(oracle, w, assign)
        rows = {"nich": 4, "gp": 3, "bb": 2, "dd": 8}[m]
        got = f.download_caches(rows)
        exp = cases.oracle_caches(oracle, want)
        if m == "dd":
            assert np.array_equal(got, exp[:-1] - exp[-1][None, :])
        elif m == "nich":
            np.testing.assert_allclose(got, exp, rtol=3e-4, atol=3e-4)
        else:
            np.testing.assert_allclose(got, exp, rtol=2e-6, atol=2e-6)
    # scoring after the batched update sees the new caches (stream ordering through the ready events)
    sizes = ws[0]["sizes"]
    prior = dev(oracle.py_prior(synth.PY_ALPHA, synth.PY_D, sizes))
    scores = torch.empty((n, G), device="cuda")
    sel = [0, 1, 3, 4]  # gp, bb features only: exact statistics, tight comparison
    ctx.score_batch([feats[i] for i in sel], [cols[i] for i in sel], n, prior, scores)
    after = [_expected_after_add(oracle, ws[i], assign) for i in sel]
    exp = cases.oracle_scores(oracle, after, prior=prior.cpu().numpy())
    got = scores.cpu().numpy()
    # the gp envelope of tests/test_gpu_parity.py: one ulp of the cached score[g] survives the cancellation
    env = sum(6e-7 * (1.0 + np.abs(oracle.gp_caches(w["shared"], w["count"], w["sum"])[0])) for w in after if w["model"] == "gp")
    assert np.all(np.abs(got - exp) <= 3e-6 * (1 + np.abs(exp)) + env[None, :])


def test_add_rows_batch_host_buffers(ctx, oracle):
    """the host-buffer entry a reference-side binding calls: same result as the device-pointer entry"""
    from distributions_b200 import capi
    G, n = 17, 2500
    ws = [synth.gp(31, G, n), synth.bb(32, G, n), synth.nich(33, G, n)]
    ids = [capi.GP, capi.BB, capi.NICH]
    assign = np.random.default_rng(9).integers(0, G, n).astype(np.int32)
    a = [ctx.feature(i).update_all(w) for i, w in zip(ids, ws)]
    b = [ctx.feature(i).update_all(w) for i, w in zip(ids, ws)]
    ctx.add_rows_batch_host(a, [w["values"] for w in ws], assign)
    ctx.add_rows_batch(b, [dev(w["values"].astype(capi.COLUMN_DTYPE[i])) for i, w in zip(ids, ws)], dev(assign), n)
    for fa, fb, nb, rows in zip(a, b, (8 * G, 8 * G, 12 * G), (3, 2, 4)):
        assert np.array_equal(fa.download_stats(nb)[:4 * G], fb.download_stats(nb)[:4 * G])  # integer arrays exact
        np.testing.assert_allclose(fa.download_caches(rows), fb.download_caches(rows), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("skew", [0.0, 0.85])
@pytest.mark.parametrize("name,G,n", [("nich", 37, 5003), ("gp", 21, 5001), ("bb", 13, 5002), ("dd", 29, 5000), ("dpd", 19, 4999)])
def test_remove_rows_matches_sequential_remove_value(ctx, oracle, name, G, n, skew):
    """batched Group::remove_value: start from groups that contain the rows, remove them, compare with the
    reference's one-at-a-time remove_value (oracle restatement) and with the statistics before the add."""
    from distributions_b200 import capi
    ids = {"dd": capi.DD, "dpd": capi.DPD, "bb": capi.BB, "gp": capi.GP, "nich": capi.NICH}
    kw = dict(dim=16) if name == "dd" else (dict(V=100, other_frac=0.05) if name == "dpd" else {})
    w = getattr(synth, name)(6000 + G, G, n, **kw)
    rng = np.random.default_rng(G + 1)
    assign = rng.integers(0, G, n).astype(np.int32)
    assign = np.where(rng.random(n) < skew, np.where(rng.random(n) < 0.7, 3, 5), assign).astype(np.int32)
    assign[::13] = -1
    keep = assign >= 0
    full = _expected_after_add(oracle, dict(w, values=w["values"][keep]), assign[keep])  # groups holding the rows
    full["values"] = w["values"]
    f = ctx.feature(ids[name]).update_all(full)
    col = dev(w["values"].astype(capi.COLUMN_DTYPE[ids[name]]))
    ctx.remove_rows_batch([f], [col], dev(assign), n)
    if name == "nich":
        raw = f.download_stats(12 * G)
        cnt, mean, ctv = raw[:4 * G].view(np.int32), raw[4 * G:8 * G].view(np.float32), raw[8 * G:].view(np.float32)
        assert np.array_equal(cnt, w["count"])
        # the reference's sequential removal, one value at a time
        seq_mean, seq_ctv = np.empty(G, np.float32), np.empty(G, np.float32)
        for g in range(G):
            xs = w["values"][keep][assign[keep] == g]
            c, m, v = oracle.nich_group_update(-1, int(full["count"][g]), float(full["mean"][g]), float(full["ctv"][g]), xs)
            assert c == w["count"][g]
            seq_mean[g], seq_ctv[g] = m, v
        scale = 1.0 + np.abs(w["values"]).max()
        # both are float32 subtractions of nearly equal sums: compare at the accuracy of the statistics held
        assert np.all(np.abs(mean - seq_mean) <= 2e-5 * scale)
        assert np.all(np.abs(ctv - seq_ctv) <= 1e-4 * (np.abs(full["ctv"]) + 1))
        assert np.all(np.abs(mean - w["mean"]) <= 2e-5 * scale)
        assert np.all(np.abs(ctv - w["ctv"]) <= 1e-4 * (np.abs(full["ctv"]) + 1))
    elif name == "gp":
        raw = f.download_stats(8 * G).view(np.uint32)
        assert np.array_equal(raw[:G], w["count"]) and np.array_equal(raw[G:], w["sum"])
    elif name == "bb":
        raw = f.download_stats(8 * G).view(np.int32)
        assert np.array_equal(raw[:G], w["heads"]) and np.array_equal(raw[G:], w["tails"])
    else:
        dim = 16 if name == "dd" else 100
        raw = f.download_stats(4 * G * dim).view(np.int32).reshape(G, dim)
        assert np.array_equal(raw, w["counts"])
    if name != "nich":  # caches are a pure function of the (exact) integer statistics
        rows = {"gp": 3, "bb": 2, "dd": 16, "dpd": 101}[name]
        got = f.download_caches(rows)
        exp = cases.oracle_caches(oracle, w)
        if name in ("dd", "dpd"):
            assert np.array_equal(got, exp[:-1] - exp[-1][None, :])
        else:
            np.testing.assert_allclose(got, exp, rtol=2e-6, atol=2e-6)


def test_remove_rows_empties_group(ctx, oracle):
    """a group emptied by the batch is reset exactly (count = mean = ctv = 0), as nich.hpp:152-156"""
    from distributions_b200 import capi
    G, n = 5, 64
    w = synth.nich(3, G, n)
    w["count"][:] = 0
    w["mean"][:] = 0
    w["ctv"][:] = 0
    assign = (np.arange(n) % 2).astype(np.int32)  # groups 0 and 1 only
    f = ctx.feature(capi.NICH).update_all(w)
    col = dev(w["values"])
    f.add_rows(col, dev(assign), n)
    ctx.remove_rows_batch([f], [col], dev(assign), n)
    raw = f.download_stats(12 * G)
    assert not raw.any()
    np.testing.assert_array_equal(f.download_caches(4), cases.oracle_caches(oracle, w))


def test_row_shard_exchange_with_count_tables(ctx, oracle):
    """the row-shard exchange on one device: two 'ranks' accumulate their halves of the rows into two exchange buffers,
    the buffers are summed (the all-reduce) and merged -- pooled models and the dd / dpd count tables in one buffer;
    the result must equal add_rows_batch over all rows, bit for bit, and a -1 merge must restore the original"""
    from distributions_b200 import capi
    G, n = 17, 4001
    names = ["gp", "dd", "nich", "dpd", "bb"]
    ids = {"bb": capi.BB, "gp": capi.GP, "nich": capi.NICH, "dd": capi.DD, "dpd": capi.DPD}
    kw = {"dd": dict(dim=8), "dpd": dict(V=53)}
    ws = [getattr(synth, m)(1300 + i, G, n, **kw.get(m, {})) for i, m in enumerate(names)]
    assign = np.random.default_rng(77).integers(0, G, n).astype(np.int32)
    cols_h = [w["values"].astype(capi.COLUMN_DTYPE[ids[m]]) for m, w in zip(names, ws)]
    stat_bytes = {"gp": 8 * G, "dd": 4 * G * 8, "nich": 12 * G, "dpd": 4 * G * 53, "bb": 8 * G}

    single = [ctx.feature(ids[m]).update_all(w) for m, w in zip(names, ws)]
    before = [f.download_stats(stat_bytes[m]) for m, f in zip(names, single)]
    ctx.add_rows_batch(single, [dev(c) for c in cols_h], dev(assign), n)

    shard = [ctx.feature(ids[m]).update_all(w) for m, w in zip(names, ws)]
    nd = ctx.rows_xchg_doubles(shard)
    assert nd == 3 * 4 * G + G * 8 + G * 53
    total = torch.zeros(nd, dtype=torch.float64, device="cuda")
    for lo, hi in ((0, n // 3), (n // 3, n)):
        x = torch.full((nd,), 123.0, dtype=torch.float64, device="cuda")  # stale contents must not leak
        ctx.rows_accumulate(shard, [dev(c[lo:hi]) for c in cols_h], dev(assign[lo:hi]), hi - lo, x)
        total += x
    # an empty shard contributes zeros
    x = torch.full((nd,), 5.0, dtype=torch.float64, device="cuda")
    ctx.rows_accumulate(shard, [dev(c[:1]) for c in cols_h], dev(assign[:1]), 0, x)
    assert float(x.abs().sum()) == 0.0
    ctx.rows_merge(shard, total, +1)
    for m, a, b in zip(names, single, shard):
        sa, sb = a.download_stats(stat_bytes[m]), b.download_stats(stat_bytes[m])
        if m == "nich":  # the pooled merge is the same arithmetic on the same sums up to the order of the double additions
            np.testing.assert_allclose(sa.view(np.float32)[G:], sb.view(np.float32)[G:], rtol=1e-6, atol=1e-6)
            assert np.array_equal(sa.view(np.int32)[:G], sb.view(np.int32)[:G])
        else:
            assert np.array_equal(sa, sb), m
        rows = {"nich": 4, "gp": 3, "bb": 2, "dd": 8, "dpd": 53 + 1}[m]
        ca, cb = a.download_caches(rows), b.download_caches(rows)
        if m == "nich":
            np.testing.assert_allclose(ca, cb, rtol=1e-5, atol=1e-5)
        else:
            assert np.array_equal(ca, cb), m
    ctx.rows_merge(shard, total, -1)
    for m, f, b0 in zip(names, shard, before):
        if m != "nich":
            assert np.array_equal(f.download_stats(stat_bytes[m]), b0), m
