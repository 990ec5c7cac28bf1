"""NormalInverseWishart statistics on the device (SURVEY.md §8f ranks 1 and 2 for niw): batched Group::add_value /
remove_value (niw.hpp:247-276) as a per-group SYRK, Group::score_data (niw.hpp:296-308) over the resident statistics,
the row-shard exchange block, and the wire round trip.  The oracle restates the reference's one-value-at-a-time float
updates; the device sums are the correctly rounded ones, so statistics agree to the reference's own accumulation error."""
import os

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


@pytest.fixture(scope="module")
def golden_niw():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "niw_golden.npz"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def stats_of(f, G, d):
    raw = f.download_stats(4 * G * (1 + d + d * d))
    count = raw[:4 * G].view(np.int32).copy()
    sum_x = raw[4 * G:4 * G * (1 + d)].view(np.float32).reshape(G, d).copy()
    sum_xxT = raw[4 * G * (1 + d):].view(np.float32).reshape(G, d, d).copy()
    return count, sum_x, sum_xxT


def assignments(rng, G, n, skew):
    a = rng.integers(0, G - 1, n).astype(np.int32)  # the last group receives nothing
    if skew:
        a[rng.random(n) < skew] = 1                   # one group takes most rows: many slices of one group
    return a


def expected_after(oracle, w, values, assign, sign):
    """the reference's way: Group::add_value / remove_value one value at a time, in float"""
    G, d = w["count"].size, w["mu"].size
    count, sum_x, sum_xxT = w["count"].copy(), w["sum_x"].copy(), w["sum_xxT"].copy()
    for g in range(G):
        rows = values[assign == g]
        if len(rows):
            count[g], sum_x[g], sum_xxT[g] = oracle.niw_group_update(sign, int(count[g]), sum_x[g], sum_xxT[g].reshape(d, d), rows)
    return count, sum_x, sum_xxT


def scale_of(w, values, assign):
    """per group: sum |x x^T| over everything the group has seen -- the magnitude float accumulation errors scale with"""
    G, d = w["count"].size, w["mu"].size
    s = np.abs(w["sum_xxT"]).reshape(G, -1).max(axis=1) + 1.0
    for g in range(G):
        rows = values[assign == g].astype(np.float64)
        if len(rows):
            s[g] += np.abs(rows[:, :, None] * rows[:, None, :]).sum(0).max()
    return s


@pytest.mark.parametrize("d,G,n,skew", [(3, 7, 2001, 0.0), (32, 40, 6001, 0.0), (32, 12, 5000, 0.8), (8, 300, 4000, 0.0)])
def test_niw_add_rows_matches_sequential_add_value(ctx, oracle, d, G, n, skew):
    from distributions_b200 import capi
    w = synth.niw(3100 + d + G, G, n, d=d)
    rng = np.random.default_rng(5)
    assign = assignments(rng, G, n, skew)
    f = ctx.feature(capi.NIW).update_all(w)
    ctx.add_rows_batch([f], [dev(w["values"])], dev(assign), n)
    count, sum_x, sum_xxT = stats_of(f, G, d)
    ecount, esum_x, esum_xxT = expected_after(oracle, w, w["values"], assign, +1)
    assert np.array_equal(count, ecount)
    s = scale_of(w, w["values"], assign)
    # the reference's sequential float sums carry up to count * 2^-24 relative error; the device sums are exact to 1 ulp
    tol = (ecount + 8.0) * 6e-8 * s
    assert np.all(np.abs(sum_xxT - esum_xxT).reshape(G, -1).max(axis=1) <= tol)
    assert np.all(np.abs(sum_x - esum_x).max(axis=1) <= tol)
    # exact check against float64 sums
    for g in range(G):
        rows = w["values"][assign == g].astype(np.float64)
        np.testing.assert_allclose(sum_xxT[g], w["sum_xxT"][g] + rows.T @ rows, rtol=0, atol=2e-7 * s[g])
    # scoring after the update sees the new records (same stream, ready event): compare with a feature loaded from the
    # correctly rounded float64 sums (what the device holds, up to an ulp) -- tight -- and with one loaded from the
    # reference's sequential float sums, whose accumulation error the posterior's cancellation amplifies -- loose
    xs, xxs = w["sum_x"].astype(np.float64), w["sum_xxT"].astype(np.float64)
    for g in range(G):
        rows = w["values"][assign == g].astype(np.float64)
        xs[g] += rows.sum(0)
        xxs[g] += rows.T @ rows
    exact = ctx.feature(capi.NIW).update_all(dict(w, count=ecount, sum_x=xs.astype(np.float32), sum_xxT=xxs.astype(np.float32)))
    seq = ctx.feature(capi.NIW).update_all(dict(w, count=ecount, sum_x=esum_x, sum_xxT=esum_xxT))
    m = min(n, 512)
    a, b, c = (torch.empty((m, G), device="cuda") for _ in range(3))
    vals = dev(w["values"][:m])
    ctx.score_batch([f], [vals], m, None, a)
    ctx.score_batch([exact], [vals], m, None, b)
    ctx.score_batch([seq], [vals], m, None, c)
    torch.cuda.synchronize()
    assert float((a - b).abs().max()) <= 5e-4
    assert float((a - c).abs().max()) <= 5e-2


def test_niw_remove_rows_and_emptied_group(ctx, oracle):
    from distributions_b200 import capi
    d, G, n = 32, 9, 3000
    w = synth.niw(3300, G, n, d=d)
    w["count"][:] = 0
    w["sum_x"][:] = 0
    w["sum_xxT"][:] = 0
    rng = np.random.default_rng(6)
    assign = assignments(rng, G, n, 0.0)
    f = ctx.feature(capi.NIW).update_all(w)
    vals, a_dev = dev(w["values"]), dev(assign)
    ctx.add_rows_batch([f], [vals], a_dev, n)
    keep = assign != 2                      # everything of group 2 leaves again: the group is emptied
    gone = np.nonzero(~keep)[0]
    ctx.remove_rows_batch([f], [dev(w["values"][gone])], dev(assign[gone]), len(gone))
    count, sum_x, sum_xxT = stats_of(f, G, d)
    assert count[2] == 0 and not sum_x[2].any() and not sum_xxT[2].any()   # exact zeros, as Group::init
    for g in range(G):
        rows = w["values"][keep & (assign == g)].astype(np.float64)
        assert count[g] == len(rows)
        if len(rows):
            np.testing.assert_allclose(sum_xxT[g], rows.T @ rows, rtol=0, atol=1e-6 * np.abs(rows.T @ rows).max() + 1e-3)
            np.testing.assert_allclose(sum_x[g], rows.sum(0), rtol=0, atol=1e-6 * np.abs(rows).sum(0).max() + 1e-4)
    # zero rows: nothing changes
    before = f.download_stats(4 * G * (1 + d + d * d)).copy()
    ctx.add_rows_batch([f], [vals], a_dev, 0)
    assert np.array_equal(f.download_stats(4 * G * (1 + d + d * d)), before)


def test_niw_add_rows_skips_negative_ids_and_many_groups(ctx, oracle):
    """negative / out-of-range group ids are skipped (as in every add_rows path); G beyond one scan thread per group"""
    from distributions_b200 import capi
    d, G, n = 4, 2500, 9000
    w = synth.niw(3500, G, n, d=d)
    rng = np.random.default_rng(9)
    assign = rng.integers(0, G, n).astype(np.int32)
    assign[::7] = -1
    assign[3::11] = G + 5
    f = ctx.feature(capi.NIW).update_all(w)
    ctx.add_rows_batch([f], [dev(w["values"])], dev(assign), n)
    count, sum_x, sum_xxT = stats_of(f, G, d)
    valid = (assign >= 0) & (assign < G)
    assert np.array_equal(count, w["count"] + np.bincount(assign[valid], minlength=G).astype(np.int32))
    xs = w["sum_x"].astype(np.float64)
    np.add.at(xs, assign[valid], w["values"][valid].astype(np.float64))
    np.testing.assert_allclose(sum_x, xs, rtol=1e-6, atol=1e-4)
    xx = w["sum_xxT"].astype(np.float64)
    v = w["values"][valid].astype(np.float64)
    np.add.at(xx, assign[valid], v[:, :, None] * v[:, None, :])
    np.testing.assert_allclose(sum_xxT, xx, rtol=1e-6, atol=1e-3)


def test_niw_row_shard_exchange(ctx, oracle):
    """two 'ranks' on one device: accumulate halves -> summed exchange block -> merge == add_rows over all rows"""
    from distributions_b200 import capi
    d, G, n = 32, 11, 4001
    w = synth.niw(3400, G, n, d=d)
    wn = synth.nich(3401, G, n)
    assign = assignments(np.random.default_rng(8), G, n, 0.0)
    single = [ctx.feature(capi.NIW).update_all(w), ctx.feature(capi.NICH).update_all(wn)]
    shard = [ctx.feature(capi.NIW).update_all(w), ctx.feature(capi.NICH).update_all(wn)]
    cols = [w["values"], wn["values"].astype(np.float32)]
    ctx.add_rows_batch(single, [dev(c) for c in cols], dev(assign), n)
    nd = ctx.rows_xchg_doubles(shard)
    assert nd == 4 * G + G * (1 + d + d * d)
    total = torch.zeros(nd, dtype=torch.float64, device="cuda")
    for lo, hi in ((0, 1500), (1500, n)):
        x = torch.full((nd,), 7.0, dtype=torch.float64, device="cuda")
        ctx.rows_accumulate(shard, [dev(c[lo:hi]) for c in cols], dev(assign[lo:hi]), hi - lo, x)
        total += x
    ctx.rows_merge(shard, total, +1)
    ca, xa, ma = stats_of(single[0], G, d)
    cb, xb, mb = stats_of(shard[0], G, d)
    assert np.array_equal(ca, cb)
    np.testing.assert_allclose(xa, xb, rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(ma, mb, rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("name", cases.NIW_GOLDEN_CASES)
def test_niw_score_data_golden_and_oracle(ctx, oracle, golden_niw, name):
    """score_data over the resident statistics: == the C restatement of niw.hpp:296-308 (same fast_lgamma / fast_log,
    determinants in double on both sides) and, through it, the reference's exact-math Python (dbg/models/niw.py:202-217)"""
    from distributions_b200 import capi
    c = cases.niw_golden_case(golden_niw, name)
    d = c["mu"].size
    f = ctx.feature(capi.NIW).update_all(c)
    packed = np.concatenate([[c["kappa"], c["nu"]], c["mu"], c["psi"].ravel()]).astype(np.float32)
    # a second grid point with other hyper-parameters
    other = packed.copy()
    other[0], other[1] = packed[0] * 1.7, packed[1] + 2.5
    other[2:2 + d] += 0.25
    got = f.score_data_grid(np.stack([packed, other]))
    want0 = oracle.niw_score_data(c["mu"], c["kappa"], c["psi"], c["nu"], c["count"], c["sum_x"], c["sum_xxT"])
    want1 = oracle.niw_score_data(other[2:2 + d], float(other[0]), c["psi"], float(other[1]), c["count"], c["sum_x"], c["sum_xxT"])
    for g, wnt in zip(got, (want0, want1)):
        # term by term the same float expression; the sum over groups is accumulated in double here, in float there
        assert abs(float(g) - wnt) <= 3e-6 * abs(wnt) + 1e-4 * c["count"].size + cases.LOG_STEP * (c["nu"] + c["count"].sum() + d)
    ref = float(np.sum(c["score_data"]))
    assert abs(float(got[0]) - ref) <= 1e-3 * (1 + abs(float(got[0])) + abs(ref))
    corr, tol = cases.niw_score_data_tolerance(oracle, c)
    assert abs(float(got[0]) - (ref + corr)) <= tol + 3e-6 * abs(ref)


def test_niw_wire_round_trip(ctx):
    """update_all_wire == update_all for niw (the golden messages come from the reference's schema), and the resident
    statistics dump back to the same Group messages"""
    from distributions_b200 import capi
    wire = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_golden.npz"))
    seed, G, kw = cases.WIRE["niw"]
    w = synth.niw(seed, G, 8, **kw)
    d = kw["d"]
    lens = wire["niw_group_lens"]
    blob = wire["niw_groups"].tobytes()
    offs = np.concatenate([[0], np.cumsum(lens)])
    g_msgs = [blob[offs[i]:offs[i + 1]] for i in range(lens.size)]
    a = ctx.feature(capi.NIW).update_all(w)
    b = ctx.feature(capi.NIW).update_all_wire(wire["niw_shared"].tobytes(), g_msgs)
    nb = 4 * G * (1 + d + d * d)
    assert np.array_equal(a.download_stats(nb), b.download_stats(nb))
    sa, sb = torch.empty((8, G), device="cuda"), torch.empty((8, G), device="cuda")
    vals = dev(w["values"])
    ctx.score_batch([a], [vals], 8, None, sa)
    ctx.score_batch([b], [vals], 8, None, sb)
    assert torch.equal(sa, sb)
    assert b.dump_groups_wire() == g_msgs
