"""BetaNegativeBinomial on the device against the oracle restatement (pinned to the
compiled reference, tests/test_oracle.py) and the committed reference outputs (tests/golden/rank3_golden.npz).

Tolerances, as for GammaPoisson (tests/test_gpu_parity.py): caches 2e-6 relative; scores 3e-6 * (1 + |ref|)
plus the cancellation envelope of score[g] + lgamma(beta) - lgamma(beta + alpha[g]) -- the three terms are
each rounded at their own magnitude before they cancel: 6e-7 * (1 + |score[g]| + |lgamma(beta)| + |lgamma(beta+alpha)|);
assignments identical except explained near-ties."""
import numpy as np
import pytest
from scipy.special import gammaln

import cases
from distributions_b200 import synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
EPS_TIE = 2e-5


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


def envelope(cache, values):
    score, post_beta, alpha = cache
    beta = post_beta[None, :] + values[:, None].astype(np.float64)
    return 6e-7 * (1 + np.abs(score)[None, :] + np.abs(gammaln(beta)) + np.abs(gammaln(beta + alpha[None, :])))


def run(ctx, w, n, prior, sample=True):
    from distributions_b200 import capi
    f = ctx.feature(capi.BNB).update_all(w)
    G = w["sizes"].size
    scores = torch.full((n, G), 777.0, device="cuda")
    assign = torch.full((n,), -5, device="cuda", dtype=torch.int32)
    ctx.score_sample_batch([f], [dev(w["values"][:n])], n, dev(prior), dev(w["u"][:n]), assign, scores)
    torch.cuda.synchronize()
    return f, assign.cpu().numpy(), scores.cpu().numpy()


@pytest.mark.parametrize("G,r,n", [(23, 1, 500), (100, 3, 700), (1000, 2, 300), (5, 40, 200)])
def test_bnb_scores_and_samples_match_oracle(ctx, oracle, G, r, n):
    w = synth.bnb(7100 + G, G, n, r=r)
    prior = oracle.py_prior(synth.PY_ALPHA, synth.PY_D, w["sizes"])
    f, assign, scores = run(ctx, w, n, prior)
    cache = oracle.bnb_caches(w["shared"], w["count"], w["sum"])
    got_cache = f.download_caches(3)
    np.testing.assert_allclose(got_cache, cache, rtol=2e-6, atol=2e-6 * (1 + np.abs(cache[0]).max()))
    want = cases.oracle_scores(oracle, [w], prior=prior)
    assert np.all(np.abs(scores - want) <= 3e-6 * (1 + np.abs(want)) + envelope(cache, w["values"][:n]))
    a_orc = oracle.sample_rows(scores.copy(), w["u"][:n])
    assert cases.explained_mismatch(scores.astype(np.float64), w["u"][:n], assign, a_orc, EPS_TIE).all()
    assert (assign != a_orc).mean() < 0.01


def test_bnb_golden_reference(ctx, oracle, golden_rank3):
    gd = golden_rank3
    for key, (seed, G, N, rr) in cases.BNB_GOLDEN.items():
        w = synth.bnb(seed, G, N, r=rr)
        w["u"] = gd[key + "_u"]  # the uniforms the reference consumed
        prior = oracle.py_prior(synth.PY_ALPHA, synth.PY_D, w["sizes"])
        f, assign, scores = run(ctx, w, N, prior)
        want = gd[key + "_scores"]
        assert np.all(np.abs(scores - want) <= 3e-6 * (1 + np.abs(want)) + envelope(gd[key + "_caches"], w["values"][:N]))
        ok = cases.explained_mismatch(scores.astype(np.float64), w["u"], assign, gd[key + "_assign"], EPS_TIE)
        assert ok.all()
        got = f.score_data_grid(gd[key + "_sd_grid"])
        for i, sh in enumerate(gd[key + "_sd_grid"]):
            _, scale, _ = oracle.score_data(w, sh)
            assert abs(got[i] - gd[key + "_sd_out"][i]) <= cases.accum_tol(cases.score_data_terms(w), scale)


def test_bnb_in_a_crosscat_kind_and_batched_add(ctx, oracle):
    """bnb next to gp / bb / nich in one kind; add_value / remove_value keep its (integer) statistics exact"""
    from distributions_b200 import capi
    G, n = 19, 1500
    ws = [synth.bnb(1, G, n, r=2), synth.gp(2, G, n), synth.bb(3, G, n), synth.nich(4, G, n), synth.bnb(5, G, n, r=5)]
    ids = [capi.BNB, capi.GP, capi.BB, capi.NICH, capi.BNB]
    feats = [ctx.feature(i).update_all(w) for i, w in zip(ids, ws)]
    cols = [dev(w["values"].astype(capi.COLUMN_DTYPE[i])) for i, w in zip(ids, ws)]
    prior = oracle.py_prior(synth.PY_ALPHA, synth.PY_D, ws[0]["sizes"])
    scores = torch.empty((n, G), device="cuda")
    ctx.score_batch(feats, cols, n, dev(prior), scores)
    want = cases.oracle_scores(oracle, ws, prior=prior)
    env = envelope(oracle.bnb_caches(ws[0]["shared"], ws[0]["count"], ws[0]["sum"]), ws[0]["values"]) + \
        envelope(oracle.bnb_caches(ws[4]["shared"], ws[4]["count"], ws[4]["sum"]), ws[4]["values"]) + \
        (6e-7 * (1.0 + np.abs(oracle.gp_caches(ws[1]["shared"], ws[1]["count"], ws[1]["sum"])[0])))[None, :] + \
        (1e-6 * np.abs(oracle.nich_caches(ws[3]["shared"], ws[3]["count"], ws[3]["mean"], ws[3]["ctv"])[1]))[None, :]
    assert np.all(np.abs(scores.cpu().numpy() - want) <= 5e-6 * (1 + np.abs(want)) + env)
    # batched add then remove: statistics return to the start exactly, caches to the oracle's
    assign = np.random.default_rng(8).integers(0, G, n).astype(np.int32)
    f, w = feats[0], ws[0]
    ctx.add_rows_batch([f], [cols[0]], dev(assign), n)
    raw = f.download_stats(8 * G).view(np.uint32)
    cnt = w["count"] + np.bincount(assign, minlength=G).astype(np.uint32)
    sm = w["sum"] + np.bincount(assign, weights=w["values"].astype(np.float64), minlength=G).astype(np.uint32)
    assert np.array_equal(raw[:G], cnt) and np.array_equal(raw[G:], sm)
    exp = oracle.bnb_caches(w["shared"], cnt, sm)
    np.testing.assert_allclose(f.download_caches(3), exp, rtol=2e-6, atol=2e-6 * (1 + np.abs(exp[0]).max()))
    ctx.remove_rows_batch([f], [cols[0]], dev(assign), n)
    raw = f.download_stats(8 * G).view(np.uint32)
    assert np.array_equal(raw[:G], w["count"]) and np.array_equal(raw[G:], w["sum"])
