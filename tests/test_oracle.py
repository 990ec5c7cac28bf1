"""Pins the C restatement (oracle/oracle.c) to the reference: against the committed golden vectors
(outputs of the compiled unmodified reference, tests/golden/make_golden.py) and, where
oracle/_ref is available, against the live reference on larger seeded sweeps.

Tolerances and why (all measured, see DESIGN.md "numerics"):
  * fast_log: the reference's gcc -O3 -ffast-math build fills its 2^14-entry table through libmvec's
    vector log2f (<= 4 ulp), the restatement uses scalar log2f as the source says -> <= 3e-7 relative.
  * fast_exp: the same build re-associates (1 + (x - r*b)) into ((x + 1) - r*b) -> <= 5e-6 relative.
  * fast_log is a STEP function of its argument (2^-14 relative steps, 6.1e-5 in the log): two
    evaluations of the same cell whose argument differs by one ulp (FMA contraction, -ffast-math
    re-association) can land on adjacent table entries.  Any score built from
    coeff * fast_log(arg) therefore carries an irreducible |coeff| * 6.2e-5 envelope; for nich,
    coeff = log_coeff = -(nu'+1)/2.
"""
import numpy as np
import pytest

import cases
from distributions_b200 import synth

LOG_STEP = 6.2e-5


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / (1.0 + np.abs(a) + np.abs(b))
    return np.where(a == b, 0.0, e)  # identical values (including +-inf) agree


# ------------------------------------------------------------------------------------ numerics
def test_numerics_golden(oracle, golden):
    x, y = golden["num_log_x"], golden["num_log_y"]
    assert np.all(np.abs(oracle.fast_log(x) - y) <= 3e-7 * (1 + np.abs(y)))
    x, y = golden["num_exp_x"], golden["num_exp_y"]
    got = oracle.fast_exp(x)
    norm = (y > 1.2e-38)  # above the denormal range the two agree relatively
    assert np.all(np.abs(got[norm] - y[norm]) <= 5e-6 * y[norm])
    assert np.all(np.abs(got[~norm] - y[~norm]) <= 1e-37)
    x, y = golden["num_lgamma_x"], golden["num_lgamma_y"]
    assert np.all(relerr(oracle.fast_lgamma(x), y) <= 1e-6)
    x, y = golden["num_lgamma_nu_x"], golden["num_lgamma_nu_y"]
    assert np.all(relerr(oracle.fast_lgamma_nu(x), y) <= 2e-6)
    x, y = golden["num_logfact_x"], golden["num_logfact_y"]
    assert np.all(relerr(oracle.fast_log_factorial(x), y) <= 1e-6)


def test_numerics_anchors(oracle):
    """SURVEY.md §8(c) sanity anchors measured from the compiled reference."""
    f = np.float32
    np.testing.assert_allclose(oracle.fast_log([0.0, -2.0, 2.5, np.inf]),
                               f([-88.029694, 0.693147, 0.9162907, 88.722839]), rtol=3e-7)
    np.testing.assert_allclose(oracle.fast_lgamma([3.0, 6.0, 100.5]), f([0.693153799, 4.78750277, 361.435608]), rtol=1e-6)
    np.testing.assert_allclose(oracle.fast_lgamma_nu([1.0, 5.0]), f([-0.559260964, 0.406583905]), rtol=2e-6)
    np.testing.assert_allclose(oracle.fast_log_factorial([70]), f([230.439178]), rtol=1e-6)
    assert oracle.fast_exp([0.0])[0] == 1.0
    np.testing.assert_allclose(oracle.fast_exp([-87.3]), f([1.21925e-38]), rtol=1e-5)


def test_numerics_live(oracle, ref):
    x = np.concatenate([np.logspace(-37, 38, 400001), -np.logspace(-5, 5, 1001)]).astype(np.float32)
    a, b = oracle.fast_log(x), ref.fast_log(x)
    assert np.all(np.abs(a - b) <= 3e-7 * (1 + np.abs(b)))
    x = np.linspace(-87.0, 0.49, 400001).astype(np.float32)
    a, b = oracle.fast_exp(x), ref.fast_exp(x)
    assert np.all(np.abs(a - b) <= 5e-6 * b)
    x = np.logspace(-4, 9.63, 400001).astype(np.float32)
    assert np.all(relerr(oracle.fast_lgamma(x), ref.fast_lgamma(x)) <= 1e-6)
    assert np.all(relerr(oracle.fast_lgamma_nu(x), ref.fast_lgamma_nu(x)) <= 2e-6)


# --------------------------------------------------------------------------------------- prior
def test_prior_golden(oracle, golden):
    """PitmanYor Mixture.score_value == score_add_value (test_clustering.py:242-327), all five
    reference EXAMPLES x empty_group_count in {1, 10}."""
    for j in range(5):
        for empties in (1, 10):
            key = "prior_%d_%d" % (j, empties)
            sizes = golden[key + "_sizes"]
            alpha, d = [float(v) for v in golden[key + "_alpha_d"]]
            got = oracle.py_prior(alpha, d, sizes)
            np.testing.assert_allclose(got, golden[key + "_out"], atol=2e-6, rtol=0)
            # the reference's own cross-check, at its own tolerance (tests/util.py:42)
            assert np.all(relerr(got, golden[key + "_add_value"]) <= 1e-3)
            nonempty, total = int((sizes > 0).sum()), int(sizes.sum())
            sav = np.array([oracle.py_score_add_value(alpha, d, int(s), nonempty, total, empties) for s in sizes])
            np.testing.assert_allclose(sav, golden[key + "_add_value"], atol=2e-6, rtol=0)


def test_prior_anchors(oracle):
    assert abs(oracle.py_score_add_value(1.0, 0.2, 3, 5, 20, 1) - (-2.01491833)) < 2e-6
    assert abs(oracle.py_score_add_value(1.0, 0.2, 0, 5, 20, 2) - (-3.04452634)) < 2e-6


# --------------------------------------------------------------------------------------- models
def _envelope(name, w, oracle):
    """per-group absolute envelope caused by fast_log's step structure (see module docstring)."""
    if name == "nich":
        return LOG_STEP * np.abs(oracle.nich_caches(w["shared"], w["count"], w["mean"], w["ctv"])[1])
    if name == "gp":
        # score[g] ~ -lgamma(post_alpha) and fast_lgamma(post_alpha + v) cancel in fp32
        # (gp.cc:61-66): a few ulp of |score[g]| survive any re-association of the four-term sum
        return 6e-7 * (1.0 + np.abs(oracle.gp_caches(w["shared"], w["count"], w["sum"])[0]))
    return np.zeros(w["sizes"].size)


@pytest.mark.parametrize("name", list(cases.SMALL))
def test_model_scores_golden(oracle, golden, name):
    cfg = cases.SMALL[name]
    w = cases.make(name, **cfg)
    prior = oracle.py_prior(synth.PY_ALPHA, synth.PY_D, w["sizes"])
    got = cases.oracle_scores(oracle, [w], prior=prior)
    want = golden["%s_scores" % name]
    tol = 2e-6 * (1 + np.abs(want)) + _envelope(name, w, oracle)[None, :]
    assert np.all(np.abs(got - want) <= tol)
    if name != "gp":  # (gp's envelope is plain cancellation noise, present in every cell)
        # most cells agree far below the step envelope
        assert np.mean(np.abs(got - want) <= 2e-6 * (1 + np.abs(want))) > 0.98


@pytest.mark.parametrize("name", list(cases.SMALL))
def test_model_accumulate_and_group_parity_golden(oracle, golden, name):
    """Mixture.score_value ACCUMULATES onto noise and equals per-group Group.score_value and
    score_value_group (reference test_models.py:537-594, TOL 1e-3 relative)."""
    cfg = cases.SMALL[name]
    w = cases.make(name, **cfg)
    noise = golden["%s_noise" % name]
    vals = w["values"][:8]
    if name == "dpd":
        vals = cases.dpd_rows(w, vals)
    acc = noise.copy()
    oracle.score_rows(cases.MODEL_ID[name], cases.oracle_caches(oracle, w), vals, acc)
    want = golden["%s_accum" % name]
    assert np.all(np.abs(acc - want) <= 3e-6 * (1 + np.abs(want)) + _envelope(name, w, oracle)[None, :])
    row0 = acc[0] - noise[0]
    for key in ("%s_group_scores", "%s_mixture_group_scores"):
        assert np.all(relerr(row0, golden[key % name]) <= 1e-3)


@pytest.mark.parametrize("name", ["nich", "gp", "bb"])
def test_caches_golden(oracle, golden, name):
    w = cases.make(name, **cases.SMALL[name])
    got = cases.oracle_caches(oracle, w)
    want = golden["%s_caches" % name]
    assert np.all(relerr(got, want) <= 2e-6)


def test_crosscat_golden(oracle, golden):
    G, N = 17, 64
    cc = synth.crosscat(201, G, N, n_gp=3, n_bb=3)
    extra = []
    for f in range(2):
        w = synth.nich(300 + f, G, N)
        w["count"] = cc["sizes"].copy()
        w["mean"][cc["sizes"] == 0] = 0
        w["ctv"][cc["sizes"] == 0] = 0
        extra.append(w)
    feats = cc["features"] + extra
    prior = oracle.py_prior(synth.PY_ALPHA, synth.PY_D, cc["sizes"])
    got = cases.oracle_scores(oracle, feats, prior=prior)
    want = golden["crosscat_scores"]
    env = sum(_envelope("nich", w, oracle) for w in extra)
    assert np.all(np.abs(got - want) <= 3e-6 * (1 + np.abs(want)) + env[None, :])


# -------------------------------------------------------------------------------------- sampler
@pytest.mark.parametrize("G", [1, 2, 5, 100, 1024])
def test_sampler_golden(oracle, golden, G):
    """Same scores, same uniforms -> same indices, except documented near-ties: rows where u*total
    lies within eps*total of a CDF boundary (eps covers fast_exp's 5e-6 build envelope)."""
    s = golden["sampler_%d_scores" % G]
    u = golden["sampler_%d_u" % G]
    want = golden["sampler_%d_assign" % G]
    got = oracle.sample_rows(s.copy(), u)
    ok = cases.explained_mismatch(s.astype(np.float64), u, got, want, eps=1e-5)
    assert ok.all()
    assert np.mean(got == want) >= 0.97


@pytest.mark.parametrize("name", list(cases.SMALL))
def test_score_sample_golden(oracle, golden, name):
    s = golden["%s_scores" % name]
    u = golden["%s_u" % name]
    want = golden["%s_assign" % name]
    got = oracle.sample_rows(s.copy(), u)
    assert cases.explained_mismatch(s.astype(np.float64), u, got, want, eps=1e-5).all()


def test_sampler_live_large(oracle, ref):
    rng = np.random.default_rng(5)
    s = (rng.standard_normal((20000, 128)) * 3).astype(np.float32)
    lik = s.copy()
    u, want = ref.sample_rows(11, lik)
    got = oracle.sample_rows(s.copy(), u)
    assert cases.explained_mismatch(s.astype(np.float64), u, got, want, eps=1e-5).all()
    assert np.mean(got != want) < 1e-3


# ---------------------------------------------------------------------------- group bookkeeping
def test_group_updates_golden(oracle, golden):
    vals = golden["nich_group_values"]
    trace = golden["nich_group_trace"]
    st = (0, 0.0, 0.0)
    got = []
    for v in vals:
        st = oracle.nich_group_update(+1, *st, [v]); got.append(st)
    for v in vals[:20]:
        st = oracle.nich_group_update(-1, *st, [v]); got.append(st)
    got = np.array(got)
    assert np.array_equal(got[:, 0], trace[:, 0])
    np.testing.assert_allclose(got[:, 1:], trace[:, 1:], rtol=2e-5, atol=2e-5)
    vals = golden["gp_group_values"]
    trace = golden["gp_group_trace"]
    st = (0, 0, 0.0)
    got = []
    for v in vals:
        st = oracle.gp_group_update(+1, *st, [v]); got.append(st)
    for v in vals[:20]:
        st = oracle.gp_group_update(-1, *st, [v]); got.append(st)
    got = np.array(got)
    assert np.array_equal(got[:, :2], trace[:, :2])
    np.testing.assert_allclose(got[:, 2], trace[:, 2], rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------ NIW
def _niw_numpy64(w, n):
    """float64 exact-math restatement of dbg/models/niw.py:103-123,192-200 + dbg/random.py:113-131."""
    from scipy.special import gammaln
    d = w["mu"].size
    G = w["sizes"].size
    out = np.zeros((n, G))
    mu0, kappa0, psi0, nu0 = w["mu"].astype(np.float64), w["kappa"], w["psi"].astype(np.float64), w["nu"]
    for g in range(G):
        cnt = float(w["count"][g])
        sx = w["sum_x"][g].astype(np.float64)
        sxx = w["sum_xxT"][g].astype(np.float64)
        xbar = sx / cnt if cnt else np.zeros(d)
        mu_n = kappa0 / (kappa0 + cnt) * mu0 + cnt / (kappa0 + cnt) * xbar
        kappa_n, nu_n = kappa0 + cnt, nu0 + cnt
        diff = xbar - mu0
        C = sxx - np.outer(sx, xbar) - np.outer(xbar, sx) + cnt * np.outer(xbar, xbar)
        psi_n = psi0 + C + kappa0 * cnt / (kappa0 + cnt) * np.outer(diff, diff)
        dof = nu_n - d + 1.0
        sigma = psi_n * (kappa_n + 1.0) / (kappa_n * dof)
        sign, logdet = np.linalg.slogdet(sigma)
        sinv = np.linalg.inv(sigma)
        x = w["values"][:n].astype(np.float64) - mu_n
        q = np.einsum("ni,ij,nj->n", x, sinv, x)
        out[:, g] = (gammaln(dof / 2 + d / 2) - gammaln(dof / 2) - 0.5 * logdet - d / 2 * (np.log(dof) + np.log(np.pi))
                     - 0.5 * (dof + d) * np.log1p(q / dof))
    return out


@pytest.mark.parametrize("d", [2, 3, 8, 32])
def test_niw_vs_exact_math(oracle, d):
    """NIW restatement vs exact float64 math at the reference's cross-flavour tolerance
    (test_model_flavors.py:56-116: 1e-3 relative to 1+|a|+|b|)."""
    w = synth.niw(400 + d, 9, 40, d=d)
    got = oracle.niw_score_rows(w["mu"], w["kappa"], w["psi"], w["nu"], w["count"], w["sum_x"], w["sum_xxT"],
                                w["values"][:40], np.zeros((40, 9), np.float32))
    want = _niw_numpy64(w, 40)
    assert np.all(relerr(got, want) <= 1e-3)


def test_niw_1d_equals_nich(oracle, ref):
    """NIW(d=1) score_value == NICH with sigmasq = psi/nu (reference test_normal_models.py:34-100),
    here against the REFERENCE's nich output."""
    rng = np.random.default_rng(3)
    G, N = 7, 50
    sizes = np.array([5, 9, 1, 30, 2, 12, 0], np.int32)
    mu, kappa, psi, nu = 0.3, 2.0, 1.7, 3.0
    xs = [rng.normal(rng.normal(0, 2), 1.0, s).astype(np.float32) for s in sizes]
    sum_x = np.array([x.astype(np.float64).sum() for x in xs], np.float32).reshape(G, 1)
    sum_xx = np.array([(x.astype(np.float64) ** 2).sum() for x in xs], np.float32).reshape(G, 1, 1)
    mean = np.array([x.mean(dtype=np.float64) if len(x) else 0 for x in xs], np.float32)
    ctv = np.array([((x.astype(np.float64) - x.mean(dtype=np.float64)) ** 2).sum() if len(x) else 0 for x in xs], np.float32)
    vals = (2 * rng.standard_normal(N)).astype(np.float32)
    k = ref.kind(G, None)
    k.add_nich([mu, kappa, psi / nu, nu], sizes, mean, ctv)
    want = k.score_rows([vals], N, with_prior=False)
    got = oracle.niw_score_rows([mu], kappa, [[psi]], nu, sizes, sum_x, sum_xx, vals.reshape(N, 1),
                                np.zeros((N, G), np.float32))
    # nich goes through fast_lgamma_nu (cubic fit, abs error up to 1.5e-2 for small nu), niw through
    # two fast_lgamma calls: they agree to the approximation error of the former, not to 1e-3.
    assert np.all(relerr(got, want) <= 5e-3)


def test_niw_1d_reference_kat(oracle, ref):
    """The reference's own KAT (test_normal_models.py:34-84): mu=30, kappa=0.3, psi=2, nu=3, data
    {4, 54, 3, -12, 7, 10} minus 54 and -12, score_value at 32 and -0.1, tolerance 1e-3."""
    data = np.array([4.0, 3.0, 7.0, 10.0])
    n = len(data)
    mean = np.float32(data.mean())
    ctv = np.float32(((data - data.mean()) ** 2).sum())
    vals = np.array([32.0, -0.1], np.float32)
    k = ref.kind(1, None)
    k.add_nich([30.0, 0.3, 2.0 / 3.0, 3.0], [n], [mean], [ctv])
    want = k.score_rows([vals], 2, with_prior=False)
    got = oracle.niw_score_rows([30.0], 0.3, [[2.0]], 3.0, [n], np.float32([[data.sum()]]),
                                np.float32([[[(data ** 2).sum()]]]), vals.reshape(2, 1), np.zeros((2, 1), np.float32))
    assert np.all(relerr(got, want) <= 1e-3)


# ------------------------------------------------------------------------------------------------
# score_data: the restatement against the compiled reference.  Both sum fp32 terms in
# group order; the -ffast-math reference build may re-associate (dd sums a vector, dpd iterates a hash
# map), so the comparison is relative to sum |term|.
@pytest.mark.parametrize("name,G", [("nich", 50), ("gp", 37), ("bb", 21), ("dd", 40), ("dpd", 12)])
def test_score_data_matches_reference(oracle, ref, name, G):
    kw = dict(dim=16) if name == "dd" else (dict(V=60, other_frac=0.05) if name == "dpd" else {})
    w = getattr(synth, name)(8100 + G, G, 10, **kw)
    k = ref.kind(G, w["sizes"], synth.PY_ALPHA, synth.PY_D)
    f = cases.ref_add_feature(k, w)
    grid = cases.shared_grid(w, 9, seed=G)
    per_point = k.score_data_grid(f, grid, use_grid=False)
    gridded = k.score_data_grid(f, grid, use_grid=True)
    for i in range(grid.shape[0]):
        got, scale, got64 = oracle.score_data(w, grid[i])
        tol = cases.accum_tol(cases.score_data_terms(w), scale)
        assert abs(got - per_point[i]) <= tol, (name, i, got, per_point[i], scale)
        assert abs(got64 - per_point[i]) <= tol, (name, i, got64, per_point[i], scale)
        # score_data_grid == score_data up to the accumulation order (dd's incremental _update)
        assert abs(got - gridded[i]) <= 4 * tol, (name, i, got, gridded[i], scale)


def test_score_data_skips_empty_groups_and_own_shared(oracle, ref):
    w = synth.nich(5, 9, 4)
    w["count"][2] = 0
    k = ref.kind(9, w["sizes"], synth.PY_ALPHA, synth.PY_D)
    f = cases.ref_add_feature(k, w)
    got, scale, _ = oracle.score_data(w)
    want = k.score_data_grid(f, np.asarray(w["shared"], np.float32)[None, :], use_grid=False)[0]
    assert abs(got - want) <= cases.accum_tol(cases.score_data_terms(w), scale)


def test_score_data_oracle_matches_golden(oracle, golden_score_data):
    """runs without the reference (GPU box / fresh checkout): the restatement against the committed outputs"""
    gd = golden_score_data
    for name, (seed, G, kw, n_grid) in cases.SCORE_DATA.items():
        w = getattr(synth, name)(seed, G, 10, **kw)
        grid = cases.shared_grid(w, n_grid, seed=seed)
        assert np.array_equal(grid, gd["sd_%s_grid" % name])
        for i in range(n_grid):
            got, scale, got64 = oracle.score_data(w, grid[i])
            tol = cases.accum_tol(cases.score_data_terms(w), scale)
            assert abs(got - gd["sd_%s_out" % name][i]) <= tol, (name, i)
            assert abs(got64 - gd["sd_%s_out" % name][i]) <= tol, (name, i)


# ------------------------------------------------------------------------------------------------
# LowEntropy clustering prior
def _low_entropy_cases():
    rng = np.random.default_rng(11)
    for dataset_size, G, empties in [(1000, 17, 1), (1000, 40, 3), (100000, 64, 1), (50, 6, 2)]:
        sizes = rng.integers(1, max(2, dataset_size // (2 * G)), G).astype(np.int32)
        sizes[rng.choice(G, empties, replace=False)] = 0
        yield dataset_size, sizes
    yield 2000000, np.array([20000, 5, 0, 1, 12345], np.int32)  # group_size > very_large branch


def test_low_entropy_prior_matches_reference(oracle, ref):
    for dataset_size, sizes in _low_entropy_cases():
        got = oracle.low_entropy_prior(dataset_size, sizes)
        want = ref.low_entropy_prior(dataset_size, sizes)
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-6 * (1 + np.abs(want).max()))


# ------------------------------------------------------------------------------------------------
# BetaNegativeBinomial: caches, score_value, score_data against the compiled reference
@pytest.mark.parametrize("G,r", [(29, 1), (64, 3), (7, 40)])
def test_bnb_matches_reference(oracle, ref, G, r):
    n = 300
    w = synth.bnb(7000 + G, G, n, r=r)
    k = ref.kind(G, w["sizes"], synth.PY_ALPHA, synth.PY_D)
    f = cases.ref_add_feature(k, w)
    cache = oracle.bnb_caches(w["shared"], w["count"], w["sum"])
    want_cache = k.scorer_caches(f)
    # fast_lgamma's polynomial is evaluated in double in both: caches agree to an ulp of the summands
    np.testing.assert_allclose(cache, want_cache, rtol=2e-6, atol=2e-6 * (1 + np.abs(want_cache[0]).max()))
    got = cases.oracle_scores(oracle, [w], prior=oracle.py_prior(synth.PY_ALPHA, synth.PY_D, w["sizes"]))
    want = k.score_rows([w["values"]], n)
    # score[g] + lgamma(beta) - lgamma(beta + alpha[g]) cancels: envelope 6e-7 * (1 + |score[g]| + |lgamma terms|)
    env = 6e-7 * (1 + np.abs(want_cache[0])[None, :] + 2 * np.abs(want).max())
    assert np.all(np.abs(got - want) <= 3e-6 * (1 + np.abs(want)) + env)
    grid = cases.shared_grid(w, 7, seed=G)
    per_point = k.score_data_grid(f, grid, use_grid=False)
    for i in range(grid.shape[0]):
        sd, scale, sd64 = oracle.score_data(w, grid[i])
        assert abs(sd - per_point[i]) <= cases.accum_tol(cases.score_data_terms(w), scale)


def test_rank3_oracle_matches_golden(oracle, golden_rank3):
    """bnb + LowEntropy restatements against committed reference outputs (no reference needed)"""
    gd = golden_rank3
    for key, (seed, G, N, rr) in cases.BNB_GOLDEN.items():
        w = synth.bnb(seed, G, N, r=rr)
        cache = oracle.bnb_caches(w["shared"], w["count"], w["sum"])
        want_cache = gd[key + "_caches"]
        np.testing.assert_allclose(cache, want_cache, rtol=2e-6, atol=2e-6 * (1 + np.abs(want_cache[0]).max()))
        got = cases.oracle_scores(oracle, [w], prior=oracle.py_prior(synth.PY_ALPHA, synth.PY_D, w["sizes"]))
        want = gd[key + "_scores"]
        env = 6e-7 * (1 + np.abs(want_cache[0])[None, :] + 2 * np.abs(want).max())
        assert np.all(np.abs(got - want) <= 3e-6 * (1 + np.abs(want)) + env)
        a = oracle.sample_rows(want.copy(), gd[key + "_u"])
        assert np.array_equal(a, gd[key + "_assign"])
        # Mixture::score_value_group == Group::score_value up to the cached / uncached evaluation order
        np.testing.assert_allclose(gd[key + "_group_scores"], gd[key + "_mixture_group_scores"], rtol=0, atol=2e-3)
        for i, sh in enumerate(gd[key + "_sd_grid"]):
            sd, scale, _ = oracle.score_data(w, sh)
            assert abs(sd - gd[key + "_sd_out"][i]) <= cases.accum_tol(cases.score_data_terms(w), scale)
    for i, (dataset_size, sizes) in enumerate(_low_entropy_cases()):
        assert int(gd["le_%d_dataset_size" % i][0]) == dataset_size and np.array_equal(gd["le_%d_sizes" % i], sizes)
        want = gd["le_%d_out" % i]
        np.testing.assert_allclose(oracle.low_entropy_prior(dataset_size, sizes), want, rtol=0, atol=2e-6 * (1 + np.abs(want).max()))


# ------------------------------------------------------------------------------------------------
# NIW pinned to the reference's own exact-math Python flavour (imported by tests/golden/make_golden_niw.py)
@pytest.fixture(scope="module")
def golden_niw():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "niw_golden.npz"))


@pytest.mark.parametrize("name", cases.NIW_GOLDEN_CASES)
def test_niw_oracle_matches_reference_python(oracle, golden_niw, name):
    c = cases.niw_golden_case(golden_niw, name)

    def score(c, values):
        n = values.shape[0]
        return oracle.niw_score_rows(c["mu"], c["kappa"], c["psi"], c["nu"], c["count"], c["sum_x"], c["sum_xxT"],
                                     np.ascontiguousarray(values, np.float32), np.zeros((n, c["count"].size), np.float32))

    cases.check_niw_golden(score, oracle, c)


@pytest.mark.parametrize("name", cases.NIW_GOLDEN_CASES)
def test_niw_score_data_oracle_matches_reference_python(oracle, golden_niw, name):
    """Group.score_data of the reference's exact-math Python (dbg/models/niw.py:202-217) summed over the groups vs the
    C restatement of niw.hpp:296-308: at the reference's cross-flavour bar, and tightly once the KNOWN deviations of
    fast_lgamma are taken out"""
    c = cases.niw_golden_case(golden_niw, name)
    got = oracle.niw_score_data(c["mu"], c["kappa"], c["psi"], c["nu"], c["count"], c["sum_x"], c["sum_xxT"])
    want = float(np.sum(c["score_data"]))
    assert abs(got - want) <= 1e-3 * (1 + abs(got) + abs(want))
    corr, tol = cases.niw_score_data_tolerance(oracle, c)
    assert abs(got - (want + corr)) <= tol + 2e-6 * abs(want), (got, want, corr, tol)


def test_niw_group_update_oracle_matches_reference_python(oracle, golden_niw):
    """Group.add_value one value at a time (niw.hpp:247-255) reproduces the statistics the reference's Python accumulated"""
    c = cases.niw_golden_case(golden_niw, "ex1")
    d = c["mu"].size
    vals = c["values"]
    cnt, sx, sxx = oracle.niw_group_update(+1, 0, np.zeros(d, np.float32), np.zeros((d, d), np.float32), vals)
    assert cnt == len(vals)
    np.testing.assert_allclose(sx, vals.astype(np.float64).sum(0), rtol=1e-5)
    np.testing.assert_allclose(sxx, vals.astype(np.float64).T @ vals.astype(np.float64), rtol=1e-5, atol=1e-4)
    cnt, sx, sxx = oracle.niw_group_update(-1, cnt, sx, sxx, vals)
    assert cnt == 0 and np.abs(sx).max() < 1e-3 and np.abs(sxx).max() < 1e-2


def test_niw_golden_covers_reference_examples(golden_niw):
    """the fixtures include the reference's EXAMPLES verbatim (dbg/models/niw.py:39-102 = lp/models/niw.pyx EXAMPLES)"""
    assert golden_niw["ex0_values"].shape == (7, 2) and golden_niw["ex1_values"].shape == (9, 3) and golden_niw["ex2_values"].shape == (9, 4)
    np.testing.assert_array_equal(golden_niw["ex0_values"][0], [1.0, 2.0])
    assert float(golden_niw["ex1_kappa"]) == 7.5 and float(golden_niw["ex2_nu"]) == 10.0
    assert golden_niw["d32_values"].shape == (96, 32) and golden_niw["d32_count"][-1] == 0
