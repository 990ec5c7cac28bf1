"""Shared helpers for the parity tests: build the reference-side / oracle-side objects from a
distributions_b200.synth workload dict.  Test infrastructure (uses oracle/)."""
import numpy as np

from distributions_b200 import synth
from oracle.pyoracle import BB, BNB, DD, DPD, GP, NICH

MODEL_ID = {"dd": DD, "dpd": DPD, "bb": BB, "gp": GP, "nich": NICH, "bnb": BNB}

# The small configurations stored in tests/golden (seed, G, N).  G values are deliberately ragged
# (not multiples of 32) and include G=1 (a single, empty group).
SMALL = {
    "nich": dict(seed=101, G=37, N=96),
    "gp": dict(seed=102, G=21, N=96),
    "bb": dict(seed=103, G=13, N=96),
    "dd": dict(seed=104, G=29, N=96, dim=16),
    "dpd": dict(seed=105, G=19, N=96, V=100, other_frac=0.1),
}


def make(model, **kw):
    kw = dict(kw)
    return getattr(synth, model)(kw.pop("seed"), kw.pop("G"), kw.pop("N"), **kw)


def ref_add_feature(kind, w):
    """Add workload w as a feature of a reference RefKind."""
    m = w["model"]
    if m == "nich":
        return kind.add_nich(w["shared"], w["count"], w["mean"], w["ctv"])
    if m == "gp":
        return kind.add_gp(w["shared"], w["count"], w["sum"], w.get("log_prod"))
    if m == "bb":
        return kind.add_bb(w["shared"], w["heads"], w["tails"])
    if m == "bnb":
        return kind.add_bnb(w["shared"], w["count"], w["sum"])
    if m == "dd":
        return kind.add_dd(w["alphas"], w["counts"])
    if m == "dpd":
        return kind.add_dpd(w["gamma"], w["alpha"], w["beta0"], w["keys"], w["betas"], w["counts"])
    raise ValueError(m)


def oracle_caches(o, w):
    m = w["model"]
    if m == "nich":
        return o.nich_caches(w["shared"], w["count"], w["mean"], w["ctv"])
    if m == "gp":
        return o.gp_caches(w["shared"], w["count"], w["sum"])
    if m == "bb":
        return o.bb_caches(w["shared"], w["heads"], w["tails"])
    if m == "bnb":
        return o.bnb_caches(w["shared"], w["count"], w["sum"])
    if m == "dd":
        return o.dd_caches(w["alphas"], w["counts"])
    if m == "dpd":
        return o.dpd_caches(w["alpha"], w["beta0"], w["betas"], w["counts"])
    raise ValueError(m)


def dpd_rows(w, values=None):
    """dense table row of each dpd value: index into keys, or V for OTHER / unknown."""
    values = w["values"] if values is None else values
    keys = w["keys"]
    order = np.argsort(keys)
    pos = np.searchsorted(keys[order], values)
    pos = np.clip(pos, 0, keys.size - 1)
    hit = keys[order][pos] == values
    return np.where(hit, order[pos], keys.size).astype(np.uint32)


def oracle_scores(o, feats, n=None, prior=None):
    """prior (or zeros) + sum over feature workloads, via the C restatement."""
    G = feats[0]["sizes"].size
    n = feats[0]["values"].shape[0] if n is None else n
    scores = np.zeros((n, G), np.float32)
    if prior is not None:
        scores += prior[None, :]
    for w in feats:
        vals = w["values"][:n]
        if w["model"] == "dpd":
            vals = dpd_rows(w, vals)
        o.score_rows(MODEL_ID[w["model"]], oracle_caches(o, w), vals, scores)
    return scores


def explained_mismatch(scores64, u, a, b, eps):
    """True where indices a != b are explained by u*total lying within eps*total of a CDF
    boundary between them (a 'near-tie', SURVEY.md §7 hard parts)."""
    s = scores64 - scores64.max(axis=1, keepdims=True)
    lik = np.exp(s)
    cdf = np.cumsum(lik, axis=1)
    total = cdf[:, -1]
    t = u.astype(np.float64) * total
    lo = np.minimum(a, b)
    hi = np.maximum(a, b)
    ok = np.ones(len(u), bool)
    for i in np.nonzero(a != b)[0]:
        # every boundary cdf[lo..hi-1] must be within eps*total of t
        seg = cdf[i, lo[i]:hi[i]]
        ok[i] = np.all(np.abs(seg - t[i]) <= eps * total[i])
    return ok


# BetaNegativeBinomial golden cases: key -> (seed, G, N, r); tests/golden/make_golden_rank3.py
BNB_GOLDEN = {"bnb_a": (9301, 23, 96, 1), "bnb_b": (9302, 70, 64, 4)}

# protobuf wire fixtures: model -> (seed, G, synth kwargs); tests/golden/make_golden_wire.py
WIRE = {
    "nich": (9401, 9, {}),
    "gp": (9402, 11, {}),
    "bnb": (9403, 7, dict(r=3)),
    "bb": (9404, 10, {}),
    "dd": (9405, 8, dict(dim=16)),
    "dpd": (9406, 6, dict(V=40, other_frac=0.05)),
    "niw": (9407, 5, dict(d=4)),
}

# score_data golden cases: model -> (seed, G, synth kwargs, grid points); tests/golden/make_golden_score_data.py
SCORE_DATA = {
    "nich": (9101, 70, {}, 12),
    "gp": (9102, 45, {}, 12),
    "bb": (9103, 33, {}, 12),
    "dd": (9104, 52, dict(dim=16), 12),
    "dpd": (9105, 14, dict(V=80, other_frac=0.05), 8),
}


def shared_grid(w, n_grid, seed=0):
    """n_grid hyper-parameter settings around workload w's Shared, packed as score_data_grid takes them:
    nich (mu, kappa, sigmasq, nu); gp (alpha, inv_beta); bb (alpha, beta); dd alphas[dim]; dpd (alpha).
    Consecutive dd settings differ in a few alphas only (the reference's incremental _update path)."""
    rng = np.random.default_rng(seed)
    m = w["model"]
    if m == "nich":
        g = np.tile(np.asarray(w["shared"], np.float32), (n_grid, 1))
        g[:, 0] += rng.normal(0, 1, n_grid)
        g[:, 1] *= rng.uniform(0.3, 3, n_grid)
        g[:, 2] *= rng.uniform(0.3, 3, n_grid)
        g[:, 3] *= rng.uniform(0.3, 3, n_grid)
    elif m in ("gp", "bb"):
        g = np.tile(np.asarray(w["shared"], np.float32), (n_grid, 1)) * rng.uniform(0.3, 3, (n_grid, 2))
    elif m == "bnb":  # (alpha, beta); r stays the feature's
        g = np.tile(np.asarray(w["shared"][:2], np.float32), (n_grid, 1)) * rng.uniform(0.3, 3, (n_grid, 2))
    elif m == "dd":
        g = np.tile(np.asarray(w["alphas"], np.float32), (n_grid, 1))
        for i in range(1, n_grid):
            g[i] = g[i - 1]
            idx = rng.integers(0, g.shape[1], 2)
            g[i, idx] = rng.uniform(0.1, 2.0, 2)
    elif m == "dpd":
        g = (w["alpha"] * rng.uniform(0.3, 3, (n_grid, 1)))
    else:
        raise ValueError(m)
    return np.ascontiguousarray(g, dtype=np.float32)


def score_data_terms(w):
    """number of fp32 terms MixtureDataScorer::score_data adds for workload w"""
    m = w["model"]
    G = w["sizes"].size
    if m == "nich":
        return 4 * G
    if m == "gp":
        return 3 * G
    if m == "bnb":
        return 2 * G
    if m == "bb":
        return G
    return (w["counts"].shape[1] + 1) * G


def accum_tol(n_terms, scale):
    """rounding of an fp32 accumulation of n_terms terms with sum |term| = scale, in whatever order:
    2 eps32 * sqrt(n_terms) * scale (a random-walk bound; the worst case is n_terms * eps32 * scale)"""
    return 2.4e-7 * np.sqrt(n_terms) * scale + 1e-5


# ---- NIW against the reference's own exact-math Python (tests/golden/make_golden_niw.py) --------------------
NIW_GOLDEN_CASES = ("ex0", "ex1", "ex2", "d32", "d2b", "d3b")
LOG_STEP = 6.2e-5  # fast_log's table step in the log (special.hpp:57-67)


def niw_golden_case(gd, name):
    """one fixture as the float32 arguments of the C-ABI / oracle + the float64 reference outputs"""
    g = lambda k: gd["%s_%s" % (name, k)]  # noqa: E731
    return dict(mu=g("mu").astype(np.float32), kappa=float(g("kappa")), psi=g("psi").astype(np.float32), nu=float(g("nu")),
                count=g("count").astype(np.int32), sum_x=g("sum_x").astype(np.float32), sum_xxT=g("sum_xxT").astype(np.float32),
                values=g("values").astype(np.float32), scores=g("scores"), post_mu=g("post_mu"), const_scores=g("const_scores"),
                score_data=g("score_data"))


def niw_score_data_tolerance(o, c):
    """|implementation - reference python| allowed for sum_g Group.score_data: the implementation evaluates the
    reference C++ expression (niw.hpp:296-308) with fast_lgamma / fast_log, whose deviations from lgamma / log are
    known from the pinned oracle functions -- so return (correction, tolerance): want + correction is what an
    implementation with exact linear algebra would give, up to fast_log(det)'s table step and float rounding"""
    from scipy.special import gammaln
    d = c["mu"].size
    fl = lambda x: np.asarray(o.fast_lgamma(np.asarray(x, np.float32)), np.float64)  # noqa: E731
    corr, tol = 0.0, 0.0
    for n in c["count"].astype(np.float64):
        for a in (0.5 * (c["nu"] + n), 0.5 * c["nu"]):
            j = np.arange(1, d + 1)
            arg = a + 0.5 * (1 - j)
            dev = float(np.sum(fl(arg) - gammaln(arg)))
            corr += dev if a != 0.5 * c["nu"] else -dev
            tol += 3e-6 * float(np.sum(np.abs(gammaln(arg)))) + 1e-6 * d
        # three fast_log terms, coefficients nu / 2, post.nu / 2, d / 2: a table step each (biased low, not corrected)
        tol += LOG_STEP * (0.5 * c["nu"] + 0.5 * (c["nu"] + n) + 0.5 * d)
        tol += 1e-5 * (0.5 * (c["nu"] + n))  # det of a float32 d x d posterior: relative 1e-5 in the log
    return corr, tol


def check_niw_golden(score_fn, o, c):
    """score_fn(case, values[n][d] float32) -> scores [n][G] of the implementation under test (prior = 0).
    `o` supplies the pinned fast_lgamma / fast_log (their deviation from lgamma / log is KNOWN, so the checks
    below are tight on everything the golden pins: posterior, Sigma^-1, det Sigma).

      1. absolute, at the reference's own cross-flavour bar 1e-3 * (1 + |a| + |b|) (test_model_flavors.py:56-116)
      2. score at the posterior mean (quadratic form 0) == exact constant corrected by the known
         fast_lgamma / fast_log(dof) deviations; what is left is -0.5 * (fast_log(det) - log det): one table step
      3. differences between rows within a group cancel the constant: they pin (x - mu')^T Sigma^-1 (x - mu')
         through coeff * [log(1 + q1/dof) - log(1 + q2/dof)] to two table steps
    """
    from scipy.special import gammaln
    d = c["mu"].size
    want = c["scores"]
    got = np.asarray(score_fn(c, c["values"]), np.float64)
    err = np.abs(got - want) / (1 + np.abs(got) + np.abs(want))
    assert np.all(err <= 1e-3), ("absolute", err.max())
    dof = (c["nu"] + c["count"] - d + 1.0).astype(np.float64)
    a, b = 0.5 * (dof + d), 0.5 * dof
    fl = lambda x: np.asarray(o.fast_lgamma(np.asarray(x, np.float32)), np.float64)  # noqa: E731
    corr = (fl(a) - gammaln(a)) - (fl(b) - gammaln(b)) - 0.5 * d * (np.asarray(o.fast_log(dof.astype(np.float32)), np.float64) - np.log(dof))
    got_c = np.asarray(score_fn(c, c["post_mu"].astype(np.float32)), np.float64)
    got_c = np.diagonal(got_c)  # row g = group g's own posterior mean
    # float32 posterior mean: q ~ |delta|^2 / sigma, negligible; det: float32 LU / Cholesky of a d x d matrix
    tol_c = 0.5 * LOG_STEP + 2e-6 * d + 3e-6 * (np.abs(c["const_scores"]) + np.abs(gammaln(a)) + np.abs(gammaln(b)))
    assert np.all(np.abs(got_c - (c["const_scores"] + corr)) <= tol_c), ("constant", np.abs(got_c - (c["const_scores"] + corr)).max())
    coeff = 0.5 * (dof + d)
    dg, dw = got - got[:1], want - want[:1]
    # the quadratic form is evaluated in float32 from float32 statistics: relative 1e-5 of |coeff * log-term| each
    tol_d = coeff[None, :] * 2 * LOG_STEP + 2e-5 * (np.abs(want + 0 * dw) * 0 + np.abs(dw)) + 4e-5 * coeff[None, :]
    assert np.all(np.abs(dg - dw) <= tol_d), ("differences", (np.abs(dg - dw) - tol_d).max())
