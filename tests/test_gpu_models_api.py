"""The Python-3 mirror of the reference's lp Mixture classes (distributions_b200/models.py) driven through
the reference's own test choreography (distributions/tests/test_models.py:498-594 test_mixture_runs /
test_mixture_score, doc/overview.rst:185-202) and checked against the oracle."""
import zlib

import numpy as np
import pytest

import cases
from distributions_b200 import models, synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
LOG_STEP = 6.2e-5


@pytest.fixture(scope="module")
def ctx():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from distributions_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _value(rng, name):
    if name == "nich":
        return float(np.float32(rng.normal(0, 3)))
    if name in ("gp", "bnb"):
        return int(rng.poisson(6))
    if name == "bb":
        return bool(rng.random() < 0.4)
    if name == "dpd":
        return int(rng.integers(0, 100))
    if name == "niw":
        return (rng.normal(0, 2, 2) + np.array([1.0, -0.5])).astype(np.float32)
    return int(rng.integers(0, 16))


def _workload(name, shared, groups):
    """oracle-side description of the mirror's current groups"""
    sizes = np.ones(len(groups), np.int32)
    if name == "nich":
        return dict(model=name, sizes=sizes, shared=np.array([shared.mu, shared.kappa, shared.sigmasq, shared.nu], np.float32),
                    count=np.array([g.count for g in groups], np.int32), mean=np.array([g.mean for g in groups], np.float32),
                    ctv=np.array([g.count_times_variance for g in groups], np.float32))
    if name == "gp":
        return dict(model=name, sizes=sizes, shared=np.array([shared.alpha, shared.inv_beta], np.float32),
                    count=np.array([g.count for g in groups], np.uint32), sum=np.array([g.sum for g in groups], np.uint32),
                    log_prod=np.zeros(len(groups), np.float32))
    if name == "bnb":
        return dict(model=name, sizes=sizes, shared=np.array([shared.alpha, shared.beta, shared.r], np.float32),
                    count=np.array([g.count for g in groups], np.uint32), sum=np.array([g.sum for g in groups], np.uint32))
    if name == "bb":
        return dict(model=name, sizes=sizes, shared=np.array([shared.alpha, shared.beta], np.float32),
                    heads=np.array([g.heads for g in groups], np.int32), tails=np.array([g.tails for g in groups], np.int32))
    if name == "dpd":
        return dict(model=name, sizes=sizes, gamma=shared.gamma, alpha=shared.alpha, beta0=shared.beta0, keys=shared.values,
                    betas=shared.betas, counts=np.array([g.dense(shared) for g in groups], np.int32).reshape(len(groups), -1))
    if name == "niw":
        G, d = len(groups), shared.dim
        return dict(model=name, sizes=sizes, mu=shared.mu, kappa=shared.kappa, psi=shared.psi, nu=shared.nu,
                    count=np.array([g.count for g in groups], np.int32),
                    sum_x=np.array([g.sum_x for g in groups], np.float32).reshape(G, d),
                    sum_xxT=np.array([g.sum_xxT for g in groups], np.float32).reshape(G, d, d))
    return dict(model=name, sizes=sizes, alphas=shared.alphas, counts=np.array([g.counts for g in groups], np.int32))


def _oracle_rows(oracle, name, w, values):
    """oracle scores [n][G] of the values against the workload's groups (no prior)"""
    n, G = len(values), w["sizes"].size
    if name == "niw":
        want = np.zeros((n, G), np.float32)
        oracle.niw_score_rows(w["mu"], w["kappa"], w["psi"], w["nu"], w["count"], w["sum_x"], w["sum_xxT"],
                              np.ascontiguousarray(np.array(values, np.float32).reshape(n, -1)), want)
        return want
    ww = dict(w)
    ww["values"] = np.array(values)
    return cases.oracle_scores(oracle, [ww])


def _tol(name, want):
    extra = 25 * LOG_STEP if name in ("nich", "niw") else (1e-4 if name in ("gp", "bnb") else 0.0)
    if name == "niw":  # float32 rank-1 statistics -> float32 Cholesky: the reference's own cross-flavour bar
        return 1e-3 * (1 + 2 * np.abs(want)) + extra
    return 4e-6 * (1 + np.abs(want)) + extra


@pytest.mark.parametrize("name", sorted(models.MODELS))
def test_mixture_choreography(ctx, oracle, name):
    model = models.MODELS[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)  # deterministic across processes (str hash is salted)
    shared = model.Shared(r=3) if name == "bnb" else model.Shared()
    mixture = model.Mixture(ctx)
    for _ in range(6):  # groups filled before init (benchmarks/mixture.cc:88-99)
        g = model.Group()
        g.init(shared)
        for _ in range(int(rng.integers(1, 9))):
            g.add_value(shared, _value(rng, name))
        mixture.append(g)
    mixture.init(shared)

    def check():
        v = _value(rng, name)
        w = _workload(name, shared, mixture.groups)
        want = _oracle_rows(oracle, name, w, [v])
        noise = rng.standard_normal(len(mixture)).astype(np.float32)
        scores = noise.copy()
        mixture.score_value(shared, v, scores)  # accumulates (test_models.py:552-557)
        assert np.all(np.abs((scores - noise) - want[0]) <= _tol(name, want[0]) + 4e-7 * np.abs(noise))
        gid = int(rng.integers(0, len(mixture)))
        assert abs(mixture.score_value_group(shared, gid, v) - want[0][gid]) <= _tol(name, want[0][gid:gid + 1])[0]

    check()
    history = [[] for _ in range(len(mixture))]
    for step in range(30):
        op = rng.random()
        if op < 0.5:
            gid = int(rng.integers(0, len(mixture)))
            v = _value(rng, name)
            mixture.add_value(shared, gid, v)
            history[gid].append(v)
        elif op < 0.7:
            cand = [i for i, h in enumerate(history) if h]
            if cand:
                gid = int(rng.choice(cand))
                mixture.remove_value(shared, gid, history[gid].pop())
        elif op < 0.85:
            mixture.add_group(shared)
            history.append([])
        elif len(mixture) > 3:
            gid = int(rng.integers(0, len(mixture)))
            mixture.remove_group(shared, gid)
            history[gid] = history[-1]
            history.pop()
        check()
    # batched entries: add_values == the same add_value calls one by one
    n = 80
    vals = [_value(rng, name) for _ in range(n)]
    gids = rng.integers(0, len(mixture), n).astype(np.int32)
    if True:
        expect = [model.Group().load(g.dump()) for g in mixture.groups]
        for g in expect:
            if name in ("dd", "dpd"):
                g.counts = g.counts.copy()
        for gid, v in zip(gids, vals):
            expect[gid].add_value(shared, v)
        mixture.add_values(shared, np.array(vals), gids)
        for got, exp in zip(mixture.groups, expect):
            for k, _ in got.FIELDS:
                if name == "dpd":
                    assert getattr(got, k) == getattr(exp, k), (name, k)
                    continue
                a, b = np.asarray(getattr(got, k), np.float64), np.asarray(getattr(exp, k), np.float64)
                assert np.allclose(a, b, rtol=1e-5, atol=1e-4), (name, k, a, b)
        check()
    if name == "niw":  # niw.hpp:296-308 over the device-resident statistics vs the C restatement
        w = _workload(name, shared, mixture.groups)
        want = oracle.niw_score_data(w["mu"], w["kappa"], w["psi"], w["nu"], w["count"], w["sum_x"], w["sum_xxT"])
        assert abs(mixture.score_data(shared) - want) <= 3e-6 * abs(want) + 1e-3
    elif name != "gp":  # gp's score_data needs Group::log_prod, which this mirror does not carry
        w = _workload(name, shared, mixture.groups)
        _, scale, want64 = oracle.score_data(w)
        assert abs(mixture.score_data(shared) - want64) <= 2e-7 * scale + 1e-5
    u = rng.random(n, dtype=np.float32)
    assign, scores = mixture.score_values(shared, np.array(vals), None, u, want_scores=True)
    want = _oracle_rows(oracle, name, _workload(name, shared, mixture.groups), vals)
    assert np.all(np.abs(scores - want) <= _tol(name, want))
    a_orc = oracle.sample_rows(scores.copy(), u)
    assert cases.explained_mismatch(scores.astype(np.float64), u, assign, a_orc, 2e-5).all()
