"""score_data_grid on the device against the oracle restatement of
MixtureDataScorer::score_data (pinned to the compiled reference in tests/test_oracle.py).

Every term is the reference's fp32 expression; the reference adds the terms in group order in fp32, the
device in double.  Two checks: (1) against the oracle's own terms summed in double the device agrees to
2e-7 * sum|term| (term-level parity: only lgammaf below 2.5 differs by an ulp between glibc and CUDA);
(2) against the reference's fp32 accumulation within that accumulation's rounding, cases.accum_tol =
2 eps32 * sqrt(n_terms) * sum|term|."""
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


def _ids():
    from distributions_b200 import capi
    return {"dd": capi.DD, "dpd": capi.DPD, "bb": capi.BB, "gp": capi.GP, "nich": capi.NICH}


@pytest.mark.parametrize("name,G", [("nich", 50), ("nich", 3000), ("gp", 37), ("bb", 21), ("bb", 2500), ("dd", 40), ("dpd", 12),
                                    ("dpd", 300)])
def test_score_data_grid_matches_oracle(ctx, oracle, name, G):
    kw = dict(dim=16) if name == "dd" else (dict(V=200, other_frac=0.05) if name == "dpd" else {})
    w = getattr(synth, name)(8100 + G, G, 10, **kw)
    if name == "nich":
        w["count"][G // 3] = 0  # an empty group contributes nothing (nich.hpp:275)
    f = ctx.feature(_ids()[name]).update_all(w)
    grid = cases.shared_grid(w, 17, seed=G)
    got = f.score_data_grid(grid)
    for i in range(grid.shape[0]):
        want32, scale, want64 = oracle.score_data(w, grid[i])
        assert abs(got[i] - want64) <= 2e-7 * scale + 1e-5, (name, i, got[i], want64, scale)
        # the reference's fp32 group-order accumulation drifts from the exact sum of its own terms (up to
        # eps32 * n_terms * scale for same-sign terms: 7.2 of 91285 at dpd G=300); the device stays as close
        # to it as the exact sum does
        assert abs(got[i] - want32) <= 1.5 * abs(want32 - want64) + 2e-7 * scale + 1e-5, (name, i, got[i], want32, want64)


def test_score_data_grid_golden_reference(ctx, golden_score_data):
    """committed outputs of the compiled reference's Mixture::score_data (tests/golden/make_golden_score_data.py)"""
    gd = golden_score_data
    for name, (seed, G, kw, n_grid) in cases.SCORE_DATA.items():
        w = getattr(synth, name)(seed, G, 10, **kw)
        f = ctx.feature(_ids()[name]).update_all(w)
        got = f.score_data_grid(gd["sd_%s_grid" % name])
        tol = cases.accum_tol(cases.score_data_terms(w), gd["sd_%s_scale" % name])
        assert np.all(np.abs(got - gd["sd_%s_out" % name]) <= tol), name
        # the reference's own score_data_grid (dd: incremental updates) stays within its accumulation envelope
        assert np.all(np.abs(got - gd["sd_%s_out_grid" % name]) <= 4 * tol), name


def test_score_data_after_add_rows(ctx, oracle):
    """the statistics score_data reads are the ones batched add_value maintains"""
    from distributions_b200 import capi
    G, n = 31, 4000
    w = synth.nich(77, G, n)
    assign = np.random.default_rng(1).integers(0, G, n).astype(np.int32)
    f = ctx.feature(capi.NICH).update_all(w)
    col = torch.from_numpy(w["values"]).cuda()
    f.add_rows(col, torch.from_numpy(assign).cuda(), n)
    raw = f.download_stats(12 * G)
    w2 = dict(w, count=raw[:4 * G].view(np.int32).copy(), mean=raw[4 * G:8 * G].view(np.float32).copy(),
              ctv=raw[8 * G:].view(np.float32).copy())
    grid = cases.shared_grid(w, 5, seed=3)
    got = f.score_data_grid(grid)
    for i in range(5):
        _, scale, want64 = oracle.score_data(w2, grid[i])
        assert abs(got[i] - want64) <= 2e-7 * scale + 1e-5


def test_score_data_gp_needs_log_prod(ctx):
    from distributions_b200 import capi
    w = synth.gp(3, 8, 10)
    lp = w.pop("log_prod")
    f = ctx.feature(capi.GP).update_all(w)
    with pytest.raises(RuntimeError):
        f.score_data_grid(np.array([[1.0, 1.0]], np.float32))
    f.set_log_prod(lp)
    assert np.isfinite(f.score_data_grid(np.array([[1.0, 1.0]], np.float32))[0])
    f.add_rows(torch.zeros(10, dtype=torch.int32).cuda().view(torch.int32), torch.zeros(10, dtype=torch.int32).cuda(), 10)
    with pytest.raises(RuntimeError):  # log_prod is stale after a batched update
        f.score_data_grid(np.array([[1.0, 1.0]], np.float32))
