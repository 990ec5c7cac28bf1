"""Clustering priors on the device against the oracle (pinned to the compiled reference in
tests/test_oracle.py): the LowEntropy vector.  The device evaluates the reference's
fp32 expressions with the literal fast_log table: bit-exact against the oracle."""
import numpy as np
import pytest

from test_oracle import _low_entropy_cases

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


def test_low_entropy_prior(ctx, oracle):
    for dataset_size, sizes in _low_entropy_cases():
        want = oracle.low_entropy_prior(dataset_size, sizes)
        got = ctx.prior_low_entropy_host(dataset_size, sizes)
        assert np.array_equal(got, want), (dataset_size, np.abs(got - want).max())
        sd = torch.from_numpy(sizes).cuda()
        out = torch.empty(sizes.size, device="cuda")
        ctx.prior_low_entropy_dev(dataset_size, sizes.size, sd, out)
        assert np.array_equal(out.cpu().numpy(), want)


def test_low_entropy_prior_drives_sampling(ctx, oracle):
    """the vector is a drop-in for the `prior` argument of the score entries"""
    from distributions_b200 import capi, synth
    import cases
    G, n = 23, 2000
    w = synth.nich(4, G, n)
    prior = oracle.low_entropy_prior(100000, w["sizes"])
    f = ctx.feature(capi.NICH).update_all(w)
    scores = torch.empty((n, G), device="cuda")
    assign = torch.empty(n, device="cuda", dtype=torch.int32)
    ctx.score_sample_batch([f], [torch.from_numpy(w["values"]).cuda()], n, torch.from_numpy(prior).cuda(),
                           torch.from_numpy(w["u"]).cuda(), assign, scores)
    exp = cases.oracle_scores(oracle, [w], prior=prior)
    coeff = np.abs(oracle.nich_caches(w["shared"], w["count"], w["mean"], w["ctv"])[1])[None, :]
    assert np.all(np.abs(scores.cpu().numpy() - exp) <= 3e-6 * (1 + np.abs(exp)) + 1e-6 * coeff)
