"""Host-side Group bookkeeping of the Python mirror (distributions_b200/models.py) against the reference's
Group::add_value / remove_value as restated by the oracle (pinned to the compiled reference in
tests/test_oracle.py): the fp32 Welford updates of nich must match bit for bit, the integer models exactly.
No device needed."""
import numpy as np

from distributions_b200 import models


def test_nich_group_arithmetic_is_the_references(oracle):
    rng = np.random.default_rng(0)
    shared = models.nich.Shared()
    g = models.nich.Group()
    g.init(shared)
    state = (0, 0.0, 0.0)
    held = []
    for step in range(400):
        if held and rng.random() < 0.35:
            v = held.pop(int(rng.integers(0, len(held))))
            g.remove_value(shared, v)
            state = oracle.nich_group_update(-1, state[0], state[1], state[2], [v])
        else:
            v = float(np.float32(rng.normal(3.0, 10.0)))
            held.append(v)
            g.add_value(shared, v)
            state = oracle.nich_group_update(+1, state[0], state[1], state[2], [v])
        assert g.count == state[0]
        assert np.float32(g.mean).view(np.uint32) == np.float32(state[1]).view(np.uint32), (step, g.mean, state[1])
        assert np.float32(g.count_times_variance).view(np.uint32) == np.float32(state[2]).view(np.uint32), (step,)


def test_count_sum_groups_wrap_like_uint32(oracle):
    shared = models.gp.Shared()
    g = models.gp.Group()
    g.init(shared)
    state = (0, 0, 0.0)
    for v in [5, 0, 4000000000, 4000000000, 17]:
        g.add_value(shared, v)
        state = oracle.gp_group_update(+1, state[0], state[1], state[2], [v])
        assert (g.count, g.sum) == (state[0], state[1])
    for v in [4000000000, 5]:
        g.remove_value(shared, v)
        state = oracle.gp_group_update(-1, state[0], state[1], state[2], [v])
        assert (g.count, g.sum) == (state[0], state[1])


def test_bb_dd_groups_and_containers():
    sb = models.bb.Shared()
    g = models.bb.Group()
    g.init(sb)
    for v in (True, False, False, True, True):
        g.add_value(sb, v)
    g.remove_value(sb, True)
    assert (g.heads, g.tails) == (2, 2)
    sd = models.dd.Shared(dim=4)
    d = models.dd.Group()
    d.init(sd)
    for v in (0, 3, 3, 1):
        d.add_value(sd, v)
    d.remove_value(sd, 3)
    assert d.counts.tolist() == [1, 1, 0, 1]
    assert models.nich.Shared(mu=1.5).dump() == {"mu": 1.5, "kappa": 1.0, "sigmasq": 1.0, "nu": 1.0}
    assert set(models.MODELS) == {"nich", "gp", "bnb", "bb", "dd", "dpd", "niw"}
    # dpd: sparse counters over a fixed value set (dpd.hpp:157-215); niw: rank-1 statistics (niw.hpp:247-276)
    sp = models.dpd.Shared(values=[5, 9, 42], betas=[0.2, 0.3, 0.4])
    assert abs(sp.beta0 - 0.1) < 1e-6
    p = models.dpd.Group()
    p.init(sp)
    for v in (5, 42, 42, 9):
        p.add_value(sp, v)
    p.remove_value(sp, 9)
    assert p.counts == {5: 1, 42: 2} and p.dense(sp).tolist() == [1, 0, 2]
    import pytest
    with pytest.raises(AssertionError):
        p.add_value(sp, 7)  # unknown value, dpd.hpp:193
    sw = models.niw.Shared(dim=3)
    assert sw.nu == 4.0 and sw.psi.shape == (3, 3)
    w = models.niw.Group()
    w.init(sw)
    w.add_value(sw, [1.0, 2.0, 3.0])
    w.add_value(sw, [0.5, 0.0, -1.0])
    w.remove_value(sw, [1.0, 2.0, 3.0])
    assert w.count == 1 and w.sum_x.tolist() == [0.5, 0.0, -1.0] and w.sum_xxT[2, 2] == 1.0


def test_clustering_driver_bookkeeping():
    """MixtureDriver semantics (mixture.hpp:59-122): there is always an empty group, adding to it appends a
    fresh one, emptying a group swaps the last group into its place"""
    m = models.PitmanYor.Mixture(ctx=None)
    model = models.PitmanYor(alpha=1.0, d=0.1)
    m.init(model, [3, 0, 2])
    assert m.sample_size == 5 and m.empty_groupids == {1}
    assert m.add_value(model, 0) is False and m.counts == [4, 0, 2]
    assert m.add_value(model, 1) is True                      # the empty group got its first value ...
    assert m.counts == [4, 1, 2, 0] and m.empty_groupids == {3}  # ... so a fresh empty group is appended
    assert m.remove_value(model, 1) is True                   # emptied: the last group (the empty one) moves in
    assert m.counts == [4, 0, 2] and m.empty_groupids == {1} and m.sample_size == 6
    assert m.remove_value(model, 2, 2) is True                # emptied group is the last one: just dropped
    assert m.counts == [4, 0] and m.empty_groupids == {1}
    le = models.LowEntropy.Mixture(ctx=None)
    le.init(models.LowEntropy(100), [0, 7])
    assert le.add_value(models.LowEntropy(100), 0, 2) is True and le.counts == [2, 7, 0] and le.empty_groupids == {2}


def test_mixture_id_tracker():
    """MixtureIdTracker (mixture.hpp:460-521): packed ids move under remove_group's swap-with-last, global ids never do"""
    from distributions_b200.models import MixtureIdTracker
    t = MixtureIdTracker(4)
    assert [t.packed_to_global(i) for i in range(4)] == [0, 1, 2, 3]
    t.remove_group(1)  # the last group (global 3) moves into packed slot 1
    assert t.packed_size() == 3 and t.global_size() == 4
    assert [t.packed_to_global(i) for i in range(3)] == [0, 3, 2]
    assert t.global_to_packed(3) == 1 and t.global_to_packed(2) == 2
    t.add_group()
    assert t.packed_to_global(3) == 4 and t.global_size() == 5
    t.remove_group(3)  # removing the last packed id moves nothing
    assert [t.packed_to_global(i) for i in range(3)] == [0, 3, 2]
    import pytest
    with pytest.raises(AssertionError):
        t.global_to_packed(1)  # stale global id
    rng = __import__("numpy").random.default_rng(3)
    live = {t.packed_to_global(i) for i in range(t.packed_size())}
    for _ in range(200):  # random churn keeps the two maps inverse of each other
        if rng.random() < 0.5 or t.packed_size() < 2:
            t.add_group()
            live.add(t.global_size() - 1)
        else:
            p = int(rng.integers(0, t.packed_size()))
            live.discard(t.packed_to_global(p))
            t.remove_group(p)
        assert {t.packed_to_global(i) for i in range(t.packed_size())} == live
        assert all(t.global_to_packed(t.packed_to_global(i)) == i for i in range(t.packed_size()))
