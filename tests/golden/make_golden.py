#!/usr/bin/env python
"""Generate tests/golden/hotpath_golden.npz from the compiled UNMODIFIED reference (oracle/_ref).

Run in the container (needs /root/reference or a prebuilt oracle/_ref):
    python tests/golden/make_golden.py
The reference stores no golden vectors for this path, so these are outputs of the
reference's own code on seeded inputs; the inputs are regenerated from distributions_b200.synth
by the tests (same seeds) and are also stored here for the numerics sweeps.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from distributions_b200 import synth  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402


def main():
    r = Ref()
    out = {}
    rng = np.random.default_rng(7)
    # ---- numerics sweeps (special.hpp / fmath.hpp); anchors of SURVEY.md §8c included
    xs = {
        "log": np.concatenate([np.logspace(-38, 38, 1500), rng.uniform(1, 2, 500), -np.logspace(-3, 3, 20),
                               [0.0, -2.0, 2.5, np.inf, 1.0, 2.0]]).astype(np.float32),
        "exp": np.concatenate([np.linspace(-104, 0.49, 1500), -np.logspace(-6, 1.9, 500),
                               [0.0, -87.3, -87.5, -88.0, -100.0]]).astype(np.float32),
        "lgamma": np.concatenate([np.logspace(-4, 9.6, 1800), np.arange(1, 200) * 0.5,
                                  [2.4999, 2.5, 3.0, 6.0, 100.5, 4294967040.0]]).astype(np.float32),
        "lgamma_nu": np.concatenate([np.logspace(-3, 9.6, 1800), [0.0624, 0.0625, 1.0, 5.0]]).astype(np.float32),
    }
    for i, name in enumerate(["log", "exp", "lgamma", "lgamma_nu"]):
        out["num_%s_x" % name] = xs[name]
        out["num_%s_y" % name] = r.vec(i, xs[name])
    n = np.concatenate([np.arange(0, 200), [1000, 65535, 1 << 20, (1 << 24) + 1, 0xFFFFFFFE, 0xFFFFFFFF]]).astype(np.uint32)
    out["num_logfact_x"] = n
    out["num_logfact_y"] = r.fast_log_factorial(n)

    # ---- Pitman-Yor prior: PitmanYor.EXAMPLES (lp/clustering.pyx:211-217) x random sizes
    examples = [(1.0, 0.0), (1.0, 0.1), (10.0, 0.1), (1.0, 0.9), (0.1, 0.5)]
    for j, (alpha, d) in enumerate(examples):
        for empties in (1, 10):
            sizes = np.concatenate([rng.integers(1, 60, 23), np.zeros(empties, np.int64)]).astype(np.int32)
            rng.shuffle(sizes)
            key = "prior_%d_%d" % (j, empties)
            out[key + "_sizes"] = sizes
            out[key + "_alpha_d"] = np.array([alpha, d], np.float32)
            out[key + "_out"] = r.kind(sizes.size, sizes, alpha, d).prior()
            nonempty = int((sizes > 0).sum())
            total = int(sizes.sum())
            out[key + "_add_value"] = np.array(
                [r.py_score_add_value(alpha, d, int(s), nonempty, total, empties) for s in sizes], np.float32)

    # ---- single-feature mixtures: prior + Mixture::score_value + sample_from_scores_overwrite
    for name, cfg in cases.SMALL.items():
        w = cases.make(name, **cfg)
        G, N = cfg["G"], cfg["N"]
        k = r.kind(G, w["sizes"], synth.PY_ALPHA, synth.PY_D)
        cases.ref_add_feature(k, w)
        u, assign, scores = k.score_sample_rows([w["values"]], N, seed=2024)
        out["%s_scores" % name] = scores
        out["%s_u" % name] = u
        out["%s_assign" % name] = assign
        # accumulate semantic (test_models.py:552-557): pre-loaded noise, no prior
        noise = rng.standard_normal((8, G)).astype(np.float32)
        out["%s_noise" % name] = noise
        out["%s_accum" % name] = k.score_rows([w["values"]], 8, with_prior=False, scores=noise.copy())
        # Mixture == per-group Group::score_value == score_value_group (test_models.py:537-594)
        v0 = w["values"][0]
        out["%s_group_scores" % name] = k.group_scores(0, v0, 0)
        out["%s_mixture_group_scores" % name] = k.group_scores(0, v0, 1)
        if name in ("nich", "gp", "bb"):
            out["%s_caches" % name] = k.scorer_caches(0)

    # ---- a small cross-cat kind: 3 gp + 3 bb + 2 nich features, one partition
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
    k = r.kind(G, cc["sizes"], synth.PY_ALPHA, synth.PY_D)
    for w in feats:
        cases.ref_add_feature(k, w)
    u, assign, scores = k.score_sample_rows([w["values"] for w in feats], N, seed=77)
    out["crosscat_scores"] = scores
    out["crosscat_u"] = u
    out["crosscat_assign"] = assign

    # ---- sampler alone on awkward score rows (ties, -inf-like, G=1, huge spread)
    for G in (1, 2, 5, 100, 1024):
        s = (rng.standard_normal((40, G)) * rng.choice([0.1, 3.0, 40.0], (40, 1))).astype(np.float32)
        s[0, :] = 0.0                      # exact ties
        s[1, :] = -1000.0                  # all equal and very negative
        if G > 2:
            s[2, :] = -200.0
            s[2, G // 2] = 0.0             # one dominant group, the rest underflow
        lik = s.copy()
        u, assign = r.sample_rows(99 + G, lik)
        out["sampler_%d_scores" % G] = s
        out["sampler_%d_u" % G] = u
        out["sampler_%d_assign" % G] = assign

    # ---- Group::add_value / remove_value sequences (nich.hpp:125-165, gp.hpp:109-135)
    vals = (3 * rng.standard_normal(25)).astype(np.float32)
    st = (0, 0.0, 0.0)
    trace = []
    for v in vals:
        st = r.nich_group_update(+1, *st, [v]); trace.append(st)
    for v in vals[:20]:
        st = r.nich_group_update(-1, *st, [v]); trace.append(st)
    out["nich_group_values"] = vals
    out["nich_group_trace"] = np.array(trace, np.float64)
    vals = rng.poisson(30, 25).astype(np.uint32); vals[3] = 100
    st = (0, 0, 0.0)
    trace = []
    for v in vals:
        st = r.gp_group_update(+1, *st, [v]); trace.append(st)
    for v in vals[:20]:
        st = r.gp_group_update(-1, *st, [v]); trace.append(st)
    out["gp_group_values"] = vals
    out["gp_group_trace"] = np.array(trace, np.float64)

    path = os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
