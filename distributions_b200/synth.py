"""Seeded synthetic workloads for the parity tests and bench.py (numpy only, no CUDA, no oracle).

Recipe follows SURVEY.md §8(d): G-1 populated groups (the last one empty so the empty-group branch
of the Pitman-Yor prior, clustering.hpp:97-100, is exercised), ~40 values per populated group drawn
from the model's own generator, PitmanYor(alpha=1, d=0.1) on the group sizes, u ~ U[0,1) float32.
Hyper-parameters are the reference's Shared::EXAMPLE() values (nich.hpp:87-94, gp.hpp:75-80,
bb.hpp:70-75, dd.hpp:78-85, dpd.hpp:141-152, niw.hpp:160-170).
"""
import numpy as np

PY_ALPHA = 1.0
PY_D = 0.1


def _sizes(rng, G, per_group=40):
    sizes = rng.multinomial(per_group * max(G - 1, 1), np.full(max(G - 1, 1), 1.0 / max(G - 1, 1))).astype(np.int32)
    sizes = np.maximum(sizes, 1)
    if G > 1:
        sizes = np.concatenate([sizes, np.zeros(1, np.int32)])
    return sizes


def nich(seed, G, N):
    """NormalInverseChiSq: Shared::EXAMPLE() mu=0,kappa=1,sigmasq=1,nu=1; values ~ 3*N(0,1)."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, G)
    centers = rng.normal(0.0, 3.0, G)
    count = sizes.copy()
    mean = np.zeros(G, np.float32)
    ctv = np.zeros(G, np.float32)
    for g in range(G):
        if count[g]:
            x = rng.normal(centers[g], 1.0, count[g]).astype(np.float32)
            mean[g] = x.mean(dtype=np.float64)
            ctv[g] = ((x.astype(np.float64) - x.mean(dtype=np.float64)) ** 2).sum()
    return dict(model="nich", shared=np.array([0.0, 1.0, 1.0, 1.0], np.float32), sizes=sizes,
                count=count, mean=mean, ctv=ctv,
                values=(3.0 * rng.standard_normal(N)).astype(np.float32),
                u=rng.random(N, dtype=np.float32))


def gp(seed, G, N, lam=5.0):
    """GammaPoisson: EXAMPLE() alpha=1, inv_beta=1; values ~ Poisson(5)."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, G)
    rates = rng.gamma(2.0, lam / 2.0, G)
    count = sizes.astype(np.uint32)
    sum_ = np.array([rng.poisson(rates[g], sizes[g]).sum() if sizes[g] else 0 for g in range(G)], np.uint32)
    # Group::log_prod = sum of log(value!) over the group's values (gp.hpp:109-116); only score_data reads it.
    # Synthesised from the sums without drawing from rng (the committed golden vectors pin the stream).
    log_prod = (0.5 * count * np.log1p(sum_ / np.maximum(count, 1)) ** 2).astype(np.float32)
    return dict(model="gp", shared=np.array([1.0, 1.0], np.float32), sizes=sizes, count=count, sum=sum_, log_prod=log_prod,
                values=rng.poisson(lam, N).astype(np.uint32), u=rng.random(N, dtype=np.float32))


def bnb(seed, G, N, r=3):
    """BetaNegativeBinomial: alpha=1, beta=1 (EXAMPLE(), bnb.hpp:79-85) with r failures; values ~ NegBin(r, p_g)."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, G)
    ps = rng.beta(2.0, 2.0, G) * 0.8 + 0.1
    count = sizes.astype(np.uint32)
    sum_ = np.array([rng.negative_binomial(r, ps[g], sizes[g]).sum() if sizes[g] else 0 for g in range(G)], np.uint32)
    return dict(model="bnb", shared=np.array([1.0, 1.0, r], np.float32), sizes=sizes, count=count, sum=sum_,
                values=rng.negative_binomial(r, 0.4, N).astype(np.uint32), u=rng.random(N, dtype=np.float32))


def bb(seed, G, N, p=0.3):
    """BetaBernoulli: EXAMPLE() alpha=0.5, beta=2; values ~ Bernoulli(0.3)."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, G)
    ps = rng.beta(0.5, 2.0, G)
    heads = np.array([rng.binomial(sizes[g], ps[g]) for g in range(G)], np.int32)
    tails = sizes - heads
    return dict(model="bb", shared=np.array([0.5, 2.0], np.float32), sizes=sizes, heads=heads, tails=tails,
                values=(rng.random(N) < p).astype(np.uint8), u=rng.random(N, dtype=np.float32))


def dd(seed, G, N, dim=16):
    """DirichletDiscrete<dim>: EXAMPLE() alphas=0.5; values ~ U{0..dim-1}."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, G)
    counts = np.zeros((G, dim), np.int32)
    for g in range(G):
        if sizes[g]:
            counts[g] = rng.multinomial(sizes[g], rng.dirichlet(np.full(dim, 0.5)))
    return dict(model="dd", alphas=np.full(dim, 0.5, np.float32), sizes=sizes, counts=counts,
                values=rng.integers(0, dim, N).astype(np.int32), u=rng.random(N, dtype=np.float32))


def dpd(seed, G, N, V=4096, other_frac=0.0, zipf=None):
    """DirichletProcessDiscrete: V known values (keys 0..V-1), betas=1/V, beta0=0, alpha=0.5, gamma=1
    (pattern of dpd.hpp:141-152).  values ~ U{0..V-1} (or Zipf), a fraction other_frac = OTHER()."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, G)
    counts = np.zeros((G, V), np.int32)
    for g in range(G):
        if sizes[g]:
            idx = rng.integers(0, V, sizes[g])
            np.add.at(counts[g], idx, 1)
    if zipf:
        w = 1.0 / np.arange(1, V + 1) ** zipf
        values = rng.choice(V, size=N, p=w / w.sum()).astype(np.uint32)
    else:
        values = rng.integers(0, V, N).astype(np.uint32)
    if other_frac > 0:
        values[rng.random(N) < other_frac] = 0xFFFFFFFF
    return dict(model="dpd", gamma=1.0, alpha=0.5, beta0=0.0, keys=np.arange(V, dtype=np.uint32),
                betas=np.full(V, 1.0 / V, np.float32), sizes=sizes, counts=counts, values=values,
                u=rng.random(N, dtype=np.float32))


def niw(seed, G, N, d=32):
    """NormalInverseWishart: mu=0, kappa=1, psi=I, nu=d+2; group data ~ N(m_g, I), m_g ~ 3 N(0, I)."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, G)
    centers = 3.0 * rng.standard_normal((G, d))
    sum_x = np.zeros((G, d), np.float32)
    sum_xxT = np.zeros((G, d, d), np.float32)
    for g in range(G):
        if sizes[g]:
            x = (centers[g] + rng.standard_normal((sizes[g], d))).astype(np.float32).astype(np.float64)
            sum_x[g] = x.sum(0)
            sum_xxT[g] = x.T @ x
    which = rng.integers(0, G, N)
    values = (centers[which] + rng.standard_normal((N, d))).astype(np.float32)
    return dict(model="niw", mu=np.zeros(d, np.float32), kappa=1.0, psi=np.eye(d, dtype=np.float32),
                nu=float(d + 2), sizes=sizes, count=sizes.copy(), sum_x=sum_x, sum_xxT=sum_xxT,
                values=values, u=rng.random(N, dtype=np.float32))


def crosscat(seed, G, N, n_gp=128, n_bb=128):
    """One cross-cat kind: n_gp GammaPoisson + n_bb BetaBernoulli features sharing one partition."""
    rng = np.random.default_rng(seed)
    sizes = _sizes(rng, G)
    feats = []
    for f in range(n_gp):
        w = gp(seed * 1000 + f, G, N)
        # all features of a kind share the partition, hence the group sizes
        w["count"] = sizes.astype(np.uint32)
        w["sum"] = np.array([rng.poisson(5.0 * sizes[g]) if sizes[g] else 0 for g in range(G)], np.uint32)
        w["sizes"] = sizes
        feats.append(w)
    for f in range(n_bb):
        w = bb(seed * 1000 + n_gp + f, G, N)
        ps = rng.beta(0.5, 2.0, G)
        w["heads"] = np.array([rng.binomial(sizes[g], ps[g]) for g in range(G)], np.int32)
        w["tails"] = sizes - w["heads"]
        w["sizes"] = sizes
        feats.append(w)
    return dict(model="crosscat", sizes=sizes, features=feats, u=rng.random(N, dtype=np.float32))
