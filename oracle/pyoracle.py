"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  Oracle  -- oracle/liboracle.so, the plain-C restatement (oracle/oracle.c); always available.
  Ref     -- oracle/_ref/libref_shim.so, the compiled UNMODIFIED reference behind oracle/ref_shim.cc;
             available where it was built (container) or travelled to (GPU box).
"""
import ctypes
import os

import numpy as np

from . import build as _build

c_f = ctypes.c_float
c_i = ctypes.c_int
c_sz = ctypes.c_size_t
c_p = ctypes.c_void_p
c_u64 = ctypes.c_uint64
c_i32 = ctypes.c_int32

DD, DPD, BB, GP, NICH, NIW, BNB = 0, 1, 2, 3, 4, 5, 6
COL_DTYPE = {DD: np.int32, DPD: np.uint32, BB: np.uint8, GP: np.uint32, NICH: np.float32, NIW: np.float32, BNB: np.uint32}


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Oracle:
    """The plain-C restatement."""

    def __init__(self):
        path = _build.build_oracle()
        L = ctypes.CDLL(path)
        self.L = L
        L.orc_vec.argtypes = [c_i, c_sz, c_p, c_p]
        L.orc_py_score_add_value.restype = c_f
        L.orc_py_score_add_value.argtypes = [c_f, c_f, c_i32, c_i32, c_i32, c_i32]
        L.orc_py_prior.argtypes = [c_f, c_f, c_sz, c_p, c_p]
        L.orc_nich_caches.argtypes = [c_p, c_sz, c_p, c_p, c_p, c_p]
        L.orc_gp_caches.argtypes = [c_p, c_sz, c_p, c_p, c_p]
        L.orc_bb_caches.argtypes = [c_p, c_sz, c_p, c_p, c_p]
        L.orc_bnb_caches.argtypes = [c_f, c_f, ctypes.c_uint32, c_sz, c_p, c_p, c_p]
        L.orc_bnb_score_rows.argtypes = [c_sz, c_p, c_sz, c_p, c_p]
        L.orc_bnb_score_data.restype = c_f
        L.orc_bnb_score_data.argtypes = [c_f, c_f, ctypes.c_uint32, c_sz, c_p, c_p, c_p]
        L.orc_dd_caches.argtypes = [c_i, c_p, c_sz, c_p, c_p]
        L.orc_dpd_caches.argtypes = [c_f, c_f, c_sz, c_p, c_sz, c_p, c_p]
        L.orc_nich_score_rows.argtypes = [c_sz, c_p, c_sz, c_p, c_p]
        L.orc_gp_score_rows.argtypes = [c_sz, c_p, c_sz, c_p, c_p]
        L.orc_bb_score_rows.argtypes = [c_sz, c_p, c_sz, c_p, c_p]
        L.orc_dd_score_rows.argtypes = [c_i, c_sz, c_p, c_sz, c_p, c_p]
        L.orc_dpd_score_rows.argtypes = [c_sz, c_sz, c_p, c_sz, c_p, c_p]
        L.orc_niw_score_rows.argtypes = [c_i, c_p, c_f, c_p, c_f, c_sz, c_p, c_p, c_p, c_sz, c_p, c_p]
        L.orc_sample_rows.argtypes = [c_sz, c_sz, c_p, c_p, c_p]
        L.orc_nich_group_update.argtypes = [c_i, c_p, c_p, c_p, c_p, c_sz]
        L.orc_gp_group_update.argtypes = [c_i, c_p, c_p, c_p, c_p, c_sz]
        L.orc_bench_nich.restype = ctypes.c_double
        L.orc_bench_nich.argtypes = [c_sz, c_p, c_p, c_sz, c_p, c_p, c_p, c_i]
        L.orc_init()

    # numerics ------------------------------------------------------------------------------
    def vec(self, fn, x):
        x = np.ascontiguousarray(x)
        if fn == 4:
            x = x.astype(np.uint32).view(np.float32)
        else:
            x = x.astype(np.float32)
        out = np.empty(x.shape, dtype=np.float32)
        self.L.orc_vec(fn, x.size, _ptr(x), _ptr(out))
        return out

    def fast_log(self, x): return self.vec(0, x)
    def fast_exp(self, x): return self.vec(1, x)
    def fast_lgamma(self, x): return self.vec(2, x)
    def fast_lgamma_nu(self, x): return self.vec(3, x)
    def fast_log_factorial(self, n): return self.vec(4, n)

    # prior ---------------------------------------------------------------------------------
    def py_score_add_value(self, alpha, d, group_size, nonempty, sample_size, empty_count=1):
        return self.L.orc_py_score_add_value(alpha, d, group_size, nonempty, sample_size, empty_count)

    def py_prior(self, alpha, d, sizes):
        sizes = _i32(sizes)
        out = np.empty(sizes.size, dtype=np.float32)
        self.L.orc_py_prior(alpha, d, sizes.size, _ptr(sizes), _ptr(out))
        return out

    def low_entropy_prior(self, dataset_size, sizes):
        sizes = _i32(sizes)
        out = np.empty(sizes.size, dtype=np.float32)
        self.L.orc_low_entropy_prior(int(dataset_size), ctypes.c_size_t(sizes.size), _ptr(sizes), _ptr(out))
        return out

    # caches --------------------------------------------------------------------------------
    def nich_caches(self, shared, count, mean, ctv):
        shared, count, mean, ctv = _f32(shared), _i32(count), _f32(mean), _f32(ctv)
        out = np.empty((4, count.size), dtype=np.float32)
        self.L.orc_nich_caches(_ptr(shared), count.size, _ptr(count), _ptr(mean), _ptr(ctv), _ptr(out))
        return out

    def gp_caches(self, shared, count, sum_):
        shared, count, sum_ = _f32(shared), _u32(count), _u32(sum_)
        out = np.empty((3, count.size), dtype=np.float32)
        self.L.orc_gp_caches(_ptr(shared), count.size, _ptr(count), _ptr(sum_), _ptr(out))
        return out

    def bnb_caches(self, shared, count, sum_):
        """shared = (alpha, beta, r); rows: score, post_beta, alpha"""
        count, sum_ = _u32(count), _u32(sum_)
        out = np.empty((3, count.size), dtype=np.float32)
        self.L.orc_bnb_caches(float(shared[0]), float(shared[1]), int(shared[2]), count.size, _ptr(count), _ptr(sum_), _ptr(out))
        return out

    def bb_caches(self, shared, heads, tails):
        shared, heads, tails = _f32(shared), _i32(heads), _i32(tails)
        out = np.empty((2, heads.size), dtype=np.float32)
        self.L.orc_bb_caches(_ptr(shared), heads.size, _ptr(heads), _ptr(tails), _ptr(out))
        return out

    def dd_caches(self, alphas, counts):
        alphas, counts = _f32(alphas), _i32(counts)
        G, dim = counts.shape
        out = np.empty((dim + 1, G), dtype=np.float32)
        self.L.orc_dd_caches(dim, _ptr(alphas), G, _ptr(counts), _ptr(out))
        return out

    def dpd_caches(self, alpha, beta0, betas, counts):
        betas, counts = _f32(betas), _i32(counts)
        G, V = counts.shape
        out = np.empty((V + 2, G), dtype=np.float32)
        self.L.orc_dpd_caches(alpha, beta0, V, _ptr(betas), G, _ptr(counts), _ptr(out))
        return out

    # scoring (accumulate into scores[n][G]) ------------------------------------------------
    def score_rows(self, model, cache, values, scores):
        assert scores.dtype == np.float32 and scores.flags.c_contiguous
        n, G = scores.shape
        cache = _f32(cache)
        values = np.ascontiguousarray(values, dtype=COL_DTYPE[model])
        if model == NICH:
            self.L.orc_nich_score_rows(G, _ptr(cache), n, _ptr(values), _ptr(scores))
        elif model == GP:
            self.L.orc_gp_score_rows(G, _ptr(cache), n, _ptr(values), _ptr(scores))
        elif model == BB:
            self.L.orc_bb_score_rows(G, _ptr(cache), n, _ptr(values), _ptr(scores))
        elif model == BNB:
            self.L.orc_bnb_score_rows(G, _ptr(cache), n, _ptr(values), _ptr(scores))
        elif model == DD:
            self.L.orc_dd_score_rows(cache.shape[0] - 1, G, _ptr(cache), n, _ptr(values), _ptr(scores))
        elif model == DPD:
            self.L.orc_dpd_score_rows(cache.shape[0] - 2, G, _ptr(cache), n, _ptr(values), _ptr(scores))
        else:
            raise ValueError(model)
        return scores

    def niw_score_rows(self, mu, kappa, psi, nu, count, sum_x, sum_xxT, values, scores):
        mu, psi, count, sum_x, sum_xxT = _f32(mu), _f32(psi), _i32(count), _f32(sum_x), _f32(sum_xxT)
        values = _f32(values)
        n, G = scores.shape
        d = mu.size
        self.L.orc_niw_score_rows(d, _ptr(mu), kappa, _ptr(psi), nu, G, _ptr(count), _ptr(sum_x),
                                  _ptr(sum_xxT), n, _ptr(values), _ptr(scores))
        return scores

    def niw_score_data(self, mu, kappa, psi, nu, count, sum_x, sum_xxT):
        """sum over the groups of Group::score_data (niw.hpp:296-308)"""
        mu, psi, count, sum_x, sum_xxT = _f32(mu), _f32(psi), _i32(count), _f32(sum_x), _f32(sum_xxT)
        self.L.orc_niw_score_data.restype = ctypes.c_float
        self.L.orc_niw_score_data.argtypes = [c_i, c_p, c_f, c_p, c_f, c_sz, c_p, c_p, c_p]
        return float(self.L.orc_niw_score_data(mu.size, _ptr(mu), kappa, _ptr(psi), nu, count.size, _ptr(count), _ptr(sum_x), _ptr(sum_xxT)))

    def niw_group_update(self, sign, count, sum_x, sum_xxT, values):
        """Group::add_value (+1) / remove_value (-1) of the rows `values` [n][d], sequentially in float; returns the new statistics"""
        sum_x, sum_xxT, values = _f32(sum_x).copy(), _f32(sum_xxT).copy(), _f32(values)
        cnt = np.array([count], np.int32)
        self.L.orc_niw_group_update.restype = None
        self.L.orc_niw_group_update.argtypes = [c_i, c_i, c_p, c_p, c_p, c_sz, c_p]
        self.L.orc_niw_group_update(sign, sum_x.size, _ptr(cnt), _ptr(sum_x), _ptr(sum_xxT), values.shape[0], _ptr(values))
        return int(cnt[0]), sum_x, sum_xxT

    def sample_rows(self, scores, u):
        """scores [n][G] is overwritten with likelihoods (reference semantic). Returns assign."""
        assert scores.dtype == np.float32 and scores.flags.c_contiguous
        n, G = scores.shape
        u = _f32(u)
        assign = np.empty(n, dtype=np.int32)
        self.L.orc_sample_rows(n, G, _ptr(scores), _ptr(u), _ptr(assign))
        return assign

    def nich_group_update(self, op, count, mean, ctv, values):
        c, m, v = c_i32(count), c_f(mean), c_f(ctv)
        values = _f32(values)
        self.L.orc_nich_group_update(op, ctypes.byref(c), ctypes.byref(m), ctypes.byref(v), _ptr(values), values.size)
        return c.value, m.value, v.value

    def gp_group_update(self, op, count, sum_, log_prod, values):
        c, s, lp = ctypes.c_uint32(count), ctypes.c_uint32(sum_), c_f(log_prod)
        values = _u32(values)
        self.L.orc_gp_group_update(op, ctypes.byref(c), ctypes.byref(s), ctypes.byref(lp), _ptr(values), values.size)
        return c.value, s.value, lp.value

    def score_data(self, w, shared=None):
        """MixtureDataScorer::score_data of workload dict `w` (tests/cases.py layout) under `shared`
        (default: the workload's own).  Returns (score, sum |term|, the same fp32 terms summed in double)."""
        a = (ctypes.c_double * 2)(0, 0)
        m = w["model"]
        L = self.L
        for fn in ("orc_nich_score_data", "orc_gp_score_data", "orc_bb_score_data", "orc_dd_score_data", "orc_dpd_score_data"):
            getattr(L, fn).restype = c_f
        if m == "nich":
            sh = _f32(w["shared"] if shared is None else shared)
            c, mean, ctv = _i32(w["count"]), _f32(w["mean"]), _f32(w["ctv"])
            r = L.orc_nich_score_data(_ptr(sh), ctypes.c_size_t(c.size), _ptr(c), _ptr(mean), _ptr(ctv), a)
        elif m == "gp":
            sh = _f32(w["shared"] if shared is None else shared)
            c, sm, lp = _u32(w["count"]), _u32(w["sum"]), _f32(w["log_prod"])
            r = L.orc_gp_score_data(_ptr(sh), ctypes.c_size_t(c.size), _ptr(c), _ptr(sm), _ptr(lp), a)
        elif m == "bnb":
            sh = np.asarray(w["shared"] if shared is None else shared, np.float64)
            c, sm = _u32(w["count"]), _u32(w["sum"])
            r = L.orc_bnb_score_data(float(sh[0]), float(sh[1]), int(w["shared"][2]), c.size, _ptr(c), _ptr(sm), a)
        elif m == "bb":
            sh = _f32(w["shared"] if shared is None else shared)
            h, t = _i32(w["heads"]), _i32(w["tails"])
            r = L.orc_bb_score_data(_ptr(sh), ctypes.c_size_t(h.size), _ptr(h), _ptr(t), a)
        elif m == "dd":
            al = _f32(w["alphas"] if shared is None else shared)
            c = _i32(w["counts"])
            r = L.orc_dd_score_data(int(al.size), _ptr(al), ctypes.c_size_t(c.shape[0]), _ptr(c), a)
        elif m == "dpd":
            alpha = float(w["alpha"] if shared is None else np.ravel(shared)[0])
            b, c = _f32(w["betas"]), _i32(w["counts"])
            r = L.orc_dpd_score_data(c_f(alpha), ctypes.c_size_t(b.size), _ptr(b), ctypes.c_size_t(c.shape[0]), _ptr(c), a)
        else:
            raise ValueError(m)
        return float(r), float(a[0]), float(a[1])

    def bench_nich(self, cache, prior, values, u, n_threads):
        cache, prior, values, u = _f32(cache), _f32(prior), _f32(values), _f32(u)
        assign = np.empty(values.size, dtype=np.int32)
        secs = self.L.orc_bench_nich(prior.size, _ptr(cache), _ptr(prior), values.size, _ptr(values), _ptr(u),
                                     _ptr(assign), n_threads)
        return secs, assign


class Ref:
    """The compiled, unmodified reference behind oracle/ref_shim.cc."""

    @staticmethod
    def available():
        return _build.build_ref() is not None

    def __init__(self):
        path = _build.build_ref()
        if path is None:
            raise RuntimeError("oracle/_ref/libref_shim.so is not built and /root/reference is absent")
        L = ctypes.CDLL(path)
        self.L = L
        L.refshim_vec.argtypes = [c_i, c_sz, c_p, c_p]
        L.refshim_py_score_add_value.restype = c_f
        L.refshim_py_score_add_value.argtypes = [c_f, c_f, c_i32, c_i32, c_i32, c_i32]
        L.refshim_kind_score_data_grid.restype = ctypes.c_int
        L.refshim_kind_score_data_grid.argtypes = [c_p, ctypes.c_int, ctypes.c_size_t, c_p, ctypes.c_size_t, ctypes.c_int, c_p]
        L.refshim_kind_create.restype = c_p
        L.refshim_kind_create.argtypes = [c_sz, c_p, c_f, c_f]
        L.refshim_kind_destroy.argtypes = [c_p]
        L.refshim_kind_add_nich.argtypes = [c_p, c_p, c_p, c_p, c_p]
        L.refshim_kind_add_gp.argtypes = [c_p, c_p, c_p, c_p, c_p]
        L.refshim_kind_add_bb.argtypes = [c_p, c_p, c_p, c_p]
        L.refshim_kind_add_bnb.argtypes = [c_p, c_f, c_f, ctypes.c_uint32, c_p, c_p]
        L.refshim_kind_add_dd.argtypes = [c_p, c_i, c_p, c_p]
        L.refshim_kind_add_dpd.argtypes = [c_p, c_f, c_f, c_f, c_sz, c_p, c_p, c_p]
        L.refshim_kind_prior.argtypes = [c_p, c_p]
        L.refshim_kind_score_rows.argtypes = [c_p, c_p, c_sz, c_sz, c_i, c_p]
        L.refshim_kind_group_scores.argtypes = [c_p, c_i, c_p, c_i, c_p]
        L.refshim_kind_scorer_caches.argtypes = [c_p, c_i, c_p]
        L.refshim_sample_rows.argtypes = [c_u64, c_sz, c_sz, c_p, c_p, c_p]
        L.refshim_kind_score_sample_rows.argtypes = [c_p, c_p, c_sz, c_sz, c_u64, c_p, c_p, c_p]
        L.refshim_kind_bench.restype = ctypes.c_double
        L.refshim_kind_bench.argtypes = [c_p, c_p, c_sz, c_i, c_u64, c_p]
        L.refshim_nich_group_update.argtypes = [c_i, c_p, c_p, c_p, c_p, c_sz]
        L.refshim_gp_group_update.argtypes = [c_i, c_p, c_p, c_p, c_p, c_sz]

    def vec(self, fn, x):
        x = np.ascontiguousarray(x)
        if fn == 4:
            x = x.astype(np.uint32).view(np.float32)
        else:
            x = x.astype(np.float32)
        out = np.empty(x.shape, dtype=np.float32)
        self.L.refshim_vec(fn, x.size, _ptr(x), _ptr(out))
        return out

    def fast_log(self, x): return self.vec(0, x)
    def fast_exp(self, x): return self.vec(1, x)
    def fast_lgamma(self, x): return self.vec(2, x)
    def fast_lgamma_nu(self, x): return self.vec(3, x)
    def fast_log_factorial(self, n): return self.vec(4, n)

    def py_score_add_value(self, alpha, d, group_size, nonempty, sample_size, empty_count=1):
        return self.L.refshim_py_score_add_value(alpha, d, group_size, nonempty, sample_size, empty_count)

    def kind(self, G, group_sizes=None, alpha=1.0, d=0.0):
        return RefKind(self, G, group_sizes, alpha, d)

    def sample_rows(self, seed, scores):
        """Runs sample_from_scores_overwrite per row; scores overwritten with likelihoods.
        Returns (u, assign): u[i] is the uniform the reference drew for row i."""
        assert scores.dtype == np.float32 and scores.flags.c_contiguous
        n, G = scores.shape
        u = np.empty(n, dtype=np.float32)
        assign = np.empty(n, dtype=np.int32)
        self.L.refshim_sample_rows(seed, n, G, _ptr(scores), _ptr(u), _ptr(assign))
        return u, assign

    def nich_group_update(self, op, count, mean, ctv, values):
        c, m, v = c_i32(count), c_f(mean), c_f(ctv)
        values = _f32(values)
        self.L.refshim_nich_group_update(op, ctypes.byref(c), ctypes.byref(m), ctypes.byref(v), _ptr(values), values.size)
        return c.value, m.value, v.value

    def low_entropy_prior(self, dataset_size, sizes):
        sizes = _i32(sizes)
        out = np.empty(sizes.size, dtype=np.float32)
        self.L.refshim_low_entropy_prior.argtypes = [ctypes.c_int32, ctypes.c_size_t, c_p, c_p]
        self.L.refshim_low_entropy_prior(int(dataset_size), sizes.size, _ptr(sizes), _ptr(out))
        return out

    def gp_group_update(self, op, count, sum_, log_prod, values):
        c, s, lp = ctypes.c_uint32(count), ctypes.c_uint32(sum_), c_f(log_prod)
        values = _u32(values)
        self.L.refshim_gp_group_update(op, ctypes.byref(c), ctypes.byref(s), ctypes.byref(lp), _ptr(values), values.size)
        return c.value, s.value, lp.value


class RefKind:
    """One partition (G groups) with a PitmanYor prior and F reference feature mixtures."""

    def __init__(self, ref, G, group_sizes, alpha, d):
        self.L = ref.L
        self.G = G
        sizes = _i32(group_sizes) if group_sizes is not None else None
        self.h = self.L.refshim_kind_create(G, _ptr(sizes), alpha, d)
        self.models = []

    def __del__(self):
        if getattr(self, "h", None):
            self.L.refshim_kind_destroy(self.h)
            self.h = None

    def add_nich(self, shared, count, mean, ctv):
        shared, count, mean, ctv = _f32(shared), _i32(count), _f32(mean), _f32(ctv)
        self.models.append(NICH)
        return self.L.refshim_kind_add_nich(self.h, _ptr(shared), _ptr(count), _ptr(mean), _ptr(ctv))

    def add_gp(self, shared, count, sum_, log_prod=None):
        shared, count, sum_ = _f32(shared), _u32(count), _u32(sum_)
        lp = _f32(log_prod) if log_prod is not None else None
        self.models.append(GP)
        return self.L.refshim_kind_add_gp(self.h, _ptr(shared), _ptr(count), _ptr(sum_), _ptr(lp))

    def add_bnb(self, shared, count, sum_):
        count, sum_ = _u32(count), _u32(sum_)
        self.models.append(BNB)
        return self.L.refshim_kind_add_bnb(self.h, float(shared[0]), float(shared[1]), int(shared[2]), _ptr(count), _ptr(sum_))

    def add_bb(self, shared, heads, tails):
        shared, heads, tails = _f32(shared), _i32(heads), _i32(tails)
        self.models.append(BB)
        return self.L.refshim_kind_add_bb(self.h, _ptr(shared), _ptr(heads), _ptr(tails))

    def add_dd(self, alphas, counts):
        alphas, counts = _f32(alphas), _i32(counts)
        assert counts.shape == (self.G, alphas.size)
        self.models.append(DD)
        return self.L.refshim_kind_add_dd(self.h, alphas.size, _ptr(alphas), _ptr(counts))

    def add_dpd(self, gamma, alpha, beta0, keys, betas, counts):
        keys, betas, counts = _u32(keys), _f32(betas), _i32(counts)
        assert counts.shape == (self.G, keys.size)
        self.models.append(DPD)
        return self.L.refshim_kind_add_dpd(self.h, gamma, alpha, beta0, keys.size, _ptr(keys), _ptr(betas), _ptr(counts))

    def _cols(self, columns):
        cols = [np.ascontiguousarray(c, dtype=COL_DTYPE[m]) for c, m in zip(columns, self.models)]
        arr = (c_p * len(cols))(*[c.ctypes.data for c in cols])
        return cols, arr

    def prior(self):
        out = np.empty(self.G, dtype=np.float32)
        self.L.refshim_kind_prior(self.h, _ptr(out))
        return out

    def score_rows(self, columns, n, with_prior=True, scores=None, row0=0):
        cols, arr = self._cols(columns)
        if scores is None:
            scores = np.zeros((n, self.G), dtype=np.float32)
        self.L.refshim_kind_score_rows(self.h, arr, row0, n, 1 if with_prior else 0, _ptr(scores))
        return scores

    def score_data_grid(self, f, shareds, use_grid=True):
        """reference Mixture::score_data_grid (or score_data per point) of feature f; shareds [n_grid][stride]"""
        shareds = np.ascontiguousarray(np.atleast_2d(shareds), dtype=np.float32)
        out = np.empty(shareds.shape[0], dtype=np.float32)
        rc = self.L.refshim_kind_score_data_grid(self.h, f, ctypes.c_size_t(shareds.shape[0]), _ptr(shareds),
                                                 ctypes.c_size_t(shareds.shape[1]), 1 if use_grid else 0, _ptr(out))
        assert rc == 0
        return out

    def group_scores(self, f, value, which=0):
        v = np.asarray([value], dtype=COL_DTYPE[self.models[f]])
        out = np.empty(self.G, dtype=np.float32)
        self.L.refshim_kind_group_scores(self.h, f, _ptr(v), which, _ptr(out))
        return out

    def scorer_caches(self, f):
        rows = {NICH: 4, GP: 3, BB: 2, BNB: 3}[self.models[f]]
        out = np.empty((rows, self.G), dtype=np.float32)
        rc = self.L.refshim_kind_scorer_caches(self.h, f, _ptr(out))
        assert rc == 0
        return out

    def score_sample_rows(self, columns, n, seed, want_scores=True, row0=0):
        cols, arr = self._cols(columns)
        u = np.empty(n, dtype=np.float32)
        assign = np.empty(n, dtype=np.int32)
        scores = np.empty((n, self.G), dtype=np.float32) if want_scores else None
        self.L.refshim_kind_score_sample_rows(self.h, arr, row0, n, seed, _ptr(u), _ptr(assign), _ptr(scores))
        return u, assign, scores

    def bench(self, columns, n, n_threads, seed=1):
        cols, arr = self._cols(columns)
        assign = np.empty(n, dtype=np.int32)
        secs = self.L.refshim_kind_bench(self.h, arr, n, n_threads, seed, _ptr(assign))
        return secs, assign
