// oracle/_ref shim: a C-ABI wrapper around the UNMODIFIED reference implementation.
//
// TEST/BENCH INFRASTRUCTURE ONLY.  This file is compiled (by oracle/build.py) together with the
// reference's own sources where they lie under /root/reference/src, against the reference's own
// headers under /root/reference/include, into oracle/_ref/libref_shim.so.  Nothing under
// distributions_b200/ links or loads it; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do, and only as the checker / the CPU arm.
//
// Every routine below just *calls* the reference:
//   numerics          distributions::fast_log / fast_exp / fast_lgamma / fast_lgamma_nu /
//                     fast_log_factorial                   (include/distributions/special.hpp)
//   prior             Clustering<int>::PitmanYor::{score_add_value, Mixture}
//                                                          (include/distributions/clustering.hpp:58-239)
//   feature mixtures  Model::Mixture::{init, score_value, score_value_group}, Group::score_value
//                                                          (include/distributions/mixture.hpp:340-450,
//                                                           models/{dd,dpd,bb,gp,nich}.hpp, src/models/*.cc)
//   sampler           sample_from_scores_overwrite         (include/distributions/random.hpp:360-366)
//
// The uniform consumed by the reference's sampler is captured the way SURVEY.md App. A describes:
// copy the rng, draw sample_unif01 from the copy, then let the reference consume the original.

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <distributions/clustering.hpp>
#include <distributions/mixture.hpp>
#include <distributions/models/bb.hpp>
#include <distributions/models/dd.hpp>
#include <distributions/models/dpd.hpp>
#include <distributions/models/bnb.hpp>
#include <distributions/models/gp.hpp>
#include <distributions/models/nich.hpp>
#include <distributions/random.hpp>
#include <distributions/special.hpp>
#include <distributions/vector.hpp>

namespace D = distributions;
typedef D::Clustering<int32_t>::PitmanYor PitmanYor;
typedef D::DirichletDiscrete<256> DD;
typedef D::DirichletProcessDiscrete DPD;
typedef D::BetaBernoulli BB;
typedef D::GammaPoisson GP;
typedef D::NormalInverseChiSq NICH;
typedef D::BetaNegativeBinomial BNB;

namespace {

enum Kind { K_DD = 0, K_DPD = 1, K_BB = 2, K_GP = 3, K_NICH = 4, K_BNB = 6 };

struct Feature {
    int kind;
    DD::Shared dd_shared;     DD::Mixture dd;
    DPD::Shared dpd_shared;   DPD::Mixture dpd;
    BB::Shared bb_shared;     BB::Mixture bb;
    GP::Shared gp_shared;     GP::Mixture gp;
    NICH::Shared nich_shared; NICH::Mixture nich;
    BNB::Shared bnb_shared;   BNB::Mixture bnb;
};

struct RefKind {
    size_t G;
    bool has_prior;
    PitmanYor py;
    PitmanYor::Mixture prior;
    std::vector<std::shared_ptr<Feature>> feats;

    // one row: (optional) prior overwrite, then every feature accumulates  
    void score_row(const void * const * cols, size_t row, bool with_prior,
                   D::VectorFloat & scores, D::rng_t & rng) const {
        if (with_prior) {
            prior.score_value(py, scores);
        }
        for (size_t f = 0; f < feats.size(); ++f) {
            const Feature & ft = *feats[f];
            switch (ft.kind) {
                case K_DD: {
                    int v = static_cast<const int32_t *>(cols[f])[row];
                    ft.dd.score_value(ft.dd_shared, v, scores, rng);
                } break;
                case K_DPD: {
                    uint32_t v = static_cast<const uint32_t *>(cols[f])[row];
                    ft.dpd.score_value(ft.dpd_shared, v, scores, rng);
                } break;
                case K_BB: {
                    bool v = static_cast<const uint8_t *>(cols[f])[row] != 0;
                    ft.bb.score_value(ft.bb_shared, v, scores, rng);
                } break;
                case K_GP: {
                    uint32_t v = static_cast<const uint32_t *>(cols[f])[row];
                    ft.gp.score_value(ft.gp_shared, v, scores, rng);
                } break;
                case K_NICH: {
                    float v = static_cast<const float *>(cols[f])[row];
                    ft.nich.score_value(ft.nich_shared, v, scores, rng);
                } break;
                case K_BNB: {
                    uint32_t v = static_cast<const uint32_t *>(cols[f])[row];
                    ft.bnb.score_value(ft.bnb_shared, v, scores, rng);
                } break;
            }
        }
    }
};

RefKind * K(void * p) { return static_cast<RefKind *>(p); }

}  // namespace

extern "C" {

// ----------------------------------------------------------------------------------------------
// numerics: fn 0 fast_log, 1 fast_exp, 2 fast_lgamma, 3 fast_lgamma_nu, 4 fast_log_factorial
// (input reinterpreted as uint32 for fn 4)
void refshim_vec(int fn, size_t n, const float * in, float * out) {
    for (size_t i = 0; i < n; ++i) {
        switch (fn) {
            case 0: out[i] = D::fast_log(in[i]); break;
            case 1: out[i] = D::fast_exp(in[i]); break;
            case 2: out[i] = D::fast_lgamma(in[i]); break;
            case 3: out[i] = D::fast_lgamma_nu(in[i]); break;
            case 4: {
                uint32_t k;
                memcpy(&k, in + i, 4);
                out[i] = D::fast_log_factorial(k);
            } break;
            default: out[i] = 0.f;
        }
    }
}

// raw coefficient tables of the reference, for oracle/gen_tables.py
// which 0: lgamma_approx_coeff5 (33*6), 1: lgamma_nu_func_approx_coeff3 (18*4), 2: log_factorial_table (64)
size_t refshim_table(int which, float * out, size_t cap) {
    const float * src = nullptr;
    size_t n = 0;
    switch (which) {
        case 0: src = D::detail::lgamma_approx_coeff5; n = 33 * 6; break;
        case 1: src = D::detail::lgamma_nu_func_approx_coeff3; n = 18 * 4; break;
        case 2: src = D::detail::log_factorial_table; n = 64; break;
        default: return 0;
    }
    for (size_t i = 0; i < n && i < cap; ++i) out[i] = src[i];
    return n;
}

float refshim_py_score_add_value(float alpha, float d, int32_t group_size, int32_t nonempty,
                                 int32_t sample_size, int32_t empty_count) {
    PitmanYor py;
    py.alpha = alpha;
    py.d = d;
    return py.score_add_value(group_size, nonempty, sample_size, empty_count);
}

// ----------------------------------------------------------------------------------------------
// LowEntropy clustering: Mixture = MixtureDriver<LowEntropy> (clustering.hpp:303), score_value overwrites out[G]
void refshim_low_entropy_prior(int32_t dataset_size, size_t G, const int32_t * group_sizes, float * out) {
    typedef D::Clustering<int32_t>::LowEntropy LowEntropy;
    LowEntropy model;
    model.dataset_size = dataset_size;
    LowEntropy::Mixture mixture;
    mixture.counts().assign(group_sizes, group_sizes + G);
    mixture.init(model);
    D::VectorFloat scores(G);
    mixture.score_value(model, scores);
    for (size_t g = 0; g < G; ++g) out[g] = scores[g];
}

// a "kind": one partition (G groups) shared by F feature mixtures + a PitmanYor prior on its sizes

void * refshim_kind_create(size_t G, const int32_t * group_sizes, float alpha, float d) {
    RefKind * k = new RefKind();
    k->G = G;
    k->has_prior = (group_sizes != nullptr);
    k->py.alpha = alpha;
    k->py.d = d;
    if (k->has_prior) {
        k->prior.counts().assign(group_sizes, group_sizes + G);
        k->prior.init(k->py);
    }
    return k;
}

void refshim_kind_destroy(void * p) { delete K(p); }

int refshim_kind_add_nich(void * p, const float * shared4, const int32_t * count,
                          const float * mean, const float * ctv) {
    RefKind * k = K(p);
    D::rng_t rng;
    auto ft = std::make_shared<Feature>();
    ft->kind = K_NICH;
    ft->nich_shared.mu = shared4[0];
    ft->nich_shared.kappa = shared4[1];
    ft->nich_shared.sigmasq = shared4[2];
    ft->nich_shared.nu = shared4[3];
    ft->nich.groups().resize(k->G);
    for (size_t g = 0; g < k->G; ++g) {
        auto & grp = ft->nich.groups(g);
        grp.count = count[g];
        grp.mean = mean[g];
        grp.count_times_variance = ctv[g];
    }
    ft->nich.init(ft->nich_shared, rng);
    k->feats.push_back(ft);
    return static_cast<int>(k->feats.size()) - 1;
}

int refshim_kind_add_gp(void * p, const float * shared2, const uint32_t * count,
                        const uint32_t * sum, const float * log_prod) {
    RefKind * k = K(p);
    D::rng_t rng;
    auto ft = std::make_shared<Feature>();
    ft->kind = K_GP;
    ft->gp_shared.alpha = shared2[0];
    ft->gp_shared.inv_beta = shared2[1];
    ft->gp.groups().resize(k->G);
    for (size_t g = 0; g < k->G; ++g) {
        auto & grp = ft->gp.groups(g);
        grp.count = count[g];
        grp.sum = sum[g];
        grp.log_prod = log_prod ? log_prod[g] : 0.f;
    }
    ft->gp.init(ft->gp_shared, rng);
    k->feats.push_back(ft);
    return static_cast<int>(k->feats.size()) - 1;
}

int refshim_kind_add_bnb(void * p, float alpha, float beta, uint32_t r, const uint32_t * count,
                         const uint32_t * sum) {
    RefKind * k = K(p);
    D::rng_t rng;
    auto ft = std::make_shared<Feature>();
    ft->kind = K_BNB;
    ft->bnb_shared.alpha = alpha;
    ft->bnb_shared.beta = beta;
    ft->bnb_shared.r = r;
    ft->bnb.groups().resize(k->G);
    for (size_t g = 0; g < k->G; ++g) {
        auto & grp = ft->bnb.groups(g);
        grp.count = count[g];
        grp.sum = sum[g];
    }
    ft->bnb.init(ft->bnb_shared, rng);
    k->feats.push_back(ft);
    return static_cast<int>(k->feats.size()) - 1;
}

int refshim_kind_add_bb(void * p, const float * shared2, const int32_t * heads,
                        const int32_t * tails) {
    RefKind * k = K(p);
    D::rng_t rng;
    auto ft = std::make_shared<Feature>();
    ft->kind = K_BB;
    ft->bb_shared.alpha = shared2[0];
    ft->bb_shared.beta = shared2[1];
    ft->bb.groups().resize(k->G);
    for (size_t g = 0; g < k->G; ++g) {
        auto & grp = ft->bb.groups(g);
        grp.heads = heads[g];
        grp.tails = tails[g];
    }
    ft->bb.init(ft->bb_shared, rng);
    k->feats.push_back(ft);
    return static_cast<int>(k->feats.size()) - 1;
}

// counts: [G][dim] row-major
int refshim_kind_add_dd(void * p, int dim, const float * alphas, const int32_t * counts) {
    RefKind * k = K(p);
    if (dim > 256) return -1;
    D::rng_t rng;
    auto ft = std::make_shared<Feature>();
    ft->kind = K_DD;
    ft->dd_shared.dim = dim;
    for (int v = 0; v < dim; ++v) ft->dd_shared.alphas[v] = alphas[v];
    ft->dd.groups().resize(k->G);
    for (size_t g = 0; g < k->G; ++g) {
        auto & grp = ft->dd.groups(g);
        grp.init(ft->dd_shared, rng);
        for (int v = 0; v < dim; ++v) {
            grp.counts[v] = counts[g * dim + v];
            grp.count_sum += counts[g * dim + v];
        }
    }
    ft->dd.init(ft->dd_shared, rng);
    k->feats.push_back(ft);
    return static_cast<int>(k->feats.size()) - 1;
}

// keys/betas: the V known values and their stick weights; counts: [G][V] dense, column v <-> keys[v]
int refshim_kind_add_dpd(void * p, float gamma, float alpha, float beta0, size_t V,
                         const uint32_t * keys, const float * betas, const int32_t * counts) {
    RefKind * k = K(p);
    D::rng_t rng;
    auto ft = std::make_shared<Feature>();
    ft->kind = K_DPD;
    ft->dpd_shared.gamma = gamma;
    ft->dpd_shared.alpha = alpha;
    ft->dpd_shared.beta0 = beta0;
    for (size_t v = 0; v < V; ++v) {
        ft->dpd_shared.betas.add(keys[v], betas[v]);
        ft->dpd_shared.counts.add(keys[v]);
    }
    ft->dpd.groups().resize(k->G);
    for (size_t g = 0; g < k->G; ++g) {
        auto & grp = ft->dpd.groups(g);
        grp.init(ft->dpd_shared, rng);
        for (size_t v = 0; v < V; ++v) {
            int32_t c = counts[g * V + v];
            if (c) grp.counts.add(keys[v], c);
        }
    }
    ft->dpd.init(ft->dpd_shared, rng);
    k->feats.push_back(ft);
    return static_cast<int>(k->feats.size()) - 1;
}

// prior vector alone: CachedMixture::score_value (overwrites)
void refshim_kind_prior(void * p, float * out) {
    RefKind * k = K(p);
    D::VectorFloat scores(k->G, 12345.f);  // noise: pins the overwrite semantic
    k->prior.score_value(k->py, scores);
    memcpy(out, scores.data(), k->G * sizeof(float));
}

// scores[n][G]: with_prior ? prior + sum_f : (in-place accumulate of sum_f onto what is there)
void refshim_kind_score_rows(void * p, const void * const * cols, size_t row0, size_t n,
                             int with_prior, float * scores) {
    RefKind * k = K(p);
    D::rng_t rng;
    D::VectorFloat buf(k->G);
    for (size_t i = 0; i < n; ++i) {
        memcpy(buf.data(), scores + i * k->G, k->G * sizeof(float));
        k->score_row(cols, row0 + i, with_prior != 0, buf, rng);
        memcpy(scores + i * k->G, buf.data(), k->G * sizeof(float));
    }
}

// per-group Group::score_value / Mixture::score_value_group for one value (test_mixture_score style)
// which: 0 = Group::score_value, 1 = Mixture::score_value_group
void refshim_kind_group_scores(void * p, int f, const void * value, int which, float * out) {
    RefKind * k = K(p);
    D::rng_t rng;
    const Feature & ft = *k->feats[f];
    for (size_t g = 0; g < k->G; ++g) {
        switch (ft.kind) {
            case K_DD: {
                int v = *static_cast<const int32_t *>(value);
                out[g] = which ? ft.dd.score_value_group(ft.dd_shared, g, v, rng)
                               : ft.dd.groups(g).score_value(ft.dd_shared, v, rng);
            } break;
            case K_DPD: {
                uint32_t v = *static_cast<const uint32_t *>(value);
                out[g] = which ? ft.dpd.score_value_group(ft.dpd_shared, g, v, rng)
                               : ft.dpd.groups(g).score_value(ft.dpd_shared, v, rng);
            } break;
            case K_BB: {
                bool v = *static_cast<const uint8_t *>(value) != 0;
                out[g] = which ? ft.bb.score_value_group(ft.bb_shared, g, v, rng)
                               : ft.bb.groups(g).score_value(ft.bb_shared, v, rng);
            } break;
            case K_GP: {
                uint32_t v = *static_cast<const uint32_t *>(value);
                out[g] = which ? ft.gp.score_value_group(ft.gp_shared, g, v, rng)
                               : ft.gp.groups(g).score_value(ft.gp_shared, v, rng);
            } break;
            case K_NICH: {
                float v = *static_cast<const float *>(value);
                out[g] = which ? ft.nich.score_value_group(ft.nich_shared, g, v, rng)
                               : ft.nich.groups(g).score_value(ft.nich_shared, v, rng);
            } break;
            case K_BNB: {
                uint32_t v = *static_cast<const uint32_t *>(value);
                out[g] = which ? ft.bnb.score_value_group(ft.bnb_shared, g, v, rng)
                               : ft.bnb.groups(g).score_value(ft.bnb_shared, v, rng);
            } break;
        }
    }
}

// Scorer::init fields per group = the reference's SoA caches.
// nich: out[4][G] = score, log_coeff, precision, mean ; gp: out[3][G] = score, post_alpha, score_coeff
// bb: out[2][G] = heads_score, tails_score
int refshim_kind_scorer_caches(void * p, int f, float * out) {
    RefKind * k = K(p);
    D::rng_t rng;
    const Feature & ft = *k->feats[f];
    const size_t G = k->G;
    for (size_t g = 0; g < G; ++g) {
        switch (ft.kind) {
            case K_NICH: {
                NICH::Scorer s;
                s.init(ft.nich_shared, ft.nich.groups(g), rng);
                out[0 * G + g] = s.score;
                out[1 * G + g] = s.log_coeff;
                out[2 * G + g] = s.precision;
                out[3 * G + g] = s.mean;
            } break;
            case K_GP: {
                GP::Scorer s;
                s.init(ft.gp_shared, ft.gp.groups(g), rng);
                out[0 * G + g] = s.score;
                out[1 * G + g] = s.post_alpha;
                out[2 * G + g] = s.score_coeff;
            } break;
            case K_BB: {
                BB::Scorer s;
                s.init(ft.bb_shared, ft.bb.groups(g), rng);
                out[0 * G + g] = s.heads_score;
                out[1 * G + g] = s.tails_score;
            } break;
            case K_BNB: {  // out[3][G] = score, post_beta, alpha
                BNB::Scorer s;
                s.init(ft.bnb_shared, ft.bnb.groups(g), rng);
                out[0 * G + g] = s.score;
                out[1 * G + g] = s.post_beta;
                out[2 * G + g] = s.alpha;
            } break;
            default: return -1;
        }
    }
    return 0;
}

// the sampler alone on caller-supplied scores: rows of G floats, overwritten with likelihoods.
// u_out[i] = the uniform the reference consumed for row i.
void refshim_sample_rows(uint64_t seed, size_t n, size_t G, float * scores, float * u_out,
                         int32_t * assign_out) {
    D::rng_t rng(seed);
    D::VectorFloat buf(G);
    for (size_t i = 0; i < n; ++i) {
        memcpy(buf.data(), scores + i * G, G * sizeof(float));
        D::rng_t copy = rng;
        u_out[i] = D::sample_unif01(copy);
        assign_out[i] = static_cast<int32_t>(D::sample_from_scores_overwrite(rng, buf));
        memcpy(scores + i * G, buf.data(), G * sizeof(float));
    }
}

// full row step: prior -> features -> sample_from_scores_overwrite; optional copy of the scores
void refshim_kind_score_sample_rows(void * p, const void * const * cols, size_t row0, size_t n,
                                    uint64_t seed, float * u_out, int32_t * assign_out,
                                    float * scores_out) {
    RefKind * k = K(p);
    D::rng_t rng(seed);
    D::VectorFloat buf(k->G);
    for (size_t i = 0; i < n; ++i) {
        k->score_row(cols, row0 + i, true, buf, rng);
        if (scores_out) memcpy(scores_out + i * k->G, buf.data(), k->G * sizeof(float));
        D::rng_t copy = rng;
        u_out[i] = D::sample_unif01(copy);
        assign_out[i] = static_cast<int32_t>(D::sample_from_scores_overwrite(rng, buf));
    }
}

// CPU arm for bench.py: the same row step over rows [0, n) split into n_threads contiguous
// shards, each thread working on its own deep copy of the mixtures (the library itself is
// single-threaded; rows are independent given frozen statistics).  Returns wall seconds.
double refshim_kind_bench(void * p, const void * const * cols, size_t n, int n_threads,
                          uint64_t seed, int32_t * assign_out) {
    RefKind * k = K(p);
    if (n_threads < 1) n_threads = 1;
    std::vector<std::unique_ptr<RefKind>> copies;
    for (int t = 0; t < n_threads; ++t) {
        std::unique_ptr<RefKind> c(new RefKind(*k));
        for (auto & f : c->feats) f = std::make_shared<Feature>(*f);  // deep copy
        copies.push_back(std::move(c));
    }
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (int t = 0; t < n_threads; ++t) {
        size_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
        RefKind * c = copies[t].get();
        threads.emplace_back([=]() {
            D::rng_t rng(seed + 7919u * t);
            D::VectorFloat buf(c->G);
            for (size_t i = lo; i < hi; ++i) {
                c->score_row(cols, i, true, buf, rng);
                assign_out[i] = static_cast<int32_t>(D::sample_from_scores_overwrite(rng, buf));
            }
        });
    }
    for (auto & th : threads) th.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ----------------------------------------------------------------------------------------------
// MixtureSlave::score_data_grid / score_data (mixture.hpp:427-438) of feature f over n_grid hyper-parameter
// settings.  shareds: n_grid packed Shareds, `stride` floats each -- nich (mu, kappa, sigmasq, nu); gp (alpha,
// inv_beta); bb (alpha, beta); dd alphas[dim]; dpd (alpha) with gamma / beta0 / betas of the feature's Shared.
// use_grid != 0 calls score_data_grid (dd: the incremental _update path, dd.hpp:259-289), else score_data per point.
int refshim_kind_score_data_grid(void * p, int f, size_t n_grid, const float * shareds, size_t stride,
                                 int use_grid, float * out) {
    RefKind * k = K(p);
    D::rng_t rng;
    Feature & ft = *k->feats[f];
    D::VectorFloat scores(n_grid);
    switch (ft.kind) {
        case K_NICH: {
            std::vector<NICH::Shared> grid(n_grid, ft.nich_shared);
            for (size_t i = 0; i < n_grid; ++i) {
                const float * s = shareds + i * stride;
                grid[i].mu = s[0]; grid[i].kappa = s[1]; grid[i].sigmasq = s[2]; grid[i].nu = s[3];
            }
            if (use_grid) ft.nich.score_data_grid(grid, scores, rng);
            else for (size_t i = 0; i < n_grid; ++i) scores[i] = ft.nich.score_data(grid[i], rng);
        } break;
        case K_GP: {
            std::vector<GP::Shared> grid(n_grid, ft.gp_shared);
            for (size_t i = 0; i < n_grid; ++i) {
                grid[i].alpha = shareds[i * stride]; grid[i].inv_beta = shareds[i * stride + 1];
            }
            if (use_grid) ft.gp.score_data_grid(grid, scores, rng);
            else for (size_t i = 0; i < n_grid; ++i) scores[i] = ft.gp.score_data(grid[i], rng);
        } break;
        case K_BB: {
            std::vector<BB::Shared> grid(n_grid, ft.bb_shared);
            for (size_t i = 0; i < n_grid; ++i) {
                grid[i].alpha = shareds[i * stride]; grid[i].beta = shareds[i * stride + 1];
            }
            if (use_grid) ft.bb.score_data_grid(grid, scores, rng);
            else for (size_t i = 0; i < n_grid; ++i) scores[i] = ft.bb.score_data(grid[i], rng);
        } break;
        case K_DD: {
            std::vector<DD::Shared> grid(n_grid, ft.dd_shared);
            for (size_t i = 0; i < n_grid; ++i)
                for (int v = 0; v < ft.dd_shared.dim; ++v) grid[i].alphas[v] = shareds[i * stride + v];
            if (use_grid) ft.dd.score_data_grid(grid, scores, rng);
            else for (size_t i = 0; i < n_grid; ++i) scores[i] = ft.dd.score_data(grid[i], rng);
        } break;
        case K_BNB: {  // packed (alpha, beta); r of the feature's Shared
            std::vector<BNB::Shared> grid(n_grid, ft.bnb_shared);
            for (size_t i = 0; i < n_grid; ++i) {
                grid[i].alpha = shareds[i * stride]; grid[i].beta = shareds[i * stride + 1];
            }
            if (use_grid) ft.bnb.score_data_grid(grid, scores, rng);
            else for (size_t i = 0; i < n_grid; ++i) scores[i] = ft.bnb.score_data(grid[i], rng);
        } break;
        case K_DPD: {
            std::vector<DPD::Shared> grid(n_grid, ft.dpd_shared);
            for (size_t i = 0; i < n_grid; ++i) grid[i].alpha = shareds[i * stride];
            if (use_grid) ft.dpd.score_data_grid(grid, scores, rng);
            else for (size_t i = 0; i < n_grid; ++i) scores[i] = ft.dpd.score_data(grid[i], rng);
        } break;
        default: return -1;
    }
    for (size_t i = 0; i < n_grid; ++i) out[i] = scores[i];
    return 0;
}

// Group::add_value / remove_value (host-side bookkeeping the product mirrors)
// op: +1 add, -1 remove; state arrays are in/out
void refshim_nich_group_update(int op, int32_t * count, float * mean, float * ctv,
                               const float * values, size_t n) {
    D::rng_t rng;
    NICH::Shared shared = NICH::Shared::EXAMPLE();
    NICH::Group g;
    g.count = *count; g.mean = *mean; g.count_times_variance = *ctv;
    for (size_t i = 0; i < n; ++i) {
        if (op > 0) g.add_value(shared, values[i], rng);
        else g.remove_value(shared, values[i], rng);
    }
    *count = g.count; *mean = g.mean; *ctv = g.count_times_variance;
}

void refshim_gp_group_update(int op, uint32_t * count, uint32_t * sum, float * log_prod,
                             const uint32_t * values, size_t n) {
    D::rng_t rng;
    GP::Shared shared = GP::Shared::EXAMPLE();
    GP::Group g;
    g.count = *count; g.sum = *sum; g.log_prod = *log_prod;
    for (size_t i = 0; i < n; ++i) {
        if (op > 0) g.add_value(shared, values[i], rng);
        else g.remove_value(shared, values[i], rng);
    }
    *count = g.count; *sum = g.sum; *log_prod = g.log_prod;
}

}  // extern "C"
