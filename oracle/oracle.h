/* oracle/oracle.h -- CPU restatement of the reference's mixture-scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the checker.  Nothing under distributions_b200/ may include, link or load
 * anything from oracle/.
 *
 * Parity status: PINNED for dd / dpd / bb / gp / nich / PitmanYor prior / sampler / numerics --
 * tests/test_oracle_vs_reference.py checks every function against the compiled, unmodified
 * reference (oracle/_ref/libref_shim.so) and against committed golden vectors generated from it
 * (tests/golden/, tests/golden/make_golden.py).  NIW: "parity unpinned" for the linear algebra
 * (the reference needs Eigen, absent here) -- pinned only through the d=1
 * NIW == NICH identity (reference test_normal_models.py:34-100) and a float64 numpy restatement
 * of dbg/random.py:113-131.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef ORACLE_H_
#define ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void orc_init(void); /* builds the fast_log / fast_exp tables; idempotent, called lazily */

/* ---- numerics: special.hpp:53-89,114-171,208-214,239-273; vendor/fmath.hpp:138-163,438-459 */
float orc_fast_log(float x);
float orc_fast_exp(float x);
float orc_fast_lgamma(float y);
float orc_fast_lgamma_nu(float nu);
float orc_fast_log_factorial(uint32_t n);
/* fn: 0 log, 1 exp, 2 lgamma, 3 lgamma_nu, 4 log_factorial (input bits as uint32) */
void orc_vec(int fn, size_t n, const float *in, float *out);

/* ---- clustering prior: clustering.hpp:81-104 (score_add_value), :151-161,:195-230 (CachedMixture) */
float orc_py_score_add_value(float alpha, float d, int32_t group_size, int32_t nonempty,
                             int32_t sample_size, int32_t empty_count);
void orc_py_prior(float alpha, float d, size_t G, const int32_t *group_sizes, float *out);

/* ---- per-model cache rebuild (MixtureValueScorer::update_all) from raw group statistics */
/* nich.hpp:58-69,239-250 ; cache = [4][G]: score, log_coeff, precision, mean */
void orc_nich_caches(const float shared[4], size_t G, const int32_t *count, const float *mean,
                     const float *ctv, float *cache);
/* gp.hpp:56-61,198-207 ; cache = [3][G]: score, post_alpha, score_coeff */
void orc_gp_caches(const float shared[2], size_t G, const uint32_t *count, const uint32_t *sum,
                   float *cache);
/* bb.hpp:276-292 ; cache = [2][G]: heads, tails */
/* BetaNegativeBinomial: caches bnb.hpp:57-63,200-211 as cache[3][G] = score, post_beta,
 * alpha; score_value bnb.hpp:308-319 (accumulates) */
void orc_bnb_caches(float alpha, float beta, uint32_t r, size_t G, const uint32_t *count, const uint32_t *sum,
                    float *cache);
void orc_bnb_score_rows(size_t G, const float *cache, size_t n, const uint32_t *values, float *scores);
void orc_bb_caches(const float shared[2], size_t G, const int32_t *heads, const int32_t *tails,
                   float *cache);
/* dd.hpp:399-421 ; counts [G][dim] ; cache = [dim+1][G]: rows 0..dim-1 = scores_[v], row dim = shift */
void orc_dd_caches(int dim, const float *alphas, size_t G, const int32_t *counts, float *cache);
/* dpd.hpp:471-497,517-543 ; counts [G][V] dense ; cache = [V+2][G]: rows 0..V-1 = scores_[v],
 * row V = the "unseen / OTHER" row (constant fast_log(alpha*beta0)), row V+1 = shift */
void orc_dpd_caches(float alpha, float beta0, size_t V, const float *betas, size_t G,
                    const int32_t *counts, float *cache);

/* ---- Mixture::score_value over rows: scores[n][G] += model term (ACCUMULATE, mixture.hpp:416-425) */
void orc_nich_score_rows(size_t G, const float *cache, size_t n, const float *values, float *scores);
void orc_gp_score_rows(size_t G, const float *cache, size_t n, const uint32_t *values, float *scores);
void orc_bb_score_rows(size_t G, const float *cache, size_t n, const uint8_t *values, float *scores);
void orc_dd_score_rows(int dim, size_t G, const float *cache, size_t n, const int32_t *values,
                       float *scores);
/* rows: dense row index per value, in [0, V) for known values, V for OTHER/unseen */
void orc_dpd_score_rows(size_t V, size_t G, const float *cache, size_t n, const uint32_t *rows,
                        float *scores);

/* ---- NIW (niw.hpp:82-103,353-360; random.hpp:160-185).  psi, sum_xxT row-major [d][d] */
void orc_niw_score_rows(int d, const float *mu, float kappa, const float *psi, float nu, size_t G,
                        const int32_t *count, const float *sum_x, const float *sum_xxT, size_t n,
                        const float *values /*[n][d]*/, float *scores /*[n][G] accumulate*/);

/* ---- sampler: random.cc:94-106 + random.hpp:315-333,360-366, u supplied by the caller.
 * scores rows are overwritten with the likelihoods, as the reference does. */
float orc_scores_to_likelihoods(size_t G, float *scores);
int32_t orc_sample_from_likelihoods(size_t G, const float *lik, float total, float u);
void orc_sample_rows(size_t n, size_t G, float *scores, const float *u, int32_t *assign);

/* ---- Group::add_value / remove_value restatements (host bookkeeping): nich.hpp:125-165, gp.hpp:109-135 */
void orc_nich_group_update(int op, int32_t *count, float *mean, float *ctv, const float *values, size_t n);
void orc_gp_group_update(int op, uint32_t *count, uint32_t *sum, float *log_prod, const uint32_t *values,
                         size_t n);

/* ---- timing helper for bench.py's cpu_baseline "port" leg: rows [0,n) of a single-feature nich
 * mixture, prior + score + sample, split over n_threads.  Returns wall seconds. */
double orc_bench_nich(size_t G, const float *cache, const float *prior, size_t n, const float *values,
                      const float *u, int32_t *assign, int n_threads);

/* ---- LowEntropy clustering prior: clustering.hpp:265-293,318-327 through MixtureDriver::score_value
 * (mixture.hpp:123-141) */
float orc_low_entropy_score_add_value(int32_t dataset_size, int32_t group_size, int32_t sample_size,
                                      int32_t empty_group_count);
void orc_low_entropy_prior(int32_t dataset_size, size_t G, const int32_t *sizes, float *out);

/* ---- score_data (log marginal likelihood of all groups): nich.hpp:262-288,
 * gp.hpp:220-241, bb.hpp:207-229, dd.hpp:250-324, dpd.hpp:344-374.  Returns the reference's fp32 group-order
 * accumulation; abs_sum (nullable, 2 doubles) receives sum |term| and the same fp32 terms summed in double. */
float orc_nich_score_data(const float sh[4], size_t G, const int32_t *count, const float *mean,
                          const float *ctv, double *abs_sum);
float orc_gp_score_data(const float sh[2], size_t G, const uint32_t *count, const uint32_t *sum,
                        const float *log_prod, double *abs_sum);
float orc_bnb_score_data(float alpha, float beta, uint32_t r, size_t G, const uint32_t *count, const uint32_t *sum,
                         double *abs_sum); /* bnb.hpp:221-243 */
float orc_bb_score_data(const float sh[2], size_t G, const int32_t *heads, const int32_t *tails, double *abs_sum);
float orc_dd_score_data(int dim, const float *alphas, size_t G, const int32_t *counts, double *abs_sum);
float orc_dpd_score_data(float alpha, size_t V, const float *betas, size_t G, const int32_t *counts, double *abs_sum);

#ifdef __cplusplus
}
#endif
#endif /* ORACLE_H_ */
