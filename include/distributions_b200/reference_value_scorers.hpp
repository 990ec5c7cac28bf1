// reference_value_scorers.hpp -- the file a maintainer of forcedotcom/distributions adds to switch a model's
// Mixture over to the B200 path: ValueScorer classes for the third template argument of
//   template<class Model, class DataScorer, class ValueScorer> struct MixtureSlave    (mixture.hpp:340-344)
// that forward to the C-ABI of dist_b200.h.  Unlike the rest of this repository it INCLUDES THE REFERENCE'S HEADERS:
// it compiles only against a reference tree (-I<reference>/include) and is exercised by
// tests/cpp/dropin_mixture_slave.cc, which instantiates the reference's own MixtureSlave with these scorers and runs
// the reference's test_mixture_score choreography (distributions/tests/test_models.py:537-594) next to the stock
// FastMixture in the same binary.
//
// One-line switch per model, e.g. models/nich.hpp:47-49:
//   typedef MixtureSlave<Model, MixtureDataScorer, B200NichValueScorer> FastMixture;
#pragma once
#include <dist_b200.h>

#include <distributions/mixture.hpp>
#include <distributions/models/dd.hpp>
#include <distributions/models/nich.hpp>

#include <vector>

namespace distributions {

// shared plumbing: one context + one feature (= one device-side MixtureValueScorer)
struct B200ScorerBase {
    B200ScorerBase(int model) {
        DIST_ASSERT(dist_b200_ctx_create(0, &ctx_) == DIST_B200_OK, "no B200 context (is a CUDA device present?)");
        DIST_ASSERT(dist_b200_feature_create(ctx_, model, &f_) == DIST_B200_OK, dist_b200_last_error(ctx_));
    }
    ~B200ScorerBase() {
        dist_b200_feature_destroy(f_);
        dist_b200_ctx_destroy(ctx_);
    }
    B200ScorerBase(const B200ScorerBase &) = delete;
    B200ScorerBase & operator=(const B200ScorerBase &) = delete;

    // mixture.hpp:361-375: packed_add of a fresh group / packed_remove = swap-with-last
    void add_group_impl() { check(dist_b200_feature_add_group(f_, nullptr)); }
    void remove_group_impl(size_t groupid) { check(dist_b200_feature_remove_group(f_, static_cast<int>(groupid), nullptr)); }

    // mixture.hpp:416-425: per-value score_value ACCUMULATES into the caller's buffer
    void score_value_impl(const void * value, AlignedFloats scores_accum) const {
        check(dist_b200_score_value_host(ctx_, f_, value, scores_accum.data()));
    }

    // the NEW batched entry: rows against frozen statistics, prior vector from PitmanYor::Mixture::score_value
    // (clustering.hpp:195-208), uniforms in place of sample_unif01(rng) (random.hpp:47-50)
    void score_sample_values(const void * values, size_t n, const float * prior, const float * u, int32_t * assign,
                             float * scores /* [n][G] or null */) const {
        const dist_b200_feature * feats[1] = {f_};
        const void * cols[1] = {values};
        check(dist_b200_score_sample_batch_host(ctx_, feats, 1, cols, n, prior, u, assign, scores));
    }

    void check(int rc) const { DIST_ASSERT(rc == DIST_B200_OK, dist_b200_last_error(ctx_)); }
    dist_b200_ctx * ctx_ = nullptr;
    dist_b200_feature * f_ = nullptr;
};

struct B200NichValueScorer : MixtureSlaveValueScorerMixin<NormalInverseChiSq>, B200ScorerBase {
    typedef NormalInverseChiSq::Shared Shared;
    typedef NormalInverseChiSq::Group Group;
    typedef NormalInverseChiSq::Value Value;

    B200NichValueScorer() : B200ScorerBase(DIST_B200_NICH) {}

    void resize(const Shared &, size_t) {}
    // nich.hpp:344-352
    void update_all(const Shared & s, const std::vector<Group> & groups, rng_t &) {
        const size_t G = groups.size();
        std::vector<int32_t> count(G);
        std::vector<float> mean(G), ctv(G);
        for (size_t g = 0; g < G; ++g) {  // Group = {count, mean, count_times_variance}, nich.hpp:98-101
            count[g] = groups[g].count;
            mean[g] = groups[g].mean;
            ctv[g] = groups[g].count_times_variance;
        }
        const float shared[4] = {s.mu, s.kappa, s.sigmasq, s.nu};
        check(dist_b200_nich_update_all(f_, shared, static_cast<int>(G), count.data(), mean.data(), ctv.data(), nullptr));
    }
    // nich.hpp:312-342
    void update_group(const Shared &, size_t groupid, const Group & g, rng_t &) {
        struct { int32_t count; float mean; float ctv; } st = {static_cast<int32_t>(g.count), g.mean, g.count_times_variance};
        check(dist_b200_feature_update_group(f_, static_cast<int>(groupid), &st, nullptr));
    }
    void add_value(const Shared & s, size_t gid, const Group & g, const Value &, rng_t & r) { update_group(s, gid, g, r); }
    void remove_value(const Shared & s, size_t gid, const Group & g, const Value &, rng_t & r) { update_group(s, gid, g, r); }
    void add_group(const Shared &, rng_t &) { add_group_impl(); }
    void remove_group(const Shared &, size_t gid) { remove_group_impl(gid); }

    void score_value(const Shared &, const std::vector<Group> &, const Value & value, AlignedFloats scores_accum, rng_t &) const {
        score_value_impl(&value, scores_accum);
    }
    float score_value_group(const Shared & s, const std::vector<Group> & gs, size_t gid, const Value & v, rng_t & r) const {
        VectorFloat tmp(gs.size(), 0.f);
        score_value(s, gs, v, tmp, r);
        return tmp[gid];
    }
};

template<int max_dim>
struct B200DdValueScorer : MixtureSlaveValueScorerMixin<DirichletDiscrete<max_dim>>, B200ScorerBase {
    typedef typename DirichletDiscrete<max_dim>::Shared Shared;
    typedef typename DirichletDiscrete<max_dim>::Group Group;
    typedef typename DirichletDiscrete<max_dim>::Value Value;

    B200DdValueScorer() : B200ScorerBase(DIST_B200_DD) {}

    void resize(const Shared &, size_t) {}
    // dd.hpp:399-421
    void update_all(const Shared & s, const std::vector<Group> & groups, rng_t &) {
        const size_t G = groups.size();
        std::vector<int32_t> counts(G * s.dim);
        for (size_t g = 0; g < G; ++g)
            for (int v = 0; v < s.dim; ++v) counts[g * s.dim + v] = groups[g].counts[v];  // Group::counts, dd.hpp:89-92
        check(dist_b200_dd_update_all(f_, s.dim, s.alphas, static_cast<int>(G), counts.data(), nullptr));
    }
    // dd.hpp:381-397,458-467
    void update_group(const Shared & s, size_t groupid, const Group & g, rng_t &) {
        int32_t counts[max_dim];
        for (int v = 0; v < s.dim; ++v) counts[v] = g.counts[v];
        check(dist_b200_feature_update_group(f_, static_cast<int>(groupid), counts, nullptr));
    }
    void add_value(const Shared & s, size_t gid, const Group & g, const Value &, rng_t & r) { update_group(s, gid, g, r); }
    void remove_value(const Shared & s, size_t gid, const Group & g, const Value &, rng_t & r) { update_group(s, gid, g, r); }
    void add_group(const Shared &, rng_t &) { add_group_impl(); }
    void remove_group(const Shared &, size_t gid) { remove_group_impl(gid); }

    void score_value(const Shared &, const std::vector<Group> &, const Value & value, AlignedFloats scores_accum, rng_t &) const {
        const int32_t v = value;
        score_value_impl(&v, scores_accum);
    }
    float score_value_group(const Shared & s, const std::vector<Group> & gs, size_t gid, const Value & v, rng_t & r) const {
        VectorFloat tmp(gs.size(), 0.f);
        score_value(s, gs, v, tmp, r);
        return tmp[gid];
    }
};

}  // namespace distributions
