// dropin_mixture_slave.cc -- the real plug point, compiled: the REFERENCE'S OWN
//   MixtureSlave<Model, MixtureDataScorer, ValueScorer>            (include/distributions/mixture.hpp:340-450)
// instantiated with the B200 ValueScorers of include/distributions_b200/reference_value_scorers.hpp, run through the
// choreography of the reference's test_mixture_score / test_mixture_runs (distributions/tests/test_models.py:507-594)
// NEXT TO the stock FastMixture in the same binary: init, score_value (accumulate semantic, with noise in the
// buffer), score_value_group, add_value / remove_value with sampling from the scores, add_group / remove_group with
// the packed swap-with-last ids -- every score vector is compared with the stock mixture's and with the per-group
// Group::score_value loop.  Also reports the latency of one per-value score_value through each scorer and the batched
// entry against the stock per-value loop.
//
// Built in the container only (needs /root/reference/include and the reference objects of oracle/_ref) by
// oracle/build.py into oracle/_ref/dropin_mixture_slave; the binary travels to the GPU box, where
// tests/test_dropin.py runs it.  Exit code 0 = every check passed.
#include <distributions_b200/reference_value_scorers.hpp>

#include <distributions/clustering.hpp>
#include <distributions/random.hpp>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <vector>

using namespace distributions;

static int failures = 0;
#define CHECK(cond, ...)                        \
    do {                                        \
        if (!(cond)) {                          \
            ++failures;                         \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);           \
            std::printf("\n");                  \
        }                                       \
    } while (0)

// the reference's own cross-implementation bar (distributions/tests/util.py:42,114-120)
static bool close(float a, float b, float tol = 1e-3f) { return std::fabs(a - b) <= tol * (1.f + std::fabs(a) + std::fabs(b)); }
// tight bar between the B200 scorer and the stock FastMixture (same caches, same fast math): nich carries the fast_log
// table step |log_coeff| * 6.2e-5; dd is a table gather and must be bit-identical
template<class Group> static float tight_tol(const Group &) { return 0.f; }
static float tight_tol(const NormalInverseChiSq::Group & g) { return 4e-6f + 6.2e-5f * 0.5f * (g.count + 2.f); }  // |log_coeff| = (nu' + 1) / 2

template<class Model, class B200Scorer>
struct Harness {
    typedef typename Model::Shared Shared;
    typedef typename Model::Group Group;
    typedef typename Model::Value Value;
    typedef typename Model::Mixture Stock;                                                        // FastMixture
    typedef MixtureSlave<Model, typename Model::MixtureDataScorer, B200Scorer> B200Mixture;       // the drop-in

    Shared shared;
    Stock stock;
    B200Mixture b200;
    rng_t rng;
    const char * name;

    Harness(const char * n, const Shared & s) : shared(s), rng(20240), name(n) {}

    void check_score_value(const Value & value, const char * phase) {
        const size_t G = stock.groups().size();
        CHECK(b200.groups().size() == G, "%s %s: group counts differ", name, phase);
        VectorFloat expected(G), a_stock(G), a_b200(G), noise(G);
        for (size_t g = 0; g < G; ++g) {
            expected[g] = stock.groups(g).score_value(shared, value, rng);  // the per-group fallback loop, mixture.hpp:321-337
            noise[g] = sample_unif01(rng) * 4.f - 2.f;
            a_stock[g] = a_b200[g] = noise[g];
        }
        stock.score_value(shared, value, a_stock, rng);  // ACCUMULATES
        b200.score_value(shared, value, a_b200, rng);
        for (size_t g = 0; g < G; ++g) {
            const float s = a_stock[g] - noise[g], b = a_b200[g] - noise[g];
            CHECK(close(b, expected[g]), "%s %s: score_value g=%zu b200 %.7g vs Group::score_value %.7g", name, phase, g, b, expected[g]);
            const float tight = tight_tol(stock.groups(g)) + 2e-6f * (std::fabs(noise[g]) + std::fabs(s));
            CHECK(std::fabs(b - s) <= tight, "%s %s: score_value g=%zu b200 %.9g vs stock FastMixture %.9g", name, phase, g, b, s);
            const float one = b200.score_value_group(shared, g, value, rng);
            CHECK(close(one, expected[g]), "%s %s: score_value_group g=%zu %.7g vs %.7g", name, phase, g, one, expected[g]);
        }
    }

    template<class MakeValue>
    void run(size_t n_values, MakeValue make_value) {
        std::vector<Value> values;
        for (size_t i = 0; i < n_values; ++i) values.push_back(make_value(rng));
        // one group per value, like test_mixture_score; both mixtures get identical groups
        stock.groups().resize(n_values);
        b200.groups().resize(n_values);
        for (size_t i = 0; i < n_values; ++i) {
            stock.groups(i).init(shared, rng);
            stock.groups(i).add_value(shared, values[i], rng);
            b200.groups(i) = stock.groups(i);
        }
        stock.init(shared, rng);
        b200.init(shared, rng);
        for (const Value & v : values) check_score_value(v, "init");

        // adding: sample a group from the B200 scores, add to both
        std::vector<size_t> groupids;
        for (const Value & v : values) {
            check_score_value(v, "adding");
            VectorFloat scores(stock.groups().size(), 0.f);
            b200.score_value(shared, v, scores, rng);
            const size_t gid = sample_from_scores_overwrite(rng, scores);
            stock.add_value(shared, gid, v, rng);
            b200.add_value(shared, gid, v, rng);
            groupids.push_back(gid);
        }
        // removing
        for (size_t i = 0; i < values.size(); ++i) {
            stock.remove_value(shared, groupids[i], values[i], rng);
            b200.remove_value(shared, groupids[i], values[i], rng);
            check_score_value(values[i], "removing");
        }
        // group churn with packed ids (swap-with-last), as in test_mixture_runs
        stock.remove_group(shared, 0);
        b200.remove_group(shared, 0);
        stock.remove_group(shared, stock.groups().size() - 1);
        b200.remove_group(shared, b200.groups().size() - 1);
        stock.add_group(shared, rng);
        b200.add_group(shared, rng);
        for (const Value & v : values) {
            check_score_value(v, "churn");
            VectorFloat scores(stock.groups().size(), 0.f);
            b200.score_value(shared, v, scores, rng);
            const size_t gid = sample_from_scores_overwrite(rng, scores);
            stock.add_value(shared, gid, v, rng);
            b200.add_value(shared, gid, v, rng);
        }
        for (const Value & v : values) check_score_value(v, "final");
        stock.validate(shared);
        b200.validate(shared);

        // per-value latency: what a sequential Gibbs loop through the drop-in pays per score_value call
        VectorFloat scores(stock.groups().size(), 0.f);
        const int iters = 2000;
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < iters; ++i) stock.score_value(shared, values[i % values.size()], scores, rng);
        auto t1 = std::chrono::steady_clock::now();
        for (int i = 0; i < iters; ++i) b200.score_value(shared, values[i % values.size()], scores, rng);
        auto t2 = std::chrono::steady_clock::now();
        const double us_stock = std::chrono::duration<double, std::micro>(t1 - t0).count() / iters;
        const double us_b200 = std::chrono::duration<double, std::micro>(t2 - t1).count() / iters;
        std::printf("LATENCY %s G=%zu per-value score_value: stock FastMixture %.3f us, B200 ValueScorer %.1f us "
                    "(zero-copy value and scores: 1 launch + 1 synchronise per call)\n", name, stock.groups().size(), us_stock, us_b200);
    }
};

int main() {
    {
        Harness<NormalInverseChiSq, B200NichValueScorer> h("nich", NormalInverseChiSq::Shared::EXAMPLE());
        h.run(40, [&](rng_t & r) { return sample_normal(r, 0.f, 4.f); });
    }
    {
        typedef DirichletDiscrete<16> DD;
        Harness<DD, B200DdValueScorer<16>> h("dd16", DD::Shared::EXAMPLE());
        h.run(40, [&](rng_t & r) { return static_cast<DD::Value>(sample_int(r, 0, 15)); });
    }
    {   // the batched entry against the stock per-value loop: nich, G = 100 populated groups, 50 000 rows
        typedef NormalInverseChiSq Model;
        Model::Shared shared = Model::Shared::EXAMPLE();
        rng_t rng(7);
        const size_t G = 100, N = 50000;
        Model::Mixture stock;
        stock.groups().resize(G);
        for (size_t g = 0; g < G; ++g) {
            stock.groups(g).init(shared, rng);
            const float center = sample_normal(rng, 0.f, 9.f);
            for (int i = 0; i < 30 && g + 1 < G; ++i) stock.groups(g).add_value(shared, sample_normal(rng, center, 1.f), rng);
        }
        stock.init(shared, rng);
        // prior vector from the reference's own PitmanYor CachedMixture (clustering.hpp:195-208)
        Clustering<int>::PitmanYor py;
        py.alpha = 1.f;
        py.d = 0.1f;
        Clustering<int>::PitmanYor::Mixture driver;
        for (size_t g = 0; g < G; ++g) driver.counts().push_back(stock.groups(g).count);
        driver.init(py);
        VectorFloat prior(G, 0.f);
        driver.score_value(py, prior);
        std::vector<float> values(N), u(N);
        for (size_t n = 0; n < N; ++n) values[n] = sample_normal(rng, 0.f, 9.f);
        rng_t draw = rng;
        for (size_t n = 0; n < N; ++n) u[n] = sample_unif01(draw);  // the uniforms the reference's sampler will consume
        std::vector<int32_t> want(N), got(N);
        VectorFloat scores(G);
        auto t0 = std::chrono::steady_clock::now();
        for (size_t n = 0; n < N; ++n) {  // examples/mixture/main.py:236-244 per row
            driver.score_value(py, scores);                       // overwrite with the prior
            stock.score_value(shared, values[n], scores, rng);    // accumulate
            want[n] = static_cast<int32_t>(sample_from_scores_overwrite(rng, scores));
        }
        auto t1 = std::chrono::steady_clock::now();
        // (MixtureSlave keeps its value_scorer_ private: the batched entry is called on a scorer of its own here; in the
        // reference tree it is a one-line forwarding method next to MixtureSlave::score_value)
        B200NichValueScorer batch;
        batch.update_all(shared, stock.groups(), rng);
        batch.score_sample_values(values.data(), N, prior.data(), u.data(), got.data(), nullptr);
        auto t2 = std::chrono::steady_clock::now();
        size_t same = 0;
        for (size_t n = 0; n < N; ++n) same += want[n] == got[n];
        std::printf("BATCH nich G=%zu N=%zu: identical indices %zu / %zu; stock per-value loop %.1f ms, B200 batched entry %.2f ms (cold)\n",
                    G, N, same, N, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                    std::chrono::duration<double, std::milli>(t2 - t1).count());
        CHECK(same >= N - N / 500, "batched entry: only %zu of %zu indices match the reference's own sampler", same, N);
    }
    std::printf(failures ? "DROPIN FAILED: %d checks\n" : "DROPIN OK%.0d\n", failures);
    return failures ? 1 : 0;
}
