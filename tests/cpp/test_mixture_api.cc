// Drives the C++ host mirror (include/distributions_b200/mixture.hpp) through the reference's
// Mixture choreography (doc/overview.rst:130-202, tests/test_models.py:537-594): init, add_value,
// remove_value, add_group, remove_group, per-value score_value (accumulate), PitmanYor prior
// (overwrite), and the new batched entry.  Reads a script of operations, prints every score vector;
// tests/test_cpp_api.py replays the same script with the oracle and compares.
#include <distributions_b200/mixture.hpp>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

using namespace distributions_b200;

template <class T>
static void print_vec(const char * tag, const std::vector<T> & v) {
    std::printf("%s", tag);
    for (auto x : v) std::printf(" %.9g", static_cast<double>(x));
    std::printf("\n");
}

template <class Model, class Parse>
static void run_script(std::shared_ptr<Context> ctx, std::istream & in, const typename Model::Shared & shared, Parse parse) {
    rng_t rng;
    Mixture<Model> mixture(ctx);
    PitmanYor py{1.0f, 0.1f};
    PitmanYor::Mixture driver(ctx);
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string op;
        ls >> op;
        if (op == "end") break;
        if (op == "groups") {  // groups <G>: start with G empty groups
            size_t G;
            ls >> G;
            mixture.groups().resize(G);
            for (auto & g : mixture.groups()) g.init(shared, rng);
            driver.counts().assign(G, 0);
        } else if (op == "fill") {  // fill <gid> <value>: Group::add_value before init (benchmarks/mixture.cc:88-99)
            size_t gid;
            ls >> gid;
            auto v = parse(ls);
            mixture.groups(gid).add_value(shared, v, rng);
            driver.counts()[gid] += 1;
        } else if (op == "init") {
            mixture.init(shared, rng);
            driver.init(py);
        } else if (op == "add") {  // doc/overview.rst:185-195
            size_t gid;
            ls >> gid;
            auto v = parse(ls);
            const bool added = driver.add_value(py, gid);
            mixture.add_value(shared, gid, v, rng);
            if (added) mixture.add_group(shared, rng);
        } else if (op == "remove") {  // doc/overview.rst:196-202
            size_t gid;
            ls >> gid;
            auto v = parse(ls);
            const bool removed = driver.remove_value(py, gid);
            mixture.remove_value(shared, gid, v, rng);
            if (removed) mixture.remove_group(shared, gid);
        } else if (op == "score") {  // prior overwrite, then the slave accumulates
            auto v = parse(ls);
            std::vector<float> scores(mixture.groups().size(), 12345.f);
            driver.score_value(py, Floats(scores));
            print_vec("prior", scores);
            mixture.score_value(shared, v, Floats(scores), rng);
            print_vec("scores", scores);
            std::vector<float> one(1, mixture.score_value_group(shared, 0, v, rng));
            print_vec("group0", one);
        } else if (op == "batch") {  // batch <n> then n values, then n uniforms
            size_t n;
            ls >> n;
            std::vector<typename detail::WireValue<typename Model::Value>::type> vals(n);
            std::vector<float> u(n);
            std::getline(in, line);
            std::istringstream vs(line);
            for (auto & v : vals) v = parse(vs);
            std::getline(in, line);
            std::istringstream us(line);
            for (auto & x : u) us >> x;
            const size_t G = mixture.groups().size();
            std::vector<float> prior(G), scores(n * G);
            std::vector<int32_t> assign(n);
            driver.score_value(py, Floats(prior));
            mixture.score_values(shared, vals.data(), n, prior.data(), u.data(), assign.data(), scores.data());
            print_vec("batch_assign", assign);
            print_vec("batch_scores", scores);
            // the batched entry must equal the per-value path row by row up to the association of the
            // prior (the fused kernel folds it into the group cache: (prior + score) + c*L instead of
            // prior + (score + c*L)), i.e. a couple of ulp
            bool same = true;
            for (size_t i = 0; i < n && i < 8; ++i) {
                std::vector<float> row(prior);
                mixture.score_value(shared, static_cast<typename Model::Value>(vals[i]), Floats(row), rng);
                for (size_t g = 0; g < G; ++g) {
                    const float a = row[g], b = scores[i * G + g];
                    const float mag = (a < 0 ? -a : a) + 1.f;
                    same = same && ((a > b ? a - b : b - a) <= 1e-6f * mag);
                }
            }
            std::printf("batch_matches_per_value %d\n", same ? 1 : 0);
        } else if (op == "scoredata") {  // Mixture::score_data under the script's Shared
            std::vector<float> one(1, mixture.score_data(shared, rng));
            print_vec("score_data", one);
        } else if (op == "addbatch") {  // addbatch <n> then n values, then n packed group ids (non-empty groups)
            size_t n;
            ls >> n;
            std::vector<typename detail::WireValue<typename Model::Value>::type> vals(n);
            std::vector<int32_t> gids(n);
            std::getline(in, line);
            std::istringstream vs(line);
            for (auto & v : vals) v = parse(vs);
            std::getline(in, line);
            std::istringstream gs(line);
            for (auto & g : gids) gs >> g;
            for (auto g : gids) driver.add_value(py, static_cast<size_t>(g));
            mixture.add_values(shared, vals.data(), gids.data(), n);
        }
    }
}

// dpd / niw through the mirror: after every incremental operation (update_group via add_value / remove_value,
// add_group, remove_group with its swap-with-last) the per-value scores must equal those of a FRESH mixture that
// update_all()s the same groups -- bit for bit for dpd (table gathers), to rounding for niw.  MixtureIdTracker follows
// the same group churn.
template <class Model, class MakeValue>
static int incremental_equals_update_all(std::shared_ptr<Context> ctx, const typename Model::Shared & shared, MakeValue make_value,
                                         float tol, const char * name) {
    rng_t rng;
    Mixture<Model> mixture(ctx);
    MixtureIdTracker ids;
    unsigned state = 12345u;
    auto next = [&]() { state = state * 1664525u + 1013904223u; return state >> 8; };
    mixture.groups().resize(5);
    for (auto & g : mixture.groups()) g.init(shared, rng);
    for (int i = 0; i < 20; ++i) mixture.groups(next() % 5).add_value(shared, make_value(next()), rng);
    mixture.init(shared, rng);
    ids.init(5);
    int bad = 0;
    std::vector<std::pair<size_t, typename Model::Value>> history;
    for (int step = 0; step < 40; ++step) {
        const unsigned op = next() % 10;
        const size_t G = mixture.groups().size();
        if (op < 5) {
            const size_t gid = next() % G;
            const auto v = make_value(next());
            mixture.add_value(shared, gid, v, rng);
            history.push_back(std::make_pair(static_cast<size_t>(ids.packed_to_global(gid)), v));
        } else if (op < 7 && !history.empty()) {
            const auto h = history.back();
            history.pop_back();
            mixture.remove_value(shared, ids.global_to_packed(h.first), h.second, rng);
        } else if (op < 9) {
            mixture.add_group(shared, rng);
            ids.add_group();
        } else if (G > 3) {
            const size_t gid = next() % G;
            const unsigned global = ids.packed_to_global(gid);
            for (size_t i = history.size(); i-- > 0;)
                if (history[i].first == global) history.erase(history.begin() + i);
            mixture.remove_group(shared, gid);
            ids.remove_group(gid);
        }
        Mixture<Model> fresh(ctx);
        fresh.groups() = mixture.groups();
        fresh.init(shared, rng);
        const auto v = make_value(next());
        std::vector<float> a(mixture.groups().size(), 0.f), b(a);
        mixture.score_value(shared, v, Floats(a), rng);
        fresh.score_value(shared, v, Floats(b), rng);
        for (size_t g = 0; g < a.size(); ++g) {
            const float d = a[g] > b[g] ? a[g] - b[g] : b[g] - a[g];
            const float mag = (b[g] < 0 ? -b[g] : b[g]) + 1.f;
            if (!(d <= tol * mag)) {
                ++bad;
                std::printf("selfcheck %s step %d group %zu: incremental %.9g vs update_all %.9g\n", name, step, g, a[g], b[g]);
            }
        }
        if (ids.packed_size() != mixture.groups().size()) ++bad;
    }
    return bad;
}

int main(int argc, char ** argv) {
    if (argc < 2) return 2;
    std::ifstream in(argv[1]);
    try {
        auto ctx = std::make_shared<Context>(0);
        std::string line;
        while (std::getline(in, line)) {
            if (line == "model nich") {
                std::printf("model nich\n");
                run_script<NormalInverseChiSq>(ctx, in, NormalInverseChiSq::Shared::EXAMPLE(),
                                               [](std::istream & s) { float v; s >> v; return v; });
            } else if (line == "model gp") {
                std::printf("model gp\n");
                run_script<GammaPoisson>(ctx, in, GammaPoisson::Shared::EXAMPLE(),
                                         [](std::istream & s) { uint32_t v; s >> v; return v; });
            } else if (line == "model bnb") {
                std::printf("model bnb\n");
                BetaNegativeBinomial::Shared sh = BetaNegativeBinomial::Shared::EXAMPLE();
                sh.r = 3;
                run_script<BetaNegativeBinomial>(ctx, in, sh, [](std::istream & s) { uint32_t v; s >> v; return v; });
            } else if (line == "lowentropy") {  // lowentropy <dataset_size> <G> then G group sizes: the prior vector
                std::getline(in, line);
                std::istringstream ls(line);
                int32_t dataset_size;
                size_t G;
                ls >> dataset_size >> G;
                LowEntropy model{dataset_size};
                LowEntropy::Mixture driver(ctx);
                driver.counts().resize(G);
                for (auto & c : driver.counts()) ls >> c;
                driver.init(model);
                std::vector<float> scores(G, 0.f);
                driver.score_value(model, Floats(scores));
                print_vec("low_entropy", scores);
            } else if (line == "model bb") {
                std::printf("model bb\n");
                run_script<BetaBernoulli>(ctx, in, BetaBernoulli::Shared::EXAMPLE(),
                                          [](std::istream & s) { int v; s >> v; return v != 0; });
            } else if (line == "selfcheck") {
                int bad = incremental_equals_update_all<DirichletProcessDiscrete>(
                    ctx, DirichletProcessDiscrete::Shared::EXAMPLE(), [](unsigned r) { return static_cast<uint32_t>(r % 100); }, 0.f, "dpd");
                typedef NormalInverseWishart<3> Niw;
                Niw::Shared sh = Niw::Shared::EXAMPLE();
                sh.nu = 6.f;
                bad += incremental_equals_update_all<Niw>(ctx, sh, [](unsigned r) {
                    Niw::Value v;
                    for (int i = 0; i < 3; ++i) v.x[i] = static_cast<float>((r >> (5 * i)) % 32) * 0.25f - 4.f;
                    return v;
                }, 1e-5f, "niw3");
                {   // niw: batched add_values (per-group SYRK on the device) == the same add_value calls one by one, and
                    // score_data (niw.hpp:296-308) over the device-resident statistics agrees between the two
                    rng_t rng;
                    Mixture<Niw> batched(ctx), single(ctx);
                    batched.groups().resize(4);
                    for (auto & g : batched.groups()) g.init(sh, rng);
                    single.groups() = batched.groups();
                    batched.init(sh, rng);
                    single.init(sh, rng);
                    std::vector<Niw::Value> vals(60);
                    std::vector<int32_t> gid(60);
                    unsigned state = 777u;
                    for (size_t i = 0; i < vals.size(); ++i) {
                        state = state * 1664525u + 1013904223u;
                        for (int k = 0; k < 3; ++k) vals[i].x[k] = static_cast<float>((state >> (8 + 5 * k)) % 32) * 0.25f - 4.f;
                        gid[i] = static_cast<int32_t>((state >> 4) % 3);  // the last group stays empty
                        single.add_value(sh, gid[i], vals[i], rng);
                    }
                    batched.add_values(sh, vals.data(), gid.data(), vals.size());
                    for (size_t g = 0; g < 4; ++g) {
                        if (batched.groups(g).count != single.groups(g).count) ++bad;
                        for (int k = 0; k < 9; ++k) {
                            const float d = batched.groups(g).sum_xxT[k] - single.groups(g).sum_xxT[k];
                            if (!(d < 1e-3f && d > -1e-3f)) ++bad;
                        }
                    }
                    const float sa = batched.score_data(sh, rng), sb = single.score_data(sh, rng);
                    if (!(sa == sa) || !(sa - sb < 1e-2f && sb - sa < 1e-2f) || !(sa < 0.f)) ++bad;
                }
                std::printf(bad ? "selfcheck FAILED %d\n" : "selfcheck ok%.0d\n", bad);
            } else if (line == "model dd") {
                std::printf("model dd\n");
                run_script<DirichletDiscrete<16>>(ctx, in, DirichletDiscrete<16>::Shared::EXAMPLE(),
                                                  [](std::istream & s) { int v; s >> v; return v; });
            }
        }
    } catch (const std::exception & e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
