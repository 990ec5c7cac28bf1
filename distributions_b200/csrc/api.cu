// api.cu -- the C-ABI of include/dist_b200.h: contexts, feature (MixtureValueScorer) lifecycle,
// dispatch of the hot-path entry points onto the sm_100a kernels.  No CPU compute path exists here:
// every scoring / sampling call ends in a kernel launch or an error code.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>

#include "common.cuh"
#include "ref_tables.inc"

using namespace distb200;

namespace {

cudaStream_t as_stream(void *s) { return static_cast<cudaStream_t>(s); }

int ensure_scratch(dist_b200_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return DIST_B200_OK;
    if (ctx->scratch_dev) {
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        DISTB200_CUDA(ctx, cudaFree(ctx->scratch_dev));
        ctx->scratch_dev = nullptr;
        ctx->scratch_bytes = 0;
    }
    const size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 20);
    DISTB200_CUDA(ctx, cudaMalloc(&ctx->scratch_dev, want));
    ctx->scratch_bytes = want;
    return DIST_B200_OK;
}

int ensure_scores_scratch(dist_b200_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scores_scratch_bytes) return DIST_B200_OK;
    if (ctx->scores_scratch) {
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        DISTB200_CUDA(ctx, cudaFree(ctx->scores_scratch));
        ctx->scores_scratch = nullptr;
        ctx->scores_scratch_bytes = 0;
    }
    DISTB200_CUDA(ctx, cudaMalloc(&ctx->scores_scratch, bytes));
    ctx->scores_scratch_bytes = bytes;
    return DIST_B200_OK;
}

int ensure_pinned(dist_b200_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_bytes) return DIST_B200_OK;
    if (ctx->pinned) {
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        DISTB200_CUDA(ctx, cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
    }
    const size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 20);
    DISTB200_CUDA(ctx, cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_bytes = want;
    return DIST_B200_OK;
}


// per-group floats of the hot layout
size_t group_floats(const dist_b200_feature *f) {
    switch (f->model) {
        case DIST_B200_DD: return static_cast<size_t>(f->dim);
        default: return 4;
    }
}

// device-side raw statistics: number of arrays and 4-byte elements per group of array a
int stat_arrays(const dist_b200_feature *f) {
    switch (f->model) {
        case DIST_B200_NICH: return 3;
        case DIST_B200_GP: case DIST_B200_BB: case DIST_B200_BNB: return 2;
        case DIST_B200_DD: return 1;
        default: return 0;
    }
}
size_t stat_elems(const dist_b200_feature *f, int) { return f->model == DIST_B200_DD ? static_cast<size_t>(f->dim) : 1; }
uint32_t *stat_ptr(const dist_b200_feature *f, int a) {
    size_t off = 0;
    for (int b = 0; b < a; ++b) off += stat_elems(f, b) * f->capacity;
    return f->stats + off;
}
// (re)allocate the statistics arrays for the feature's current capacity, keeping the first old_G groups
int ensure_stats(dist_b200_feature *f, int old_capacity, int old_G, bool keep) {
    dist_b200_ctx *ctx = f->ctx;
    const int na = stat_arrays(f);
    if (na == 0) return DIST_B200_OK;
    size_t words = 0;
    for (int a = 0; a < na; ++a) words += stat_elems(f, a) * f->capacity;
    if (f->stats && old_capacity == f->capacity && words <= f->stats_words) return DIST_B200_OK;
    uint32_t *fresh = nullptr;
    DISTB200_CUDA(ctx, cudaMalloc(&fresh, words * 4));
    DISTB200_CUDA(ctx, cudaMemset(fresh, 0, words * 4));
    if (f->stats) {
        if (keep && old_capacity > 0) {
            size_t off_old = 0, off_new = 0;
            for (int a = 0; a < na; ++a) {
                const size_t e = stat_elems(f, a);
                DISTB200_CUDA(ctx, cudaMemcpy(fresh + off_new, f->stats + off_old, e * std::min(old_G, f->capacity) * 4, cudaMemcpyDeviceToDevice));
                off_old += e * old_capacity;
                off_new += e * f->capacity;
            }
        }
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        DISTB200_CUDA(ctx, cudaFree(f->stats));
    }
    f->stats = fresh;
    f->stats_words = words;
    return DIST_B200_OK;
}
// copy `n` groups of statistics (device pointer, e.g. the freshly uploaded scratch) into array a from group g0
int mirror_stats(dist_b200_feature *f, int a, const void *src_dev, int g0, int n, cudaStream_t s) {
    if (!f->stats || n <= 0) return DIST_B200_OK;
    const size_t e = stat_elems(f, a);
    DISTB200_CUDA(f->ctx, cudaMemcpyAsync(stat_ptr(f, a) + e * g0, src_dev, e * n * 4, cudaMemcpyDeviceToDevice, s));
    return DIST_B200_OK;
}

// (re)allocate the hot-layout buffer for `G` groups, zero-filled, padded to 128 groups so that the
// score kernel may stage whole register tiles
int ensure_params(dist_b200_feature *f, int G, bool keep) {
    dist_b200_ctx *ctx = f->ctx;
    size_t bytes;
    const int cap = static_cast<int>(round_up(static_cast<size_t>(std::max(G, 1)), 128));
    if (f->model == DIST_B200_DPD) {
        bytes = sizeof(float) * static_cast<size_t>(f->dim + 2) * G;  // table rows + OTHER row + shift row
    } else {
        bytes = sizeof(float) * group_floats(f) * cap;
    }
    const int old_capacity = f->capacity, old_G = f->G;
    if (bytes <= f->params_bytes && (f->model == DIST_B200_DPD || cap <= f->capacity))
        return ensure_stats(f, f->stats ? old_capacity : 0, old_G, keep);
    void *fresh = nullptr;
    const size_t want = f->model == DIST_B200_DPD ? bytes : sizeof(float) * group_floats(f) * round_up(cap + cap / 2, 128);
    DISTB200_CUDA(ctx, cudaMalloc(&fresh, want));
    DISTB200_CUDA(ctx, cudaMemset(fresh, 0, want));
    if (f->params) {
        if (keep) DISTB200_CUDA(ctx, cudaMemcpy(fresh, f->params, std::min(want, f->params_bytes), cudaMemcpyDeviceToDevice));
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        DISTB200_CUDA(ctx, cudaFree(f->params));
    }
    f->params = fresh;
    f->params_bytes = want;
    f->capacity = f->model == DIST_B200_DPD ? G : static_cast<int>(want / (sizeof(float) * group_floats(f)));
    if (f->model == DIST_B200_NICH) {
        float *aux = nullptr;
        DISTB200_CUDA(ctx, cudaMalloc(&aux, sizeof(float) * f->capacity));
        DISTB200_CUDA(ctx, cudaMemset(aux, 0, sizeof(float) * f->capacity));
        if (f->aux) {
            if (keep) DISTB200_CUDA(ctx, cudaMemcpy(aux, f->aux, sizeof(float) * std::min(f->capacity, f->G), cudaMemcpyDeviceToDevice));
            DISTB200_CUDA(ctx, cudaFree(f->aux));
        }
        f->aux = aux;
    }
    return ensure_stats(f, old_capacity, old_G, keep);
}

// copy host arrays into consecutive, 256-byte aligned regions of the context scratch
struct Upload {
    dist_b200_ctx *ctx;
    cudaStream_t s;
    size_t off = 0;
    int err = DIST_B200_OK;
    template <class T>
    const T *put(const T *host, size_t n) {
        if (err) return nullptr;
        char *dst = static_cast<char *>(ctx->scratch_dev) + off;
        cudaError_t e = cudaMemcpyAsync(dst, host, n * sizeof(T), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) {
            err = fail(ctx, DIST_B200_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e));
            return nullptr;
        }
        off += round_up(n * sizeof(T), 256);
        return reinterpret_cast<const T *>(dst);
    }
};

// dpd storage: table (V + 2) x capacity floats (dense stride G inside it), statistics counts[capacity][V] | betas[V]
uint32_t *dpd_betas(const dist_b200_feature *f) { return f->stats + static_cast<size_t>(f->capacity) * f->dim; }
int dpd_reserve(dist_b200_feature *f, int G, int V, bool keep) {
    dist_b200_ctx *ctx = f->ctx;
    const int need = std::max(G, 1);
    const bool fits = f->params && f->stats && need <= f->capacity &&
                      f->params_bytes >= sizeof(float) * static_cast<size_t>(V + 2) * f->capacity &&
                      f->stats_words >= static_cast<size_t>(f->capacity) * V + V;
    if (fits) return DIST_B200_OK;
    const int cap = static_cast<int>(round_up(static_cast<size_t>(need) + need / 4, 8));
    void *params = nullptr;
    uint32_t *stats = nullptr;
    const size_t pbytes = sizeof(float) * static_cast<size_t>(V + 2) * cap, words = static_cast<size_t>(cap) * V + V;
    DISTB200_CUDA(ctx, cudaMalloc(&params, pbytes));
    DISTB200_CUDA(ctx, cudaMalloc(&stats, words * 4));
    DISTB200_CUDA(ctx, cudaMemset(stats, 0, words * 4));
    DISTB200_CUDA(ctx, cudaDeviceSynchronize());
    if (keep && f->stats && f->G > 0) {  // counts rows are contiguous; betas move with the capacity
        DISTB200_CUDA(ctx, cudaMemcpy(stats, f->stats, sizeof(int32_t) * static_cast<size_t>(f->G) * V, cudaMemcpyDeviceToDevice));
        DISTB200_CUDA(ctx, cudaMemcpy(stats + static_cast<size_t>(cap) * V, dpd_betas(f), sizeof(float) * V, cudaMemcpyDeviceToDevice));
    }
    if (f->params) DISTB200_CUDA(ctx, cudaFree(f->params));
    if (f->stats) DISTB200_CUDA(ctx, cudaFree(f->stats));
    f->params = params;
    f->params_bytes = pbytes;
    f->stats = stats;
    f->stats_words = words;
    f->capacity = cap;
    return DIST_B200_OK;
}

// dpd caches: the canonical value-major table (dpd.hpp:471-497) + the scratch table_rows.cu re-lays it out into per call
int dpd_rebuild(dist_b200_feature *f, const float *betas_dev, const int32_t *counts_dev, cudaStream_t s) {
    dist_b200_ctx *ctx = f->ctx;
    int rc = launch_dpd_prep(ctx, f->alpha, f->beta0, f->dim, betas_dev, f->G, counts_dev, static_cast<float *>(f->params), s);
    if (rc || f->G < 1) return rc;
    const size_t need = table_hot_floats(f->dim + 1, f->G);
    if (need > f->dpd_hot_floats) {
        if (f->dpd_hot) {
            DISTB200_CUDA(ctx, cudaDeviceSynchronize());
            DISTB200_CUDA(ctx, cudaFree(f->dpd_hot));
            f->dpd_hot = nullptr;
            f->dpd_hot_floats = 0;
        }
        DISTB200_CUDA(ctx, cudaMalloc(&f->dpd_hot, need * sizeof(float)));
        f->dpd_hot_floats = need;
    }
    return DIST_B200_OK;
}

bool check_feature(const dist_b200_feature *f, int model) { return f && f->ctx && f->model == model; }

// cache mutations are asynchronous on the caller's stream: mark their completion ...
int mark_ready(dist_b200_feature *f, int rc, cudaStream_t s) {
    if (rc != DIST_B200_OK) return rc;
    DISTB200_CUDA(f->ctx, cudaEventRecord(f->ready, s));
    // the statistics were staged through the context's scratch buffer: drain the stream so the next
    // host-side upload (possibly on another stream) cannot overwrite it under the prep kernel
    DISTB200_CUDA(f->ctx, cudaStreamSynchronize(s));
    return DIST_B200_OK;
}
// ... and make a scoring stream wait for them (no-op when it is the same stream)
int wait_ready(dist_b200_ctx *ctx, const dist_b200_feature *const *features, int F, cudaStream_t s) {
    for (int f = 0; f < F; ++f)
        if (features[f] && features[f]->ready) DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, features[f]->ready, 0));
    return DIST_B200_OK;
}

}  // namespace

extern "C" {

int dist_b200_abi_version(void) { return DIST_B200_ABI_VERSION; }

int dist_b200_ctx_create(int device, dist_b200_ctx **out) {
    if (!out) return DIST_B200_ERR_INVALID;
    *out = nullptr;
    dist_b200_ctx *ctx = new (std::nothrow) dist_b200_ctx();
    if (!ctx) return DIST_B200_ERR_CUDA;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        delete ctx;
        return DIST_B200_ERR_CUDA;
    }
    // numerics tables.  fast_log's table is rebuilt the way the reference builds it
    // (special.cc:35-44: log2 of 1 + i / 2^14, float argument); the polynomial / factorial tables are
    // the reference's data (ref_tables.inc, generated by oracle/gen_tables.py).
    const size_t n_log = 1 << 14, n_lg = 33 * kLgammaRowStride, n_nu = 18 * 4, n_lf = 64;
    std::vector<float> host(n_log + n_lg + n_nu + n_lf, 0.f);
    for (size_t i = 0; i < n_log; ++i) {
        const float scaled = static_cast<float>(i) * static_cast<float>(1 << 9);
        const float v = static_cast<float>(1.0 + static_cast<double>(scaled / static_cast<float>(1 << 23)));
        host[i] = log2f(v);
    }
    for (int c = 0; c < 33; ++c)
        for (int k = 0; k < 6; ++k) std::memcpy(&host[n_log + c * kLgammaRowStride + k], &kLgammaCoeff5Bits[c * 6 + k], 4);
    std::memcpy(&host[n_log + n_lg], kLgammaNuCoeff3Bits, sizeof(float) * n_nu);
    std::memcpy(&host[n_log + n_lg + n_nu], kLogFactorialBits, sizeof(float) * n_lf);
    e = cudaMalloc(&ctx->tables_storage, host.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->tables_storage, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        if (ctx->tables_storage) cudaFree(ctx->tables_storage);
        delete ctx;
        return DIST_B200_ERR_CUDA;
    }
    ctx->tables.log2_table = ctx->tables_storage;
    ctx->tables.lgamma5 = ctx->tables_storage + n_log;
    ctx->tables.lgamma_nu3 = ctx->tables_storage + n_log + n_lg;
    ctx->tables.log_factorial = ctx->tables_storage + n_log + n_lg + n_nu;
    *out = ctx;
    return DIST_B200_OK;
}

void dist_b200_ctx_destroy(dist_b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->tables_storage) cudaFree(ctx->tables_storage);
    if (ctx->scratch_dev) cudaFree(ctx->scratch_dev);
    if (ctx->add_acc) cudaFree(ctx->add_acc);
    if (ctx->add_done) cudaEventDestroy(ctx->add_done);
    if (ctx->scores_scratch) cudaFree(ctx->scores_scratch);
    if (ctx->xpack) cudaFree(ctx->xpack);
    if (ctx->bbt) cudaFree(ctx->bbt);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->own_stream2) cudaStreamDestroy(ctx->own_stream2);
    if (ctx->ev) cudaEventDestroy(ctx->ev);
    delete ctx;
}

int dist_b200_ctx_set_option(dist_b200_ctx *ctx, int option, int value) {
    if (!ctx || option < 0 || option >= DIST_B200_OPT_COUNT_) return DIST_B200_ERR_INVALID;
    ctx->opt[option] = value;
    return DIST_B200_OK;
}

const char *dist_b200_last_error(const dist_b200_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
int dist_b200_sm_count(const dist_b200_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

int dist_b200_feature_create(dist_b200_ctx *ctx, int model, dist_b200_feature **out) {
    if (!ctx || !out) return DIST_B200_ERR_INVALID;
    *out = nullptr;
    if (model < DIST_B200_DD || model > DIST_B200_BNB) return fail(ctx, DIST_B200_ERR_INVALID, "unknown model");
    dist_b200_feature *f = new (std::nothrow) dist_b200_feature();
    if (!f) return fail(ctx, DIST_B200_ERR_CUDA, "out of host memory");
    f->ctx = ctx;
    f->model = model;
    if (cudaEventCreateWithFlags(&f->ready, cudaEventDisableTiming) != cudaSuccess) {
        delete f;
        return fail(ctx, DIST_B200_ERR_CUDA, "cudaEventCreate failed");
    }
    *out = f;
    return DIST_B200_OK;
}

void dist_b200_feature_destroy(dist_b200_feature *f) {
    if (!f) return;
    cudaDeviceSynchronize();
    if (f->params) cudaFree(f->params);
    if (f->aux) cudaFree(f->aux);
    if (f->gp_table) cudaFree(f->gp_table);
    if (f->keys_dev) cudaFree(f->keys_dev);
    if (f->key_rows_dev) cudaFree(f->key_rows_dev);
    if (f->dpd_hot) cudaFree(f->dpd_hot);
    if (f->cdf_buf) cudaFree(f->cdf_buf);
    if (f->niw_buf) cudaFree(f->niw_buf);
    if (f->niw_tc) cudaFree(f->niw_tc);
    if (f->niw_stats) cudaFree(f->niw_stats);
    if (f->niw_shared_dev) cudaFree(f->niw_shared_dev);
    if (f->stats) cudaFree(f->stats);
    if (f->alphas_dev) cudaFree(f->alphas_dev);
    if (f->log_prod_dev) cudaFree(f->log_prod_dev);
    if (f->ready) cudaEventDestroy(f->ready);
    delete f;
}

int dist_b200_feature_model(const dist_b200_feature *f) { return f ? f->model : -1; }
int dist_b200_feature_groups(const dist_b200_feature *f) { return f ? f->G : 0; }

int dist_b200_nich_update_all(dist_b200_feature *f, const float shared[4], int G, const int32_t *count,
                              const float *mean, const float *ctv, void *stream) {
    if (!check_feature(f, DIST_B200_NICH) || !shared || G < 0 || (G && (!count || !mean || !ctv))) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    std::memcpy(f->shared, shared, sizeof(float) * 4);
    int rc = ensure_params(f, G, false);
    if (rc) return rc;
    if ((rc = ensure_scratch(ctx, 3 * round_up(sizeof(float) * G, 256) + 256))) return rc;
    Upload up{ctx, as_stream(stream)};
    const int32_t *c = up.put(count, G);
    const float *m = up.put(mean, G);
    const float *v = up.put(ctv, G);
    if (up.err) return up.err;
    f->G = G;
    if ((rc = mirror_stats(f, 0, c, 0, G, as_stream(stream))) || (rc = mirror_stats(f, 1, m, 0, G, as_stream(stream))) ||
        (rc = mirror_stats(f, 2, v, 0, G, as_stream(stream))))
        return rc;
    return mark_ready(f, launch_nich_prep(ctx, f->shared, G, 0, G, c, m, v, static_cast<float4 *>(f->params), f->aux, as_stream(stream)), as_stream(stream));
}

int dist_b200_gp_update_all(dist_b200_feature *f, const float shared[2], int G, const uint32_t *count,
                            const uint32_t *sum, void *stream) {
    if (!check_feature(f, DIST_B200_GP) || !shared || G < 0 || (G && (!count || !sum))) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    f->shared[0] = shared[0];
    f->shared[1] = shared[1];
    int rc = ensure_params(f, G, false);
    if (rc) return rc;
    if ((rc = ensure_scratch(ctx, 2 * round_up(sizeof(float) * G, 256) + 256))) return rc;
    Upload up{ctx, as_stream(stream)};
    const uint32_t *c = up.put(count, G);
    const uint32_t *sm = up.put(sum, G);
    if (up.err) return up.err;
    f->G = G;
    f->gp_table_dirty = true;
    f->log_prod_valid = false;
    if ((rc = mirror_stats(f, 0, c, 0, G, as_stream(stream))) || (rc = mirror_stats(f, 1, sm, 0, G, as_stream(stream)))) return rc;
    return mark_ready(f, launch_gp_prep(ctx, f->shared, 0, G, c, sm, static_cast<float4 *>(f->params), as_stream(stream)), as_stream(stream));
}

int dist_b200_bnb_update_all(dist_b200_feature *f, const float shared[2], uint32_t r, int G, const uint32_t *count,
                             const uint32_t *sum, void *stream) {
    if (!check_feature(f, DIST_B200_BNB) || !shared || G < 0 || (G && (!count || !sum))) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    f->shared[0] = shared[0];
    f->shared[1] = shared[1];
    f->shared[2] = static_cast<float>(r);  // the reference converts r to float wherever it enters (bnb.hpp:59,208)
    int rc = ensure_params(f, G, false);
    if (rc) return rc;
    if ((rc = ensure_scratch(ctx, 2 * round_up(sizeof(float) * G, 256) + 256))) return rc;
    Upload up{ctx, as_stream(stream)};
    const uint32_t *c = up.put(count, G);
    const uint32_t *sm = up.put(sum, G);
    if (up.err) return up.err;
    f->G = G;
    if ((rc = mirror_stats(f, 0, c, 0, G, as_stream(stream))) || (rc = mirror_stats(f, 1, sm, 0, G, as_stream(stream)))) return rc;
    return mark_ready(f, launch_bnb_prep(ctx, f->shared, 0, G, c, sm, static_cast<float4 *>(f->params), as_stream(stream)), as_stream(stream));
}

int dist_b200_bb_update_all(dist_b200_feature *f, const float shared[2], int G, const int32_t *heads,
                            const int32_t *tails, void *stream) {
    if (!check_feature(f, DIST_B200_BB) || !shared || G < 0 || (G && (!heads || !tails))) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    f->shared[0] = shared[0];
    f->shared[1] = shared[1];
    int rc = ensure_params(f, G, false);
    if (rc) return rc;
    if ((rc = ensure_scratch(ctx, 2 * round_up(sizeof(float) * G, 256) + 256))) return rc;
    Upload up{ctx, as_stream(stream)};
    const int32_t *h = up.put(heads, G);
    const int32_t *t = up.put(tails, G);
    if (up.err) return up.err;
    f->G = G;
    if ((rc = mirror_stats(f, 0, h, 0, G, as_stream(stream))) || (rc = mirror_stats(f, 1, t, 0, G, as_stream(stream)))) return rc;
    return mark_ready(f, launch_bb_prep(ctx, f->shared, 0, G, h, t, static_cast<float4 *>(f->params), as_stream(stream)), as_stream(stream));
}

int dist_b200_dd_update_all(dist_b200_feature *f, int dim, const float *alphas, int G, const int32_t *counts,
                            void *stream) {
    if (!check_feature(f, DIST_B200_DD) || dim < 1 || !alphas || G < 0 || (G && !counts)) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (dim > 256) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "dd: dim > 256");
    if (dim != f->dim) {  // layout changes: drop the old buffer
        if (f->params) {
            DISTB200_CUDA(ctx, cudaDeviceSynchronize());
            DISTB200_CUDA(ctx, cudaFree(f->params));
        }
        f->params = nullptr;
        f->params_bytes = 0;
        f->capacity = 0;
    }
    f->dim = dim;
    f->alphas.assign(alphas, alphas + dim);
    float alpha_sum = 0;  // dd.hpp:404-407, sequential float sum
    for (int v = 0; v < dim; ++v) alpha_sum += alphas[v];
    f->alpha_sum = alpha_sum;
    int rc = ensure_params(f, G, false);
    if (rc) return rc;
    if ((rc = ensure_scratch(ctx, round_up(sizeof(float) * dim, 256) + round_up(sizeof(int32_t) * static_cast<size_t>(G) * dim, 256) + 256))) return rc;
    Upload up{ctx, as_stream(stream)};
    const float *a = up.put(alphas, dim);
    const int32_t *c = up.put(counts, static_cast<size_t>(G) * dim);
    if (up.err) return up.err;
    f->G = G;
    if (!f->alphas_dev) DISTB200_CUDA(ctx, cudaMalloc(&f->alphas_dev, sizeof(float) * 256));
    DISTB200_CUDA(ctx, cudaMemcpyAsync(f->alphas_dev, a, sizeof(float) * dim, cudaMemcpyDeviceToDevice, as_stream(stream)));
    if ((rc = mirror_stats(f, 0, c, 0, G, as_stream(stream)))) return rc;
    return mark_ready(f, launch_dd_prep(ctx, dim, a, alpha_sum, 0, G, c, static_cast<float *>(f->params), as_stream(stream)), as_stream(stream));
}

int dist_b200_dpd_update_all(dist_b200_feature *f, float alpha, float beta0, int V, const uint32_t *keys,
                             const float *betas, int G, const int32_t *counts, void *stream) {
    if (!check_feature(f, DIST_B200_DPD) || V < 1 || !keys || !betas || G < 0 || (G && !counts)) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    f->alpha = alpha;
    f->beta0 = beta0;
    // value -> table row: identity when keys are 0..V-1, else sorted keys + binary search on device
    bool dense = true;
    for (int v = 0; v < V; ++v) dense = dense && keys[v] == static_cast<uint32_t>(v);
    if (V != f->dim || !f->keys_dev || !std::equal(keys, keys + V, f->keys.begin()) ) {
        std::vector<int> order(V);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int x, int y) { return keys[x] < keys[y]; });
        std::vector<uint32_t> sorted(V);
        for (int i = 0; i < V; ++i) sorted[i] = keys[order[i]];
        for (int i = 1; i < V; ++i)
            if (sorted[i] == sorted[i - 1]) return fail(ctx, DIST_B200_ERR_INVALID, "dpd: duplicate key");
        if (f->keys_dev) {
            DISTB200_CUDA(ctx, cudaDeviceSynchronize());
            DISTB200_CUDA(ctx, cudaFree(f->keys_dev));
            DISTB200_CUDA(ctx, cudaFree(f->key_rows_dev));
            f->keys_dev = nullptr;
            f->key_rows_dev = nullptr;
        }
        DISTB200_CUDA(ctx, cudaMalloc(&f->keys_dev, sizeof(uint32_t) * V));
        DISTB200_CUDA(ctx, cudaMalloc(&f->key_rows_dev, sizeof(int) * V));
        DISTB200_CUDA(ctx, cudaMemcpy(f->keys_dev, sorted.data(), sizeof(uint32_t) * V, cudaMemcpyHostToDevice));
        DISTB200_CUDA(ctx, cudaMemcpy(f->key_rows_dev, order.data(), sizeof(int) * V, cudaMemcpyHostToDevice));
        f->keys.assign(keys, keys + V);
    }
    f->keys_dense = dense;
    f->dim = V;
    f->alphas.assign(betas, betas + V);
    int rc = dpd_reserve(f, G, V, false);
    if (rc) return rc;
    if ((rc = ensure_scratch(ctx, round_up(sizeof(float) * V, 256) + round_up(sizeof(int32_t) * static_cast<size_t>(G) * V, 256) + 256))) return rc;
    Upload up{ctx, as_stream(stream)};
    const float *b = up.put(betas, V);
    const int32_t *c = up.put(counts, static_cast<size_t>(G) * V);
    if (up.err) return up.err;
    f->G = G;
    DISTB200_CUDA(ctx, cudaMemcpyAsync(f->stats, c, sizeof(int32_t) * static_cast<size_t>(G) * V, cudaMemcpyDeviceToDevice, as_stream(stream)));
    DISTB200_CUDA(ctx, cudaMemcpyAsync(dpd_betas(f), b, sizeof(float) * V, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return mark_ready(f, dpd_rebuild(f, b, c, as_stream(stream)), as_stream(stream));
}

static int niw_reserve(dist_b200_feature *f, int G, int keep);
static int32_t *niw_count(const dist_b200_feature *f) { return reinterpret_cast<int32_t *>(f->niw_stats); }
static float *niw_sum_x(const dist_b200_feature *f) { return reinterpret_cast<float *>(f->niw_stats) + f->niw_cap; }
static float *niw_sum_xxT(const dist_b200_feature *f) {
    return reinterpret_cast<float *>(f->niw_stats) + f->niw_cap + static_cast<size_t>(f->niw_cap) * f->dim;
}
// records of groups [g0, g0 + n) (and the tensor-core images) from the resident statistics
static int niw_rebuild(dist_b200_feature *f, int g0, int n, cudaStream_t s) {
    dist_b200_ctx *ctx = f->ctx;
    const int d = f->dim;
    const size_t dd = static_cast<size_t>(d) * d;
    const size_t rec = static_cast<size_t>(niw_padded_dim(d)) * (niw_padded_dim(d) + 1) + 4;
    int rc = launch_niw_prep(ctx, d, f->niw_shared_dev, f->kappa, f->niw_shared_dev + d, f->nu, n, niw_count(f) + g0,
                             niw_sum_x(f) + static_cast<size_t>(g0) * d, niw_sum_xxT(f) + g0 * dd, f->niw_buf + rec * g0, s);
    if (rc == DIST_B200_OK && d == 32 && f->niw_tc && f->G > 0) rc = launch_niw_tc_prep(ctx, f->G, f->niw_buf, f->niw_tc, s);
    return rc;
}

int dist_b200_niw_update_all(dist_b200_feature *f, int d, const float *mu, float kappa, const float *psi,
                             float nu, int G, const int32_t *count, const float *sum_x, const float *sum_xxT,
                             void *stream) {
    if (!check_feature(f, DIST_B200_NIW) || d < 1 || !mu || !psi || G < 0 || (G && (!count || !sum_x || !sum_xxT)))
        return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (d > 32) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "niw: d > 32");
    if (!(kappa > 0.f) || !(nu > static_cast<float>(d) - 1.f))
        return fail(ctx, DIST_B200_ERR_INVALID, "niw: need kappa > 0 and nu > d - 1 (niw.hpp:121,132)");
    if (f->niw_stats && f->dim != d) {  // the statistics layout depends on d
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        DISTB200_CUDA(ctx, cudaFree(f->niw_stats));
        f->niw_stats = nullptr;
        f->niw_cap = 0;
    }
    f->dim = d;
    {
        int rcr = niw_reserve(f, G, 0);
        if (rcr) return rcr;
    }
    const size_t dd = static_cast<size_t>(d) * d;
    int rc = ensure_scratch(ctx, round_up(4 * d, 256) + round_up(4 * dd, 256) + round_up(4 * static_cast<size_t>(G), 256) +
                                     round_up(4 * static_cast<size_t>(G) * d, 256) + round_up(4 * G * dd, 256) + 256);
    if (rc) return rc;
    Upload up{ctx, as_stream(stream)};
    const float *mu_d = up.put(mu, d);
    const float *psi_d = up.put(psi, dd);
    const int32_t *cnt_d = up.put(count, G);
    const float *sx_d = up.put(sum_x, static_cast<size_t>(G) * d);
    const float *sxx_d = up.put(sum_xxT, static_cast<size_t>(G) * dd);
    if (up.err) return up.err;
    f->dim = d;
    f->G = G;
    f->kappa = kappa;
    f->nu = nu;
    f->mu.assign(mu, mu + d);
    f->psi.assign(psi, psi + dd);
    cudaStream_t s = as_stream(stream);
    DISTB200_CUDA(ctx, cudaMemcpyAsync(f->niw_shared_dev, mu_d, sizeof(float) * d, cudaMemcpyDeviceToDevice, s));
    DISTB200_CUDA(ctx, cudaMemcpyAsync(f->niw_shared_dev + d, psi_d, sizeof(float) * dd, cudaMemcpyDeviceToDevice, s));
    if (G > 0) {
        DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_count(f), cnt_d, sizeof(int32_t) * G, cudaMemcpyDeviceToDevice, s));
        DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_sum_x(f), sx_d, sizeof(float) * G * d, cudaMemcpyDeviceToDevice, s));
        DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_sum_xxT(f), sxx_d, sizeof(float) * G * dd, cudaMemcpyDeviceToDevice, s));
    }
    return mark_ready(f, niw_rebuild(f, 0, G, s), s);
}

// NormalInverseWishart: one group's record (posterior, whitening matrix, constants) from its raw statistics
// {int32 count; float sum_x[d]; float sum_xxT[d][d]} (niw.hpp:187-190, :247-276), then the tensor-core block records
static int niw_update_group(dist_b200_feature *f, int groupid, const void *stats, cudaStream_t s) {
    dist_b200_ctx *ctx = f->ctx;
    const int d = f->dim;
    const size_t dd = static_cast<size_t>(d) * d;
    if (f->mu.size() != static_cast<size_t>(d) || f->psi.size() != dd) return fail(ctx, DIST_B200_ERR_STATE, "niw update_group: call update_all first");
    if (!f->niw_stats || groupid >= f->niw_cap) return fail(ctx, DIST_B200_ERR_STATE, "niw update_group: call update_all first");
    int rc = ensure_scratch(ctx, round_up(4 * d, 256) + round_up(4 * dd, 256) + 512);
    if (rc) return rc;
    const char *p = static_cast<const char *>(stats);
    Upload up{ctx, s};
    const int32_t *cnt_d = up.put(reinterpret_cast<const int32_t *>(p), 1);
    const float *sx_d = up.put(reinterpret_cast<const float *>(p + 4), d);
    const float *sxx_d = up.put(reinterpret_cast<const float *>(p + 4 + 4 * d), dd);
    if (up.err) return up.err;
    DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, f->ready, 0));
    DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_count(f) + groupid, cnt_d, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_sum_x(f) + static_cast<size_t>(groupid) * d, sx_d, sizeof(float) * d, cudaMemcpyDeviceToDevice, s));
    DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_sum_xxT(f) + groupid * dd, sxx_d, sizeof(float) * dd, cudaMemcpyDeviceToDevice, s));
    return mark_ready(f, niw_rebuild(f, groupid, 1, s), s);
}

// (re)allocate the niw record buffers for G groups, keeping the first `keep` records
static int niw_reserve(dist_b200_feature *f, int G, int keep) {
    dist_b200_ctx *ctx = f->ctx;
    const int d = f->dim;
    const size_t rec = static_cast<size_t>(niw_padded_dim(d)) * (niw_padded_dim(d) + 1) + 4;
    const size_t bytes = sizeof(float) * rec * std::max(G, 1);
    if (bytes > f->niw_bytes) {
        const size_t want = bytes + bytes / 4;
        float *fresh = nullptr;
        DISTB200_CUDA(ctx, cudaMalloc(&fresh, want));
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        if (f->niw_buf) {
            if (keep > 0) DISTB200_CUDA(ctx, cudaMemcpy(fresh, f->niw_buf, sizeof(float) * rec * keep, cudaMemcpyDeviceToDevice));
            DISTB200_CUDA(ctx, cudaFree(f->niw_buf));
        }
        f->niw_buf = fresh;
        f->niw_bytes = want;
    }
    if (!f->niw_shared_dev) DISTB200_CUDA(ctx, cudaMalloc(&f->niw_shared_dev, sizeof(float) * (32 + 32 * 32)));
    if (G > f->niw_cap || !f->niw_stats) {
        const int cap = std::max(G, 1) + std::max(G, 1) / 4 + 8;
        const size_t dd = static_cast<size_t>(d) * d;
        uint32_t *fresh = nullptr;
        DISTB200_CUDA(ctx, cudaMalloc(&fresh, sizeof(uint32_t) * static_cast<size_t>(cap) * (1 + d + dd)));
        DISTB200_CUDA(ctx, cudaMemset(fresh, 0, sizeof(uint32_t) * static_cast<size_t>(cap) * (1 + d + dd)));
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        if (f->niw_stats) {
            if (keep > 0) {
                DISTB200_CUDA(ctx, cudaMemcpy(fresh, niw_count(f), sizeof(int32_t) * keep, cudaMemcpyDeviceToDevice));
                DISTB200_CUDA(ctx, cudaMemcpy(fresh + cap, niw_sum_x(f), sizeof(float) * keep * d, cudaMemcpyDeviceToDevice));
                DISTB200_CUDA(ctx, cudaMemcpy(fresh + cap + static_cast<size_t>(cap) * d, niw_sum_xxT(f), sizeof(float) * keep * dd, cudaMemcpyDeviceToDevice));
            }
            DISTB200_CUDA(ctx, cudaFree(f->niw_stats));
        }
        f->niw_stats = fresh;
        f->niw_cap = cap;
    }
    if (d == 32) {
        const size_t tc_bytes = sizeof(float) * niw_tc_floats(std::max(G, 1));
        if (tc_bytes > f->niw_tc_bytes) {
            if (f->niw_tc) {
                DISTB200_CUDA(ctx, cudaDeviceSynchronize());
                DISTB200_CUDA(ctx, cudaFree(f->niw_tc));
                f->niw_tc = nullptr;
            }
            DISTB200_CUDA(ctx, cudaMalloc(&f->niw_tc, tc_bytes + tc_bytes / 4));
            f->niw_tc_bytes = tc_bytes + tc_bytes / 4;
        }
    }
    return DIST_B200_OK;
}

int dist_b200_feature_update_group(dist_b200_feature *f, int groupid, const void *stats, void *stream) {
    if (!f || !f->ctx || !stats) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (groupid < 0 || groupid >= f->G) return fail(ctx, DIST_B200_ERR_INVALID, "update_group: bad groupid");
    int rc = ensure_scratch(ctx, 4096 + sizeof(int32_t) * 256);
    if (rc) return rc;
    cudaStream_t s = as_stream(stream);
    DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, f->ready, 0));
    Upload up{ctx, s};
    switch (f->model) {
        case DIST_B200_NICH: {
            const char *p = static_cast<const char *>(stats);
            const int32_t *c = up.put(reinterpret_cast<const int32_t *>(p), 1);
            const float *m = up.put(reinterpret_cast<const float *>(p + 4), 1);
            const float *v = up.put(reinterpret_cast<const float *>(p + 8), 1);
            if (up.err) return up.err;
            if ((rc = mirror_stats(f, 0, c, groupid, 1, s)) || (rc = mirror_stats(f, 1, m, groupid, 1, s)) ||
                (rc = mirror_stats(f, 2, v, groupid, 1, s)))
                return rc;
            return mark_ready(f, launch_nich_prep(ctx, f->shared, f->G, groupid, 1, c, m, v, static_cast<float4 *>(f->params), f->aux, s), s);
        }
        case DIST_B200_GP: {
            const uint32_t *p = static_cast<const uint32_t *>(stats);
            const uint32_t *c = up.put(p, 1);
            const uint32_t *sm = up.put(p + 1, 1);
            if (up.err) return up.err;
            f->gp_table_dirty = true;
            f->log_prod_valid = false;
            if ((rc = mirror_stats(f, 0, c, groupid, 1, s)) || (rc = mirror_stats(f, 1, sm, groupid, 1, s))) return rc;
            return mark_ready(f, launch_gp_prep(ctx, f->shared, groupid, 1, c, sm, static_cast<float4 *>(f->params), s), s);
        }
        case DIST_B200_BNB: {
            const uint32_t *p = static_cast<const uint32_t *>(stats);
            const uint32_t *c = up.put(p, 1);
            const uint32_t *sm = up.put(p + 1, 1);
            if (up.err) return up.err;
            if ((rc = mirror_stats(f, 0, c, groupid, 1, s)) || (rc = mirror_stats(f, 1, sm, groupid, 1, s))) return rc;
            return mark_ready(f, launch_bnb_prep(ctx, f->shared, groupid, 1, c, sm, static_cast<float4 *>(f->params), s), s);
        }
        case DIST_B200_BB: {
            const int32_t *p = static_cast<const int32_t *>(stats);
            const int32_t *h = up.put(p, 1);
            const int32_t *t = up.put(p + 1, 1);
            if (up.err) return up.err;
            if ((rc = mirror_stats(f, 0, h, groupid, 1, s)) || (rc = mirror_stats(f, 1, t, groupid, 1, s))) return rc;
            return mark_ready(f, launch_bb_prep(ctx, f->shared, groupid, 1, h, t, static_cast<float4 *>(f->params), s), s);
        }
        case DIST_B200_DD: {
            const float *a = up.put(f->alphas.data(), f->dim);
            const int32_t *c = up.put(static_cast<const int32_t *>(stats), f->dim);
            if (up.err) return up.err;
            if ((rc = mirror_stats(f, 0, c, groupid, 1, s))) return rc;
            return mark_ready(f, launch_dd_prep(ctx, f->dim, a, f->alpha_sum, groupid, 1, c, static_cast<float *>(f->params), s), s);
        }
        case DIST_B200_DPD: {  // stats: the group's dense counts[V], in update_all's key order
            const int V = f->dim;
            if ((rc = ensure_scratch(ctx, round_up(sizeof(int32_t) * V, 256) + 256))) return rc;
            Upload up2{ctx, s};
            const int32_t *c = up2.put(static_cast<const int32_t *>(stats), V);
            if (up2.err) return up2.err;
            int32_t *row = reinterpret_cast<int32_t *>(f->stats) + static_cast<size_t>(groupid) * V;
            DISTB200_CUDA(ctx, cudaMemcpyAsync(row, c, sizeof(int32_t) * V, cudaMemcpyDeviceToDevice, s));
            return mark_ready(f, launch_dpd_update_group(ctx, f->alpha, f->beta0, V, reinterpret_cast<const float *>(dpd_betas(f)), f->G, groupid,
                                                         row, static_cast<float *>(f->params), s), s);
        }
        case DIST_B200_NIW: return niw_update_group(f, groupid, stats, s);
        default:
            return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "update_group: unknown model");
    }
}

int dist_b200_feature_add_group(dist_b200_feature *f, void *stream) {
    if (!f || !f->ctx) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (f->model == DIST_B200_DPD) {
        // a fresh empty group: one more zero counts row; the dense table's stride is G, so it is rebuilt (O(V G))
        if (f->G < 1 && !f->stats) return fail(ctx, DIST_B200_ERR_STATE, "add_group: call update_all first");
        int rc = dpd_reserve(f, f->G + 1, f->dim, true);
        if (rc) return rc;
        cudaStream_t s = as_stream(stream);
        DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, f->ready, 0));
        DISTB200_CUDA(ctx, cudaMemsetAsync(f->stats + static_cast<size_t>(f->G) * f->dim, 0, sizeof(int32_t) * f->dim, s));
        f->G += 1;
        return mark_ready(f, dpd_rebuild(f, reinterpret_cast<const float *>(dpd_betas(f)), reinterpret_cast<const int32_t *>(f->stats), s), s);
    }
    if (f->model == DIST_B200_NIW) {
        if (f->mu.empty()) return fail(ctx, DIST_B200_ERR_STATE, "add_group: call update_all first");
        int rc = niw_reserve(f, f->G + 1, f->G);
        if (rc) return rc;
        f->G += 1;
        std::vector<float> zeros(1 + f->dim + static_cast<size_t>(f->dim) * f->dim, 0.f);  // Group::init: count = 0, sums = 0
        return niw_update_group(f, f->G - 1, zeros.data(), as_stream(stream));
    }
    int rc = ensure_params(f, f->G + 1, true);
    if (rc) return rc;
    f->G += 1;
    // a fresh group is Group::init (all-zero statistics) followed by update_group (mixture.hpp:361-369)
    int32_t zeros[256];
    std::memset(zeros, 0, sizeof(zeros));
    return dist_b200_feature_update_group(f, f->G - 1, zeros, stream);
}

int dist_b200_feature_remove_group(dist_b200_feature *f, int groupid, void *stream) {
    if (!f || !f->ctx) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (groupid < 0 || groupid >= f->G) return fail(ctx, DIST_B200_ERR_INVALID, "remove_group: bad groupid");
    if (f->model == DIST_B200_DPD) {  // packed_remove on the counts rows, then the dense table again
        cudaStream_t s = as_stream(stream);
        DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, f->ready, 0));
        const int last = f->G - 1, V = f->dim;
        if (groupid != last)
            DISTB200_CUDA(ctx, cudaMemcpyAsync(f->stats + static_cast<size_t>(groupid) * V, f->stats + static_cast<size_t>(last) * V,
                                               sizeof(int32_t) * V, cudaMemcpyDeviceToDevice, s));
        f->G = last;
        return mark_ready(f, dpd_rebuild(f, reinterpret_cast<const float *>(dpd_betas(f)), reinterpret_cast<const int32_t *>(f->stats), s), s);
    }
    if (f->model == DIST_B200_NIW) {
        cudaStream_t s = as_stream(stream);
        DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, f->ready, 0));
        const size_t rec = static_cast<size_t>(niw_padded_dim(f->dim)) * (niw_padded_dim(f->dim) + 1) + 4;
        const int last = f->G - 1;
        if (groupid != last) {
            const int d = f->dim;
            const size_t dd = static_cast<size_t>(d) * d;
            DISTB200_CUDA(ctx, cudaMemcpyAsync(f->niw_buf + rec * groupid, f->niw_buf + rec * last, sizeof(float) * rec, cudaMemcpyDeviceToDevice, s));
            DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_count(f) + groupid, niw_count(f) + last, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
            DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_sum_x(f) + static_cast<size_t>(groupid) * d, niw_sum_x(f) + static_cast<size_t>(last) * d,
                                               sizeof(float) * d, cudaMemcpyDeviceToDevice, s));
            DISTB200_CUDA(ctx, cudaMemcpyAsync(niw_sum_xxT(f) + groupid * dd, niw_sum_xxT(f) + last * dd, sizeof(float) * dd, cudaMemcpyDeviceToDevice, s));
        }
        f->G = last;
        int rc = DIST_B200_OK;
        if (f->dim == 32 && f->niw_tc && f->G > 0) rc = launch_niw_tc_prep(ctx, f->G, f->niw_buf, f->niw_tc, s);
        return mark_ready(f, rc, s);
    }
    const size_t gb = sizeof(float) * group_floats(f);
    char *base = static_cast<char *>(f->params);
    const int last = f->G - 1;
    if (groupid != last)  // packed_remove: move the last group into the hole (vector.hpp:47-51)
        DISTB200_CUDA(ctx, cudaMemcpyAsync(base + gb * groupid, base + gb * last, gb, cudaMemcpyDeviceToDevice, as_stream(stream)));
    DISTB200_CUDA(ctx, cudaMemsetAsync(base + gb * last, 0, gb, as_stream(stream)));
    f->gp_table_dirty = true;
    f->log_prod_valid = false;
    if (f->stats && groupid != last) {
        for (int a = 0; a < stat_arrays(f); ++a) {
            const size_t e = stat_elems(f, a);
            DISTB200_CUDA(ctx, cudaMemcpyAsync(stat_ptr(f, a) + e * groupid, stat_ptr(f, a) + e * last, e * 4, cudaMemcpyDeviceToDevice,
                                               as_stream(stream)));
        }
    }
    if (f->aux && groupid != last)
        DISTB200_CUDA(ctx, cudaMemcpyAsync(f->aux + groupid, f->aux + last, sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
    f->G = last;
    return mark_ready(f, DIST_B200_OK, as_stream(stream));
}

// batched Group::add_value over many features of one kind.  Everything is enqueued on `stream`: the
// accumulators live in a dedicated context buffer guarded by an event (not the upload scratch), so the
// call returns without draining the stream; later scoring calls order themselves behind the features'
// `ready` events.
// phase: kRowsBoth = accumulate + merge (single GPU); kRowsAccumulate = accumulate and pack into `xchg`
// ([n_features][4][G] doubles) for an all-reduce; kRowsMerge = merge from the (all-reduced) `xchg`
enum { kRowsBoth = 0, kRowsAccumulate = 1, kRowsMerge = 2 };

static bool pooled_model(int model) {
    return model == DIST_B200_NICH || model == DIST_B200_GP || model == DIST_B200_BB || model == DIST_B200_BNB;
}

static int rows_batch(dist_b200_ctx *ctx, dist_b200_feature *const *features, int n_features,
                      const void *const *columns_dev, const int32_t *assign_dev, size_t n_rows, int sign, void *stream,
                      int phase = kRowsBoth, double *xchg = nullptr) {
    if (!ctx) return DIST_B200_ERR_INVALID;
    if (n_features < 0 || (n_features && !features) || (phase != kRowsMerge && ((n_features && !columns_dev) || !assign_dev)))
        return fail(ctx, DIST_B200_ERR_INVALID, "add_rows: null argument");
    if (phase != kRowsBoth) {
        if (!xchg) return fail(ctx, DIST_B200_ERR_INVALID, "rows exchange: null exchange buffer");
        for (int i = 0; i < n_features; ++i) {
            const dist_b200_feature *f = features[i];
            if (!f) return fail(ctx, DIST_B200_ERR_INVALID, "rows exchange: null feature");
            if (f->G != features[0]->G) return fail(ctx, DIST_B200_ERR_INVALID, "rows exchange: features disagree on the number of groups");
        }
    }
    cudaStream_t s = as_stream(stream);
    size_t acc_need = 0, table_tmp = 0, niw_tmp = 0;  // pooled accumulators | (exchange) one count table of delta counts | niw work area
    int n_pooled = 0;
    for (int i = 0; i < n_features; ++i) {
        dist_b200_feature *f = features[i];
        if (!f || f->ctx != ctx || (phase != kRowsMerge && !columns_dev[i])) return fail(ctx, DIST_B200_ERR_INVALID, "add_rows: bad feature / column");
        if (f->G < 1 || !(f->model == DIST_B200_NIW ? f->niw_stats : f->stats)) return fail(ctx, DIST_B200_ERR_STATE, "add_rows: call update_all first");
        if (pooled_model(f->model)) {
            acc_need = std::max(acc_need, add_rows_acc_bytes(f->G) * std::min(n_features, kAddBatch));
            ++n_pooled;
        } else if (f->model == DIST_B200_NIW) {
            if (phase != kRowsMerge) niw_tmp = std::max(niw_tmp, niw_add_rows_bytes(f->G, f->dim, n_rows));
        } else if (phase == kRowsAccumulate) {
            table_tmp = std::max(table_tmp, round_up(sizeof(int32_t) * static_cast<size_t>(f->G) * f->dim, 256));
        }
    }
    acc_need = round_up(acc_need, 256);
    const size_t pooled_bytes = acc_need;
    acc_need += table_tmp + niw_tmp;
    // exchange layout: the pooled features' [4][G] blocks first (list order), then the count tables (list order)
    size_t table_off = phase == kRowsBoth ? 0 : static_cast<size_t>(n_pooled) * 4 * features[0]->G;
    if (acc_need > ctx->add_acc_bytes) {
        if (ctx->add_acc) {
            DISTB200_CUDA(ctx, cudaDeviceSynchronize());
            DISTB200_CUDA(ctx, cudaFree(ctx->add_acc));
            ctx->add_acc = nullptr;
            ctx->add_acc_bytes = 0;
        }
        DISTB200_CUDA(ctx, cudaMalloc(&ctx->add_acc, acc_need));
        ctx->add_acc_bytes = acc_need;
    }
    if (!ctx->add_done) DISTB200_CUDA(ctx, cudaEventCreateWithFlags(&ctx->add_done, cudaEventDisableTiming));
    DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, ctx->add_done, 0));  // the previous batch may have run on another stream
    int rc = wait_ready(ctx, features, n_features, s);
    if (rc) return rc;

    AddBatch b{};
    b.sign = sign;
    b.N = n_rows;
    b.assign = assign_dev;
    b.acc = static_cast<char *>(ctx->add_acc);
    int flushed = 0;  // pooled features already processed = index of b.d[0] in the exchange buffer
    auto flush = [&]() -> int {
        if (b.n == 0) return DIST_B200_OK;
        b.acc_stride = add_rows_acc_bytes(b.G);
        b.xchg = phase == kRowsMerge ? xchg : nullptr;
        b.xchg_first = flushed;
        int r = DIST_B200_OK;
        if (phase != kRowsMerge) r = launch_add_rows_pooled(ctx, b, s);
        if (!r && phase == kRowsAccumulate) r = launch_pack_accumulators(ctx, b, xchg, s);
        if (!r && phase != kRowsAccumulate) r = launch_merge_prep_batch(ctx, b, s);
        flushed += b.n;
        b.n = 0;
        return r;
    };
    // count tables (dd / dpd).  Single GPU: the rows go straight into the feature's table.  Exchange: accumulate = this
    // rank's delta counts as doubles into the feature's block of xchg, merge = the summed deltas into the table.
    auto table_counts = [&](dist_b200_feature *f, int i) -> int {
        const size_t cells = static_cast<size_t>(f->G) * f->dim;
        int r = DIST_B200_OK;
        if (phase == kRowsBoth) {
            r = launch_add_rows_counts(ctx, f, columns_dev[i], assign_dev, n_rows, sign, s);
        } else if (phase == kRowsAccumulate) {
            int32_t *tmp = reinterpret_cast<int32_t *>(static_cast<char *>(ctx->add_acc) + pooled_bytes);
            DISTB200_CUDA(ctx, cudaMemsetAsync(tmp, 0, sizeof(int32_t) * cells, s));
            r = launch_add_rows_counts(ctx, f, columns_dev[i], assign_dev, n_rows, +1, s, tmp);
            if (!r) r = launch_counts_to_doubles(ctx, tmp, xchg + table_off, cells, s);
        } else {
            r = launch_merge_counts(ctx, reinterpret_cast<int32_t *>(f->stats), xchg + table_off, cells, sign, s);
        }
        table_off += cells;
        return r;
    };
    for (int i = 0; i < n_features; ++i) {
        dist_b200_feature *f = features[i];
        const int G = f->G;
        switch (f->model) {
            case DIST_B200_NICH:
            case DIST_B200_GP:
            case DIST_B200_BB:
            case DIST_B200_BNB: {
                if (b.n == kAddBatch || (b.n && b.G != G))
                    if ((rc = flush())) return rc;
                b.G = G;
                AddDesc &d = b.d[b.n++];
                d.column = columns_dev ? columns_dev[i] : nullptr;
                d.st0 = stat_ptr(f, 0);
                d.st1 = stat_ptr(f, 1);
                d.st2 = f->model == DIST_B200_NICH ? stat_ptr(f, 2) : nullptr;
                d.params = static_cast<float4 *>(f->params);
                d.aux = f->aux;
                for (int k = 0; k < 4; ++k) d.shared[k] = f->shared[k];
                d.model = f->model;
                if (f->model == DIST_B200_GP && phase != kRowsAccumulate) {
                    f->gp_table_dirty = true;
                    f->log_prod_valid = false;  // Group::log_prod is not maintained by the batched update
                }
            } break;
            case DIST_B200_DD:
                if (!f->alphas_dev) return fail(ctx, DIST_B200_ERR_STATE, "add_rows: dd alphas not resident (update_all first)");
                if ((rc = table_counts(f, i))) return rc;
                if (phase == kRowsAccumulate) break;
                if ((rc = launch_dd_prep(ctx, f->dim, f->alphas_dev, f->alpha_sum, 0, G, reinterpret_cast<const int32_t *>(stat_ptr(f, 0)),
                                         static_cast<float *>(f->params), s)))
                    return rc;
                break;
            case DIST_B200_DPD:
                if ((rc = table_counts(f, i))) return rc;
                if (phase == kRowsAccumulate) break;
                if ((rc = dpd_rebuild(f, reinterpret_cast<const float *>(dpd_betas(f)), reinterpret_cast<const int32_t *>(f->stats), s)))
                    return rc;
                break;
            case DIST_B200_NIW: {  // per-group SYRK into a [G][1 + d + d^2] double block (niw_stats.cu), then the records again
                const int d = f->dim;
                const size_t per = 1 + static_cast<size_t>(d) + static_cast<size_t>(d) * d;
                char *area = static_cast<char *>(ctx->add_acc) + pooled_bytes + table_tmp;
                double *acc = phase == kRowsBoth ? reinterpret_cast<double *>(area) : xchg + table_off;
                void *work = area + round_up(sizeof(double) * G * per, 256);
                if (phase != kRowsMerge && (rc = launch_niw_accumulate(ctx, G, d, columns_dev[i], assign_dev, n_rows, acc, work, s))) return rc;
                table_off += G * per;
                if (phase == kRowsAccumulate) break;
                if ((rc = launch_niw_apply(ctx, G, d, sign, acc, niw_count(f), niw_sum_x(f), niw_sum_xxT(f), s))) return rc;
                if ((rc = niw_rebuild(f, 0, G, s))) return rc;
            } break;
            default: return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "add_rows: unsupported model");
        }
    }
    if ((rc = flush())) return rc;
    if (phase != kRowsAccumulate)
        for (int i = 0; i < n_features; ++i) DISTB200_CUDA(ctx, cudaEventRecord(features[i]->ready, s));
    DISTB200_CUDA(ctx, cudaEventRecord(ctx->add_done, s));
    return DIST_B200_OK;
}

// Row-sharded update (one process per GPU): every rank accumulates its own rows, the [n_features][4][G] double
// buffers are summed over the ranks (NCCL all-reduce by the caller), every rank merges the global sums into
// its replica of the statistics -- all replicas stay identical.
int dist_b200_rows_accumulate(dist_b200_ctx *ctx, dist_b200_feature *const *features, int n_features,
                              const void *const *columns_dev, const int32_t *assign_dev, size_t n_rows, double *xchg_dev,
                              void *stream) {
    return rows_batch(ctx, features, n_features, columns_dev, assign_dev, n_rows, +1, stream, kRowsAccumulate, xchg_dev);
}

int dist_b200_rows_xchg_doubles(dist_b200_feature *const *features, int n_features, size_t *n_doubles) {
    if (n_features < 0 || (n_features && !features) || !n_doubles) return DIST_B200_ERR_INVALID;
    size_t n = 0;
    for (int i = 0; i < n_features; ++i) {
        const dist_b200_feature *f = features[i];
        if (!f) return DIST_B200_ERR_INVALID;
        n += pooled_model(f->model) ? static_cast<size_t>(4) * f->G
             : f->model == DIST_B200_NIW ? static_cast<size_t>(f->G) * (1 + f->dim + static_cast<size_t>(f->dim) * f->dim)
                                         : static_cast<size_t>(f->G) * f->dim;
    }
    *n_doubles = n;
    return DIST_B200_OK;
}

int dist_b200_rows_merge(dist_b200_ctx *ctx, dist_b200_feature *const *features, int n_features, const double *xchg_dev,
                         int sign, void *stream) {
    if (sign != 1 && sign != -1) return ctx ? fail(ctx, DIST_B200_ERR_INVALID, "rows_merge: sign must be +1 or -1") : DIST_B200_ERR_INVALID;
    return rows_batch(ctx, features, n_features, nullptr, nullptr, 0, sign, stream, kRowsMerge, const_cast<double *>(xchg_dev));
}

int dist_b200_add_rows_batch(dist_b200_ctx *ctx, dist_b200_feature *const *features, int n_features,
                             const void *const *columns_dev, const int32_t *assign_dev, size_t n_rows, void *stream) {
    return rows_batch(ctx, features, n_features, columns_dev, assign_dev, n_rows, +1, stream);
}

int dist_b200_remove_rows_batch(dist_b200_ctx *ctx, dist_b200_feature *const *features, int n_features,
                                const void *const *columns_dev, const int32_t *assign_dev, size_t n_rows, void *stream) {
    return rows_batch(ctx, features, n_features, columns_dev, assign_dev, n_rows, -1, stream);
}

// bytes of one value of the feature's column: bool as uint8, niw a row of d floats, everything else 4 bytes
static size_t column_bytes(const dist_b200_feature *f) {
    return f->model == DIST_B200_BB ? 1 : f->model == DIST_B200_NIW ? 4 * static_cast<size_t>(f->dim) : 4;
}

// host buffers: stage columns + assignments through the context scratch, run the device batch, drain
static int rows_batch_host(dist_b200_ctx *ctx, dist_b200_feature *const *features, int n_features,
                           const void *const *columns_host, const int32_t *assign_host, size_t n_rows, int sign) {
    if (!ctx) return DIST_B200_ERR_INVALID;
    if (n_features < 1 || !features || !columns_host || !assign_host) return fail(ctx, DIST_B200_ERR_INVALID, "add_rows_host: null argument");
    if (n_rows == 0) return DIST_B200_OK;
    std::vector<size_t> off(n_features);
    size_t total = 0;
    for (int i = 0; i < n_features; ++i) {
        if (!features[i] || !columns_host[i]) return fail(ctx, DIST_B200_ERR_INVALID, "add_rows_host: null feature / column");
        off[i] = total;
        total += round_up(column_bytes(features[i]) * n_rows, 256);
    }
    const size_t assign_off = total;
    total += round_up(sizeof(int32_t) * n_rows, 256);
    int rc = ensure_scratch(ctx, total + 256);
    if (rc) return rc;
    cudaStream_t s = ctx->own_stream;
    char *dev = static_cast<char *>(ctx->scratch_dev);
    std::vector<const void *> cols(n_features);
    for (int i = 0; i < n_features; ++i) {
        const size_t vb = column_bytes(features[i]);
        DISTB200_CUDA(ctx, cudaMemcpyAsync(dev + off[i], columns_host[i], vb * n_rows, cudaMemcpyHostToDevice, s));
        cols[i] = dev + off[i];
    }
    DISTB200_CUDA(ctx, cudaMemcpyAsync(dev + assign_off, assign_host, sizeof(int32_t) * n_rows, cudaMemcpyHostToDevice, s));
    rc = rows_batch(ctx, features, n_features, cols.data(), reinterpret_cast<const int32_t *>(dev + assign_off), n_rows, sign, s);
    DISTB200_CUDA(ctx, cudaStreamSynchronize(s));  // the scratch is free for the next upload
    return rc;
}

int dist_b200_add_rows_batch_host(dist_b200_ctx *ctx, dist_b200_feature *const *features, int n_features,
                                  const void *const *columns_host, const int32_t *assign_host, size_t n_rows) {
    return rows_batch_host(ctx, features, n_features, columns_host, assign_host, n_rows, +1);
}

int dist_b200_remove_rows_batch_host(dist_b200_ctx *ctx, dist_b200_feature *const *features, int n_features,
                                     const void *const *columns_host, const int32_t *assign_host, size_t n_rows) {
    return rows_batch_host(ctx, features, n_features, columns_host, assign_host, n_rows, -1);
}

int dist_b200_feature_add_rows(dist_b200_feature *f, const void *column_dev, const int32_t *assign_dev, size_t n_rows,
                               void *stream) {
    if (!f || !f->ctx) return DIST_B200_ERR_INVALID;
    return dist_b200_add_rows_batch(f->ctx, &f, 1, &column_dev, assign_dev, n_rows, stream);
}

// ---- score_data_grid (next row: hyper-parameter inference over the same device-resident statistics) ----
static size_t shared_stride(const dist_b200_feature *f) {
    switch (f->model) {
        case DIST_B200_NICH: return 4;
        case DIST_B200_GP: case DIST_B200_BB: case DIST_B200_BNB: return 2;
        case DIST_B200_DD: return static_cast<size_t>(f->dim);
        case DIST_B200_DPD: return 1;
        case DIST_B200_NIW: return 2 + static_cast<size_t>(f->dim) + static_cast<size_t>(f->dim) * f->dim;  // kappa, nu, mu[d], psi[d][d]
        default: return 0;
    }
}

int dist_b200_gp_set_log_prod(dist_b200_feature *f, const float *log_prod_host, void *stream) {
    if (!check_feature(f, DIST_B200_GP) || !log_prod_host) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (f->G < 1) return fail(ctx, DIST_B200_ERR_STATE, "gp_set_log_prod: call update_all first");
    if (f->log_prod_dev && f->log_prod_cap < f->capacity) {
        DISTB200_CUDA(ctx, cudaDeviceSynchronize());
        DISTB200_CUDA(ctx, cudaFree(f->log_prod_dev));
        f->log_prod_dev = nullptr;
    }
    if (!f->log_prod_dev) {
        DISTB200_CUDA(ctx, cudaMalloc(&f->log_prod_dev, sizeof(float) * f->capacity));
        f->log_prod_cap = f->capacity;
    }
    cudaStream_t s = as_stream(stream);
    DISTB200_CUDA(ctx, cudaMemcpyAsync(f->log_prod_dev, log_prod_host, sizeof(float) * f->G, cudaMemcpyHostToDevice, s));
    DISTB200_CUDA(ctx, cudaStreamSynchronize(s));  // pageable source
    f->log_prod_valid = true;
    return DIST_B200_OK;
}

int dist_b200_score_data_grid(dist_b200_feature *f, const float *shareds_dev, size_t n_grid, size_t stride,
                              float *out_dev, void *stream) {
    if (!f || !f->ctx) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (!shareds_dev || !out_dev) return fail(ctx, DIST_B200_ERR_INVALID, "score_data_grid: null argument");
    if (f->G < 1 || !(f->model == DIST_B200_NIW ? f->niw_stats : f->stats)) return fail(ctx, DIST_B200_ERR_STATE, "score_data_grid: call update_all first");
    if (stride < shared_stride(f)) return fail(ctx, DIST_B200_ERR_INVALID, "score_data_grid: stride shorter than the model's packed Shared");
    if (f->model == DIST_B200_GP && !f->log_prod_valid)
        return fail(ctx, DIST_B200_ERR_STATE, "score_data_grid: gp needs Group::log_prod (dist_b200_gp_set_log_prod) after the last statistics change");
    if (n_grid == 0) return DIST_B200_OK;
    cudaStream_t s = as_stream(stream);
    DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, f->ready, 0));
    // the double accumulators live next to the add_value accumulators (same event guards both)
    const size_t need = sizeof(double) * n_grid;
    if (need > ctx->add_acc_bytes) {
        if (ctx->add_acc) {
            DISTB200_CUDA(ctx, cudaDeviceSynchronize());
            DISTB200_CUDA(ctx, cudaFree(ctx->add_acc));
            ctx->add_acc = nullptr;
            ctx->add_acc_bytes = 0;
        }
        DISTB200_CUDA(ctx, cudaMalloc(&ctx->add_acc, need));
        ctx->add_acc_bytes = need;
    }
    if (!ctx->add_done) DISTB200_CUDA(ctx, cudaEventCreateWithFlags(&ctx->add_done, cudaEventDisableTiming));
    DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, ctx->add_done, 0));
    if (f->model == DIST_B200_NIW) {
        DISTB200_CUDA(ctx, cudaMemsetAsync(ctx->add_acc, 0, sizeof(double) * n_grid, s));
        int rcn = launch_niw_score_data(ctx, f->G, f->dim, niw_count(f), niw_sum_x(f), niw_sum_xxT(f), shareds_dev, n_grid, stride,
                                        static_cast<double *>(ctx->add_acc), s);
        if (!rcn) rcn = launch_score_data_finish(ctx, n_grid, static_cast<const double *>(ctx->add_acc), out_dev, s);
        if (rcn) return rcn;
        DISTB200_CUDA(ctx, cudaEventRecord(ctx->add_done, s));
        return DIST_B200_OK;
    }
    const uint32_t *st0, *st1 = nullptr, *st2 = nullptr;
    const float *betas = nullptr;
    if (f->model == DIST_B200_DPD) {
        st0 = f->stats;
        betas = reinterpret_cast<const float *>(dpd_betas(f));
    } else {
        st0 = stat_ptr(f, 0);
        if (stat_arrays(f) > 1) st1 = stat_ptr(f, 1);
        if (stat_arrays(f) > 2) st2 = stat_ptr(f, 2);
    }
    int rc = launch_score_data(ctx, f, st0, st1, st2, betas, f->log_prod_dev, shareds_dev, n_grid, stride,
                               static_cast<double *>(ctx->add_acc), out_dev, s);
    if (rc) return rc;
    DISTB200_CUDA(ctx, cudaEventRecord(ctx->add_done, s));
    return DIST_B200_OK;
}

int dist_b200_score_data_grid_host(dist_b200_feature *f, const float *shareds_host, size_t n_grid, size_t stride,
                                   float *out_host) {
    if (!f || !f->ctx) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (!shareds_host || !out_host) return fail(ctx, DIST_B200_ERR_INVALID, "score_data_grid: null argument");
    if (n_grid == 0) return DIST_B200_OK;
    const size_t in_bytes = round_up(sizeof(float) * n_grid * stride, 256);
    int rc = ensure_scratch(ctx, in_bytes + sizeof(float) * n_grid + 256);
    if (rc) return rc;
    cudaStream_t s = ctx->own_stream;
    char *dev = static_cast<char *>(ctx->scratch_dev);
    DISTB200_CUDA(ctx, cudaMemcpyAsync(dev, shareds_host, sizeof(float) * n_grid * stride, cudaMemcpyHostToDevice, s));
    rc = dist_b200_score_data_grid(f, reinterpret_cast<const float *>(dev), n_grid, stride, reinterpret_cast<float *>(dev + in_bytes), s);
    if (rc == DIST_B200_OK)
        DISTB200_CUDA(ctx, cudaMemcpyAsync(out_host, dev + in_bytes, sizeof(float) * n_grid, cudaMemcpyDeviceToHost, s));
    DISTB200_CUDA(ctx, cudaStreamSynchronize(s));
    return rc;
}

// ---- protobuf wire format (schema.proto) -> update_all --------------------------
int dist_b200_wire_decode(dist_b200_ctx *ctx, int model, const void *shared_msg, size_t shared_len,
                          const void *const *group_msgs, const size_t *group_lens, int G, float *shared_out,
                          size_t shared_cap, uint32_t *keys_out, size_t keys_cap, uint32_t *stats_out, size_t stats_cap,
                          size_t counts_out[3]) {
    if (!counts_out) return DIST_B200_ERR_INVALID;  // ctx may be null: decoding needs no device
    WireFeature w;
    int rc = wire_decode(ctx, model, shared_msg, shared_len, group_msgs, group_lens, G, w);
    if (rc) return rc;
    counts_out[0] = w.shared.size();
    counts_out[1] = w.keys.size();
    counts_out[2] = w.stats.size();
    if (w.shared.size() > shared_cap || w.keys.size() > keys_cap || w.stats.size() > stats_cap)
        return fail(ctx, DIST_B200_ERR_INVALID, "wire_decode: output buffer too small (sizes returned in counts_out)");
    if (shared_out && !w.shared.empty()) std::memcpy(shared_out, w.shared.data(), sizeof(float) * w.shared.size());
    if (keys_out && !w.keys.empty()) std::memcpy(keys_out, w.keys.data(), sizeof(uint32_t) * w.keys.size());
    if (stats_out && !w.stats.empty()) std::memcpy(stats_out, w.stats.data(), sizeof(uint32_t) * w.stats.size());
    return DIST_B200_OK;
}

int dist_b200_update_all_wire(dist_b200_feature *f, const void *shared_msg, size_t shared_len,
                              const void *const *group_msgs, const size_t *group_lens, int G, void *stream) {
    if (!f || !f->ctx) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    WireFeature w;
    int rc = wire_decode(ctx, f->model, shared_msg, shared_len, group_msgs, group_lens, G, w);
    if (rc) return rc;
    const size_t g = static_cast<size_t>(G);
    const uint32_t *st = w.stats.data();
    switch (f->model) {
        case DIST_B200_NICH:
            return dist_b200_nich_update_all(f, w.shared.data(), G, reinterpret_cast<const int32_t *>(st),
                                             reinterpret_cast<const float *>(st + g), reinterpret_cast<const float *>(st + 2 * g), stream);
        case DIST_B200_GP:
            if ((rc = dist_b200_gp_update_all(f, w.shared.data(), G, st, st + g, stream))) return rc;
            return G ? dist_b200_gp_set_log_prod(f, reinterpret_cast<const float *>(st + 2 * g), stream) : DIST_B200_OK;
        case DIST_B200_BNB:
            return dist_b200_bnb_update_all(f, w.shared.data(), w.keys[0], G, st, st + g, stream);
        case DIST_B200_BB:
            return dist_b200_bb_update_all(f, w.shared.data(), G, reinterpret_cast<const int32_t *>(st),
                                           reinterpret_cast<const int32_t *>(st + g), stream);
        case DIST_B200_DD:
            return dist_b200_dd_update_all(f, w.dim, w.shared.data(), G, reinterpret_cast<const int32_t *>(st), stream);
        case DIST_B200_DPD:
            return dist_b200_dpd_update_all(f, w.shared[1], w.shared[2], w.dim, w.keys.data(), w.shared.data() + 3, G,
                                            reinterpret_cast<const int32_t *>(st), stream);
        case DIST_B200_NIW: {
            const size_t d = static_cast<size_t>(w.dim);
            return dist_b200_niw_update_all(f, w.dim, w.shared.data() + 2, w.shared[0], w.shared.data() + 2 + d, w.shared[1], G,
                                            reinterpret_cast<const int32_t *>(st), reinterpret_cast<const float *>(st + g),
                                            reinterpret_cast<const float *>(st + g + g * d), stream);
        }
        default: return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "update_all_wire: unsupported model");
    }
}

// the reference's record stream (distributions/io/stream.py:141-153): [uint32 little-endian length][message] ...
static int split_stream(dist_b200_ctx *ctx, const void *bytes, size_t len, std::vector<const void *> &msgs,
                        std::vector<size_t> &lens) {
    const uint8_t *p = static_cast<const uint8_t *>(bytes), *end = p + len;
    while (p < end) {
        if (end - p < 4) return fail(ctx, DIST_B200_ERR_INVALID, "wire stream: truncated record length");
        const size_t n = static_cast<size_t>(p[0]) | (static_cast<size_t>(p[1]) << 8) | (static_cast<size_t>(p[2]) << 16) |
                         (static_cast<size_t>(p[3]) << 24);
        p += 4;
        if (n > static_cast<size_t>(end - p)) return fail(ctx, DIST_B200_ERR_INVALID, "wire stream: record runs past the end");
        msgs.push_back(p);
        lens.push_back(n);
        p += n;
    }
    return DIST_B200_OK;
}

int dist_b200_wire_split_stream(dist_b200_ctx *ctx, const void *stream_bytes, size_t stream_len, size_t *offsets_out,
                                size_t *lens_out, size_t capacity, size_t *n_records) {
    if ((!stream_bytes && stream_len) || !n_records) return DIST_B200_ERR_INVALID;  // ctx may be null
    std::vector<const void *> msgs;
    std::vector<size_t> lens;
    int rc = split_stream(ctx, stream_bytes, stream_len, msgs, lens);
    if (rc) return rc;
    *n_records = msgs.size();
    if (msgs.size() > capacity) return fail(ctx, DIST_B200_ERR_INVALID, "wire_split_stream: output arrays too small (count returned)");
    for (size_t i = 0; i < msgs.size(); ++i) {
        if (offsets_out) offsets_out[i] = static_cast<size_t>(static_cast<const uint8_t *>(msgs[i]) - static_cast<const uint8_t *>(stream_bytes));
        if (lens_out) lens_out[i] = lens[i];
    }
    return DIST_B200_OK;
}

int dist_b200_update_all_stream(dist_b200_feature *f, const void *shared_msg, size_t shared_len, const void *stream_bytes,
                                size_t stream_len, void *stream) {
    if (!f || !f->ctx) return DIST_B200_ERR_INVALID;
    if (!stream_bytes && stream_len) return fail(f->ctx, DIST_B200_ERR_INVALID, "update_all_stream: null stream");
    std::vector<const void *> msgs;
    std::vector<size_t> lens;
    int rc = split_stream(f->ctx, stream_bytes, stream_len, msgs, lens);
    if (rc) return rc;
    if (msgs.size() > 0x7FFFFFFFull) return fail(f->ctx, DIST_B200_ERR_UNSUPPORTED, "update_all_stream: too many groups");
    return dist_b200_update_all_wire(f, shared_msg, shared_len, msgs.data(), lens.data(), static_cast<int>(msgs.size()), stream);
}

int dist_b200_prior_wire_host(dist_b200_ctx *ctx, const void *clustering_msg, size_t len, int G,
                              const int32_t *group_sizes, float *prior_host) {
    if (!ctx || !clustering_msg || !group_sizes || !prior_host || G < 1) return DIST_B200_ERR_INVALID;
    int which = 0;
    float alpha = 0, d = 0;
    uint64_t dataset_size = 0;
    int rc = wire_decode_clustering(ctx, clustering_msg, len, &which, &alpha, &d, &dataset_size);
    if (rc) return rc;
    if (which == 1) return dist_b200_prior_pitman_yor_host(ctx, alpha, d, G, group_sizes, prior_host);
    if (dataset_size < 1 || dataset_size > 0x7FFFFFFFull) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "wire: LowEntropy dataset_size out of range");
    return dist_b200_prior_low_entropy_host(ctx, static_cast<int>(dataset_size), G, group_sizes, prior_host);
}

static int emit_messages(dist_b200_ctx *ctx, const std::vector<uint8_t> &bytes, const std::vector<size_t> &lens, void *out,
                         size_t capacity, size_t *lens_out, size_t *n_bytes) {
    if (n_bytes) *n_bytes = bytes.size();
    if (bytes.size() > capacity) return fail(ctx, DIST_B200_ERR_INVALID, "wire: output buffer too small (size returned in n_bytes)");
    if (!bytes.empty()) std::memcpy(out, bytes.data(), bytes.size());
    if (lens_out) for (size_t i = 0; i < lens.size(); ++i) lens_out[i] = lens[i];
    return DIST_B200_OK;
}

int dist_b200_wire_encode_groups(dist_b200_ctx *ctx, int model, int G, int dim, const uint32_t *keys, const uint32_t *stats,
                                 size_t stats_words, void *out, size_t capacity, size_t *lens_out, size_t *n_bytes) {
    std::vector<uint8_t> bytes;
    std::vector<size_t> lens;
    int rc = wire_encode_groups(ctx, model, G, dim, keys, stats, stats_words, bytes, lens);
    if (rc) return rc;
    return emit_messages(ctx, bytes, lens, out, capacity, lens_out, n_bytes);
}

int dist_b200_wire_encode_shared(dist_b200_ctx *ctx, int model, const float *shared, size_t n_shared, const uint32_t *keys,
                                 size_t n_keys, void *out, size_t capacity, size_t *n_bytes) {
    std::vector<uint8_t> bytes;
    int rc = wire_encode_shared(ctx, model, shared, n_shared, keys, n_keys, bytes);
    if (rc) return rc;
    return emit_messages(ctx, bytes, std::vector<size_t>(), out, capacity, nullptr, n_bytes);
}

int dist_b200_feature_dump_groups_wire(dist_b200_feature *f, void *out, size_t capacity, size_t *lens_out, size_t *n_bytes,
                                       void *stream) {
    if (!f || !f->ctx) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (!(f->model == DIST_B200_NIW ? f->niw_stats : f->stats) || f->G < 1)
        return fail(ctx, DIST_B200_ERR_STATE, "dump_groups_wire: no device statistics (update_all first)");
    const size_t g = static_cast<size_t>(f->G);
    size_t words;
    switch (f->model) {
        case DIST_B200_NICH: case DIST_B200_GP: words = 3 * g; break;
        case DIST_B200_BNB: case DIST_B200_BB: words = 2 * g; break;
        case DIST_B200_DD: case DIST_B200_DPD: words = g * f->dim; break;
        case DIST_B200_NIW: words = g * (1 + f->dim + static_cast<size_t>(f->dim) * f->dim); break;
        default: return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "dump_groups_wire: unsupported model");
    }
    std::vector<uint32_t> st(words, 0);
    size_t got = 0;
    int rc = dist_b200_feature_download_stats(f, st.data(), 4 * words, &got, stream);  // synchronises
    if (rc) return rc;
    if (f->model == DIST_B200_GP) {  // Group::log_prod is a required field of the message
        if (!f->log_prod_valid) return fail(ctx, DIST_B200_ERR_STATE, "dump_groups_wire: gp log_prod is stale (dist_b200_gp_set_log_prod)");
        DISTB200_CUDA(ctx, cudaMemcpy(st.data() + 2 * g, f->log_prod_dev, 4 * g, cudaMemcpyDeviceToHost));
    }
    // dpd: f->keys is update_all's key array, in the caller's order like the mirrored counts
    std::vector<uint32_t> keys;
    if (f->model == DIST_B200_DPD) {
        if (f->keys.size() != static_cast<size_t>(f->dim)) return fail(ctx, DIST_B200_ERR_STATE, "dump_groups_wire: dpd keys missing");
        keys = f->keys;
    }
    std::vector<uint8_t> bytes;
    std::vector<size_t> lens;
    if ((rc = wire_encode_groups(ctx, f->model, f->G, f->dim, keys.data(), st.data(), words, bytes, lens))) return rc;
    return emit_messages(ctx, bytes, lens, out, capacity, lens_out, n_bytes);
}

int dist_b200_feature_download_stats(const dist_b200_feature *f, void *out_host, size_t capacity_bytes, size_t *n_bytes,
                                     void *stream) {
    if (!f || !f->ctx || !out_host) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    if (!(f->model == DIST_B200_NIW ? f->niw_stats : f->stats))
        return fail(ctx, DIST_B200_ERR_STATE, "download_stats: no device statistics (update_all first)");
    cudaStream_t s = as_stream(stream);
    DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, f->ready, 0));
    char *out = static_cast<char *>(out_host);
    size_t total = 0;
    if (f->model == DIST_B200_NIW) {  // count[G] | sum_x[G][d] | sum_xxT[G][d][d]
        const size_t g = static_cast<size_t>(f->G), d = static_cast<size_t>(f->dim);
        total = 4 * g * (1 + d + d * d);
        if (n_bytes) *n_bytes = total;
        if (total > capacity_bytes) return fail(ctx, DIST_B200_ERR_INVALID, "download_stats: buffer too small");
        DISTB200_CUDA(ctx, cudaMemcpyAsync(out, niw_count(f), 4 * g, cudaMemcpyDeviceToHost, s));
        DISTB200_CUDA(ctx, cudaMemcpyAsync(out + 4 * g, niw_sum_x(f), 4 * g * d, cudaMemcpyDeviceToHost, s));
        DISTB200_CUDA(ctx, cudaMemcpyAsync(out + 4 * g * (1 + d), niw_sum_xxT(f), 4 * g * d * d, cudaMemcpyDeviceToHost, s));
    } else if (f->model == DIST_B200_DPD) {
        total = sizeof(int32_t) * static_cast<size_t>(f->G) * f->dim;
        if (n_bytes) *n_bytes = total;
        if (total > capacity_bytes) return fail(ctx, DIST_B200_ERR_INVALID, "download_stats: buffer too small");
        DISTB200_CUDA(ctx, cudaMemcpyAsync(out, f->stats, total, cudaMemcpyDeviceToHost, s));
    } else {
        for (int a = 0; a < stat_arrays(f); ++a) total += stat_elems(f, a) * f->G * 4;
        if (n_bytes) *n_bytes = total;
        if (total > capacity_bytes) return fail(ctx, DIST_B200_ERR_INVALID, "download_stats: buffer too small");
        size_t off = 0;
        for (int a = 0; a < stat_arrays(f); ++a) {
            const size_t b = stat_elems(f, a) * f->G * 4;
            DISTB200_CUDA(ctx, cudaMemcpyAsync(out + off, stat_ptr(f, a), b, cudaMemcpyDeviceToHost, s));
            off += b;
        }
    }
    DISTB200_CUDA(ctx, cudaStreamSynchronize(s));
    return DIST_B200_OK;
}

int dist_b200_count_assignments(dist_b200_ctx *ctx, const int32_t *assign_dev, size_t n_rows, int G, int32_t *counts_dev,
                                int accumulate, void *stream) {
    if (!ctx || !assign_dev || !counts_dev || G < 1) return DIST_B200_ERR_INVALID;
    return launch_count_assignments(ctx, assign_dev, n_rows, G, counts_dev, accumulate, as_stream(stream));
}

int dist_b200_prior_pitman_yor_dev(dist_b200_ctx *ctx, float alpha, float d, int G, const int32_t *group_sizes_dev,
                                   float *prior_dev, void *stream) {
    if (!ctx || G < 1 || !group_sizes_dev || !prior_dev) return DIST_B200_ERR_INVALID;
    return launch_prior_prep(ctx, alpha, d, G, group_sizes_dev, prior_dev, as_stream(stream));
}

int dist_b200_feature_download_caches(const dist_b200_feature *f, float *out_host, size_t capacity_floats,
                                      size_t *n_floats, void *stream) {
    if (!f || !f->ctx || !out_host) return DIST_B200_ERR_INVALID;
    dist_b200_ctx *ctx = f->ctx;
    size_t rows;
    switch (f->model) {
        case DIST_B200_NICH: rows = 4; break;
        case DIST_B200_GP: rows = 3; break;
        case DIST_B200_BNB: rows = 3; break;
        case DIST_B200_BB: rows = 2; break;
        case DIST_B200_DD: rows = f->dim; break;
        case DIST_B200_DPD: rows = f->dim + 1; break;
        default: return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "download_caches: unsupported model");
    }
    const size_t n = rows * f->G;
    if (n_floats) *n_floats = n;
    if (n > capacity_floats) return fail(ctx, DIST_B200_ERR_INVALID, "download_caches: buffer too small");
    if (n == 0) return DIST_B200_OK;
    int rc = ensure_scratch(ctx, n * sizeof(float));
    if (rc) return rc;
    cudaStream_t s = as_stream(stream);
    DISTB200_CUDA(ctx, cudaStreamWaitEvent(s, f->ready, 0));  // a batched add / remove may still be rebuilding the caches
    if ((rc = launch_unpack_caches(ctx, f, static_cast<float *>(ctx->scratch_dev), s))) return rc;
    DISTB200_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->scratch_dev, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    DISTB200_CUDA(ctx, cudaStreamSynchronize(s));
    return DIST_B200_OK;
}

int dist_b200_prior_pitman_yor(dist_b200_ctx *ctx, float alpha, float d, int G, const int32_t *group_sizes,
                               float *prior_dev, void *stream) {
    if (!ctx || G < 1 || !group_sizes || !prior_dev) return DIST_B200_ERR_INVALID;
    int rc = ensure_scratch(ctx, sizeof(int32_t) * G + 256);
    if (rc) return rc;
    Upload up{ctx, as_stream(stream)};
    const int32_t *sz = up.put(group_sizes, G);
    if (up.err) return up.err;
    int rc2 = launch_prior_prep(ctx, alpha, d, G, sz, prior_dev, as_stream(stream));
    if (rc2) return rc2;
    DISTB200_CUDA(ctx, cudaStreamSynchronize(as_stream(stream)));  // scratch (group sizes) is free again
    return DIST_B200_OK;
}

// LowEntropy clustering prior (clustering.hpp:245-331 through MixtureDriver::score_value): overwrite semantics
int dist_b200_prior_low_entropy_dev(dist_b200_ctx *ctx, int dataset_size, int G, const int32_t *sizes_dev,
                                    float *prior_dev, void *stream) {
    if (!ctx || G < 1 || !sizes_dev || !prior_dev || dataset_size < 1) return DIST_B200_ERR_INVALID;
    return launch_low_entropy_prep(ctx, dataset_size, G, sizes_dev, prior_dev, as_stream(stream));
}

int dist_b200_prior_low_entropy_host(dist_b200_ctx *ctx, int dataset_size, int G, const int32_t *group_sizes,
                                     float *prior_host) {
    if (!ctx || G < 1 || !group_sizes || !prior_host || dataset_size < 1) return DIST_B200_ERR_INVALID;
    int rc = ensure_scratch(ctx, round_up(sizeof(int32_t) * G, 256) + sizeof(float) * G + 256);
    if (rc) return rc;
    cudaStream_t s = ctx->own_stream;
    Upload up{ctx, s};
    const int32_t *sz = up.put(group_sizes, G);
    if (up.err) return up.err;
    float *out = reinterpret_cast<float *>(static_cast<char *>(ctx->scratch_dev) + up.off);
    if ((rc = launch_low_entropy_prep(ctx, dataset_size, G, sz, out, s))) return rc;
    DISTB200_CUDA(ctx, cudaMemcpyAsync(prior_host, out, sizeof(float) * G, cudaMemcpyDeviceToHost, s));
    DISTB200_CUDA(ctx, cudaStreamSynchronize(s));
    return DIST_B200_OK;
}

int dist_b200_prior_pitman_yor_host(dist_b200_ctx *ctx, float alpha, float d, int G, const int32_t *group_sizes,
                                    float *prior_host) {
    if (!ctx || G < 1 || !group_sizes || !prior_host) return DIST_B200_ERR_INVALID;
    int rc = ensure_scratch(ctx, round_up(sizeof(int32_t) * G, 256) + sizeof(float) * G + 256);
    if (rc) return rc;
    cudaStream_t s = ctx->own_stream;
    Upload up{ctx, s};
    const int32_t *sz = up.put(group_sizes, G);
    if (up.err) return up.err;
    float *out = reinterpret_cast<float *>(static_cast<char *>(ctx->scratch_dev) + up.off);
    if ((rc = launch_prior_prep(ctx, alpha, d, G, sz, out, s))) return rc;
    DISTB200_CUDA(ctx, cudaMemcpyAsync(prior_host, out, sizeof(float) * G, cudaMemcpyDeviceToHost, s));
    DISTB200_CUDA(ctx, cudaStreamSynchronize(s));
    return DIST_B200_OK;
}

// ---- the hot path ---------------------------------------------------------------------------
// Describe one row-mapped feature for the score kernel.  In multi-feature lists GammaPoisson goes through
// its per-(group, value) table (rebuilt lazily, stream-ordered, after any cache update).
static int queue_gp_table(dist_b200_ctx *ctx, dist_b200_feature *f, GpTableBatch &tb, cudaStream_t s) {
    if (tb.n == kGpTableBatch) {
        int rc = launch_gp_table_batch(ctx, tb, s);
        if (rc) return rc;
        tb.n = 0;
    }
    tb.n_groups[tb.n] = f->capacity;
    tb.params[tb.n] = static_cast<const float4 *>(f->params);
    tb.table[tb.n] = f->gp_table;
    ++tb.n;
    f->gp_table_dirty = false;
    return DIST_B200_OK;
}

// stale tables are queued in `tb`; the caller launches the remainder before the score kernel
static int fill_desc(dist_b200_ctx *ctx, const dist_b200_feature *cf, const void *column, bool multi, FeatDesc &d,
                     GpTableBatch &tb, cudaStream_t s) {
    dist_b200_feature *f = const_cast<dist_b200_feature *>(cf);
    d.params = f->params;
    d.column = column;
    d.kind = f->model;
    d.vdim = f->dim;
    d.aux = nullptr;
    d.cap = f->capacity;
    d.pad_ = 0;
    if (multi && f->model == DIST_B200_GP) {
        if (f->gp_table_cap < f->capacity) {
            if (f->gp_table) {
                DISTB200_CUDA(ctx, cudaDeviceSynchronize());
                DISTB200_CUDA(ctx, cudaFree(f->gp_table));
                f->gp_table = nullptr;
            }
            DISTB200_CUDA(ctx, cudaMalloc(&f->gp_table, sizeof(float) * 2 * kGpTableX * f->capacity));  // [capacity][x] | [x][capacity]
            f->gp_table_cap = f->capacity;
            f->gp_table_dirty = true;
        }
        if (f->gp_table_dirty) {
            int rc = queue_gp_table(ctx, f, tb, s);
            if (rc) return rc;
        }
        d.params = f->gp_table;
        d.kind = kKindGpTable;
        d.vdim = kGpTableX;
        d.aux = f->params;
    }
    return DIST_B200_OK;
}

// rebuild every stale GammaPoisson value table of a multi-feature list on `s` (one launch per 128) and
// re-record the features' ready events, so that other streams order themselves behind the rebuild
static int refresh_gp_tables(dist_b200_ctx *ctx, const dist_b200_feature *const *features, int F, cudaStream_t s) {
    if (F < 2) return DIST_B200_OK;
    GpTableBatch tb;
    tb.n = 0;
    FeatDesc scratch_desc;
    bool any = false;
    for (int f = 0; f < F; ++f) {
        if (!features[f] || features[f]->model != DIST_B200_GP) continue;
        const bool stale = features[f]->gp_table_dirty || features[f]->gp_table_cap < features[f]->capacity;
        int rc = fill_desc(ctx, features[f], nullptr, true, scratch_desc, tb, s);
        if (rc) return rc;
        any = any || stale;
    }
    if (!any) return DIST_B200_OK;
    int rc = launch_gp_table_batch(ctx, tb, s);
    if (rc) return rc;
    for (int f = 0; f < F; ++f)
        if (features[f] && features[f]->model == DIST_B200_GP) DISTB200_CUDA(ctx, cudaEventRecord(features[f]->ready, s));
    return DIST_B200_OK;
}

static int score_dispatch(dist_b200_ctx *ctx, const dist_b200_feature *const *features, int F,
                          const void *const *columns, size_t N, const float *prior, const float *u,
                          int32_t *assign, float *scores, int accumulate, cudaStream_t s) {
    if (!ctx || !features || !columns || F < 1) return DIST_B200_ERR_INVALID;
    if (!assign && !scores) return fail(ctx, DIST_B200_ERR_INVALID, "score: no output requested");
    if (assign && !u) return fail(ctx, DIST_B200_ERR_INVALID, "score_sample: u is required");
    if (accumulate && prior) return fail(ctx, DIST_B200_ERR_INVALID, "score: prior is an overwrite, not valid with accumulate");
    const int G = features[0] ? features[0]->G : 0;
    if (G < 1) return fail(ctx, DIST_B200_ERR_STATE, "score: feature 0 has no groups (call update_all)");
    for (int f = 0; f < F; ++f) {
        if (!features[f] || !columns[f]) return fail(ctx, DIST_B200_ERR_INVALID, "score: null feature / column");
        if (features[f]->ctx != ctx) return fail(ctx, DIST_B200_ERR_INVALID, "score: feature belongs to another context");
        if (features[f]->G != G) return fail(ctx, DIST_B200_ERR_STATE, "score: features disagree on the number of groups");
    }
    if (N == 0) return DIST_B200_OK;
    {
        int rcw = wait_ready(ctx, features, F, s);
        if (rcw) return rcw;
    }
    // Row-mapped models (nich/gp/bb/dd) fuse into one launch.  dpd (value-major table, warp per row) and
    // niw (dense contraction) have their own kernels: alone they run directly, in mixed lists the scores
    // are materialised, every feature accumulates, and the stand-alone sampler finishes.
    auto solo = [](const dist_b200_feature *f) { return f->model == DIST_B200_DPD || f->model == DIST_B200_NIW; };
    int n_solo = 0;
    for (int f = 0; f < F; ++f) n_solo += solo(features[f]) ? 1 : 0;
    // ONE table feature (dpd / dd / bb), sampling only: every row with the same value has the same likelihood
    // vector, so it is evaluated once per distinct value (per-value CDF trees, table_rows.cu).  SURVEY 8(d)
    // "algorithmic shortcut"; DIST_B200_OPT_VALUE_CDF = 1 keeps the per-cell kernels (what bench.py reports beside it),
    // 2 = the round-2 tree search instead of the guide-table walk
    if (F == 1 && assign && !scores && !accumulate && ctx->opt[DIST_B200_OPT_VALUE_CDF] != 1 &&
        (features[0]->model == DIST_B200_DPD || features[0]->model == DIST_B200_DD || features[0]->model == DIST_B200_BB)) {
        dist_b200_feature *f = const_cast<dist_b200_feature *>(features[0]);
        const int R = f->model == DIST_B200_DPD ? f->dim + 1 : (f->model == DIST_B200_DD ? f->dim : 2);
        const size_t need = value_cdf_floats(R, G);
        if (need > f->cdf_floats) {
            if (f->cdf_buf) {
                DISTB200_CUDA(ctx, cudaDeviceSynchronize());
                DISTB200_CUDA(ctx, cudaFree(f->cdf_buf));
                f->cdf_buf = nullptr;
                f->cdf_floats = 0;
            }
            DISTB200_CUDA(ctx, cudaMalloc(&f->cdf_buf, need * sizeof(float)));
            f->cdf_floats = need;
        }
        const int rc = launch_value_cdf(ctx, f, f->cdf_buf, columns[0], N, prior, u, assign, s);
        if (rc != DIST_B200_ERR_UNSUPPORTED) return rc;
    }
    if (n_solo == 0) {
        if (F > kMaxFeatures) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "score: more than 512 features in one call");
        FeatList fl;
        fl.n = F;
        GpTableBatch tb;
        tb.n = 0;
        for (int f = 0; f < F; ++f) {
            int rc = fill_desc(ctx, features[f], columns[f], F > 1, fl.f[f], tb, s);
            if (rc) return rc;
        }
        if (int rc = launch_gp_table_batch(ctx, tb, s)) return rc;
        return launch_score_rows(ctx, fl, G, N, prior, u, assign, scores, accumulate, s);
    }
    // one NIW<32> feature, sampling only: quadratic forms, scores and sample_from_scores in one tcgen05 kernel
    if (F == 1 && features[0]->model == DIST_B200_NIW && features[0]->dim == 32 && features[0]->niw_tc && assign && !scores &&
        !accumulate && ctx->opt[DIST_B200_OPT_NIW_PATH] == 0)
        return launch_niw_tc(ctx, G, features[0]->niw_tc, columns[0], N, prior, nullptr, 0, u, assign, s);
    // one NIW feature with d <= 8, sampling only: the fused FP32 kernel (DIST_B200_OPT_NIW_PATH = 1 keeps the materialising route)
    if (F == 1 && features[0]->model == DIST_B200_NIW && features[0]->dim <= 8 && assign && !scores && !accumulate &&
        ctx->opt[DIST_B200_OPT_NIW_PATH] == 0) {
        const int rc = launch_niw_rows_small(ctx, features[0], columns[0], N, prior, u, assign, s);
        if (rc != DIST_B200_ERR_UNSUPPORTED) return rc;
    }
    if (F == 1 && features[0]->model == DIST_B200_DPD) {
        if (assign && !scores && !accumulate && ctx->opt[DIST_B200_OPT_TABLE_KERNEL] != 1) {
            const int rc = launch_table_rows(ctx, features[0], columns[0], N, prior, u, assign, s);
            if (rc != DIST_B200_ERR_UNSUPPORTED) return rc;
        }
        return launch_gather_rows(ctx, features[0], columns[0], N, prior, u, assign, scores, accumulate, s);
    }
    float *buf = scores;
    if (!buf) {
        int rc = ensure_scores_scratch(ctx, sizeof(float) * N * G);
        if (rc) return rc;
        buf = static_cast<float *>(ctx->scores_scratch);
    }
    FeatList fl;
    fl.n = 0;
    GpTableBatch tb;
    tb.n = 0;
    for (int f = 0; f < F; ++f) {
        if (solo(features[f])) continue;
        if (fl.n >= kMaxFeatures) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "score: more than 512 features in one call");
        int rc = fill_desc(ctx, features[f], columns[f], true, fl.f[fl.n++], tb, s);
        if (rc) return rc;
    }
    bool started = accumulate != 0;
    int rc;
    if ((rc = launch_gp_table_batch(ctx, tb, s))) return rc;
    if (fl.n) {
        if ((rc = launch_score_rows(ctx, fl, G, N, prior, nullptr, nullptr, buf, accumulate, s))) return rc;
        started = true;
    }
    for (int f = 0; f < F; ++f) {
        if (!solo(features[f])) continue;
        if (features[f]->model == DIST_B200_DPD)
            rc = launch_gather_rows(ctx, features[f], columns[f], N, started ? nullptr : prior, nullptr, nullptr, buf,
                                    started ? 1 : 0, s);
        else
            rc = launch_niw_scores(ctx, features[f], columns[f], N, started ? nullptr : prior, buf, started ? 1 : 0, s);
        if (rc) return rc;
        started = true;
    }
    if (assign) return launch_sample_scores(ctx, buf, N, G, u, assign, s);
    return DIST_B200_OK;
}

int dist_b200_score_batch(dist_b200_ctx *ctx, const dist_b200_feature *const *features, int n_features,
                          const void *const *columns_dev, size_t n_rows, const float *prior_dev, float *scores_dev,
                          int accumulate, void *stream) {
    if (!scores_dev) return ctx ? fail(ctx, DIST_B200_ERR_INVALID, "score_batch: scores_dev is null") : DIST_B200_ERR_INVALID;
    return score_dispatch(ctx, features, n_features, columns_dev, n_rows, prior_dev, nullptr, nullptr, scores_dev,
                          accumulate, as_stream(stream));
}

int dist_b200_score_sample_batch(dist_b200_ctx *ctx, const dist_b200_feature *const *features, int n_features,
                                 const void *const *columns_dev, size_t n_rows, const float *prior_dev,
                                 const float *u_dev, int32_t *assign_dev, float *scores_dev, void *stream) {
    if (!assign_dev) return ctx ? fail(ctx, DIST_B200_ERR_INVALID, "score_sample_batch: assign_dev is null") : DIST_B200_ERR_INVALID;
    return score_dispatch(ctx, features, n_features, columns_dev, n_rows, prior_dev, u_dev, assign_dev, scores_dev, 0,
                          as_stream(stream));
}

int dist_b200_sample_from_scores(dist_b200_ctx *ctx, const float *scores_dev, size_t n_rows, int G,
                                 const float *u_dev, int32_t *assign_dev, void *stream) {
    if (!ctx || !scores_dev || !u_dev || !assign_dev || G < 1) return DIST_B200_ERR_INVALID;
    return launch_sample_scores(ctx, scores_dev, n_rows, G, u_dev, assign_dev, as_stream(stream));
}

int dist_b200_peer_alloc(dist_b200_ctx *ctx, size_t bytes, void **dev_ptr, unsigned char handle_out[64]) {
    if (!ctx || !dev_ptr || !handle_out || bytes == 0) return DIST_B200_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DISTB200_CUDA(ctx, cudaMalloc(dev_ptr, bytes));
    DISTB200_CUDA(ctx, cudaMemset(*dev_ptr, 0, bytes));  // epoch flags start at 0
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
    if (e != cudaSuccess) {
        cudaFree(*dev_ptr);
        *dev_ptr = nullptr;
        return fail(ctx, DIST_B200_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    std::memcpy(handle_out, &h, 64);
    return DIST_B200_OK;
}

int dist_b200_peer_open(dist_b200_ctx *ctx, const unsigned char handle[64], void **dev_ptr) {
    if (!ctx || !handle || !dev_ptr) return DIST_B200_ERR_INVALID;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    DISTB200_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DIST_B200_OK;
}

int dist_b200_peer_close(dist_b200_ctx *ctx, void *dev_ptr) {
    if (!ctx || !dev_ptr) return DIST_B200_ERR_INVALID;
    DISTB200_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return DIST_B200_OK;
}

int dist_b200_peer_free(dist_b200_ctx *ctx, void *dev_ptr) {
    if (!ctx || !dev_ptr) return DIST_B200_ERR_INVALID;
    DISTB200_CUDA(ctx, cudaDeviceSynchronize());
    DISTB200_CUDA(ctx, cudaFree(dev_ptr));
    return DIST_B200_OK;
}

int dist_b200_score_push_batch(dist_b200_ctx *ctx, const dist_b200_feature *const *features, int F,
                               const void *const *columns_dev, size_t n_rows, size_t row0, const float *prior_dev,
                               void *const *slot_ptrs, int n_owners, size_t block_rows, void *stream) {
    if (!ctx || !features || !columns_dev || F < 1 || !slot_ptrs || n_owners < 1 || block_rows == 0) return DIST_B200_ERR_INVALID;
    if (n_owners > kMaxPushOwners) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "score_push: more than 16 owners");
    if (F > kMaxFeatures) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "score_push: more than 512 features in one call");
    const int G = features[0] ? features[0]->G : 0;
    if (G < 1) return fail(ctx, DIST_B200_ERR_STATE, "score_push: feature 0 has no groups");
    if ((row0 + n_rows + block_rows - 1) / block_rows > static_cast<size_t>(n_owners))
        return fail(ctx, DIST_B200_ERR_INVALID, "score_push: rows extend past the last owner's block");
    cudaStream_t s = as_stream(stream);
    int rc = wait_ready(ctx, features, F, s);
    if (rc) return rc;
    FeatList fl;
    fl.n = F;
    GpTableBatch tb;
    tb.n = 0;
    for (int f = 0; f < F; ++f) {
        if (!features[f] || !columns_dev[f] || features[f]->ctx != ctx || features[f]->G != G)
            return fail(ctx, DIST_B200_ERR_INVALID, "score_push: bad feature list");
        if (features[f]->model == DIST_B200_DPD || features[f]->model == DIST_B200_NIW)
            return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "score_push: row-mapped models only");
        if ((rc = fill_desc(ctx, features[f], columns_dev[f], true, fl.f[f], tb, s))) return rc;
    }
    if ((rc = launch_gp_table_batch(ctx, tb, s))) return rc;
    PushTargets pt{};
    pt.n = n_owners;
    pt.row0 = row0;
    pt.block_rows = block_rows;
    for (int i = 0; i < n_owners; ++i) pt.ptr[i] = static_cast<float *>(slot_ptrs[i]);
    return launch_score_rows(ctx, fl, G, n_rows, prior_dev, nullptr, nullptr, nullptr, 0, s, &pt);
}

int dist_b200_sample_from_slots(dist_b200_ctx *ctx, const float *slots_dev, int n_slots, size_t slot_stride,
                                size_t n_rows, int G, const float *u_dev, int32_t *assign_dev, void *stream) {
    if (!ctx || !slots_dev || n_slots < 1 || !u_dev || !assign_dev || G < 1) return DIST_B200_ERR_INVALID;
    return launch_sample_scores(ctx, slots_dev, n_rows, G, u_dev, assign_dev, as_stream(stream), n_slots, slot_stride);
}

static size_t value_bytes(const dist_b200_feature *f) {
    switch (f->model) {
        case DIST_B200_BB: return 1;
        case DIST_B200_NIW: return sizeof(float) * f->dim;
        default: return 4;
    }
}

int dist_b200_score_sample_batch_host(dist_b200_ctx *ctx, const dist_b200_feature *const *features, int F,
                                      const void *const *columns_host, size_t N, const float *prior_host,
                                      const float *u_host, int32_t *assign_host, float *scores_host) {
    if (!ctx || !features || !columns_host || F < 1 || F > kMaxFeatures || !u_host || !assign_host) return DIST_B200_ERR_INVALID;
    for (int f = 0; f < F; ++f)
        if (!features[f] || !columns_host[f]) return fail(ctx, DIST_B200_ERR_INVALID, "score_sample_host: null feature / column");
    const int G = features[0]->G;
    if (G < 1) return fail(ctx, DIST_B200_ERR_STATE, "score_sample_host: no groups");
    if (N == 0) return DIST_B200_OK;
    // layout of the staging area (identical in pinned host memory and on the device):
    //   columns | prior | u | assign | scores(optional, device only + host out)
    std::vector<size_t> col_off(F);
    size_t off = 0;
    for (int f = 0; f < F; ++f) {
        col_off[f] = off;
        off += round_up(value_bytes(features[f]) * N, 256);
    }
    const size_t prior_off = off;
    off += round_up(sizeof(float) * G, 256);
    const size_t u_off = off;
    off += round_up(sizeof(float) * N, 256);
    const size_t in_bytes = off;
    const size_t assign_off = off;
    off += round_up(sizeof(int32_t) * N, 256);
    const size_t scores_off = off;
    if (scores_host) off += round_up(sizeof(float) * N * G, 256);
    (void)in_bytes;
    int rc;
    if ((rc = ensure_pinned(ctx, scores_off))) return rc;
    if ((rc = ensure_scratch(ctx, off + 256))) return rc;
    char *pin = static_cast<char *>(ctx->pinned);
    char *dev = static_cast<char *>(ctx->scratch_dev);
    // A caller buffer that is already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory)
    // is copied from / to directly; pageable buffers go through the context's pinned staging area.
    auto is_pinned = [](const void *p) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return at.type == cudaMemoryTypeHost;
    };
    std::vector<char> col_pinned(F);
    for (int f = 0; f < F; ++f) col_pinned[f] = is_pinned(columns_host[f]) ? 1 : 0;
    const bool u_pinned = is_pinned(u_host), assign_pinned = is_pinned(assign_host);
    const bool scores_pinned = scores_host && is_pinned(scores_host);

    // Row chunks are pipelined over two streams: the H2D copy of chunk k+1 and the D2H copy of chunk
    // k-1 overlap the kernels of chunk k (rows are independent given frozen statistics).
    cudaStream_t st[2] = {ctx->own_stream, ctx->own_stream2};
    if ((rc = wait_ready(ctx, features, F, st[0]))) return rc;
    if ((rc = refresh_gp_tables(ctx, features, F, st[0]))) return rc;  // before the second stream starts reading them
    const float *prior_dev = nullptr;
    // Zero-copy: when every caller buffer is page-locked, the kernels read the rows straight from host memory and write
    // the assignments straight back (device-accessible under unified addressing): ONE launch, the PCIe transfers overlap
    // the math row tile by row tile with no chunking, no staging copies and no second stream.
    // DIST_B200_OPT_HOST_ZEROCOPY: 0 = when all buffers are page-locked, 1 = never, 2 = same as 0 (A/B runs).
    {
        bool all_pinned = u_pinned && assign_pinned && !scores_host;
        for (int f = 0; f < F; ++f) all_pinned = all_pinned && col_pinned[f];
        // niw keeps the staged path: its row images are built by a separate pack kernel, which would read the rows over
        // PCIe without any math to hide behind (measured: c5 e2e 6.1e10 staged vs 4.1e10 zero-copy)
        bool staged_only = ctx->opt[DIST_B200_OPT_HOST_ZEROCOPY] == 1;
        for (int f = 0; f < F; ++f) staged_only = staged_only || features[f]->model == DIST_B200_NIW;
        if (all_pinned && !staged_only) {
            std::vector<const void *> zc_cols(F);
            void *dp = nullptr;
            bool ok = true;
            for (int f = 0; f < F && ok; ++f) {
                ok = cudaHostGetDevicePointer(&dp, const_cast<void *>(columns_host[f]), 0) == cudaSuccess;
                zc_cols[f] = dp;
            }
            const float *u_zc = nullptr;
            int32_t *assign_zc = nullptr;
            if (ok) ok = cudaHostGetDevicePointer(&dp, const_cast<float *>(u_host), 0) == cudaSuccess;
            u_zc = static_cast<const float *>(dp);
            if (ok) ok = cudaHostGetDevicePointer(&dp, assign_host, 0) == cudaSuccess;
            assign_zc = static_cast<int32_t *>(dp);
            if (ok) {
                const float *prior_zc = nullptr;
                if (prior_host) {  // the prior vector is re-read per table row by the dpd / dd kernels: it goes to the device
                    std::memcpy(pin + prior_off, prior_host, sizeof(float) * G);
                    if (cudaMemcpyAsync(dev + prior_off, pin + prior_off, sizeof(float) * G, cudaMemcpyHostToDevice, st[0]) != cudaSuccess)
                        return fail(ctx, DIST_B200_ERR_CUDA, "score_sample_host: prior upload");
                    prior_zc = reinterpret_cast<const float *>(dev + prior_off);
                }
                rc = score_dispatch(ctx, features, F, zc_cols.data(), N, prior_zc, u_zc, assign_zc, nullptr, 0, st[0]);
                cudaError_t e = cudaStreamSynchronize(st[0]);
                if (rc) return rc;
                if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("score_sample_host: ") + cudaGetErrorString(e));
                return DIST_B200_OK;
            }
            cudaGetLastError();  // not mappable: fall through to the staged path
        }
    }
    if ((rc = wait_ready(ctx, features, F, st[1]))) return rc;
    if (prior_host) {
        std::memcpy(pin + prior_off, prior_host, sizeof(float) * G);
        DISTB200_CUDA(ctx, cudaMemcpyAsync(dev + prior_off, pin + prior_off, sizeof(float) * G, cudaMemcpyHostToDevice, st[0]));
        DISTB200_CUDA(ctx, cudaEventRecord(ctx->ev, st[0]));
        DISTB200_CUDA(ctx, cudaStreamWaitEvent(st[1], ctx->ev, 0));
        prior_dev = reinterpret_cast<const float *>(dev + prior_off);
    }
    const size_t min_chunk = 32768;
    // five chunks measured best at 1M rows (DIST_B200_OPT_HOST_CHUNKS overrides it for A/B runs)
    const size_t max_chunks = ctx->opt[DIST_B200_OPT_HOST_CHUNKS] > 0 ? static_cast<size_t>(ctx->opt[DIST_B200_OPT_HOST_CHUNKS]) : 5;
    size_t nchunks = std::min<size_t>(max_chunks, std::max<size_t>(1, N / min_chunk));
    // niw launches share the context's packed-row buffer and mixed dpd lists its single scores buffer: their kernels stay
    // on ONE stream in chunk order; only the H2D copies of the later chunks run ahead on the second stream
    bool one_compute_stream = false;
    for (int f = 0; f < F; ++f)
        if (features[f]->model == DIST_B200_NIW || (features[f]->model == DIST_B200_DPD && F > 1)) one_compute_stream = true;
    if (one_compute_stream && scores_host) nchunks = 1;
    // chunk boundaries: equal interior chunks, half-size first and last ones -- the first H2D copy and the
    // last D2H copy are the only transfers no kernel hides
    std::vector<size_t> bounds(1, 0);
    if (nchunks >= 4) {
        const size_t unit = round_up((N + 2 * (nchunks - 1) - 1) / (2 * (nchunks - 1)), 256);  // half an interior chunk
        size_t at = unit;
        while (at < N && bounds.size() < nchunks - 1) {
            bounds.push_back(at);
            at += 2 * unit;
        }
        const size_t last = N > unit ? (N - unit) / 256 * 256 : 0;  // 256-row aligned like every other boundary
        if (last > bounds.back()) bounds.push_back(last);
    } else {
        const size_t chunk = round_up((N + nchunks - 1) / nchunks, 256);
        for (size_t at = chunk; at < N; at += chunk) bounds.push_back(at);
    }
    bounds.push_back(N);
    std::vector<const void *> cols(F);
    float *scores_dev = scores_host ? reinterpret_cast<float *>(dev + scores_off) : nullptr;
    if (one_compute_stream && bounds.size() > 2) {
        // copies of all chunks on the second stream, one event each; the kernels follow on the first stream
        std::vector<cudaEvent_t> ready(bounds.size() - 1, nullptr);
        auto cleanup = [&]() {
            cudaStreamSynchronize(st[0]);
            cudaStreamSynchronize(st[1]);
            for (cudaEvent_t e : ready)
                if (e) cudaEventDestroy(e);
        };
        cudaError_t ce = cudaSuccess;
        for (size_t k = 0; k + 1 < bounds.size() && ce == cudaSuccess; ++k) {
            const size_t lo = bounds[k], n = bounds[k + 1] - lo;
            for (int f = 0; f < F && ce == cudaSuccess; ++f) {
                const size_t vb = value_bytes(features[f]);
                const char *src = static_cast<const char *>(columns_host[f]) + vb * lo;
                if (!col_pinned[f]) {
                    std::memcpy(pin + col_off[f] + vb * lo, src, vb * n);
                    src = pin + col_off[f] + vb * lo;
                }
                ce = cudaMemcpyAsync(dev + col_off[f] + vb * lo, src, vb * n, cudaMemcpyHostToDevice, st[1]);
            }
            const char *src = reinterpret_cast<const char *>(u_host + lo);
            if (!u_pinned) {
                std::memcpy(pin + u_off + 4 * lo, src, 4 * n);
                src = pin + u_off + 4 * lo;
            }
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(dev + u_off + 4 * lo, src, 4 * n, cudaMemcpyHostToDevice, st[1]);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventRecord(ready[k], st[1]);
        }
        for (size_t k = 0; k + 1 < bounds.size() && ce == cudaSuccess && rc == DIST_B200_OK; ++k) {
            const size_t lo = bounds[k], n = bounds[k + 1] - lo;
            if (n == 0) continue;
            ce = cudaStreamWaitEvent(st[0], ready[k], 0);
            for (int f = 0; f < F; ++f) cols[f] = dev + col_off[f] + value_bytes(features[f]) * lo;
            if (ce == cudaSuccess)
                rc = score_dispatch(ctx, features, F, cols.data(), n, prior_dev, reinterpret_cast<const float *>(dev + u_off) + lo,
                                    reinterpret_cast<int32_t *>(dev + assign_off) + lo, nullptr, 0, st[0]);
        }
        int32_t *adst = assign_pinned ? assign_host : reinterpret_cast<int32_t *>(pin + assign_off);
        if (ce == cudaSuccess && rc == DIST_B200_OK) ce = cudaMemcpyAsync(adst, dev + assign_off, 4 * N, cudaMemcpyDeviceToHost, st[0]);
        cleanup();
        if (rc) return rc;
        if (ce != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("score_sample_host: ") + cudaGetErrorString(ce));
        if (!assign_pinned) std::memcpy(assign_host, pin + assign_off, sizeof(int32_t) * N);
        return DIST_B200_OK;
    }
    for (size_t k = 0; k + 1 < bounds.size(); ++k) {
        const size_t lo = bounds[k], n = bounds[k + 1] - lo;
        if (n == 0) continue;
        cudaStream_t s = st[k & 1];
        for (int f = 0; f < F; ++f) {
            const size_t vb = value_bytes(features[f]);
            const char *src = static_cast<const char *>(columns_host[f]) + vb * lo;
            if (!col_pinned[f]) {
                std::memcpy(pin + col_off[f] + vb * lo, src, vb * n);
                src = pin + col_off[f] + vb * lo;
            }
            DISTB200_CUDA(ctx, cudaMemcpyAsync(dev + col_off[f] + vb * lo, src, vb * n, cudaMemcpyHostToDevice, s));
            cols[f] = dev + col_off[f] + vb * lo;
        }
        {
            const char *src = reinterpret_cast<const char *>(u_host + lo);
            if (!u_pinned) {
                std::memcpy(pin + u_off + 4 * lo, src, 4 * n);
                src = pin + u_off + 4 * lo;
            }
            DISTB200_CUDA(ctx, cudaMemcpyAsync(dev + u_off + 4 * lo, src, 4 * n, cudaMemcpyHostToDevice, s));
        }
        rc = score_dispatch(ctx, features, F, cols.data(), n, prior_dev, reinterpret_cast<const float *>(dev + u_off) + lo,
                            reinterpret_cast<int32_t *>(dev + assign_off) + lo, scores_dev ? scores_dev + lo * G : nullptr, 0, s);
        if (rc) {  // copies of earlier chunks into the caller's buffers may still be in flight
            cudaStreamSynchronize(st[0]);
            cudaStreamSynchronize(st[1]);
            return rc;
        }
        int32_t *adst = assign_pinned ? assign_host + lo : reinterpret_cast<int32_t *>(pin + assign_off) + lo;
        DISTB200_CUDA(ctx, cudaMemcpyAsync(adst, dev + assign_off + 4 * lo, 4 * n, cudaMemcpyDeviceToHost, s));
        if (scores_host)
            DISTB200_CUDA(ctx, cudaMemcpyAsync(scores_host + lo * G, scores_dev + lo * G, sizeof(float) * n * G,
                                               cudaMemcpyDeviceToHost, s));
    }
    (void)scores_pinned;
    DISTB200_CUDA(ctx, cudaStreamSynchronize(st[0]));
    DISTB200_CUDA(ctx, cudaStreamSynchronize(st[1]));
    if (!assign_pinned) std::memcpy(assign_host, pin + assign_off, sizeof(int32_t) * N);
    return DIST_B200_OK;
}

int dist_b200_host_register(dist_b200_ctx *ctx, void *ptr, size_t bytes) {
    if (!ctx || !ptr || bytes == 0) return DIST_B200_ERR_INVALID;
    DISTB200_CUDA(ctx, cudaSetDevice(ctx->device));
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return DIST_B200_OK;
    }
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

int dist_b200_host_unregister(dist_b200_ctx *ctx, void *ptr) {
    if (!ctx || !ptr) return DIST_B200_ERR_INVALID;
    DISTB200_CUDA(ctx, cudaSetDevice(ctx->device));
    DISTB200_CUDA(ctx, cudaDeviceSynchronize());  // no kernel may still be reading the range
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, DIST_B200_ERR_INVALID, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
    }
    return DIST_B200_OK;
}

int dist_b200_score_value_host(dist_b200_ctx *ctx, const dist_b200_feature *feature, const void *value_host,
                               float *scores_accum_host) {
    if (!ctx || !feature || !value_host || !scores_accum_host) return DIST_B200_ERR_INVALID;
    const int G = feature->G;
    if (G < 1) return fail(ctx, DIST_B200_ERR_STATE, "score_value: no groups");
    const size_t vb = value_bytes(feature);
    // zero-copy through the context's page-locked area (device-accessible under unified addressing): the kernel reads
    // the value from host memory and writes the G scores back to it, so a call is ONE launch + one synchronise instead
    // of two H2D copies, a launch and a D2H copy; the accumulation into the caller's (pageable) buffer is done here
    int rc = ensure_pinned(ctx, round_up(vb, 256) + sizeof(float) * G + 256);
    if (rc) return rc;
    char *pin = static_cast<char *>(ctx->pinned);
    float *sc = reinterpret_cast<float *>(pin + round_up(vb, 256));
    cudaStream_t s = ctx->own_stream;
    if ((rc = wait_ready(ctx, &feature, 1, s))) return rc;
    std::memcpy(pin, value_host, vb);
    const void *col = pin;
    rc = score_dispatch(ctx, &feature, 1, &col, 1, nullptr, nullptr, nullptr, sc, 0, s);
    if (rc) return rc;
    DISTB200_CUDA(ctx, cudaStreamSynchronize(s));
    for (int g = 0; g < G; ++g) scores_accum_host[g] += sc[g];  // MixtureSlave::score_value accumulates (mixture.hpp:416-425)
    return DIST_B200_OK;
}

int dist_b200_numerics_probe(dist_b200_ctx *ctx, int fn, size_t n, const float *in_dev, float *out_dev, void *stream) {
    if (!ctx || !in_dev || !out_dev) return DIST_B200_ERR_INVALID;
    return launch_numerics_probe(ctx, fn, n, in_dev, out_dev, as_stream(stream));
}

}  // extern "C"
