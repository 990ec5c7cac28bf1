// stats.cu -- batched Group::add_value on the device (first "next" row).
//
// The reference adds one value to one group at a time (nich.hpp:125-133, gp.hpp:109-116, bb.hpp:102-107,
// dd.hpp:123-130, dpd.hpp:188-196) and refreshes that group's cache entries (update_group).  After a
// batched score + sample, the N sampled assignments are folded into the device-resident group statistics
// with one segmented reduction (shared-memory accumulators per block, then global atomics), merged into
// the stored statistics, and the caches are rebuilt by the same prep code update_all uses -- the
// score -> sample -> update loop never returns to the host.  All pooled features (nich / gp / bb) of one
// kind go through ONE accumulate launch (grid.y = feature) and ONE merge + cache-rebuild launch (prep.cu).
//
//   nich : per group (m, sum x, sum x^2) in double -> (m, mean_b, ctv_b) -> merged with the stored
//          (count, mean, count_times_variance) by the pairwise formula of Group::merge (nich.hpp:167-179).
//          The reference's one-at-a-time Welford updates round differently: statistics agree to ~1e-6
//          relative, counts exactly.
//   gp   : count += m, sum += sum x (uint32, wraps like the reference's)   bb : heads / tails += counts (exact)
//   dd   : counts[g][x] += 1 (exact)                            dpd: counts[g][row(x)] += 1 (exact; OTHER skipped)
#include "common.cuh"

namespace distb200 {

constexpr int kAddThreads = 256;
constexpr int kAddSmemGroups = 2048;   // groups whose accumulators fit in shared memory (24 B / group)
constexpr int kCountSmemBins = 12288;  // int bins of a shared-memory histogram (48 KB)

size_t add_rows_acc_bytes(int G) {
    // cnt_a | cnt_b | sum_x | sum_xx, each G entries, 256-byte aligned regions
    const size_t g = static_cast<size_t>(G);
    return 2 * ((g * 4 + 255) / 256 * 256) + 2 * ((g * 8 + 255) / 256 * 256);
}

// ---------------------------------------------------------------------------------------------
// pooled models: grid (row tiles, features)
template <bool kSmem>
__global__ void __launch_bounds__(kAddThreads) add_rows_pooled_kernel(const AddBatch b) {
    extern __shared__ __align__(8) unsigned char add_smem[];
    const int G = b.G;
    const AddDesc &d = b.d[blockIdx.y];
    char *acc = b.acc + b.acc_stride * blockIdx.y;
    const size_t ci = (static_cast<size_t>(G) * 4 + 255) / 256 * 256, di = (static_cast<size_t>(G) * 8 + 255) / 256 * 256;
    int *g_ca = reinterpret_cast<int *>(acc);
    int *g_cb = reinterpret_cast<int *>(acc + ci);
    double *g_x = reinterpret_cast<double *>(acc + 2 * ci);
    double *g_xx = reinterpret_cast<double *>(acc + 2 * ci + di);
    double *s_x = reinterpret_cast<double *>(add_smem);
    double *s_xx = s_x + G;
    int *s_ca = reinterpret_cast<int *>(s_xx + G);
    int *s_cb = s_ca + G;
    const int model = d.model;
    if (kSmem) {
        for (int g = threadIdx.x; g < G; g += kAddThreads) {
            s_x[g] = 0.0;
            s_xx[g] = 0.0;
            s_ca[g] = 0;
            s_cb[g] = 0;
        }
        __syncthreads();
    }
    int *ca = kSmem ? s_ca : g_ca, *cb = kSmem ? s_cb : g_cb;
    double *ax = kSmem ? s_x : g_x, *axx = kSmem ? s_xx : g_xx;
    const size_t stride = static_cast<size_t>(gridDim.x) * kAddThreads;
    const size_t first = static_cast<size_t>(blockIdx.x) * kAddThreads + threadIdx.x;
    if (model == DIST_B200_NICH) {
        const float *col = static_cast<const float *>(d.column);
        for (size_t n = first; n < b.N; n += stride) {
            const int g = b.assign[n];
            if (g < 0 || g >= G) continue;
            const double x = static_cast<double>(col[n]);
            atomicAdd(&ca[g], 1);
            atomicAdd(&ax[g], x);
            atomicAdd(&axx[g], x * x);
        }
    } else if (model == DIST_B200_GP) {
        const uint32_t *col = static_cast<const uint32_t *>(d.column);
        for (size_t n = first; n < b.N; n += stride) {
            const int g = b.assign[n];
            if (g < 0 || g >= G) continue;
            atomicAdd(&ca[g], 1);
            atomicAdd(reinterpret_cast<unsigned int *>(&cb[g]), col[n]);
        }
    } else {  // bb
        const uint8_t *col = static_cast<const uint8_t *>(d.column);
        for (size_t n = first; n < b.N; n += stride) {
            const int g = b.assign[n];
            if (g < 0 || g >= G) continue;
            atomicAdd(col[n] != 0 ? &ca[g] : &cb[g], 1);
        }
    }
    if (kSmem) {
        __syncthreads();
        for (int g = threadIdx.x; g < G; g += kAddThreads) {
            if (s_ca[g]) atomicAdd(&g_ca[g], s_ca[g]);
            if (s_cb[g]) atomicAdd(&g_cb[g], s_cb[g]);
            if (model == DIST_B200_NICH && s_ca[g]) {
                atomicAdd(&g_x[g], s_x[g]);
                atomicAdd(&g_xx[g], s_xx[g]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dd / dpd: counts[g][row] += 1 in place; a shared-memory histogram when the whole table fits
struct CountArgs {
    int model, G, dim, keys_dense;
    size_t N;
    const void *column;
    const int32_t *assign;
    int32_t *counts;
    const uint32_t *keys;
    const int *key_rows;
};

__device__ __forceinline__ int dpd_row(const CountArgs &a, uint32_t value) {
    if (a.keys_dense) return value < static_cast<uint32_t>(a.dim) ? static_cast<int>(value) : -1;
    int lo = 0, hi = a.dim;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a.keys[mid] < value) lo = mid + 1;
        else hi = mid;
    }
    return (lo < a.dim && a.keys[lo] == value) ? a.key_rows[lo] : -1;
}

template <bool kSmem>
__global__ void __launch_bounds__(kAddThreads) add_rows_counts_kernel(const CountArgs a) {
    extern __shared__ int32_t bins[];
    const int cells = a.G * a.dim;
    if (kSmem) {
        for (int i = threadIdx.x; i < cells; i += kAddThreads) bins[i] = 0;
        __syncthreads();
    }
    int32_t *dst = kSmem ? bins : a.counts;
    for (size_t n = static_cast<size_t>(blockIdx.x) * kAddThreads + threadIdx.x; n < a.N;
         n += static_cast<size_t>(gridDim.x) * kAddThreads) {
        const int g = a.assign[n];
        if (g < 0 || g >= a.G) continue;
        int r;
        if (a.model == DIST_B200_DD) {
            r = static_cast<const int32_t *>(a.column)[n];
            if (r < 0 || r >= a.dim) continue;
        } else {
            r = dpd_row(a, static_cast<const uint32_t *>(a.column)[n]);
            if (r < 0) continue;
        }
        atomicAdd(&dst[static_cast<size_t>(g) * a.dim + r], 1);
    }
    if (kSmem) {
        __syncthreads();
        for (int i = threadIdx.x; i < cells; i += kAddThreads)
            if (bins[i]) atomicAdd(&a.counts[i], bins[i]);
    }
}

template <bool kSmem>
__global__ void __launch_bounds__(256) count_assignments_kernel(const int32_t *__restrict__ assign, size_t N, int G,
                                                                int32_t *__restrict__ counts) {
    extern __shared__ int32_t bins[];
    if (kSmem) {
        for (int i = threadIdx.x; i < G; i += 256) bins[i] = 0;
        __syncthreads();
    }
    int32_t *dst = kSmem ? bins : counts;
    for (size_t n = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; n < N; n += static_cast<size_t>(gridDim.x) * 256) {
        const int g = assign[n];
        if (g >= 0 && g < G) atomicAdd(&dst[g], 1);
    }
    if (kSmem) {
        __syncthreads();
        for (int i = threadIdx.x; i < G; i += 256)
            if (bins[i]) atomicAdd(&counts[i], bins[i]);
    }
}

// ---------------------------------------------------------------------------------------------
static unsigned row_tiles(const dist_b200_ctx *ctx, size_t N, int per_sm, int split) {
    // enough blocks to fill the machine, few enough that the per-block flush stays negligible
    const size_t want = (N + kAddThreads - 1) / kAddThreads;
    size_t cap = static_cast<size_t>(ctx->sm_count) * per_sm / (split > 0 ? split : 1);
    if (cap < 1) cap = 1;
    return static_cast<unsigned>(want < cap ? want : cap);
}

int launch_add_rows_pooled(dist_b200_ctx *ctx, const AddBatch &b, cudaStream_t s) {
    if (b.N == 0 || b.G == 0 || b.n == 0) return DIST_B200_OK;
    DISTB200_CUDA(ctx, cudaMemsetAsync(b.acc, 0, b.acc_stride * b.n, s));
    const dim3 grid(row_tiles(ctx, b.N, 8, b.n), b.n);
    if (b.G <= kAddSmemGroups) {
        add_rows_pooled_kernel<true><<<grid, kAddThreads, static_cast<size_t>(b.G) * 24, s>>>(b);
    } else {
        add_rows_pooled_kernel<false><<<grid, kAddThreads, 0, s>>>(b);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("add_rows launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

int launch_add_rows_counts(dist_b200_ctx *ctx, dist_b200_feature *f, const void *column, const int32_t *assign, size_t N,
                           cudaStream_t s) {
    if (N == 0 || f->G == 0) return DIST_B200_OK;
    CountArgs a{};
    a.model = f->model;
    a.G = f->G;
    a.dim = f->dim;
    a.keys_dense = f->keys_dense ? 1 : 0;
    a.N = N;
    a.column = column;
    a.assign = assign;
    a.counts = reinterpret_cast<int32_t *>(f->stats);  // dd: counts[G][dim]; dpd: counts[G][V]
    a.keys = f->keys_dev;
    a.key_rows = f->key_rows_dev;
    const size_t cells = static_cast<size_t>(f->G) * f->dim;
    if (cells <= kCountSmemBins) {
        add_rows_counts_kernel<true><<<row_tiles(ctx, N, 4, 1), kAddThreads, cells * 4, s>>>(a);
    } else {
        add_rows_counts_kernel<false><<<row_tiles(ctx, N, 8, 1), kAddThreads, 0, s>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("add_rows launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

int launch_count_assignments(dist_b200_ctx *ctx, const int32_t *assign, size_t N, int G, int32_t *counts, int accumulate,
                             cudaStream_t s) {
    if (!accumulate) DISTB200_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int32_t) * G, s));
    if (N == 0) return DIST_B200_OK;
    const size_t want = (N + 255) / 256;
    const size_t cap = static_cast<size_t>(ctx->sm_count) * 4;
    const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
    if (G <= kCountSmemBins) count_assignments_kernel<true><<<grid, 256, static_cast<size_t>(G) * 4, s>>>(assign, N, G, counts);
    else count_assignments_kernel<false><<<grid, 256, 0, s>>>(assign, N, G, counts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("count_assignments launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

}  // namespace distb200
