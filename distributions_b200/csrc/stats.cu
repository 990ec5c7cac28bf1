// stats.cu -- batched Group::add_value on the device (first "next" row).
//
// The reference adds one value to one group at a time (nich.hpp:125-133, gp.hpp:109-116, bb.hpp:102-107,
// dd.hpp:123-130, dpd.hpp:188-196) and refreshes that group's cache entries (update_group).  After a
// batched score + sample, the N sampled assignments are folded into the device-resident group statistics
// with one segmented reduction (shared-memory accumulators per block, then global atomics), merged into
// the stored statistics, and the caches are rebuilt by the same *_prep kernels update_all uses -- the
// score -> sample -> update loop never returns to the host.
//
//   nich : per group (m, sum x, sum x^2) in double -> (m, mean_b, ctv_b) -> merged with the stored
//          (count, mean, count_times_variance) by the pairwise formula of Group::merge (nich.hpp:167-179).
//          The reference's one-at-a-time Welford updates round differently: statistics agree to ~1e-6
//          relative, counts exactly.
//   gp   : count += m, sum += sum x (integers, exact)          bb : heads / tails += counts (exact)
//   dd   : counts[g][x] += 1 (exact)                            dpd: counts[g][row(x)] += 1 (exact; OTHER skipped)
#include "common.cuh"

namespace distb200 {

constexpr int kAddThreads = 256;
constexpr int kAddSmemGroups = 2048;  // groups whose accumulators fit in shared memory (20 B / group)

struct AddArgs {
    int model, G, dim, keys_dense;
    size_t N;
    const void *column;
    const int32_t *assign;
    // nich / gp / bb batch accumulators (context scratch, zeroed)
    int *cnt_a;       // nich: m ; gp: m ; bb: heads
    int *cnt_b;       // bb: tails
    double *sum_x;    // nich: sum x ; gp: sum x (exact in double below 2^53)
    double *sum_xx;   // nich: sum x^2
    // dd / dpd: the persistent counts themselves
    int32_t *counts;
    const uint32_t *keys;
    const int *key_rows;
};

__device__ __forceinline__ int dpd_row(const AddArgs &a, uint32_t value) {
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
__global__ void __launch_bounds__(kAddThreads) add_rows_kernel(const AddArgs a) {
    extern __shared__ __align__(8) unsigned char add_smem[];
    const int G = a.G;
    double *s_x = reinterpret_cast<double *>(add_smem);
    double *s_xx = s_x + G;
    int *s_ca = reinterpret_cast<int *>(s_xx + G);
    int *s_cb = s_ca + G;
    const bool pooled = a.model == DIST_B200_NICH || a.model == DIST_B200_GP || a.model == DIST_B200_BB;
    if (kSmem && pooled) {
        for (int g = threadIdx.x; g < G; g += kAddThreads) {
            s_x[g] = 0.0;
            s_xx[g] = 0.0;
            s_ca[g] = 0;
            s_cb[g] = 0;
        }
        __syncthreads();
    }
    for (size_t n = static_cast<size_t>(blockIdx.x) * kAddThreads + threadIdx.x; n < a.N;
         n += static_cast<size_t>(gridDim.x) * kAddThreads) {
        const int g = a.assign[n];
        if (g < 0 || g >= G) continue;
        switch (a.model) {
            case DIST_B200_NICH: {
                const double x = static_cast<double>(static_cast<const float *>(a.column)[n]);
                if (kSmem) {
                    atomicAdd(&s_ca[g], 1);
                    atomicAdd(&s_x[g], x);
                    atomicAdd(&s_xx[g], x * x);
                } else {
                    atomicAdd(&a.cnt_a[g], 1);
                    atomicAdd(&a.sum_x[g], x);
                    atomicAdd(&a.sum_xx[g], x * x);
                }
            } break;
            case DIST_B200_GP: {
                const double x = static_cast<double>(static_cast<const uint32_t *>(a.column)[n]);
                if (kSmem) {
                    atomicAdd(&s_ca[g], 1);
                    atomicAdd(&s_x[g], x);
                } else {
                    atomicAdd(&a.cnt_a[g], 1);
                    atomicAdd(&a.sum_x[g], x);
                }
            } break;
            case DIST_B200_BB: {
                const bool v = static_cast<const uint8_t *>(a.column)[n] != 0;
                if (kSmem) atomicAdd(v ? &s_ca[g] : &s_cb[g], 1);
                else atomicAdd(v ? &a.cnt_a[g] : &a.cnt_b[g], 1);
            } break;
            case DIST_B200_DD: {
                const int v = static_cast<const int32_t *>(a.column)[n];
                if (v >= 0 && v < a.dim) atomicAdd(&a.counts[static_cast<size_t>(g) * a.dim + v], 1);
            } break;
            case DIST_B200_DPD: {
                const int r = dpd_row(a, static_cast<const uint32_t *>(a.column)[n]);
                if (r >= 0) atomicAdd(&a.counts[static_cast<size_t>(g) * a.dim + r], 1);
            } break;
        }
    }
    if (kSmem && pooled) {
        __syncthreads();
        for (int g = threadIdx.x; g < G; g += kAddThreads) {
            if (s_ca[g]) atomicAdd(&a.cnt_a[g], s_ca[g]);
            if (s_cb[g]) atomicAdd(&a.cnt_b[g], s_cb[g]);
            if (s_x[g] != 0.0) atomicAdd(&a.sum_x[g], s_x[g]);
            if (s_xx[g] != 0.0) atomicAdd(&a.sum_xx[g], s_xx[g]);
        }
    }
}

// fold the batch accumulators into the stored statistics
__global__ void merge_stats_kernel(int model, int G, const int *__restrict__ cnt_a, const int *__restrict__ cnt_b,
                                   const double *__restrict__ sum_x, const double *__restrict__ sum_xx,
                                   uint32_t *__restrict__ st0, uint32_t *__restrict__ st1, uint32_t *__restrict__ st2) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    if (model == DIST_B200_NICH) {
        const int m = cnt_a[g];
        if (m == 0) return;
        int32_t *count = reinterpret_cast<int32_t *>(st0);
        float *mean = reinterpret_cast<float *>(st1), *ctv = reinterpret_cast<float *>(st2);
        const double mean_b = sum_x[g] / m;
        const double ctv_b = fmax(sum_xx[g] - m * mean_b * mean_b, 0.0);
        // Group::merge (nich.hpp:167-179), evaluated in double
        const double n = count[g], tot = n + m;
        const double delta = mean_b - static_cast<double>(mean[g]);
        const double source_part = static_cast<double>(m) / tot;
        const double cross_part = n * source_part;
        count[g] = static_cast<int32_t>(tot);
        mean[g] = static_cast<float>(static_cast<double>(mean[g]) + source_part * delta);
        ctv[g] = static_cast<float>(static_cast<double>(ctv[g]) + ctv_b + cross_part * delta * delta);
    } else if (model == DIST_B200_GP) {
        st0[g] += static_cast<uint32_t>(cnt_a[g]);
        st1[g] += static_cast<uint32_t>(static_cast<unsigned long long>(sum_x[g]));
    } else {  // bb
        reinterpret_cast<int32_t *>(st0)[g] += cnt_a[g];
        reinterpret_cast<int32_t *>(st1)[g] += cnt_b[g];
    }
}

__global__ void count_assignments_kernel(const int32_t *__restrict__ assign, size_t N, int G, int32_t *__restrict__ counts) {
    for (size_t n = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; n < N; n += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = assign[n];
        if (g >= 0 && g < G) atomicAdd(&counts[g], 1);
    }
}

// ---------------------------------------------------------------------------------------------
size_t add_rows_scratch_bytes(const dist_b200_feature *f) {
    // cnt_a | cnt_b | sum_x | sum_xx, each G entries, 256-byte aligned regions
    const size_t G = static_cast<size_t>(f->G);
    return 2 * ((G * 4 + 255) / 256 * 256) + 2 * ((G * 8 + 255) / 256 * 256);
}

int launch_add_rows(dist_b200_ctx *ctx, dist_b200_feature *f, const void *column, const int32_t *assign, size_t N,
                    void *scratch, size_t scratch_bytes, cudaStream_t s) {
    const int G = f->G;
    if (N == 0 || G == 0) return DIST_B200_OK;
    AddArgs a{};
    a.model = f->model;
    a.G = G;
    a.dim = f->dim;
    a.keys_dense = f->keys_dense ? 1 : 0;
    a.N = N;
    a.column = column;
    a.assign = assign;
    a.keys = f->keys_dev;
    a.key_rows = f->key_rows_dev;
    const bool pooled = f->model == DIST_B200_NICH || f->model == DIST_B200_GP || f->model == DIST_B200_BB;
    if (pooled) {
        if (scratch_bytes < add_rows_scratch_bytes(f)) return fail(ctx, DIST_B200_ERR_INVALID, "add_rows: scratch too small");
        const size_t ci = (static_cast<size_t>(G) * 4 + 255) / 256 * 256, di = (static_cast<size_t>(G) * 8 + 255) / 256 * 256;
        char *p = static_cast<char *>(scratch);
        a.cnt_a = reinterpret_cast<int *>(p);
        a.cnt_b = reinterpret_cast<int *>(p + ci);
        a.sum_x = reinterpret_cast<double *>(p + 2 * ci);
        a.sum_xx = reinterpret_cast<double *>(p + 2 * ci + di);
        DISTB200_CUDA(ctx, cudaMemsetAsync(scratch, 0, add_rows_scratch_bytes(f), s));
    } else {
        a.counts = reinterpret_cast<int32_t *>(f->stats);  // dd: array 0; dpd: counts[G][V]
    }
    const size_t want = (N + kAddThreads - 1) / kAddThreads;
    const size_t cap = static_cast<size_t>(ctx->sm_count) * 8;
    const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
    if (pooled && G <= kAddSmemGroups) {
        const size_t smem = static_cast<size_t>(G) * (8 + 8 + 4 + 4);
        add_rows_kernel<true><<<grid, kAddThreads, smem, s>>>(a);
    } else {
        add_rows_kernel<false><<<grid, kAddThreads, 0, s>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("add_rows launch: ") + cudaGetErrorString(e));
    if (pooled) {
        size_t o1 = static_cast<size_t>(f->capacity), o2 = 2 * static_cast<size_t>(f->capacity);
        merge_stats_kernel<<<(G + 127) / 128, 128, 0, s>>>(f->model, G, a.cnt_a, a.cnt_b, a.sum_x, a.sum_xx, f->stats, f->stats + o1,
                                                           f->stats + o2);
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("merge_stats launch: ") + cudaGetErrorString(e));
    }
    return DIST_B200_OK;
}

int launch_count_assignments(dist_b200_ctx *ctx, const int32_t *assign, size_t N, int G, int32_t *counts, int accumulate,
                             cudaStream_t s) {
    if (!accumulate) DISTB200_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int32_t) * G, s));
    if (N == 0) return DIST_B200_OK;
    const size_t want = (N + 255) / 256;
    const size_t cap = static_cast<size_t>(ctx->sm_count) * 8;
    count_assignments_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, s>>>(assign, N, G, counts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("count_assignments launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

}  // namespace distb200
