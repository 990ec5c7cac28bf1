// niw.cu -- NormalInverseWishart<d>: batched form of looping Group::score_value over groups.
//
// The reference has no NIW Mixture; a value is scored by Scorer::eval (models/niw.hpp:343-361):
//   post  = Shared::plus_group(group)                                  (niw.hpp:82-103)
//   dof   = post.nu - d + 1;  sigma = post.psi * (post.kappa + 1) / (post.kappa * dof)
//   score = score_mv_student_t(value, dof, post.mu, sigma)             (random.hpp:160-185)
// and score_mv_student_t re-derives sigma.inverse() and sigma.determinant() on EVERY call.
//
// Here the per-group work is hoisted into niw_prep_kernel (once per batch, double precision):
// posterior, Cholesky sigma = L L^T, whitening matrix W = L^-1, determinant, Student-t constants.
// Then  diff^T sigma^-1 diff = |W (x - mu')|^2, a sum of squares (non-negative, no cancellation in the
// final sum), and
//   score[n][g] = C_g - 0.5 (dof_g + d) * fast_log(1 + |W_g (x_n - mu'_g)|^2 / dof_g).
//
// niw_score_kernel (this file, FP32 CUDA-core form): rows on lanes, two rows per thread held in
// registers, group tiles (mu', W, constants) streamed through shared memory with cp.async double
// buffering and read as broadcast LDS.128; scores leave through a padded per-warp tile so that global
// stores are coalesced.  The sampler then runs over the materialised scores (gather_rows.cu).
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace distb200 {

// per-group record: mu'[DP] | W[DP][DP] row-major lower triangular (zero above the diagonal) |
// {C_g, -0.5 (dof + d) ln 2, 1 / dof, 0}
__host__ __device__ constexpr int niw_group_floats(int dp) { return dp + dp * dp + 4; }

// one block per group; d <= 32
__global__ void niw_prep_kernel(int d, int dp, const float *__restrict__ mu, float kappa,
                                const float *__restrict__ psi, float nu, const int32_t *__restrict__ count,
                                const float *__restrict__ sum_x, const float *__restrict__ sum_xxT,
                                float *__restrict__ out, NumericTables t) {
    __shared__ double S[32][33];  // sigma, then its Cholesky factor L (lower)
    __shared__ double Winv[32][33];
    __shared__ double xbar[32], diff[32];
    const int g = blockIdx.x, tid = threadIdx.x;
    const double n = static_cast<double>(count[g]);
    const float *sx = sum_x + static_cast<size_t>(g) * d;
    const float *sxx = sum_xxT + static_cast<size_t>(g) * d * d;
    const double kap = kappa, post_kappa = kap + n, post_nu = static_cast<double>(nu) + n;
    const double dof = post_nu - static_cast<double>(d) + 1.0;
    if (tid < d) {
        xbar[tid] = count[g] ? static_cast<double>(sx[tid]) / n : 0.0;
        diff[tid] = xbar[tid] - static_cast<double>(mu[tid]);
    }
    __syncthreads();
    for (int e = tid; e < d * d; e += blockDim.x) {
        const int i = e / d, j = e % d;
        const double c_n = static_cast<double>(sxx[e]) - static_cast<double>(sx[i]) * xbar[j] -
                           xbar[i] * static_cast<double>(sx[j]) + n * xbar[i] * xbar[j];
        const double post_psi = static_cast<double>(psi[e]) + c_n + kap * n / (kap + n) * diff[i] * diff[j];
        S[i][j] = post_psi * (post_kappa + 1.0) / (post_kappa * dof);
    }
    __syncthreads();
    // Cholesky (right-looking), lower triangle in place
    for (int k = 0; k < d; ++k) {
        if (tid == 0) S[k][k] = sqrt(S[k][k]);
        __syncthreads();
        if (tid > k && tid < d) S[tid][k] /= S[k][k];
        __syncthreads();
        for (int e = tid; e < d * d; e += blockDim.x) {
            const int i = e / d, j = e % d;
            if (j > k && i >= j) S[i][j] -= S[i][k] * S[j][k];
        }
        __syncthreads();
    }
    // W = L^-1 by forward substitution, one column per thread
    if (tid < d) {
        const int c = tid;
        for (int i = 0; i < d; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s -= S[i][k] * Winv[k][c];
            Winv[i][c] = (i >= c) ? s / S[i][i] : 0.0;
        }
    }
    __syncthreads();
    float *rec = out + static_cast<size_t>(g) * niw_group_floats(dp);
    for (int i = tid; i < dp; i += blockDim.x)
        rec[i] = i < d ? static_cast<float>(kap / (kap + n) * static_cast<double>(mu[i]) + n / (kap + n) * xbar[i]) : 0.f;
    for (int e = tid; e < dp * dp; e += blockDim.x) {
        const int i = e / dp, j = e % dp;
        rec[dp + e] = (i < d && j <= i) ? static_cast<float>(Winv[i][j]) : 0.f;
    }
    if (tid == 0) {
        double det = 1.0;
        for (int i = 0; i < d; ++i) det *= S[i][i] * S[i][i];
        // score_mv_student_t's constant terms, with its float/double mix (random.hpp:166-178)
        const float dof_f = static_cast<float>(dof);
        const float log_pi = 1.1447298858494002f;
        const float term1 = fast_lgamma_exact(static_cast<float>(dof_f / 2. + static_cast<float>(d) / 2.), t.lgamma5) -
                            fast_lgamma_exact(static_cast<float>(dof_f / 2.), t.lgamma5);
        const float term2 = static_cast<float>(-0.5 * fast_log_table(static_cast<float>(det), t.log2_table) -
                                               static_cast<float>(d) / 2. * (fast_log_table(dof_f, t.log2_table) + log_pi));
        float *k4 = rec + dp + dp * dp;
        k4[0] = term1 + term2;
        k4[1] = static_cast<float>(-0.5 * (dof_f + static_cast<float>(d))) * kLn2;
        k4[2] = 1.f / dof_f;
        k4[3] = 0.f;
    }
}

constexpr int kNiwThreads = 128;
constexpr int kNiwRows = 2;  // rows per thread

__device__ __forceinline__ void niw_cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}

struct NiwArgs {
    int G, d, accumulate;
    size_t N;
    const float *recs;    // [G] group records
    const float *values;  // [N][d]
    const float *prior;   // [G] or nullptr
    float *scores;        // [N][G]
};

// DP: padded dimension (4, 8, 16, 32); GT: groups per staging tile
template <int DP, int GT>
__global__ void __launch_bounds__(kNiwThreads) niw_score_kernel(const NiwArgs a) {
    constexpr int REC = niw_group_floats(DP);
    extern __shared__ __align__(16) float smem[];
    float *stage = smem;                              // [2][GT * REC]
    float *tile = smem + 2 * GT * REC;                // [warps][rows per thread][32][33]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = a.G, d = a.d;
    const int ntile_g = (G + GT - 1) / GT;
    const size_t rows_per_block = static_cast<size_t>(kNiwThreads) * kNiwRows;
    const size_t nblk = (a.N + rows_per_block - 1) / rows_per_block;

    for (size_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const size_t base = blk * rows_per_block;
        float x[kNiwRows][DP];
#pragma unroll
        for (int r = 0; r < kNiwRows; ++r) {
            size_t row = base + static_cast<size_t>(r) * kNiwThreads + tid;
            if (row >= a.N) row = a.N - 1;
            const float *src = a.values + row * d;
#pragma unroll
            for (int k = 0; k < DP; ++k) x[r][k] = k < d ? src[k] : 0.f;
        }
        __syncthreads();  // staging buffers free
        {   // stage group tile 0
            const float *src = a.recs;
            const int n = min(GT, G) * REC;
            for (int i = tid * 4; i < n; i += kNiwThreads * 4) niw_cp_async16(stage + i, src + i);
            asm volatile("cp.async.commit_group;\n" ::);
        }
        for (int gt = 0; gt < ntile_g; ++gt) {
            if (gt + 1 < ntile_g) {
                const float *src = a.recs + static_cast<size_t>(gt + 1) * GT * REC;
                const int n = min(GT, G - (gt + 1) * GT) * REC;
                float *dst = stage + ((gt + 1) & 1) * GT * REC;
                for (int i = tid * 4; i < n; i += kNiwThreads * 4) niw_cp_async16(dst + i, src + i);
                asm volatile("cp.async.commit_group;\n" ::);
                asm volatile("cp.async.wait_group 1;\n" ::);
            } else {
                asm volatile("cp.async.wait_group 0;\n" ::);
            }
            __syncthreads();
            const float *buf = stage + (gt & 1) * GT * REC;
            const int ng = min(GT, G - gt * GT);
            for (int j = 0; j < ng; ++j) {
                const float *rec = buf + j * REC;
                float z[kNiwRows][DP];
#pragma unroll
                for (int k = 0; k < DP; k += 4) {
                    const float4 m = *reinterpret_cast<const float4 *>(rec + k);
#pragma unroll
                    for (int r = 0; r < kNiwRows; ++r) {
                        z[r][k] = x[r][k] - m.x;
                        z[r][k + 1] = x[r][k + 1] - m.y;
                        z[r][k + 2] = x[r][k + 2] - m.z;
                        z[r][k + 3] = x[r][k + 3] - m.w;
                    }
                }
                float q[kNiwRows];
#pragma unroll
                for (int r = 0; r < kNiwRows; ++r) q[r] = 0.f;
                const float *W = rec + DP;
#pragma unroll
                for (int i = 0; i < DP; ++i) {
                    float y[kNiwRows];
#pragma unroll
                    for (int r = 0; r < kNiwRows; ++r) y[r] = 0.f;
#pragma unroll
                    for (int k = 0; k <= i; k += 4) {
                        const float4 w = *reinterpret_cast<const float4 *>(W + i * DP + k);
#pragma unroll
                        for (int r = 0; r < kNiwRows; ++r) {
                            y[r] = fmaf(w.x, z[r][k], y[r]);
                            if (k + 1 <= i) y[r] = fmaf(w.y, z[r][k + 1], y[r]);
                            if (k + 2 <= i) y[r] = fmaf(w.z, z[r][k + 2], y[r]);
                            if (k + 3 <= i) y[r] = fmaf(w.w, z[r][k + 3], y[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < kNiwRows; ++r) q[r] = fmaf(y[r], y[r], q[r]);
                }
                const float4 c = *reinterpret_cast<const float4 *>(rec + DP + DP * DP);
                const int g = gt * GT + j;
                const float pr = (a.prior && !a.accumulate) ? a.prior[g] : 0.f;
#pragma unroll
                for (int r = 0; r < kNiwRows; ++r) {
                    // C_g - 0.5 (dof + d) * fast_log(1 + q / dof)
                    const float arg = __fadd_rn(1.f, __fmul_rn(c.z, q[r]));
                    const float s = fmaf(c.y, fast_log2_cell(arg), c.x) + pr;
                    // park in the per-warp tile: column (g & 31), flushed every 32 groups
                    float *tw = tile + (warp * kNiwRows + r) * 32 * 33;
                    tw[lane * 33 + (g & 31)] = s;
                    if (((g & 31) == 31 || g == G - 1)) {
                        __syncwarp();
                        const int g0 = g & ~31;
                        const int gl = g0 + lane;
                        const size_t wrow0 = base + static_cast<size_t>(r) * kNiwThreads + warp * 32;
                        if (gl <= g) {
                            for (int i = 0; i < 32; ++i) {
                                const size_t rr = wrow0 + i;
                                if (rr >= a.N) break;
                                float *dst = a.scores + rr * G + gl;
                                const float v = tw[i * 33 + lane];
                                *dst = a.accumulate ? *dst + v : v;
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            __syncthreads();  // buffer (gt & 1) is refilled two tiles from now
        }
    }
}

template <int DP>
static int launch_niw_dp(dist_b200_ctx *ctx, const NiwArgs &a, cudaStream_t s) {
    constexpr int GT = DP >= 32 ? 8 : (DP >= 16 ? 16 : 32);
    const size_t smem = sizeof(float) * (2 * GT * niw_group_floats(DP) + (kNiwThreads / 32) * kNiwRows * 32 * 33);
    auto kern = niw_score_kernel<DP, GT>;
    DISTB200_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    DISTB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kNiwThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const size_t rows_per_block = static_cast<size_t>(kNiwThreads) * kNiwRows;
    const size_t nblk = (a.N + rows_per_block - 1) / rows_per_block;
    const size_t cap = static_cast<size_t>(ctx->sm_count) * per_sm;
    kern<<<static_cast<unsigned>(nblk < cap ? nblk : cap), kNiwThreads, smem, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("niw launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

int niw_padded_dim(int d) { return d <= 4 ? 4 : (d <= 8 ? 8 : (d <= 16 ? 16 : 32)); }

int launch_niw_prep(dist_b200_ctx *ctx, int d, const float *mu, float kappa, const float *psi, float nu, int G,
                    const int32_t *count, const float *sum_x, const float *sum_xxT, float *recs, cudaStream_t s) {
    if (G <= 0) return DIST_B200_OK;
    niw_prep_kernel<<<G, 64, 0, s>>>(d, niw_padded_dim(d), mu, kappa, psi, nu, count, sum_x, sum_xxT, recs, ctx->tables);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("niw prep launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

int launch_niw_scores(dist_b200_ctx *ctx, const dist_b200_feature *f, const void *values, size_t N,
                      const float *prior, float *scores, int accumulate, cudaStream_t s) {
    if (N == 0 || f->G == 0) return DIST_B200_OK;
    // d = 32: tcgen05 with split fp16 operands; DIST_B200_OPT_NIW_PATH = 1 keeps the FP32 CUDA-core kernel for A/B runs
    if (f->dim == 32 && f->niw_tc && ctx->opt[DIST_B200_OPT_NIW_PATH] == 0) {
        const int rc = launch_niw_tc(ctx, f->G, f->niw_tc, values, N, prior, scores, accumulate, nullptr, nullptr, s);
        if (rc != DIST_B200_ERR_UNSUPPORTED) return rc;
    }
    NiwArgs a{};
    a.G = f->G;
    a.d = f->dim;
    a.accumulate = accumulate;
    a.N = N;
    a.recs = f->niw_buf;
    a.values = static_cast<const float *>(values);
    a.prior = prior;
    a.scores = scores;
    switch (niw_padded_dim(f->dim)) {
        case 4: return launch_niw_dp<4>(ctx, a, s);
        case 8: return launch_niw_dp<8>(ctx, a, s);
        case 16: return launch_niw_dp<16>(ctx, a, s);
        default: return launch_niw_dp<32>(ctx, a, s);
    }
}

}  // namespace distb200
