// microbench.cu -- register-only pipe-throughput probes used as roofline denominators for the
// SFU- and FP32-bound score kernels (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
#include "common.cuh"

namespace distb200 {

// which = 0: MUFU (alternating ex2 / lg2, 8 independent chains); which = 1: FFMA (8 chains)
__global__ void __launch_bounds__(256) pipe_probe_kernel(int which, int iters, float seed, float *sink) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + 0.001f * static_cast<float>(threadIdx.x + k);
    if (which == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = mufu_ex2(a[k]);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = mufu_lg2(a[k]);
        }
    } else {
        const float m = 1.0000001f, c = 1e-9f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456f) sink[0] = s;  // keep the chains alive
}

// which = 2: L2 gather bandwidth with the access pattern of the dpd table kernel -- every warp reads random
// 2 KB rows (4 x LDG.128 per lane) of an 8 MB table that stays L2-resident; returns BYTES per second
__global__ void __launch_bounds__(256) l2_gather_probe_kernel(const float4 *__restrict__ table, int n_rows, int iters, float *sink) {
    const int lane = threadIdx.x & 31;
    unsigned state = (blockIdx.x * 8u + (threadIdx.x >> 5)) * 2654435761u + 12345u;  // warp-uniform LCG
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < iters; ++i) {
        float4 q[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // four rows in flight
            state = state * 1664525u + 1013904223u;
            const float4 *src = table + static_cast<size_t>((state >> 8) % static_cast<unsigned>(n_rows)) * 128 + lane;
#pragma unroll
            for (int k = 0; k < 4; ++k) q[r][k] = __ldg(src + 32 * k);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc.x += q[r][k].x;
                acc.y += q[r][k].y;
                acc.z += q[r][k].z;
                acc.w += q[r][k].w;
            }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

static int run_l2_gather_probe(dist_b200_ctx *ctx, double *bytes_per_s) {
    const int n_rows = 4096, iters = 512, blocks = ctx->sm_count * 8, threads = 256;
    float *table = nullptr, *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    DISTB200_CUDA(ctx, cudaMalloc(&table, static_cast<size_t>(n_rows) * 2048));
    DISTB200_CUDA(ctx, cudaMemset(table, 0, static_cast<size_t>(n_rows) * 2048));
    DISTB200_CUDA(ctx, cudaMalloc(&sink, 4));
    DISTB200_CUDA(ctx, cudaEventCreate(&e0));
    DISTB200_CUDA(ctx, cudaEventCreate(&e1));
    int rc = DIST_B200_OK;
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, ctx->own_stream);
        l2_gather_probe_kernel<<<blocks, threads, 0, ctx->own_stream>>>(reinterpret_cast<const float4 *>(table), n_rows, iters, sink);
        cudaEventRecord(e1, ctx->own_stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            rc = fail(ctx, DIST_B200_ERR_CUDA, cudaGetErrorString(e));
            break;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = static_cast<double>(blocks) * (threads / 32) * iters * 4.0 * 2048.0;
        if (rep && ms > 0) best = std::max(best, bytes / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(table);
    if (bytes_per_s) *bytes_per_s = best;
    return rc;
}

int run_pipe_probe(dist_b200_ctx *ctx, int which, double *ops_per_s) {
    if (which == 2) return run_l2_gather_probe(ctx, ops_per_s);
    const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
    int rc = DIST_B200_OK;
    float *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    DISTB200_CUDA(ctx, cudaMalloc(&sink, 4));
    DISTB200_CUDA(ctx, cudaEventCreate(&e0));
    DISTB200_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, ctx->own_stream);
        pipe_probe_kernel<<<blocks, threads, 0, ctx->own_stream>>>(which, iters, 0.5f, sink);
        cudaEventRecord(e1, ctx->own_stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            rc = fail(ctx, DIST_B200_ERR_CUDA, cudaGetErrorString(e));
            break;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = static_cast<double>(blocks) * threads * iters * 16.0;
        if (rep && ms > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (ops_per_s) *ops_per_s = best;
    return rc;
}

}  // namespace distb200

extern "C" int dist_b200_pipe_peak(dist_b200_ctx *ctx, int which, double *ops_per_s) {
    if (!ctx || !ops_per_s || which < 0 || which > 2) return DIST_B200_ERR_INVALID;
    return distb200::run_pipe_probe(ctx, which, ops_per_s);
}
