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

int run_pipe_probe(dist_b200_ctx *ctx, int which, double *ops_per_s) {
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
    if (!ctx || !ops_per_s || which < 0 || which > 1) return DIST_B200_ERR_INVALID;
    return distb200::run_pipe_probe(ctx, which, ops_per_s);
}
