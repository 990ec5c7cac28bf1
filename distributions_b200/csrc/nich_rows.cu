// nich_rows.cu -- NormalInverseChiSq, one feature, G > 128, sampling only: the c2 shape (1M rows x 1024 groups).
//
// Same algorithm as score_rows_kernel<32, nich packed> (score_rows.cuh: rows on lanes, 32-group register tiles,
// per-tile (max, sum exp) pairs merged into 16 slots, walk over the slots, re-score of the selected slot), rebuilt
// around what its ncu capture showed (profiles/r02_c2_nich_packed.txt): the MUFU pipe it is bound by was 77 % busy
// while the stall reasons were MIO throttle + short scoreboard, because the shared-memory pipe was almost as loaded
// as the MUFU pipe -- a warp-wide LDS.128 occupies it for four wavefronts even when all lanes read the same address,
// i.e. 4 cycles per (warp, cell) against MUFU's 16 cycles per (warp, cell) / 4 sub-partitions, and the re-score pass
// read with up to 32-way bank conflicts (every lane its own slot, all slots at the same bank offset).  Here
//   * every thread carries TWO rows, so a group's parameters are loaded once per two cells (half the wavefronts);
//   * the block's parameter copy is skewed by 16 bytes per slot, so the lanes of the re-score pass, which sit in
//     different slots, mostly hit different banks.
// Cell arithmetic is the packed fp32x2 form of nich.cc:59-65 (see accumulate_feature, kKindNichPacked): two groups per
// FADD2 / FMUL2 / FFMA2, every product and sum rounded as the reference's unfused expression.
#include "score_rows.cuh"

namespace distb200 {

constexpr int kNrThreads = 256;
constexpr int kNrChunk = 32;
constexpr int kNrRows = 2;  // rows per thread

struct NichRowsArgs {
    int G;
    size_t N;
    const float4 *params;  // {mean, precision, log_coeff * ln 2, score} per group (capacity padded to 128 groups)
    const float *prior;    // [G] or nullptr
    const float *values;
    const float *u;
    int32_t *assign;
};

__global__ void __launch_bounds__(kNrThreads, 2) nich_rows_kernel(const NichRowsArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int G = a.G;
    const int nchunks = (G + kNrChunk - 1) / kNrChunk;
    const int Gpad = nchunks * kNrChunk;
    const int cps = (nchunks + kSlots - 1) / kSlots;           // chunks per slot
    const int nslots = (nchunks + cps - 1) / cps;
    const int slot_groups = cps * kNrChunk;
    // caches: pair p of groups at float4 index 2 p + (2 p) / slot_groups (one float4 of skew per slot)
    float4 *caches = reinterpret_cast<float4 *>(smem);
    float2 *slots = reinterpret_cast<float2 *>(caches + Gpad + nslots);
    const int tid = threadIdx.x;

    // block-private parameter copy: prior folded into the score, padded groups -inf (they borrow group 0's mean /
    // precision / coefficient so the term stays finite-or-(-inf) exactly as for a real group), pairs interleaved
    for (int p = tid; p < Gpad / 2; p += kNrThreads) {
        float4 q[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int g = 2 * p + k;
            q[k] = a.params[g < G ? g : 0];
            q[k].w = g < G ? q[k].w + (a.prior ? a.prior[g] : 0.f) : -INFINITY;
        }
        const int at = 2 * p + (2 * p) / slot_groups;
        caches[at] = make_float4(-q[0].x, -q[1].x, q[0].y, q[1].y);
        caches[at + 1] = make_float4(q[0].z, q[1].z, q[0].w, q[1].w);
    }
    __syncthreads();

    const uint64_t one2 = f2_pack(1.f, 1.f), l2e2 = f2_pack(kLog2e, kLog2e);
    const size_t tile_rows = static_cast<size_t>(kNrThreads) * kNrRows;
    const size_t ntiles = (a.N + tile_rows - 1) / tile_rows;
    // the next tile's rows are requested while the current tile is evaluated (2 048 cells per row: the load latency --
    // HBM, or PCIe when the host entry hands the kernel page-locked host buffers -- disappears behind them)
    float xnext[kNrRows], unext[kNrRows];
    auto fetch = [&](size_t tile) {
#pragma unroll
        for (int r = 0; r < kNrRows; ++r) {
            const size_t rw = tile * tile_rows + static_cast<size_t>(r) * kNrThreads + tid;
            const size_t rr = rw < a.N ? rw : a.N - 1;  // clamp: compute on a real row, discard the result
            xnext[r] = __ldg(a.values + rr);
            unext[r] = __ldg(a.u + rr);
        }
    };
    if (blockIdx.x < ntiles) fetch(blockIdx.x);
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        size_t row[kNrRows];
        uint64_t x2[kNrRows];
        float urow[kNrRows];
#pragma unroll
        for (int r = 0; r < kNrRows; ++r) {
            row[r] = tile * tile_rows + static_cast<size_t>(r) * kNrThreads + tid;
            x2[r] = f2_pack(xnext[r], xnext[r]);
            urow[r] = unext[r];
        }
        if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);
        float slot_m[kNrRows], slot_s[kNrRows];
#pragma unroll
        for (int r = 0; r < kNrRows; ++r) {
            slot_m[r] = INFINITY;  // negated scaled max of the slot being merged
            slot_s[r] = 0.f;
        }
        // one 32-group tile for both rows: acc[r][j] = prior + score + log_coeff * fast_log(1 + precision * (x - mean)^2)
        auto score_tile = [&](const float4 *p4, int r, float (&acc)[kNrChunk]) {
#pragma unroll
            for (int j = 0; j < kNrChunk; j += 2) {
                const float4 qa = p4[j], qb = p4[j + 1];
                const uint64_t d2 = f2_add(x2[r], f2_pack(qa.x, qa.y));
                const uint64_t z2 = f2_fma(f2_mul(f2_pack(qa.z, qa.w), f2_mul(d2, d2)), one2, one2);  // unfused 1 + w (see score_rows.cuh)
                float za, zb;
                f2_unpack(z2, za, zb);
                f2_unpack(f2_fma(f2_pack(qb.x, qb.y), f2_pack(fast_log2_cell(za), fast_log2_cell(zb)), f2_pack(qb.z, qb.w)), acc[j], acc[j + 1]);
            }
        };
        for (int c = 0; c < nchunks; ++c) {
            const float4 *p4 = caches + c * kNrChunk + c / cps;
            float acc[kNrRows][kNrChunk];
#pragma unroll
            for (int j = 0; j < kNrChunk; j += 2) {  // parameters once, both rows
                const float4 qa = p4[j], qb = p4[j + 1];
                const uint64_t nm2 = f2_pack(qa.x, qa.y), pr2 = f2_pack(qa.z, qa.w), co2 = f2_pack(qb.x, qb.y), sc2 = f2_pack(qb.z, qb.w);
#pragma unroll
                for (int r = 0; r < kNrRows; ++r) {
                    const uint64_t d2 = f2_add(x2[r], nm2);
                    const uint64_t z2 = f2_fma(f2_mul(pr2, f2_mul(d2, d2)), one2, one2);
                    float za, zb;
                    f2_unpack(z2, za, zb);
                    f2_unpack(f2_fma(co2, f2_pack(fast_log2_cell(za), fast_log2_cell(zb)), sc2), acc[r][j], acc[r][j + 1]);
                }
            }
#pragma unroll
            for (int r = 0; r < kNrRows; ++r) {
                float m = acc[r][0];
#pragma unroll
                for (int j = 1; j < kNrChunk; ++j) m = fmaxf(m, acc[r][j]);
                const float nm = -m * kLog2e;  // rounded once: every later rescale is a difference of these values
                const uint64_t nm2 = f2_pack(nm, nm);
                uint64_t s2 = f2_pack(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < kNrChunk; j += 2) {
                    float ea, eb;
                    f2_unpack(f2_fma(f2_pack(acc[r][j], acc[r][j + 1]), l2e2, nm2), ea, eb);
                    s2 = f2_add(s2, f2_pack(mufu_ex2(ea), mufu_ex2(eb)));
                }
                float se, so;
                f2_unpack(s2, se, so);
                const float s = se + so;
                const float dlt = slot_m[r] - nm;
                const float e = mufu_ex2(-fabsf(dlt));
                slot_s[r] = dlt > 0.f ? fmaf(slot_s[r], e, s) : fmaf(s, e, slot_s[r]);
                slot_m[r] = fminf(slot_m[r], nm);
                if ((c + 1) % cps == 0 || c + 1 == nchunks) {
                    slots[(r * kSlots + c / cps) * kNrThreads + tid] = make_float2(slot_m[r], slot_s[r]);
                    slot_m[r] = INFINITY;
                    slot_s[r] = 0.f;
                }
            }
        }
        // per row: total over the slots, walk to the slot holding u * total, re-score that slot, walk its cells
#pragma unroll
        for (int r = 0; r < kNrRows; ++r) {
            float2 *sl = slots + (r * kSlots) * kNrThreads + tid;
            float mm = INFINITY;
            for (int k = 0; k < nslots; ++k) mm = fminf(mm, sl[k * kNrThreads].x);
            float total = 0.f;
            for (int k = 0; k < nslots; ++k) {
                const float2 ms = sl[k * kNrThreads];
                const float w = ms.y * mufu_ex2(mm - ms.x);
                sl[k * kNrThreads].y = w;
                total += w;
            }
            float t = total * urow[r];
            int sel = nslots - 1;
            for (int k = 0; k < nslots; ++k) {
                const float w = sl[k * kNrThreads].y;
                if (t <= w) {
                    sel = k;
                    break;
                }
                if (k + 1 < nslots) t -= w;
            }
            int count = 0;
            for (int cc = 0; cc < cps; ++cc) {
                const int c = sel * cps + cc;
                if (c >= nchunks) break;
                float acc[kNrChunk];
                score_tile(caches + c * kNrChunk + sel, r, acc);
#pragma unroll
                for (int j = 0; j < kNrChunk; ++j) {
                    t -= mufu_ex2(fmaf(acc[j], kLog2e, mm));
                    count += 1 - static_cast<int>(__float_as_uint(t) >> 31);  // t >= +0 continues (an exact 0: a near-tie)
                }
            }
            if (row[r] < a.N) a.assign[row[r]] = min(sel * slot_groups + count, G - 1);
        }
    }
}

// DIST_B200_ERR_UNSUPPORTED: the caller takes the generic kernel
int launch_nich_rows(dist_b200_ctx *ctx, const float4 *params, const void *column, int G, size_t N, const float *prior,
                     const float *u, int32_t *assign, cudaStream_t s) {
    if (G <= 128 || !assign) return DIST_B200_ERR_UNSUPPORTED;
    const int nchunks = (G + kNrChunk - 1) / kNrChunk;
    const int Gpad = nchunks * kNrChunk;
    const int cps = (nchunks + kSlots - 1) / kSlots;
    const int nslots = (nchunks + cps - 1) / cps;
    const size_t smem = sizeof(float4) * (Gpad + nslots) + sizeof(float2) * kNrRows * kSlots * kNrThreads;
    if (smem > 110 * 1024) return DIST_B200_ERR_UNSUPPORTED;  // two blocks per SM
    NichRowsArgs a{};
    a.G = G;
    a.N = N;
    a.params = params;
    a.prior = prior;
    a.values = static_cast<const float *>(column);
    a.u = u;
    a.assign = assign;
    DISTB200_CUDA(ctx, cudaFuncSetAttribute(nich_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    DISTB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nich_rows_kernel, kNrThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const size_t tile_rows = static_cast<size_t>(kNrThreads) * kNrRows;
    const size_t ntiles = (N + tile_rows - 1) / tile_rows;
    const size_t cap = static_cast<size_t>(ctx->sm_count) * per_sm;
    nich_rows_kernel<<<static_cast<unsigned>(ntiles < cap ? ntiles : cap), kNrThreads, smem, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("nich_rows launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

}  // namespace distb200
