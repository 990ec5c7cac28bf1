// score_rows.cu -- the row-mapped fused score(+prior)(+sample) kernel.
//
// Replaces, for N rows at once, the per-value stack of SURVEY.md §3.2:
//   PitmanYor::Mixture::score_value (clustering.hpp:195-208, overwrite with the prior vector)
//   -> Model::Mixture::score_value for every feature (accumulate; src/models/nich.cc:33-66,
//      src/models/gp.cc:32-67, bb.hpp:303-313, dd.hpp:433-445)
//   -> sample_from_scores_overwrite (random.hpp:360-366 = random.cc:94-106 + random.hpp:315-333).
//
// Mapping (B200): one ROW per lane, groups walked in register tiles of CHUNK.  The per-group caches
// of the current (feature, chunk) sit in shared memory and every lane of a warp reads the same
// group at the same time, so a cache entry is one conflict-free broadcast LDS.128 shared by 32
// rows; row values are read with coalesced loads from feature-major columns.  Everything the
// reference does in three passes over a G-float buffer (score, max/exp/sum, scan) happens in
// registers, in the reference's own left-to-right order within a chunk, so no warp shuffles and no
// [N][G] round trip through HBM are needed unless the caller asks for the scores.
//
//   G <= CHUNK  : the whole score row lives in registers; max, exp, running total and the
//                 `t -= l[i]; t <= 0` walk are the reference's loops verbatim.
//   G  > CHUNK  : per chunk (max, sum of exp) pairs are merged into at most kSlots slots kept in
//                 shared memory; the slot holding the draw is found by a walk over slots, then the
//                 warp re-scores just that slot's groups for each of its rows with groups mapped to
//                 lanes (coalesced cache reads, warp prefix scan) to find the index.
//
// Group caches are staged with cp.async: resident for the whole kernel when all features fit in
// shared memory, otherwise double-buffered per (feature, chunk) behind the compute of the previous
// feature.
#include "common.cuh"

namespace distb200 {

constexpr int kThreads = 256;
constexpr int kSlots = 16;
constexpr size_t kResidentBudget = 96 * 1024;  // bytes of group caches kept resident in smem

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ int kind_stride(const FeatDesc &fd) {  // floats of cache per group
    return fd.kind == DIST_B200_DD ? fd.vdim : 4;
}

// raw 32-bit value of row `row` of a feature column
__device__ __forceinline__ uint32_t load_value(const FeatDesc &fd, size_t row) {
    switch (fd.kind) {
        case DIST_B200_BB: return static_cast<const uint8_t *>(fd.column)[row];
        default: return static_cast<const uint32_t *>(fd.column)[row];
    }
}

// one cell, caches read through a generic pointer (phase 2 / tails): identical arithmetic to the
// tiled loop below so that both phases see the same score bits
__device__ __forceinline__ float cell_score(int kind, uint32_t xb, const float *__restrict__ p, int vdim,
                                            const float *__restrict__ coeff, const float *__restrict__ logfact) {
    switch (kind) {
        case DIST_B200_NICH: {
            const float4 q = *reinterpret_cast<const float4 *>(p);
            const float d = __uint_as_float(xb) - q.x;
            const float z = __fadd_rn(1.f, __fmul_rn(q.y, __fmul_rn(d, d)));
            return fmaf(q.z, fast_log2_cell(z), q.w);  // q.z = log_coeff * ln 2
        }
        case DIST_B200_GP: {
            const float4 q = *reinterpret_cast<const float4 *>(p);
            const float xf = static_cast<float>(xb);
            const float lf = xb < 64 ? logfact[xb] : fast_lgamma_cell(static_cast<float>(xb + 1u), coeff);
            const float lg = fast_lgamma_cell(q.x + xf, coeff);
            return fmaf(q.y, xf, (q.z + lg) - lf);
        }
        case DIST_B200_BB: {
            const float2 q = *reinterpret_cast<const float2 *>(p);
            return xb ? q.x : q.y;
        }
        default: {  // DD
            const int v = min(static_cast<int>(xb), vdim - 1);
            return p[v];
        }
    }
}

template <int CHUNK, int R>
__device__ __forceinline__ void accumulate_feature(int kind, const uint32_t (&xb)[R], const float *__restrict__ pb,
                                                   int vdim, float (&acc)[R][CHUNK],
                                                   const float *__restrict__ coeff,
                                                   const float *__restrict__ logfact) {
    switch (kind) {
        case DIST_B200_NICH: {
            // acc += score + log_coeff * fast_log(1 + precision * (v - mean)^2)   (nich.cc:59-65)
            const float4 *p4 = reinterpret_cast<const float4 *>(pb);
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const float4 q = p4[j];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float d = __uint_as_float(xb[r]) - q.x;
                    const float z = __fadd_rn(1.f, __fmul_rn(q.y, __fmul_rn(d, d)));
                    acc[r][j] += fmaf(q.z, fast_log2_cell(z), q.w);  // q.z = log_coeff * ln 2
                }
            }
        } break;
        case DIST_B200_GP: {
            // acc += score + fast_lgamma(post_alpha + v) - fast_log_factorial(v) + score_coeff * v  (gp.cc:56-66)
            const float4 *p4 = reinterpret_cast<const float4 *>(pb);
            float xf[R], lf[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                xf[r] = static_cast<float>(xb[r]);
                lf[r] = xb[r] < 64 ? logfact[xb[r]] : fast_lgamma_cell(static_cast<float>(xb[r] + 1u), coeff);
            }
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const float4 q = p4[j];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float lg = fast_lgamma_cell(q.x + xf[r], coeff);
                    acc[r][j] += fmaf(q.y, xf[r], (q.z + lg) - lf[r]);
                }
            }
        } break;
        case DIST_B200_BB: {
            // acc += value ? heads[g] : tails[g]   (bb.hpp:303-313)
            const float4 *p4 = reinterpret_cast<const float4 *>(pb);
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const float2 q = *reinterpret_cast<const float2 *>(p4 + j);
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r][j] += xb[r] ? q.x : q.y;
            }
        } break;
        default: {  // DD: acc += scores_[value][g] - scores_shift_[g], pre-subtracted table (dd.hpp:433-445)
            int v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = min(static_cast<int>(xb[r]), vdim - 1);
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r][j] += pb[j * vdim + v[r]];
            }
        } break;
    }
}

struct RowsArgs {
    int G;
    int resident;        // all caches resident in smem
    int stage_floats;    // floats per staging buffer (streaming mode)
    int accumulate;
    size_t N;
    const float *prior;
    const float *u;
    int32_t *assign;
    float *scores;
    NumericTables t;
};

template <int CHUNK, int R, bool kSample, bool kScores>
__global__ void __launch_bounds__(kThreads)
score_rows_kernel(const __grid_constant__ FeatList feats, const RowsArgs a) {
    extern __shared__ __align__(16) float smem[];
    // layout: coeff[33*8] | logfact[64] | prior[Gpad] | tile[8 warps][32][33] (kScores) |
    //         slots[kSlots][R][kThreads] float2 (kSample, multi-chunk) | caches
    const int G = a.G;
    const int nchunks = (G + CHUNK - 1) / CHUNK;
    const bool multi = nchunks > 1;
    float *coeff = smem;
    float *logfact = coeff + 33 * kLgammaRowStride;
    float *prior_s = logfact + 64;
    float *cursor = prior_s + nchunks * CHUNK;
    float *tile = cursor;
    if (kScores) cursor += (kThreads / 32) * 32 * 33;
    float2 *slots = reinterpret_cast<float2 *>(cursor);
    if (kSample && multi) cursor += 2 * kSlots * R * kThreads;
    float *caches = cursor;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = feats.n;
    const int Gpad = nchunks * CHUNK;

    for (int i = tid; i < 33 * kLgammaRowStride; i += kThreads) coeff[i] = a.t.lgamma5[i];
    if (tid < 64) logfact[tid] = a.t.log_factorial[tid];
    // the prior vector (clustering's overwrite) seeds every accumulator; padded groups get -inf
    for (int g = tid; g < Gpad; g += kThreads)
        prior_s[g] = g < G ? ((a.prior && !a.accumulate) ? a.prior[g] : 0.f) : -INFINITY;
    if (a.resident) {
        size_t off = 0;
        for (int f = 0; f < F; ++f) {
            const int n = Gpad * kind_stride(feats.f[f]);
            const float *src = static_cast<const float *>(feats.f[f].params);
            for (int i = tid * 4; i < n; i += kThreads * 4) cp_async16(caches + off + i, src + i);
            off += n;
        }
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();

    const int chunks_per_slot = (nchunks + kSlots - 1) / kSlots;
    const size_t rows_per_tile = static_cast<size_t>(kThreads) * R;
    const size_t ntiles = (a.N + rows_per_tile - 1) / rows_per_tile;

    for (size_t tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const size_t tile_base = tile_id * rows_per_tile;
        size_t row[R];
        bool valid[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            row[r] = tile_base + static_cast<size_t>(r) * kThreads + tid;
            valid[r] = row[r] < a.N;
            if (!valid[r]) row[r] = a.N - 1;  // clamp: compute on a real row, discard the result
        }
        float slot_m[R], slot_s[R];  // slot being merged (multi-chunk sampling)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            slot_m[r] = INFINITY;  // negated scaled maximum of the slot so far
            slot_s[r] = 0.f;
        }
        int result[R];

        for (int c = 0; c < nchunks; ++c) {
            const int g0 = c * CHUNK;
            if (!a.resident) {  // stage feature 0 of this chunk (buffer 0 was released by the last sync)
                const FeatDesc &fd = feats.f[0];
                const int st = kind_stride(fd);
                const float *src = static_cast<const float *>(fd.params) + static_cast<size_t>(g0) * st;
                for (int i = tid * 4; i < CHUNK * st; i += kThreads * 4) cp_async16(caches + i, src + i);
                cp_async_commit();
            }

            float acc[R][CHUNK];
#pragma unroll
            for (int j = 0; j < CHUNK; j += 4) {
                const float4 p = *reinterpret_cast<const float4 *>(prior_s + g0 + j);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r][j] = p.x;
                    acc[r][j + 1] = p.y;
                    acc[r][j + 2] = p.z;
                    acc[r][j + 3] = p.w;
                }
            }

            uint32_t xb[R];
#pragma unroll
            for (int r = 0; r < R; ++r) xb[r] = load_value(feats.f[0], row[r]);
            size_t res_off = 0;
            for (int f = 0; f < F; ++f) {
                const FeatDesc &fd = feats.f[f];
                const int st = kind_stride(fd);
                uint32_t xn[R];
                if (f + 1 < F) {
#pragma unroll
                    for (int r = 0; r < R; ++r) xn[r] = load_value(feats.f[f + 1], row[r]);
                }
                const float *pb;
                if (a.resident) {
                    pb = caches + res_off + static_cast<size_t>(g0) * st;
                    res_off += static_cast<size_t>(Gpad) * st;
                } else {
                    if (f + 1 < F) {  // prefetch the next feature's caches behind this feature's math
                        const FeatDesc &fn = feats.f[f + 1];
                        const int sn = kind_stride(fn);
                        const float *src = static_cast<const float *>(fn.params) + static_cast<size_t>(g0) * sn;
                        float *dst = caches + ((f + 1) & 1) * a.stage_floats;
                        for (int i = tid * 4; i < CHUNK * sn; i += kThreads * 4) cp_async16(dst + i, src + i);
                        cp_async_commit();
                        cp_async_wait<1>();
                    } else {
                        cp_async_wait<0>();
                    }
                    __syncthreads();
                    pb = caches + (f & 1) * a.stage_floats;
                }
                accumulate_feature<CHUNK, R>(fd.kind, xb, pb, fd.vdim, acc, coeff, logfact);
                if (!a.resident) __syncthreads();  // buffer (f&1) is rewritten two features from now
                if (f + 1 < F) {
#pragma unroll
                    for (int r = 0; r < R; ++r) xb[r] = xn[r];
                }
            }

            if (g0 + CHUNK > G) {
                // ragged last tile: padded groups carry zeroed caches, whose model terms may be inf/NaN
                // (lgamma(0)); pin them to -inf so they vanish from max / exp / the walk
#pragma unroll
                for (int j = 0; j < CHUNK; ++j) {
                    if (g0 + j >= G) {
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r][j] = -INFINITY;
                    }
                }
            }

            if (kScores) {
                // [32 rows][32 groups] transposes through a padded per-warp tile -> coalesced rows
                float *tw = tile + warp * 32 * 33;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const size_t wrow0 = tile_base + static_cast<size_t>(r) * kThreads + warp * 32;
#pragma unroll
                    for (int sb = 0; sb < CHUNK / 32; ++sb) {
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 32; ++j) tw[lane * 33 + j] = acc[r][sb * 32 + j];
                        __syncwarp();
                        const int g = g0 + sb * 32 + lane;
                        if (g < G) {
                            for (int i = 0; i < 32; ++i) {
                                const size_t rr = wrow0 + i;
                                if (rr >= a.N) break;
                                float *dst = a.scores + rr * G + g;
                                const float v = tw[i * 33 + lane];
                                *dst = a.accumulate ? *dst + v : v;
                            }
                        }
                    }
                }
            }

            if (kSample) {
                if (!multi) {
                    // scores_to_likelihoods + sample_from_likelihoods, the reference's loops verbatim
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float m = acc[r][0];
#pragma unroll
                        for (int j = 1; j < CHUNK; ++j) m = fmaxf(m, acc[r][j]);
                        const float nm = -m * kLog2e;  // exp(s - m) = 2^(s*log2e + nm): one FFMA + MUFU.EX2
                        float total = 0.f;
#pragma unroll
                        for (int j = 0; j < CHUNK; ++j) {
                            acc[r][j] = mufu_ex2(fmaf(acc[r][j], kLog2e, nm));
                            total += acc[r][j];
                        }
                        float t = total * a.u[row[r]];
                        int idx = 0;
#pragma unroll
                        for (int j = 0; j < CHUNK; ++j) {
                            t -= acc[r][j];
                            idx += (t > 0.f) ? 1 : 0;
                        }
                        result[r] = min(idx, G - 1);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float m = acc[r][0];
#pragma unroll
                        for (int j = 1; j < CHUNK; ++j) m = fmaxf(m, acc[r][j]);
                        // slots carry nm = -(max * log2 e), rounded ONCE: every later rescale is a
                        // difference of these rounded values, so chunk sums stay mutually consistent
                        const float nm = -m * kLog2e;
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < CHUNK; ++j) s += mufu_ex2(fmaf(acc[r][j], kLog2e, nm));
                        // merge into the running slot
                        const float nn = fminf(slot_m[r], nm);
                        slot_s[r] = slot_s[r] * mufu_ex2(nn - slot_m[r]) + s * mufu_ex2(nn - nm);
                        slot_m[r] = nn;
                        if ((c + 1) % chunks_per_slot == 0 || c + 1 == nchunks) {
                            slots[((c / chunks_per_slot) * R + r) * kThreads + tid] = make_float2(slot_m[r], slot_s[r]);
                            slot_m[r] = INFINITY;
                            slot_s[r] = 0.f;
                        }
                    }
                }
            }
        }  // chunks

        if (kSample && multi) {
            const int nslots = (nchunks + chunks_per_slot - 1) / chunks_per_slot;
            float M[R], tres[R];
            int slot_sel[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float mm = INFINITY;  // = -(row maximum) * log2 e
                for (int k = 0; k < nslots; ++k) mm = fminf(mm, slots[(k * R + r) * kThreads + tid].x);
                float total = 0.f;
                for (int k = 0; k < nslots; ++k) {
                    const float2 ms = slots[(k * R + r) * kThreads + tid];
                    total += ms.y * mufu_ex2(mm - ms.x);
                }
                float t = total * a.u[row[r]];
                int sel = nslots - 1;
                for (int k = 0; k < nslots; ++k) {
                    const float2 ms = slots[(k * R + r) * kThreads + tid];
                    const float w = ms.y * mufu_ex2(mm - ms.x);
                    if (t <= w) {
                        sel = k;
                        break;
                    }
                    if (k + 1 < nslots) t -= w;
                }
                M[r] = mm;
                tres[r] = t;
                slot_sel[r] = sel;
            }
            // phase 2: the warp re-scores the selected slot of each of its rows, groups on lanes
#pragma unroll
            for (int r = 0; r < R; ++r) {
                result[r] = G - 1;
                for (int i = 0; i < 32; ++i) {
                    const size_t rr = tile_base + static_cast<size_t>(r) * kThreads + warp * 32 + i;
                    if (rr >= a.N) break;  // warp-uniform
                    const float Mi = __shfl_sync(0xffffffffu, M[r], i);
                    float t = __shfl_sync(0xffffffffu, tres[r], i);
                    const int sel = __shfl_sync(0xffffffffu, slot_sel[r], i);
                    int idx = G - 1;
                    for (int gb = sel * chunks_per_slot * CHUNK; gb < G; gb += 32) {
                        const int g = gb + lane;
                        float l = 0.f;
                        if (g < G) {
                            float s = (a.prior && !a.accumulate) ? a.prior[g] : 0.f;
                            for (int f = 0; f < F; ++f) {
                                const FeatDesc &fd = feats.f[f];
                                const int st = kind_stride(fd);
                                s += cell_score(fd.kind, load_value(fd, rr),
                                                static_cast<const float *>(fd.params) + static_cast<size_t>(g) * st,
                                                fd.vdim, coeff, logfact);
                            }
                            l = mufu_ex2(fmaf(s, kLog2e, Mi));
                        }
                        float scan = l;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const float n = __shfl_up_sync(0xffffffffu, scan, o);
                            if (lane >= o) scan += n;
                        }
                        const float tot = __shfl_sync(0xffffffffu, scan, 31);
                        const unsigned hit = __ballot_sync(0xffffffffu, g < G && scan >= t);
                        if (hit) {
                            idx = gb + __ffs(hit) - 1;
                            break;
                        }
                        t -= tot;
                    }
                    if (lane == i) result[r] = idx;
                }
            }
        }
        if (kSample) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (valid[r]) a.assign[row[r]] = result[r];
        }
    }  // row tiles
}

// ---------------------------------------------------------------------------------------------
template <int CHUNK, int R, bool kSample, bool kScores>
static int launch_variant(dist_b200_ctx *ctx, const FeatList &feats, RowsArgs a, cudaStream_t s) {
    const int G = a.G;
    const int nchunks = (G + CHUNK - 1) / CHUNK;
    const int Gpad = nchunks * CHUNK;
    size_t cache_floats = 0;
    int max_stride = 4;
    for (int f = 0; f < feats.n; ++f) {
        const int st = feats.f[f].kind == DIST_B200_DD ? feats.f[f].vdim : 4;
        cache_floats += static_cast<size_t>(Gpad) * st;
        if (st > max_stride) max_stride = st;
    }
    size_t fixed = (33 * kLgammaRowStride + 64 + Gpad) * sizeof(float);
    if (kScores) fixed += (kThreads / 32) * 32 * 33 * sizeof(float);
    if (kSample && nchunks > 1) fixed += sizeof(float2) * kSlots * R * kThreads;
    a.resident = cache_floats * sizeof(float) <= kResidentBudget ? 1 : 0;
    a.stage_floats = CHUNK * max_stride;
    const size_t smem = fixed + (a.resident ? cache_floats : 2 * static_cast<size_t>(a.stage_floats)) * sizeof(float);
    if (smem > 227 * 1024) return fail(ctx, DIST_B200_ERR_UNSUPPORTED, "score_rows: group caches exceed shared memory");
    auto kern = score_rows_kernel<CHUNK, R, kSample, kScores>;
    DISTB200_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    DISTB200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const size_t rows_per_tile = static_cast<size_t>(kThreads) * R;
    const size_t ntiles = (a.N + rows_per_tile - 1) / rows_per_tile;
    const size_t max_blocks = static_cast<size_t>(ctx->sm_count) * per_sm;
    const unsigned grid = static_cast<unsigned>(ntiles < max_blocks ? ntiles : max_blocks);
    kern<<<grid, kThreads, smem, s>>>(feats, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("score_rows launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

template <int CHUNK, int R>
static int launch_modes(dist_b200_ctx *ctx, const FeatList &feats, const RowsArgs &a, cudaStream_t s) {
    const bool sample = a.assign != nullptr, scores = a.scores != nullptr;
    if (sample && scores) return launch_variant<CHUNK, R, true, true>(ctx, feats, a, s);
    if (sample) return launch_variant<CHUNK, R, true, false>(ctx, feats, a, s);
    return launch_variant<CHUNK, R, false, true>(ctx, feats, a, s);
}

int launch_score_rows(dist_b200_ctx *ctx, const FeatList &feats, int G, size_t N, const float *prior,
                      const float *u, int32_t *assign, float *scores, int accumulate, cudaStream_t s) {
    if (N == 0 || G == 0) return DIST_B200_OK;
    RowsArgs a{};
    a.G = G;
    a.N = N;
    a.prior = prior;
    a.u = u;
    a.assign = assign;
    a.scores = scores;
    a.accumulate = accumulate;
    a.t = ctx->tables;
    if (!assign && !scores) return fail(ctx, DIST_B200_ERR_INVALID, "score_rows: nothing to produce");
    // register tile: the whole row when it fits (the sampler is then the reference's loops
    // verbatim), 32-group chunks otherwise
    if (G <= 32) return launch_modes<32, 1>(ctx, feats, a, s);
    if (G <= 64) return launch_modes<64, 1>(ctx, feats, a, s);
    if (G <= 128) return launch_modes<128, 1>(ctx, feats, a, s);
    return launch_modes<32, 2>(ctx, feats, a, s);
}

}  // namespace distb200
