// niw_tc.cu -- NormalInverseWishart<32> quadratic forms on the 5th-generation tensor cores.
//
// The only dense contraction on the hot path: for every (row n, group g)
//   y = W_g x_n   (W_g = L_g^-1, d x d),    q = |y - W_g mu'_g|^2,
//   score = C_g - 0.5 (dof_g + d) fast_log(1 + q / dof_g)          (niw.hpp:353-360, random.hpp:160-185)
// Stacking the W_g of 8 groups gives a B operand of 256 rows, so one tcgen05.mma tile is
//   D[128 rows][256 = 8 groups x 32 dims] = X[128][32] * Wstack[256][32]^T        (kind::tf32, fp32 accum)
// with the accumulator in TMEM (2 x 256 columns, double buffered).  A TMEM lane is a data row and a
// 32-column slice is exactly one cell's y vector, so the epilogue is one tcgen05.ld (32x32b.x32) per
// cell followed by 32 (y-b)^2 FMAs, MUFU.LG2 and the score FFMA, all in registers.
//
// Precision: one TF32 pass rounds x and W to 10 mantissa bits, and y - b cancels |W mu'| ~ 17 down to
// |y - b| ~ 1: ~0.04 absolute on a score.  The default mode therefore splits both operands
// (x = x_hi + x_lo, W = W_hi + W_lo with x_hi, W_hi exactly representable in TF32) and accumulates
// x_hi W_hi + x_lo W_hi + x_hi W_lo in the same TMEM tile ("3xTF32", error ~2^-21), which keeps the
// scores within the fp32 kernel's tolerance.  DIST_B200_NIW_MODE=tf32 selects the single pass.
//
// Pipeline per CTA (128 threads = 4 warps = the 128 TMEM lanes; one CTA per SM, persistent over row
// tiles): the X tile is converted and laid out in shared memory once per row tile; the operand images
// of 8-group blocks, pre-laid-out by niw_tc_prep_kernel in the canonical no-swizzle K-major core-matrix
// order, stream in with cp.async.bulk + mbarrier (double buffered); one thread issues the MMAs and a
// tcgen05.commit; the epilogue of block b-1 runs while the tensor pipe works on block b.
#include <cstdlib>

#include "common.cuh"

namespace distb200 {

constexpr int kTcDim = 32;                 // d (K of the GEMM)
constexpr int kTcGroupsPerBlock = 8;       // groups per B tile (N = 256)
constexpr int kTcRows = 128;               // rows per tile (M)
constexpr int kTcThreads = 256;            // two warpgroups: both see all 128 TMEM lanes, each drains half the columns
constexpr int kTcImageFloats = 256 * 32;   // one operand image of a block: 256 x 32 fp32 = 32 KB

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// no-swizzle K-major canonical layout: core matrix = 8 rows x 16 bytes, rows 16 B apart;
// core (row group j, k core c) at byte offset (c * n_row_groups + j) * 128
__host__ __device__ __forceinline__ int core_offset_floats(int row, int k, int n_row_groups) {
    return ((k >> 2) * n_row_groups + (row >> 3)) * 32 + (row & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading (K) byte offset, bits [16,30)
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride (M/N) byte offset, bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                           // descriptor version (Blackwell)
    return d;                                                      // layout type 0 = no swizzle
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 26)) __trap();  // never hang the device on a protocol error
    }
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 64 columns = two cells per instruction, results left in flight (the caller issues tcgen05.wait::ld once)
__device__ __forceinline__ void tmem_ld64_nowait(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
}

// Blackwell packed fp32 pairs: one instruction, two lanes of the FMA pipe (halves the epilogue's issue count)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float sum2(uint64_t a) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    return lo + hi;
}

// ---------------------------------------------------------------------------------------------
// Operand images.  From the group records of niw_prep_kernel (mu'[32] | W[32][32] | consts[4]) build, per
// block of 8 groups: W_hi image (32 KB), W_lo image (32 KB) in core-matrix order; and per group
// -b = -(W mu') (32 floats) and the constants.
__global__ void niw_tc_prep_kernel(int G, int n_blocks, const float *__restrict__ recs, float *__restrict__ images,
                                   float *__restrict__ bvec, float *__restrict__ consts) {
    constexpr int REC = kTcDim + kTcDim * kTcDim + 4;
    const int blk = blockIdx.x;
    for (int e = threadIdx.x; e < 256 * 32; e += blockDim.x) {
        const int n = e >> 5, k = e & 31;          // B row n = (group in block) * 32 + i, column k
        const int g = blk * kTcGroupsPerBlock + (n >> 5), i = n & 31;
        const float w = g < G ? recs[static_cast<size_t>(g) * REC + kTcDim + i * kTcDim + k] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);  // exactly representable in TF32
        const int off = core_offset_floats(n, k, 32);
        images[static_cast<size_t>(blk) * 2 * kTcImageFloats + off] = hi;
        images[static_cast<size_t>(blk) * 2 * kTcImageFloats + kTcImageFloats + off] = w - hi;
    }
    for (int e = threadIdx.x; e < 256; e += blockDim.x) {
        const int g = blk * kTcGroupsPerBlock + (e >> 5), i = e & 31;
        double b = 0.0;
        if (g < G) {
            const float *rec = recs + static_cast<size_t>(g) * REC;
            for (int k = 0; k <= i; ++k) b += static_cast<double>(rec[kTcDim + i * kTcDim + k]) * static_cast<double>(rec[k]);
        }
        bvec[static_cast<size_t>(blk) * 256 + e] = static_cast<float>(-b);  // negated: the epilogue adds
    }
    if (threadIdx.x < kTcGroupsPerBlock * 4) {
        const int g = blk * kTcGroupsPerBlock + (threadIdx.x >> 2), c = threadIdx.x & 3;
        consts[static_cast<size_t>(blk) * 32 + threadIdx.x] = g < G ? recs[static_cast<size_t>(g) * REC + kTcDim + kTcDim * kTcDim + c] : 0.f;
    }
    (void)n_blocks;
}

struct NiwTcArgs {
    int G, n_blocks, accumulate;
    size_t N;
    const float *images;  // [n_blocks][2][256*32]
    const float *bvec;    // [n_blocks][256]
    const float *consts;  // [n_blocks][8][4]
    const float *values;  // [N][32]
    const float *prior;   // [G] or nullptr
    float *scores;        // [N][G]
};

// W-stationary schedule.  A work item is (block of 8 groups, chunk of row tiles): the CTA loads that
// block's operand images once (64 KB in split mode), then streams the chunk's row tiles through a
// double-buffered A tile and a double-buffered TMEM accumulator.  Items are ordered chunk-major so the
// CTAs running at the same time share a 1 MB chunk of rows in L2 while each keeps its own W block in
// shared memory: L2->SM traffic is n_blocks x |X| instead of n_row_tiles x |W|.
constexpr int kTcChunkTiles = 64;  // row tiles per work item (8192 rows)

template <bool kSplit>
__global__ void __launch_bounds__(kTcThreads, 1) niw_tc_kernel(const NiwTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int kImages = kSplit ? 2 : 1;
    constexpr int kAFloats = kTcRows * kTcDim;                                // one A image: 16 KB
    float *As = reinterpret_cast<float *>(smem_raw);                          // [2 buffers][kImages][16 KB]
    float *Bs = As + 2 * kImages * kAFloats;                                  // [kImages][32 KB]
    float *bs = Bs + kImages * kTcImageFloats;                                // [256] -b = -(W mu')
    float *cs = bs + 256;                                                     // [8][4] constants
    float *ps = cs + 32;                                                      // [8] prior
    uint64_t *bars = reinterpret_cast<uint64_t *>(ps + 8);                    // bfull, mma_done[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r_in_tile = tid & (kTcRows - 1);  // this thread's row of the tile = its TMEM lane
    const int wg = tid >> 7;                     // warpgroup: k-cores [4 wg, 4 wg + 4) of A, groups [4 wg, 4 wg + 4) of D
    const int nb = a.n_blocks;
    uint64_t *bfull = bars, *mma_done = bars + 1;

    if (tid == 0) {
        mbar_init(bfull, 1);
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t image_bytes = kTcImageFloats * sizeof(float);
    uint32_t full_phase = 0, done_phase[2] = {0, 0};

    const size_t ntiles = (a.N + kTcRows - 1) / kTcRows;
    const size_t nchunks = (ntiles + kTcChunkTiles - 1) / kTcChunkTiles;
    const size_t nitems = nchunks * nb;

    // this thread's half row of a tile: global -> registers
    auto load_x = [&](size_t tile, float4 (&x)[4]) {
        size_t row = tile * kTcRows + r_in_tile;
        if (row >= a.N) row = a.N - 1;
        const float4 *src = reinterpret_cast<const float4 *>(a.values + row * kTcDim) + 4 * wg;
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = __ldg(src + c);
    };
    // registers -> A buffer `ab` in core-matrix order, split into TF32-exact hi and the remainder
    auto store_x = [&](int ab, const float4 (&x)[4]) {
        float *A_hi = As + ab * kImages * kAFloats, *A_lo = A_hi + kAFloats;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float4 hi;
            hi.x = __uint_as_float(__float_as_uint(x[c].x) & 0xFFFFE000u);
            hi.y = __uint_as_float(__float_as_uint(x[c].y) & 0xFFFFE000u);
            hi.z = __uint_as_float(__float_as_uint(x[c].z) & 0xFFFFE000u);
            hi.w = __uint_as_float(__float_as_uint(x[c].w) & 0xFFFFE000u);
            const int off = core_offset_floats(r_in_tile, (4 * wg + c) * 4, kTcRows / 8);
            *reinterpret_cast<float4 *>(A_hi + off) = kSplit ? hi : x[c];
            if (kSplit)
                *reinterpret_cast<float4 *>(A_lo + off) = make_float4(x[c].x - hi.x, x[c].y - hi.y, x[c].z - hi.z, x[c].w - hi.w);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy stores -> visible to the MMA
    };

    for (size_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int blk = static_cast<int>(item % nb);
        const size_t chunk = item / nb;
        const size_t t0 = chunk * kTcChunkTiles, t1 = t0 + kTcChunkTiles < ntiles ? t0 + kTcChunkTiles : ntiles;
        const int nt = static_cast<int>(t1 - t0);
        // ---- per-item setup: W images (bulk, async), b / constants / prior of this block, A tile 0
        if (tid == 0) {
            mbar_expect_tx(bfull, kImages * image_bytes);
            for (int im = 0; im < kImages; ++im)
                bulk_g2s(Bs + im * kTcImageFloats, a.images + (static_cast<size_t>(blk) * 2 + im) * kTcImageFloats, image_bytes, bfull);
        }
        bs[tid] = a.bvec[static_cast<size_t>(blk) * 256 + tid];
        if (tid < 32) cs[tid] = a.consts[static_cast<size_t>(blk) * 32 + tid];
        if (tid < 8) {
            const int g = blk * kTcGroupsPerBlock + tid;
            ps[tid] = (g < a.G && a.prior && !a.accumulate) ? a.prior[g] : 0.f;
        }
        {
            float4 x[4];
            load_x(t0, x);
            store_x(0, x);
        }
        __syncthreads();
        if (tid == 0) {
            mbar_wait(bfull, full_phase);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
        full_phase ^= 1;
        __syncwarp();

        for (int t = 0; t <= nt; ++t) {
            const int buf = t & 1;
            float4 xn[4];
            const bool have_next = t + 1 < nt;
            if (have_next) load_x(t0 + t + 1, xn);  // in flight across the wait and the epilogue below
            if (t >= 1) {  // MMA of tile t-1 finished: accumulator ready, A buffer (t-1)&1 reusable
                mbar_wait(&mma_done[buf ^ 1], done_phase[buf ^ 1]);
                done_phase[buf ^ 1] ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            if (tid == 0 && t < nt) {
                const uint32_t d_addr = tmem_base + buf * 256;
                const uint32_t a_hi = smem_u32(As + buf * kImages * kAFloats), a_lo = a_hi + kAFloats * sizeof(float);
                const uint32_t b_hi = smem_u32(Bs), b_lo = b_hi + image_bytes;
                uint32_t acc = 0;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {  // k-step of 8 = two k-cores
                    const uint32_t ao = ks * 2 * (kTcRows / 8) * 128, bo = ks * 2 * 32 * 128;
                    umma_tf32(d_addr, make_smem_desc(a_hi + ao, (kTcRows / 8) * 128, 128), make_smem_desc(b_hi + bo, 32 * 128, 128), idesc, acc);
                    acc = 1;
                    if (kSplit) {
                        umma_tf32(d_addr, make_smem_desc(a_lo + ao, (kTcRows / 8) * 128, 128), make_smem_desc(b_hi + bo, 32 * 128, 128), idesc, 1);
                        umma_tf32(d_addr, make_smem_desc(a_hi + ao, (kTcRows / 8) * 128, 128), make_smem_desc(b_lo + bo, 32 * 128, 128), idesc, 1);
                    }
                }
                umma_commit(&mma_done[buf]);
            }
            __syncwarp();  // warp 0 reconverges before the warp-collective tcgen05.ld below
            if (t >= 1) {
                // ---- epilogue of tile t-1: TMEM lane = row, 32 columns = one group's y vector
                const int pbuf = buf ^ 1;
                const size_t row = (t0 + t - 1) * kTcRows + r_in_tile;
                constexpr int kPer = kTcGroupsPerBlock / 2;  // groups per warpgroup
                float out[kPer];
                // all four cells of this thread are pulled out of TMEM up front (2 loads of 64 columns,
                // one wait): the TMEM latency is paid once per tile, and the four sums of squares below
                // are independent instruction streams
                uint32_t yr[2][64];
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + pbuf * 256 + wg * kPer * 32;
                tmem_ld64_nowait(t_row, yr[0]);
                tmem_ld64_nowait(t_row + 64, yr[1]);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int jj = 0; jj < kPer; ++jj) {
                    const int j = wg * kPer + jj;
                    const uint32_t *y = &yr[jj >> 1][(jj & 1) * 32];
                    const float4 *b4 = reinterpret_cast<const float4 *>(bs + j * 32);
                    // |y - b|^2 with packed fp32x2 adds / fmas, four independent chains
                    uint64_t qa = 0ull, qb = 0ull;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 nb4 = b4[i];  // holds -b
                        const uint64_t d0 = add2(pack2(__uint_as_float(y[4 * i]), __uint_as_float(y[4 * i + 1])), pack2(nb4.x, nb4.y));
                        const uint64_t d1 = add2(pack2(__uint_as_float(y[4 * i + 2]), __uint_as_float(y[4 * i + 3])), pack2(nb4.z, nb4.w));
                        qa = fma2(d0, d0, qa);
                        qb = fma2(d1, d1, qb);
                    }
                    const float q = sum2(qa) + sum2(qb);
                    const float4 c = *reinterpret_cast<const float4 *>(cs + j * 4);
                    const float arg = __fadd_rn(1.f, __fmul_rn(c.z, q));
                    out[jj] = fmaf(c.y, fast_log2_cell(arg), c.x) + ps[j];
                }
                if (row < a.N) {
                    const int gbase = blk * kTcGroupsPerBlock + wg * kPer;
                    float *dst = a.scores + row * a.G + gbase;
                    if (gbase + kPer <= a.G && (a.G & 3) == 0) {
                        float4 o0 = make_float4(out[0], out[1], out[2], out[3]);
                        if (a.accumulate) {
                            const float4 p0 = *reinterpret_cast<float4 *>(dst);
                            o0 = make_float4(o0.x + p0.x, o0.y + p0.y, o0.z + p0.z, o0.w + p0.w);
                        }
                        *reinterpret_cast<float4 *>(dst) = o0;
                    } else {
#pragma unroll
                        for (int jj = 0; jj < kPer; ++jj)
                            if (gbase + jj < a.G) dst[jj] = a.accumulate ? dst[jj] + out[jj] : out[jj];
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            }
            if (have_next) store_x(buf ^ 1, xn);  // A buffer (t+1)&1 was last read by MMA(t-1), complete above
            __syncthreads();  // accumulator (t-1)&1 drained and A tile t+1 visible before the next issue
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised form.  The row tiles are first packed once into the A-operand images
// (niw_tc_pack_x_kernel: TF32-exact hi part + remainder, core-matrix order, zero padded), so that the
// main kernel moves EVERY operand with cp.async.bulk and no thread touches operand data:
//   warp 0 (one lane)  : loader  -- W images of the item, then the A images of its row tiles through a
//                                   kAStages-deep ring (a_full / a_empty mbarriers)
//   warp 1 (one lane)  : issuer  -- tcgen05.mma for tile t into TMEM buffer t & 1; tcgen05.commit frees
//                                   the A stage (a_empty) and publishes the accumulator (t_full)
//   warps 2..9         : epilogue -- tcgen05.ld, release the TMEM buffer (t_empty) as soon as the loads
//                                   have landed, then sums of squares / MUFU.LG2 / score stores
// so the tensor pipe runs tile t+1 while the epilogue drains tile t, with no block-wide barrier.
constexpr int kAStages = 3;
constexpr int kWsThreads = 320;

template <bool kSplit>
__global__ void niw_tc_pack_x_kernel(size_t N, size_t ntiles, const float *__restrict__ values, float *__restrict__ xpack) {
    constexpr int kImages = kSplit ? 2 : 1;
    constexpr int kAFloats = kTcRows * kTcDim;
    // one thread per (row, k-core): 4 floats
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= ntiles * kTcRows * 8) return;
    const size_t row = i >> 3;
    const int c = static_cast<int>(i & 7);
    const size_t tile = row / kTcRows;
    const int r = static_cast<int>(row % kTcRows);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < N) x = __ldg(reinterpret_cast<const float4 *>(values + row * kTcDim) + c);
    float4 hi;
    hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
    hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
    hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
    hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
    float *base = xpack + tile * kImages * kAFloats;
    const int off = core_offset_floats(r, c * 4, kTcRows / 8);
    *reinterpret_cast<float4 *>(base + off) = kSplit ? hi : x;
    if (kSplit) *reinterpret_cast<float4 *>(base + kAFloats + off) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

template <bool kSplit>
__global__ void __launch_bounds__(kWsThreads, 1) niw_tc_ws_kernel(const NiwTcArgs a, const float *__restrict__ xpack) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int kImages = kSplit ? 2 : 1;
    constexpr int kAFloats = kTcRows * kTcDim;
    float *As = reinterpret_cast<float *>(smem_raw);                          // [kAStages][kImages][16 KB]
    float *Bs = As + kAStages * kImages * kAFloats;                           // [kImages][32 KB]
    float *bs = Bs + kImages * kTcImageFloats;                                // [256] -b
    float *cs = bs + 256;                                                     // [8][4]
    float *ps = cs + 32;                                                      // [8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ps + 8);
    uint64_t *a_full = bars, *a_empty = bars + kAStages, *t_full = bars + 2 * kAStages, *t_empty = t_full + 2;
    uint64_t *b_full = t_empty + 2, *b_empty = b_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = a.n_blocks;
    if (tid == 0) {
        for (int i = 0; i < kAStages; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 8);  // one arrival per epilogue warp
        }
        mbar_init(b_full, 1);
        mbar_init(b_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t image_bytes = kTcImageFloats * sizeof(float), a_bytes = kAFloats * sizeof(float);

    const size_t ntiles = (a.N + kTcRows - 1) / kTcRows;
    const size_t nchunks = (ntiles + kTcChunkTiles - 1) / kTcChunkTiles;
    const size_t nitems = nchunks * nb;

    if (warp == 0) {
        // ===================== loader =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, bphase = 0;
            for (size_t item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int blk = static_cast<int>(item % nb);
                const size_t t0 = (item / nb) * kTcChunkTiles, t1 = t0 + kTcChunkTiles < ntiles ? t0 + kTcChunkTiles : ntiles;
                mbar_wait(b_empty, bphase ^ 1);  // every MMA of the previous item has finished reading Bs
                mbar_expect_tx(b_full, kImages * image_bytes);
                for (int im = 0; im < kImages; ++im)
                    bulk_g2s(Bs + im * kTcImageFloats, a.images + (static_cast<size_t>(blk) * 2 + im) * kTcImageFloats, image_bytes, b_full);
                bphase ^= 1;
                for (size_t t = t0; t < t1; ++t) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    mbar_expect_tx(&a_full[stage], kImages * a_bytes);
                    bulk_g2s(As + stage * kImages * kAFloats, xpack + t * kImages * kAFloats, kImages * a_bytes, &a_full[stage]);
                    if (++stage == kAStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, bphase = 0, tphase[2] = {0, 0};
            int tb = 0;
            for (size_t item = blockIdx.x; item < nitems; item += gridDim.x) {
                const size_t t0 = (item / nb) * kTcChunkTiles, t1 = t0 + kTcChunkTiles < ntiles ? t0 + kTcChunkTiles : ntiles;
                mbar_wait(b_full, bphase);
                bphase ^= 1;
                for (size_t t = t0; t < t1; ++t) {
                    mbar_wait(&a_full[stage], phase);
                    mbar_wait(&t_empty[tb], tphase[tb] ^ 1);  // epilogue has drained this accumulator
                    tphase[tb] ^= 1;
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t d_addr = tmem_base + tb * 256;
                    const uint32_t a_hi = smem_u32(As + stage * kImages * kAFloats), a_lo = a_hi + a_bytes;
                    const uint32_t b_hi = smem_u32(Bs), b_lo = b_hi + image_bytes;
                    uint32_t acc = 0;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t ao = ks * 2 * (kTcRows / 8) * 128, bo = ks * 2 * 32 * 128;
                        umma_tf32(d_addr, make_smem_desc(a_hi + ao, (kTcRows / 8) * 128, 128), make_smem_desc(b_hi + bo, 32 * 128, 128), idesc, acc);
                        acc = 1;
                        if (kSplit) {
                            umma_tf32(d_addr, make_smem_desc(a_lo + ao, (kTcRows / 8) * 128, 128), make_smem_desc(b_hi + bo, 32 * 128, 128), idesc, 1);
                            umma_tf32(d_addr, make_smem_desc(a_hi + ao, (kTcRows / 8) * 128, 128), make_smem_desc(b_lo + bo, 32 * 128, 128), idesc, 1);
                        }
                    }
                    umma_commit(&a_empty[stage]);  // A stage reusable once these MMAs retire
                    umma_commit(&t_full[tb]);      // accumulator ready for the epilogue
                    if (++stage == kAStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                    tb ^= 1;
                }
                umma_commit(b_empty);  // all MMAs of this item retired -> Bs may be overwritten
            }
        }
    } else {
        // ===================== epilogue (8 warps) =====================
        const int e = warp - 2;        // 0..7
        const int q = warp & 3;        // TMEM lane quarter this warp may access
        const int h = e >> 2;          // column half: groups [4h, 4h + 4)
        const int r_in_tile = q * 32 + lane;
        constexpr int kPer = kTcGroupsPerBlock / 2;
        uint32_t fphase[2] = {0, 0};
        int tb = 0;
        for (size_t item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int blk = static_cast<int>(item % nb);
            const size_t t0 = (item / nb) * kTcChunkTiles, t1 = t0 + kTcChunkTiles < ntiles ? t0 + kTcChunkTiles : ntiles;
            // per-block tables (the previous item's epilogue must be finished in all 8 warps first)
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            const int et = tid - 64;
            bs[et] = a.bvec[static_cast<size_t>(blk) * 256 + et];
            if (et < 32) cs[et] = a.consts[static_cast<size_t>(blk) * 32 + et];
            if (et < 8) {
                const int g = blk * kTcGroupsPerBlock + et;
                ps[et] = (g < a.G && a.prior && !a.accumulate) ? a.prior[g] : 0.f;
            }
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            for (size_t t = t0; t < t1; ++t) {
                mbar_wait(&t_full[tb], fphase[tb]);
                fphase[tb] ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                uint32_t yr[2][64];
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tb * 256 + h * kPer * 32;
                tmem_ld64_nowait(t_row, yr[0]);
                tmem_ld64_nowait(t_row + 64, yr[1]);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                if (lane == 0) {  // accumulator drained by this warp: one of the 8 arrivals
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&t_empty[tb])) : "memory");
                }
                tb ^= 1;
                float out[kPer];
#pragma unroll
                for (int jj = 0; jj < kPer; ++jj) {
                    const int j = h * kPer + jj;
                    const uint32_t *y = &yr[jj >> 1][(jj & 1) * 32];
                    const float4 *b4 = reinterpret_cast<const float4 *>(bs + j * 32);
                    uint64_t qa = 0ull, qb = 0ull;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 nb4 = b4[i];  // holds -b
                        const uint64_t d0 = add2(pack2(__uint_as_float(y[4 * i]), __uint_as_float(y[4 * i + 1])), pack2(nb4.x, nb4.y));
                        const uint64_t d1 = add2(pack2(__uint_as_float(y[4 * i + 2]), __uint_as_float(y[4 * i + 3])), pack2(nb4.z, nb4.w));
                        qa = fma2(d0, d0, qa);
                        qb = fma2(d1, d1, qb);
                    }
                    const float qq = sum2(qa) + sum2(qb);
                    const float4 c = *reinterpret_cast<const float4 *>(cs + j * 4);
                    const float arg = __fadd_rn(1.f, __fmul_rn(c.z, qq));
                    out[jj] = fmaf(c.y, fast_log2_cell(arg), c.x) + ps[j];
                }
                const size_t row = t * kTcRows + r_in_tile;
                if (row < a.N) {
                    const int gbase = blk * kTcGroupsPerBlock + h * kPer;
                    float *dst = a.scores + row * a.G + gbase;
                    if (gbase + kPer <= a.G && (a.G & 3) == 0) {
                        float4 o0 = make_float4(out[0], out[1], out[2], out[3]);
                        if (a.accumulate) {
                            const float4 p0 = *reinterpret_cast<float4 *>(dst);
                            o0 = make_float4(o0.x + p0.x, o0.y + p0.y, o0.z + p0.z, o0.w + p0.w);
                        }
                        *reinterpret_cast<float4 *>(dst) = o0;
                    } else {
#pragma unroll
                        for (int jj = 0; jj < kPer; ++jj)
                            if (gbase + jj < a.G) dst[jj] = a.accumulate ? dst[jj] + out[jj] : out[jj];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
}

// ---------------------------------------------------------------------------------------------
size_t niw_tc_floats(int G) {
    const size_t nb = (G + kTcGroupsPerBlock - 1) / kTcGroupsPerBlock;
    return nb * (2 * kTcImageFloats + 256 + 32);
}

int launch_niw_tc_prep(dist_b200_ctx *ctx, int G, const float *recs, float *tc_buf, cudaStream_t s) {
    const int nb = (G + kTcGroupsPerBlock - 1) / kTcGroupsPerBlock;
    if (nb == 0) return DIST_B200_OK;
    float *images = tc_buf, *bvec = images + static_cast<size_t>(nb) * 2 * kTcImageFloats, *consts = bvec + static_cast<size_t>(nb) * 256;
    niw_tc_prep_kernel<<<nb, 256, 0, s>>>(G, nb, recs, images, bvec, consts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("niw_tc_prep launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

int launch_niw_tc_scores(dist_b200_ctx *ctx, int G, const float *tc_buf, const void *values, size_t N, const float *prior,
                         float *scores, int accumulate, bool split, cudaStream_t s) {
    if (N == 0 || G == 0) return DIST_B200_OK;
    const int nb = (G + kTcGroupsPerBlock - 1) / kTcGroupsPerBlock;
    NiwTcArgs a{};
    a.G = G;
    a.n_blocks = nb;
    a.accumulate = accumulate;
    a.N = N;
    a.images = tc_buf;
    a.bvec = tc_buf + static_cast<size_t>(nb) * 2 * kTcImageFloats;
    a.consts = a.bvec + static_cast<size_t>(nb) * 256;
    a.values = static_cast<const float *>(values);
    a.prior = prior;
    a.scores = scores;
    const int images = split ? 2 : 1;
    const size_t ntiles = (N + kTcRows - 1) / kTcRows;
    const size_t nitems = ((ntiles + kTcChunkTiles - 1) / kTcChunkTiles) * static_cast<size_t>(nb);
    const unsigned grid = static_cast<unsigned>(nitems < static_cast<size_t>(ctx->sm_count) ? nitems : ctx->sm_count);
    static const bool use_ws = [] {
        const char *e = getenv("DIST_B200_NIW_WS");
        return !(e && e[0] == '0');
    }();
    cudaError_t e;
    if (use_ws) {
        // pack the rows into A-operand images once, then the warp-specialised pipeline
        const size_t xbytes = sizeof(float) * ntiles * images * kTcRows * kTcDim;
        if (xbytes > ctx->xpack_bytes) {
            if (ctx->xpack) {
                DISTB200_CUDA(ctx, cudaDeviceSynchronize());
                DISTB200_CUDA(ctx, cudaFree(ctx->xpack));
                ctx->xpack = nullptr;
                ctx->xpack_bytes = 0;
            }
            DISTB200_CUDA(ctx, cudaMalloc(&ctx->xpack, xbytes));
            ctx->xpack_bytes = xbytes;
        }
        float *xpack = static_cast<float *>(ctx->xpack);
        const size_t nthreads = ntiles * kTcRows * 8;
        const unsigned pgrid = static_cast<unsigned>((nthreads + 255) / 256);
        const size_t smem = sizeof(float) * (static_cast<size_t>(kAStages) * images * kTcRows * kTcDim +
                                             static_cast<size_t>(images) * kTcImageFloats + 256 + 32 + 8) + 256 + 1024;
        if (split) {
            niw_tc_pack_x_kernel<true><<<pgrid, 256, 0, s>>>(N, ntiles, a.values, xpack);
            e = cudaFuncSetAttribute(niw_tc_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e == cudaSuccess) niw_tc_ws_kernel<true><<<grid, kWsThreads, smem, s>>>(a, xpack);
        } else {
            niw_tc_pack_x_kernel<false><<<pgrid, 256, 0, s>>>(N, ntiles, a.values, xpack);
            e = cudaFuncSetAttribute(niw_tc_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e == cudaSuccess) niw_tc_ws_kernel<false><<<grid, kWsThreads, smem, s>>>(a, xpack);
        }
    } else {
        const size_t smem = sizeof(float) * (2 * static_cast<size_t>(images) * kTcRows * kTcDim + static_cast<size_t>(images) * kTcImageFloats +
                                             256 + 32 + 8) + 64 + 1024;
        if (split) {
            e = cudaFuncSetAttribute(niw_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e == cudaSuccess) niw_tc_kernel<true><<<grid, kTcThreads, smem, s>>>(a);
        } else {
            e = cudaFuncSetAttribute(niw_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e == cudaSuccess) niw_tc_kernel<false><<<grid, kTcThreads, smem, s>>>(a);
        }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, DIST_B200_ERR_CUDA, std::string("niw_tc launch: ") + cudaGetErrorString(e));
    return DIST_B200_OK;
}

}  // namespace distb200
