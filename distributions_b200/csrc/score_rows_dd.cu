// score_rows_dd.cu -- single-feature instantiations of score_rows_kernel for one model (see score_rows.cuh)
#include "score_rows.cuh"

namespace distb200 {

int launch_single_dd(dist_b200_ctx *ctx, const FeatList &feats, const RowsArgs &a, cudaStream_t s) {
    // G <= 128, sampling only: the max-relative, log2e-scaled table (kKindDdScaled); DIST_B200_OPT_SMALL_TILE != 0
    // keeps the generic tiers for A/B runs
    if (a.G <= 128 && a.assign && !a.scores && !a.accumulate && ctx->opt[DIST_B200_OPT_SMALL_TILE] == 0) {
        if (feats.f[0].vdim == 16 && a.G > 64) {  // DirichletDiscrete<16>: dim as a compile-time constant
            if (a.G <= 80) return launch_variant<80, kKindDdScaled16, true, false, 128>(ctx, feats, a, s);
            if (a.G <= 96) return launch_variant<96, kKindDdScaled16, true, false, 128>(ctx, feats, a, s);
            if (a.G <= 104) return launch_variant<104, kKindDdScaled16, true, false, 128>(ctx, feats, a, s);
            if (a.G <= 112) return launch_variant<112, kKindDdScaled16, true, false, 128>(ctx, feats, a, s);
            return launch_variant<128, kKindDdScaled16, true, false, 128>(ctx, feats, a, s);
        }
        if (a.G <= 32) return launch_variant<32, kKindDdScaled, true, false, 128>(ctx, feats, a, s);
        if (a.G <= 64) return launch_variant<64, kKindDdScaled, true, false, 128>(ctx, feats, a, s);
        if (a.G <= 80) return launch_variant<80, kKindDdScaled, true, false, 128>(ctx, feats, a, s);
        if (a.G <= 96) return launch_variant<96, kKindDdScaled, true, false, 128>(ctx, feats, a, s);
        if (a.G <= 112) return launch_variant<112, kKindDdScaled, true, false, 128>(ctx, feats, a, s);
        return launch_variant<128, kKindDdScaled, true, false, 128>(ctx, feats, a, s);
    }
    return launch_tiers<DIST_B200_DD>(ctx, feats, a, s);
}

}  // namespace distb200
