// caps_pass.cu -- instantiations + dispatch of the pass kernel (see caps_kernels.cuh).
#include "caps_internal.h"

namespace caps {
namespace {
// ---- kernel dispatch --------------------------------------------------------------------------
template <int DP, int SPT, int JW>
int launch_pass_t(const Plan& pl, int mode, const PassParams& pp, cudaStream_t st) {
    constexpr int K = 8;
    const size_t smem = (size_t)kPassStages * kPassIC * JW * K * DP * sizeof(float);
    dim3 grid(pl.IS, pl.JG, pl.ntg), block(32 * JW);
#define CAPS_LAUNCH_MODE(MODE)                                                                           \
    {                                                                                                    \
        auto kern = k_pass<K, DP, SPT, JW, MODE>;                                                        \
        CAPS_SET_SMEM(kern, smem);      /* per instantiation and per device */                           \
        kern<<<grid, block, smem, st>>>(pp);                                                             \
    }
    if (mode == kModeAUniform) CAPS_LAUNCH_MODE(kModeAUniform)
    else if (mode == kModeA) CAPS_LAUNCH_MODE(kModeA)
    else CAPS_LAUNCH_MODE(kModeL)
#undef CAPS_LAUNCH_MODE
    LAUNCH_CHECK();
    return 0;
}

template <int DP>
int launch_pass_d(const Plan& pl, int mode, const PassParams& pp, cudaStream_t st) {
    const int key = pl.JW * 10 + pl.SPT;
    switch (key) {
        case 11: return launch_pass_t<DP, 1, 1>(pl, mode, pp, st);
        case 12: return launch_pass_t<DP, 2, 1>(pl, mode, pp, st);
        case 41: return launch_pass_t<DP, 1, 4>(pl, mode, pp, st);
        case 42: return launch_pass_t<DP, 2, 4>(pl, mode, pp, st);
        case 81: return launch_pass_t<DP, 1, 8>(pl, mode, pp, st);
        case 82: return launch_pass_t<DP, 2, 8>(pl, mode, pp, st);
        default: break;
    }
    if constexpr (DP <= 16) {
        if (key == 14) return launch_pass_t<DP, 4, 1>(pl, mode, pp, st);
        if (key == 44) return launch_pass_t<DP, 4, 4>(pl, mode, pp, st);
        if (key == 84) return launch_pass_t<DP, 4, 8>(pl, mode, pp, st);
    }
    return fail(CAPS_E_UNSUPPORTED, "no pass kernel for JW=%d SPT=%d DP=%d", pl.JW, pl.SPT, DP);
}

}  // namespace

int launch_pass(const Plan& pl, int mode, const PassParams& pp, cudaStream_t st) {
    switch (pl.DP) {
        case 8: return launch_pass_d<8>(pl, mode, pp, st);
        case 16: return launch_pass_d<16>(pl, mode, pp, st);
        case 24: return launch_pass_d<24>(pl, mode, pp, st);
        case 32: return launch_pass_d<32>(pl, mode, pp, st);
        case 48: return launch_pass_d<48>(pl, mode, pp, st);
    }
    return fail(CAPS_E_UNSUPPORTED, "D=%d unsupported", pl.D);
}


}  // namespace caps
