// caps_grad.cu -- instantiations + dispatch of the final backward kernel (see caps_kernels.cuh).
#include "caps_internal.h"

namespace caps {
namespace {
template <int DP, int JW, int M>
int launch_grad_t(const Plan& pl, const GradParams& gp, cudaStream_t st) {
    constexpr int K = 8;
    constexpr int IT = (DP <= 16) ? 8 : 4;
    constexpr bool XREG = (M * DP <= 96);
    const size_t smem = ((size_t)2 * IT * JW * K * DP + (size_t)JW * kGradDuBatch * K * 32) * sizeof(float);
    auto kern = k_grad<K, DP, JW, IT, M, XREG>;
    CAPS_SET_SMEM(kern, smem);          // per instantiation and per device
    dim3 grid(cdiv(pl.N, IT), pl.JG), block(32 * JW);
    kern<<<grid, block, smem, st>>>(gp);
    LAUNCH_CHECK();
    return 0;
}

template <int DP, int JW>
int launch_grad_m(const Plan& pl, const GradParams& gp, cudaStream_t st) {
    switch (pl.M) {
        case 1: return launch_grad_t<DP, JW, 1>(pl, gp, st);
        case 3: return launch_grad_t<DP, JW, 3>(pl, gp, st);
        case 5: return launch_grad_t<DP, JW, 5>(pl, gp, st);
        case 7: return launch_grad_t<DP, JW, 7>(pl, gp, st);
        case 9: return launch_grad_t<DP, JW, 9>(pl, gp, st);
    }
    return fail(CAPS_E_UNSUPPORTED, "R=%d unsupported", pl.R);
}

template <int DP>
int launch_grad_j(const Plan& pl, const GradParams& gp, cudaStream_t st) {
    switch (pl.JW) {
        case 1: return launch_grad_m<DP, 1>(pl, gp, st);
        case 4: return launch_grad_m<DP, 4>(pl, gp, st);
        case 8: return launch_grad_m<DP, 8>(pl, gp, st);
    }
    return fail(CAPS_E_UNSUPPORTED, "JW=%d unsupported", pl.JW);
}

}  // namespace

int launch_grad(const Plan& pl, const GradParams& gp, cudaStream_t st) {
    switch (pl.DP) {
        case 8: return launch_grad_j<8>(pl, gp, st);
        case 16: return launch_grad_j<16>(pl, gp, st);
        case 24: return launch_grad_j<24>(pl, gp, st);
        case 32: return launch_grad_j<32>(pl, gp, st);
        case 48: return launch_grad_j<48>(pl, gp, st);
    }
    return fail(CAPS_E_UNSUPPORTED, "D=%d unsupported", pl.D);
}


}  // namespace caps
