// bc1d_inst_fp32.cu -- instantiates the Transport / ReactionDiffusion step kernels for ModeFp32
#include "bc1d_kernels.cuh"

namespace pdecg {

int launch_bc1d_fp32(const Bc1dArgs& a, cudaStream_t st) {
    return a.cfg.env == PDECG_ENV_TRANSPORT ? launch_var<PDECG_ENV_TRANSPORT, ModeFp32>(a, st)
                                            : launch_var<PDECG_ENV_REACTION_DIFFUSION, ModeFp32>(a, st);
}

}  // namespace pdecg
