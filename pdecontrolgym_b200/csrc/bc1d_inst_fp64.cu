// bc1d_inst_fp64.cu -- instantiates the Transport / ReactionDiffusion step kernels for ModeFp64
#include "bc1d_kernels.cuh"

namespace pdecg {

int launch_bc1d_fp64(const Bc1dArgs& a, cudaStream_t st) {
    return a.cfg.env == PDECG_ENV_TRANSPORT ? launch_var<PDECG_ENV_TRANSPORT, ModeFp64>(a, st)
                                            : launch_var<PDECG_ENV_REACTION_DIFFUSION, ModeFp64>(a, st);
}

}  // namespace pdecg
