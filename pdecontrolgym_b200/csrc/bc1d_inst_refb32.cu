// bc1d_inst_refb32.cu -- instantiates the Transport / ReactionDiffusion step kernels for ModeRefB32
#include "bc1d_kernels.cuh"

namespace pdecg {

int launch_bc1d_refb32(const Bc1dArgs& a, cudaStream_t st) {
    return a.cfg.env == PDECG_ENV_TRANSPORT ? launch_var<PDECG_ENV_TRANSPORT, ModeRefB32>(a, st)
                                            : launch_var<PDECG_ENV_REACTION_DIFFUSION, ModeRefB32>(a, st);
}

}  // namespace pdecg
