// bc1d_inst_refb64.cu -- instantiates the Transport / ReactionDiffusion step kernels for ModeRefB64
#include "bc1d_kernels.cuh"

namespace pdecg {

int launch_bc1d_refb64(const Bc1dArgs& a, cudaStream_t st) {
    return a.cfg.env == PDECG_ENV_TRANSPORT ? launch_var<PDECG_ENV_TRANSPORT, ModeRefB64>(a, st)
                                            : launch_var<PDECG_ENV_REACTION_DIFFUSION, ModeRefB64>(a, st);
}

}  // namespace pdecg
