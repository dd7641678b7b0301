// neuron.cu -- NeuronPDE1D (neuron growth control: a 1-D advection-diffusion PDE on a moving domain coupled to a two-state
// ODE, closed with the env's own backstepping boundary law): step kernel + C ABI (sm_100a, float64).
//
// Reference behaviour restated here (no code shared): environments1d/neuron_env.py:263-372 of lukebhan/PDEControlGym.
// The reference class cannot run as shipped (`self.atilde1` at :288 vs `aTilde1` at :78; no reset(); the default
// sensing_loc="Full" raises): the numpy oracle is pinned to the class with those defects shimmed from outside
// (tests/test_oracle_vs_reference.py::test_neuron).  step() takes NO action -- the env is a closed-loop simulation.
//
// One warp = one neuron.  The concentration-error row u (M <= 512 points, K per lane) and the ODE state Z stay in
// registers for `n_steps` explicit steps per launch; per step: three row broadcasts, one warp reduction (the trapezoid
// integral of the control law; a reduction TREE, so results agree with the reference's sequential Python sum to
// rounding, not bit for bit: tolerance 1e-10 in the tests), two halo shuffles.  Everything is built with -fmad=false,
// so the plain operators below are numpy's individually rounded operations in numpy's order.
#include "pdecg_internal.h"

namespace pdecg {

struct NeuronArgs {
    pdecg_neuron_cfg cfg;
    pdecg_neuron_buffers buf;
    int n_steps;
};

// value of point `jq` of the row, broadcast to all lanes
template <int K>
__device__ __forceinline__ double nrow_value(const double (&u)[K], int jq) {
    const int src = jq / K, slot = jq - src * K;
    double v = u[0];
#pragma unroll
    for (int k = 1; k < K; ++k) v = (slot == k) ? u[k] : v;
    return __shfl_sync(kFull, v, src);
}

template <int K>
__global__ void __launch_bounds__(kThreads) neuron_step_kernel(const __grid_constant__ NeuronArgs a) {
    __shared__ double sm_all[kWarpsPerBlock][32 * K];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + wib;
    if (env >= a.buf.n_envs) return;
    double* sm = sm_all[wib];
    const pdecg_neuron_cfg& c = a.cfg;
    const int M = c.M, j0 = lane * K;
    double* urow = a.buf.u + env * M;
    double u[K], pB[K];
    load_row_chunk<double, double, K>(u, urow, M, sm, lane);
    load_row_chunk<double, double, K>(pB, a.buf.pB, M, sm, lane);
    double Z0 = a.buf.Z[env * 2], Z1 = a.buf.Z[env * 2 + 1];
    int L = a.buf.L[env];
    double time = a.buf.time[env], ctl = a.buf.control[env];
    bool term = false;
    const double pB0 = a.buf.pB[0];
    const double dt = c.dt, dx = c.dx, D = c.D, ca = c.a, g = c.g;

    for (int s = 0; s < a.n_steps; ++s) {
        time = time + dt;                                                         // :276
        const double lt = Z1 + c.LSubS;                                           // :277
        const double dxr = dx * lt;                                               // :278
        const double uL = nrow_value<K>(u, L), uL1 = nrow_value<K>(u, L - 1), uL2 = nrow_value<K>(u, L - 2);
        const double Z0n = (c.aTilde1 * Z0 - c.beta * (((3 * uL - 4 * uL1) + uL2) / (2 * dxr))) * dt + Z0;    // :288-290
        const double Z1n = (c.RSubG * Z0) * dt + Z1;                              // :293
        // Conversion (:363-372): code length -> grid index, python round() = round half to even
        const double lt_code = c.inv_scale * (Z1n + c.LSubS);
        int Ln = static_cast<int>(rint(lt_code / dx));
        Ln = Ln < 1 ? 1 : (Ln > M - 1 ? M - 1 : Ln);
        // trapezoid integral of the control law (:300-305)
        double part = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = j0 + k;
            if (j >= 1 && j < Ln) part += (dxr * pB[k]) * u[k];
        }
        const double sum = warp_sum(part);
        const double u0 = nrow_value<K>(u, 0), u1 = nrow_value<K>(u, 1), uLn = nrow_value<K>(u, Ln);
        const double pBL = a.buf.pB[Ln], p0L = a.buf.p[2 * Ln], p1L = a.buf.p[2 * Ln + 1];
        const double mid = (((((dxr * pB0) * u0) / 2) + sum) + (((dxr / 2) * pBL) * uLn)) / D;
        const double pz = p0L * Z0n + p1L * Z1n;                                  // p[L,:] @ Z  (Z already advanced, :288-293)
        ctl = lt * (((c.coef * u0) - mid) + pz);                                  // :305
        const double ufic = u1 - ((ctl * dxr) * 2);                               // :311
        const double dxr2 = dxr * dxr;
        const double new0 = ((((D / dxr2) * ((u1 - 2 * u0) + ufic)) - ((ca / (2 * dxr)) * (u1 - ufic))) - (g * u0)) * dt + u0;   // :312
        const double hl = __shfl_up_sync(kFull, u[K - 1], 1), hr = __shfl_down_sync(kFull, u[0], 1);
        const double zo = c.RSubG * Z0;
        double n[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = j0 + k;
            const double t = u[k];
            const double tp = (k == 0) ? hl : u[k > 0 ? k - 1 : 0];
            const double tn = (k == K - 1) ? hr : u[k < K - 1 ? k + 1 : K - 1];
            double v = t;
            if (j >= 1 && j < Ln) {                                               // :315-318
                const double uxx = ((tn - 2 * t) + tp) / dxr2;
                const double ux = (tn - tp) / dxr;
                const double adv = (((static_cast<double>(j - 1) / lt) * zo) * (tn - tp)) / 2;
                v = t + dt * ((((D * uxx) + adv) - ((ca * ux) / 2)) - (g * t));
            }
            if (j == 0) v = new0;
            if (j == Ln) v = c.H0 * Z0n + c.H1 * Z1n;                             // :320  H.T @ Z
            n[k] = v;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) u[k] = n[k];
        Z0 = Z0n;
        Z1 = Z1n;
        L = Ln;
        if (time >= static_cast<double>(c.nt)) {                                  // :341-346 (terminate() zeroes the time)
            time = 0.0;
            term = true;
        }
    }
    store_row_chunk<double, K>(u, urow, M, sm, lane);
    if (lane == 0) {
        a.buf.Z[env * 2] = Z0;
        a.buf.Z[env * 2 + 1] = Z1;
        a.buf.L[env] = L;
        a.buf.time[env] = time;
        a.buf.control[env] = ctl;
        a.buf.terminated[env] = term ? 1 : 0;
    }
}

}  // namespace pdecg

using namespace pdecg;

extern "C" int pdecg_neuron_step(const pdecg_neuron_cfg* cfg, const pdecg_neuron_buffers* buf, int32_t n_steps, pdecg_stream_t stream) {
    if (!cfg || !buf) return fail(PDECG_ERR_INVALID, "null cfg/buffers");
    if (cfg->abi_version != PDECG_ABI_VERSION) return fail(PDECG_ERR_INVALID, "abi_version mismatch");
    if (cfg->M < 4) return fail(PDECG_ERR_INVALID, "need at least 4 grid points");
    if (cfg->M > 512) return fail(PDECG_ERR_UNSUPPORTED, "grid of %d points exceeds the register-resident limit of 512 points per env", cfg->M);
    if (!(cfg->dt > 0) || !(cfg->dx > 0) || !(cfg->D > 0)) return fail(PDECG_ERR_INVALID, "dt, dx, D must be positive");
    if (n_steps < 0 || buf->n_envs < 0) return fail(PDECG_ERR_INVALID, "negative step / env count");
    if (buf->n_envs == 0 || n_steps == 0) return PDECG_OK;
    if (!buf->u || !buf->Z || !buf->L || !buf->time || !buf->control || !buf->pB || !buf->p || !buf->terminated)
        return fail(PDECG_ERR_INVALID, "null device buffer");
    NeuronArgs a{*cfg, *buf, n_steps};
    const unsigned grid = static_cast<unsigned>((buf->n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int K = (cfg->M + 31) / 32;
    if (K > 8) K = (K + 1) / 2 * 2;
    switch (K) {
#define PDECG_CASE(KK) case KK: neuron_step_kernel<KK><<<grid, kThreads, 0, st>>>(a); break;
        PDECG_CASE(1) PDECG_CASE(2) PDECG_CASE(3) PDECG_CASE(4) PDECG_CASE(5) PDECG_CASE(6) PDECG_CASE(7) PDECG_CASE(8)
        PDECG_CASE(10) PDECG_CASE(12) PDECG_CASE(14) PDECG_CASE(16)
#undef PDECG_CASE
        default: return fail(PDECG_ERR_UNSUPPORTED, "unsupported points-per-lane %d", K);
    }
    count_launch();
    return check_launch("neuron_step_kernel");
}
