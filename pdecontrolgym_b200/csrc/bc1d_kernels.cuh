// bc1d_kernels.cuh -- TransportPDE1D and ReactionDiffusionPDE1D step kernel templates (sm_100a).
// Instantiated per precision mode in bc1d_inst_*.cu (parallel compilation); ABI + reset in bc1d.cu.
//
// Reference behaviour restated here (no code shared): hyperbolic.py:126-227, parabolic.py:126-221,
// base_env_1d.py:20-45, tuned_reward_1d.py:25-40, norm_reward.py:29-73 of lukebhan/PDEControlGym.
//
// One warp integrates one env.  All `substeps` explicit finite-difference updates of a control step
// run on registers; actuation (Dirichlet / Neumann, optional normalisation), sensing, the reward
// (TunedReward1D / NormReward) and terminate / truncate are fused into the same launch, so HBM
// carries one state row in and out (+ a few scalars) per env per control step.
#pragma once
#include <cstdlib>

#include "pdecg_internal.h"

namespace pdecg {

// precision / arithmetic variants ---------------------------------------------------------------
struct ModeRefB32 { using S = float;  using C = float;  static constexpr bool kFast = false; };  // f32 state, beta f32
struct ModeRefB64 { using S = float;  using C = double; static constexpr bool kFast = false; };  // f32 state, beta f64
struct ModeFp64   { using S = double; using C = double; static constexpr bool kFast = false; };
struct ModeFp32   { using S = float;  using C = float;  static constexpr bool kFast = true;  };

struct RandomIc {          // optional: u <- c * profile with c ~ U(lo, hi) from Philox (pdecg_bc1d_reset_random)
    int enabled;
    double lo, hi;
    uint64_t seed;
    int64_t env_id_offset;
    int32_t* episode;
};

struct AutoReset {         // pdecg_bc1d_step_autoreset: finished envs restart inside the step launch
    int enabled;
    int init_dtype;
    const void* init;
    int64_t init_stride;
    RandomIc rnd;
    void* final_obs;       // optional: terminal observation of the envs that restart (pdecg_bc1d_autoreset::final_obs)
};

struct Bc1dArgs {
    pdecg_bc1d_cfg cfg;
    pdecg_bc1d_buffers buf;
    int npts;
    int carry_len;
    int ring_len;   // TunedReward1D look-back ring (entries) -- see carry layout below
    int tuned_general;   // 1: ring of the last 101 per-SUBSTEP norms (any substep count; slow path)
    AutoReset ar;
    int64_t map_envs;    // batch size the lane mapping is chosen for (the whole batch, also when step_host launches chunks)
};

// carry layout (doubles, per env):
//   [0]            sum_t |u_t[-1]|              (tuned_reward_1d.py:37: np.sum(abs(uVec[:, -1])))
//   [1 .. ring]    ring of step-end L2 norms    (tuned_reward_1d.py:40: norm(uVec[t-100]))
constexpr int kLookback = 100;
constexpr int kWarpMaxPoints = 512;      // register-resident warp kernels: 32 lanes x up to 16 points
#ifndef PDECG_BC1D_MINBLOCKS_F32
#define PDECG_BC1D_MINBLOCKS_F32 20  // float32 states: <= 102 registers, 20 warps/SM (measured Transport / RD ref: 12 -> 2.43 / 0.346, 16 -> 2.43 / 0.343, 20 -> 2.36 / 0.338, 24 -> 2.78 / 0.355 ms)
#endif
#ifndef PDECG_BC1D_MINBLOCKS
#define PDECG_BC1D_MINBLOCKS 12  // x PDECG_BC1D_WARPS (1) = 12 warps per SM, <= 170 registers: measured best for the fp64 two-envs-per-warp kernel
#endif

__host__ __device__ inline int ring_len_for(int substeps) {
    return (substeps >= 1 && substeps <= kLookback && kLookback % substeps == 0) ? kLookback / substeps : 1;
}
// TunedReward1D needs ||u[t-100]||.  Fast layouts exist when row t-100 is a step boundary (substeps | 100) or lies
// inside the current step (substeps >= 100) and episodes end on a step boundary; every other configuration keeps
// the norms of the last 101 substeps in the carry and recomputes the norm after EVERY substep (general, slow).
__host__ __device__ inline bool tuned_needs_general_ring(int substeps, int nt) {
    if (substeps < 1) return false;
    if (substeps < kLookback && kLookback % substeps != 0) return true;
    return (nt - 1) % substeps != 0;
}

// ---- per-point updates -----------------------------------------------------------------------
// ReactionDiffusion, parabolic.py:143-144:
//   n_j = (p_j + F*((p_{j-1} - 2 p_j) + p_{j+1})) + ((dt*beta_j) * p_j)          c = dt*beta_j (hoisted)
// FUSED2: (p_{j-1} - 2 p_j) as ONE fma(-2, p_j, p_{j-1}).  2*p_j is exact, so the fma rounds the same exact
// value as sub(p_{j-1}, mul(2, p_j)) -- bit-identical -- unless 2*p_j overflows.  The kernel enables it per env
// only over spans of substeps for which a rigorous growth bound keeps |u| far below the overflow threshold.
template <class M, bool FUSED2 = false>
__device__ __forceinline__ typename M::S rd_point(typename M::S pl, typename M::S p, typename M::S pr,
                                                   typename M::S F, typename M::C c) {
    using S = typename M::S;
    using C = typename M::C;
    if constexpr (M::kFast) {
        S lap = (pl - S(2) * p) + pr;
        return fmaf(c, p, fmaf(F, lap, p));
    } else {
        S lap0;
        if constexpr (FUSED2) {
            if constexpr (sizeof(S) == 8) lap0 = __fma_rn(S(-2), p, pl);
            else lap0 = __fmaf_rn(S(-2), p, pl);
        } else {
            lap0 = Strict<S>::sub(pl, Strict<S>::mul(S(2), p));
        }
        S lap = Strict<S>::add(lap0, pr);
        S diff = Strict<S>::add(p, Strict<S>::mul(F, lap));
        // f32 state with f64 beta: the reaction product and the final sum are evaluated in f64
        C react = Strict<C>::mul(c, static_cast<C>(p));
        return static_cast<S>(Strict<C>::add(static_cast<C>(diff), react));
    }
}

// Transport, hyperbolic.py:146-155:
//   n_j = p_j + dt*(((p_{j+1} - p_j)/dx) + (p_0 * beta_j))        dq = (p_{j+1} - p_j)/dx, divided beforehand
template <class M>
__device__ __forceinline__ typename M::S tr_point(typename M::S p, typename M::S dq, typename M::S p0,
                                                   typename M::C dt, typename M::C beta) {
    using S = typename M::S;
    using C = typename M::C;
    if constexpr (M::kFast) {
        return fmaf(dt, fmaf(p0, beta, dq), p);
    } else {
        C e = Strict<C>::mul(static_cast<C>(p0), beta);                  // promoted to beta's dtype
        C f = Strict<C>::add(static_cast<C>(dq), e);
        C g = Strict<C>::mul(dt, f);
        return static_cast<S>(Strict<C>::add(static_cast<C>(p), g));
    }
}

// ---- actuation -------------------------------------------------------------------------------
// Boundary value written at the actuated point, evaluated in numpy's promoted dtype:
//   Dirichlet : normalize(a)                      in the action dtype A
//   Neumann   : normalize(a*dx + neighbour)       a*dx in A, sum and normalize in promote(A, S)
// (hyperbolic.py:143-145 reads the not-yet-written new row -> neighbour == 0; parabolic.py:148-150
//  reads the previous row's u[nx-1].)
template <typename S>
__device__ __forceinline__ S actuate(const Bc1dArgs& a, int64_t env, S neighbour) {
    const pdecg_bc1d_cfg& c = a.cfg;
    const bool nz = c.normalize != 0;
    if (c.action_dtype == PDECG_F32) {
        float act = static_cast<const float*>(a.buf.action)[env];
        if (c.control_type == PDECG_CONTROL_DIRICHLET)
            return static_cast<S>(normalize_action<float>(act, nz, static_cast<float>(c.max_control_value)));
        float adx = __fmul_rn(act, static_cast<float>(c.dx));
        S w = Strict<S>::add(static_cast<S>(adx), neighbour);          // promote(f32, S) == S
        return normalize_action<S>(w, nz, static_cast<S>(c.max_control_value));
    } else {
        double act = static_cast<const double*>(a.buf.action)[env];
        if (c.control_type == PDECG_CONTROL_DIRICHLET)
            return static_cast<S>(normalize_action<double>(act, nz, c.max_control_value));
        double w = __dadd_rn(__dmul_rn(act, c.dx), static_cast<double>(neighbour));
        return static_cast<S>(normalize_action<double>(w, nz, c.max_control_value));
    }
}

// value of point `jq` of the row, broadcast to all lanes
template <typename S, int K, int LPE = 32>
__device__ __forceinline__ S row_value(const S (&u)[K], int jq, unsigned m = kFull) {
    int src = jq / K, slot = jq - src * K;
    S v = u[0];
#pragma unroll
    for (int k = 1; k < K; ++k) v = (slot == k) ? u[k] : v;
    return __shfl_sync(m, v, src, LPE);
}

// Kernel variants (uniform over the batch, chosen on the host):
//   kVarFast : boundary / padding points are "frozen slots" -- their per-slot coefficients are zero, so
//              the unmodified stencil leaves them unchanged (x + 0*finite == x).  The inner loop is then
//              pure stencil arithmetic + 2 halo shuffles: no per-point selects.  The actuated value is
//              written once, after the first substep (it is constant over a control step).  The halos of
//              the segment's two edge lanes are forced to 0 so no non-finite value can wrap around.
//   kVarSafe : boundary points overwritten with selects every substep (any npts <= 32*K).
//   kVarDyn  : kVarSafe + the actuated value recomputed every substep from u[nx-1]
//              (ReactionDiffusion with Neumann control, parabolic.py:148-150).
enum { kVarFast = 0, kVarSafe = 1, kVarDyn = 2 };

template <int ENV, class M, int K, int VAR, int LPE = 32>
struct Integrator {
    using S = typename M::S;
    using C = typename M::C;
    S u[K];
    C coef[K];      // RD: dt*beta_j ; Transport: beta_j      (zero on frozen slots in kVarFast)
    S fk[K];        // RD, kVarFast: F on interior points, 0 on frozen slots
    C dk[K];        // Transport, kVarFast: dt on updated points, 0 on frozen slots
    S g;            // value of the actuated point
    ConstDiv<S> by_dx;
    S Fs, dxs;
    C dtc;
    int lane, j0, jb;       // lane = index inside the env's LPE-lane segment
    unsigned mask;          // member mask of the segment
    double bsum_dyn;

    __device__ __forceinline__ explicit Integrator(S dx) : by_dx(dx) {}

    // WIDE: both envs of a half-warp pair run this loop in lock step (see `wide` in the kernel), so the shuffles
    // may name the whole warp with a compile-time mask -- a run-time member mask costs ~6 extra instructions
    // per substep (MATCH/REDUX/VOTE convergence check emitted by the compiler).
    template <bool FUSED2, bool WIDE>
    __device__ __forceinline__ void substep(const Bc1dArgs& a, int64_t env, const S (&p)[K], S (&n)[K]) {
        const unsigned mask = WIDE ? kFull : this->mask;
        if constexpr (ENV == PDECG_ENV_REACTION_DIFFUSION) {
            if constexpr (VAR == kVarDyn) {
                g = actuate<S>(a, env, row_value<S, K, LPE>(p, jb - 1, mask));
                bsum_dyn += fabs(static_cast<double>(g));
            }
            S hl = __shfl_up_sync(mask, p[K - 1], 1, LPE);
            S hr = __shfl_down_sync(mask, p[0], 1, LPE);
            if constexpr (VAR == kVarFast) {
                hl = (lane == 0) ? S(0) : hl;          // the segment's edge lanes see 0, never a wrapped value:
                hr = (lane == LPE - 1) ? S(0) : hr;    // frozen slots stay finite unless their real neighbour is not
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const S pl = (k == 0) ? hl : p[k > 0 ? k - 1 : 0];
                const S pr = (k == K - 1) ? hr : p[k < K - 1 ? k + 1 : K - 1];
                n[k] = rd_point<M, FUSED2>(pl, p[k], pr, (VAR == kVarFast) ? fk[k] : Fs, coef[k]);
            }
            if constexpr (VAR != kVarFast) {
                if (lane == 0) n[0] = S(0);            // parabolic.py:146
            }
        } else {
            const S p0 = __shfl_sync(mask, p[0], 0, LPE);
            S hr = __shfl_down_sync(mask, p[0], 1, LPE);
            if constexpr (VAR == kVarFast) hr = (lane == LPE - 1) ? S(0) : hr;   // never a wrapped value (see RD)
            // (p_{j+1} - p_j) / dx for the whole chunk: 3-instruction exact tail per point, ONE range test
            // per substep; operands outside the window (zero, denormal, inf, ...) redo the chunk exactly
            S x[K], dq[K];
            bool ok = true;
            [[maybe_unused]] float xmin = 0x1p100f, xmax = 0.0f;  // float states: min/max |x| (3-input FMNMX3), one test
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const S pr = (k == K - 1) ? hr : p[k < K - 1 ? k + 1 : K - 1];
                if constexpr (M::kFast) {
                    dq[k] = (pr - p[k]) * by_dx.y;
                } else {
                    x[k] = Strict<S>::sub(pr, p[k]);
                    dq[k] = by_dx.fast(x[k]);
                    if constexpr (sizeof(S) == 4) {
                        xmin = fminf(xmin, fabsf(x[k]));     // NaNs drop out here: the fast tail maps NaN -> NaN
                        xmax = fmaxf(xmax, fabsf(x[k]));     // (canonical) exactly like the division does
                    } else {
                        ok = ok & by_dx.in_range(x[k]);
                    }
                }
            }
            if constexpr (!M::kFast && sizeof(S) == 4) ok = by_dx.in_range_minmax(xmin, xmax);
            if constexpr (!M::kFast) {
                if (!ok) {
#pragma unroll
                    for (int k = 0; k < K; ++k) dq[k] = by_dx.exact(x[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) n[k] = tr_point<M>(p[k], dq[k], p0, (VAR == kVarFast) ? dk[k] : dtc, coef[k]);
        }
        if constexpr (VAR != kVarFast) {
#pragma unroll
            for (int k = 0; k < K; ++k) n[k] = (j0 + k == jb) ? g : n[k];
        }
    }

    // `cnt` substeps in place (register ping-pong: no copies in the steady state)
    template <bool FUSED2, bool WIDE>
    __device__ __forceinline__ void advance_t(const Bc1dArgs& a, int64_t env, int cnt) {
        S v[K];
        int i = 0;
        for (; i + 2 <= cnt; i += 2) {
            substep<FUSED2, WIDE>(a, env, u, v);
            substep<FUSED2, WIDE>(a, env, v, u);
        }
        if (i < cnt) {
            substep<FUSED2, WIDE>(a, env, u, v);
#pragma unroll
            for (int k = 0; k < K; ++k) u[k] = v[k];
        }
    }
    // float64 only: for float32 states the re-checks every ~70 substeps cost more than the saved FMUL (measured)
    static constexpr bool kCanFuse = (ENV == PDECG_ENV_REACTION_DIFFUSION) && !M::kFast && (sizeof(S) == 8);
    static constexpr bool kCanWide = (LPE == 16);   // LPE == 32: the member mask is already a constant
    bool fuse2;     // warp-uniform, set per env per control step (see rd_point)
    bool wide;      // warp-uniform: the warp's two envs take identical control flow this step
    template <bool WIDE>
    __device__ __forceinline__ void advance_w(const Bc1dArgs& a, int64_t env, int cnt) {
        if constexpr (kCanFuse) {
            if (fuse2) { advance_t<true, WIDE>(a, env, cnt); return; }
        }
        advance_t<false, WIDE>(a, env, cnt);
    }
    __device__ __forceinline__ void advance(const Bc1dArgs& a, int64_t env, int cnt) {
        if constexpr (kCanWide) {
            if (wide) { advance_w<true>(a, env, cnt); return; }
        }
        advance_w<false>(a, env, cnt);
    }

    // write the boundary values the reference assigns explicitly
    __device__ __forceinline__ void fix_boundaries() {
#pragma unroll
        for (int k = 0; k < K; ++k) u[k] = (j0 + k == jb) ? g : u[k];
        if constexpr (ENV == PDECG_ENV_REACTION_DIFFUSION) {
            if (lane == 0) u[0] = S(0);
        }
    }
};

// ---- (re)initialise one env: `reset` of hyperbolic.py:198-227 / parabolic.py:192-221 --------------------------
// Executed by the LPE lanes of the env's segment.  `keep_outputs`: leave reward / flags alone (fused auto-reset).
template <typename S, typename I, int LPE>
__device__ __forceinline__ void reset_env(const Bc1dArgs& a, const I* __restrict__ init, int64_t init_stride,
                                          const RandomIc& rnd, int64_t env, int lane, unsigned mask, bool keep_outputs) {
    const int npts = a.npts;
    S* urow = static_cast<S*>(a.buf.u) + env * npts;
    const I* irow = init + env * init_stride;
    double scale = 1.0;
    if (rnd.enabled) {
        const int ep = rnd.episode[env];
        const double u01 = Philox4x32::uniform53(rnd.seed, static_cast<uint64_t>(rnd.env_id_offset + env), static_cast<uint32_t>(ep));
        scale = __dadd_rn(rnd.lo, __dmul_rn(__dsub_rn(rnd.hi, rnd.lo), u01));
        __syncwarp(mask);
        if (lane == 0) rnd.episode[env] = ep + 1;
    }
    auto ic = [&](int j) -> S {                        // `u[0] = init_condition` casts to the state dtype
        return rnd.enabled ? static_cast<S>(__dmul_rn(scale, static_cast<double>(irow[j]))) : static_cast<S>(irow[j]);
    };
    double ss = 0.0;
    S last = S(0), first = S(0), second = S(0), blast = S(0);
    void* const crow = a.buf.obs_copy;
    const bool c32 = a.buf.obs_copy_f32 != 0;
    // ||u_0||^2 in the summation order of a 32-lane warp (32 strided partial sums, xor tree 16-8-4-2-1) whatever
    // LPE is: a half-warp segment carries two of the partial sums per lane, so both reset paths agree bit for bit
    static_assert(LPE == 32 || LPE == 16, "reset_env: 32 or 16 lanes per env");
#pragma unroll
    for (int r = 0; r < 32 / LPE; ++r) {
        double part = 0.0;
        for (int j = lane + r * LPE; j < npts; j += 32) {
            S v = ic(j);
            urow[j] = v;
            if (crow) put_obs<S>(crow, c32, env * npts + j, v);
            double dv = static_cast<double>(v);
            part += dv * dv;
        }
        ss = (r == 0) ? part : ss + part;
    }
    ss = seg_sum<LPE>(ss, mask);
    if (npts >= 1) last = ic(npts - 1);
    if (npts >= 2) blast = ic(npts - 2);
    first = ic(0);
    if (npts >= 2) second = ic(1);
    if (lane == 0) {
        double* carry = a.buf.carry + env * a.carry_len;
        carry[0] = fabs(static_cast<double>(last));    // row 0 of uVec[:, -1]
        for (int i = 1; i < a.carry_len; ++i) carry[i] = 0.0;
        carry[1] = static_cast<double>(static_cast<S>(sqrt(ss)));   // ||u_0||: ring slot of step 0
        a.buf.time_index[env] = 0;
        if (!keep_outputs) {
            a.buf.reward[env] = 0.0;
            a.buf.terminated[env] = 0;
            a.buf.truncated[env] = 0;
        }
        const pdecg_bc1d_cfg& c = a.cfg;
        if (c.sensing_loc != PDECG_SENSE_FULL) {
            const S dxs = static_cast<S>(c.dx);
            S o;
            if (c.sensing_loc == PDECG_SENSE_COLLOCATED)
                o = (c.control_type == PDECG_CONTROL_NEUMANN) ? last : Strict<S>::div(Strict<S>::sub(last, blast), dxs);
            else
                o = (c.sensing_type == PDECG_SENSETYPE_NEUMANN) ? Strict<S>::div(Strict<S>::sub(second, first), dxs) : first;
            static_cast<S*>(a.buf.obs)[env] = o;
        }
    }
}

// fused same-step auto-reset: out of line so the (rare) restart does not weigh on the step kernel's registers
template <typename S, int LPE>
__device__ __noinline__ void autoreset_env(const Bc1dArgs& a, int64_t env, int lane, unsigned mask) {
    __syncwarp(mask);                                  // the step's own row / carry stores come first
    if (a.ar.init_dtype == PDECG_F32)
        reset_env<S, float, LPE>(a, static_cast<const float*>(a.ar.init), a.ar.init_stride, a.ar.rnd, env, lane, mask, true);
    else
        reset_env<S, double, LPE>(a, static_cast<const double*>(a.ar.init), a.ar.init_stride, a.ar.rnd, env, lane, mask, true);
}

// Warps of a block never cooperate (one env or env pair per warp, private smem slice), so the block size only
// sets the scheduling granularity of the tail wave (measured on the headline config: 4 warps 0.5333 ms,
// 2 warps 0.5327 ms, 1 warp 0.5304 ms per launch, always 12 warps resident per SM).
#ifndef PDECG_BC1D_WARPS
#define PDECG_BC1D_WARPS 1
#endif
constexpr int kStepWarps = PDECG_BC1D_WARPS;
constexpr int kStepThreads = kStepWarps * 32;

template <int ENV, class M, int K, int VAR, int LPE = 32>
__global__ void __launch_bounds__(kStepThreads, sizeof(typename M::S) == 8 ? PDECG_BC1D_MINBLOCKS : PDECG_BC1D_MINBLOCKS_F32)
bc1d_step_kernel(const __grid_constant__ Bc1dArgs a) {
    using S = typename M::S;
    using C = typename M::C;
    constexpr int EPW = 32 / LPE;                      // envs per warp
    __shared__ S sm_all[kStepWarps * EPW][LPE * K];
    const int lane = threadIdx.x & (LPE - 1);          // lane inside the env's segment
    const int seg = (threadIdx.x >> 5) * EPW + ((threadIdx.x & 31) / LPE);
    const unsigned mask = seg_mask<LPE>(threadIdx.x & 31);
    const int64_t env = static_cast<int64_t>(blockIdx.x) * (kStepWarps * EPW) + seg;
    if (env >= a.buf.n_envs) return;
    S* sm = sm_all[seg];
    const pdecg_bc1d_cfg& cfg = a.cfg;
    const int npts = a.npts;
    const int j0 = lane * K;

    Integrator<ENV, M, K, VAR, LPE> I{static_cast<S>(cfg.dx)};
    I.lane = lane;
    I.mask = mask;
    I.j0 = j0;
    I.jb = npts - 1;                               // actuated point (x = X; ghost point for RD)
    I.Fs = static_cast<S>(cfg.dt / (cfg.dx * cfg.dx));     // parabolic.py:138 (python float -> S)
    I.dxs = static_cast<S>(cfg.dx);
    I.dtc = static_cast<C>(cfg.dt);
    I.bsum_dyn = 0.0;

    S* urow = static_cast<S*>(a.buf.u) + env * npts;
    load_row_chunk<S, S, K, LPE>(I.u, urow, npts, sm, lane, mask);
    if constexpr (ENV == PDECG_ENV_TRANSPORT) {
        // padding slots (never stored, never read by valid points) hold a ramp instead of zeros so that the
        // differences (u[j+1]-u[j]) formed there are ordinary numbers and stay on the division fast path
#pragma unroll
        for (int k = 0; k < K; ++k) I.u[k] = (j0 + k >= npts) ? static_cast<S>(j0 + k - npts + 1) : I.u[k];
    }

    // per-slot coefficient: RD -> c_j = dt*beta_j (parabolic.py:144, left-assoc `dt*beta*u`);
    //                       Transport -> beta_j
    {
        const C* brow = static_cast<const C*>(a.buf.beta) + env * a.buf.beta_env_stride;
        if constexpr (sizeof(C) <= sizeof(S)) {
            load_row_chunk<C, C, K, LPE>(I.coef, brow, npts, reinterpret_cast<C*>(sm), lane, mask);
        } else {
            // f64 beta with f32 state: direct (L1-served) chunk loads; once per control step
#pragma unroll
            for (int k = 0; k < K; ++k) I.coef[k] = (j0 + k < npts) ? brow[j0 + k] : C(0);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = j0 + k;
            if constexpr (ENV == PDECG_ENV_REACTION_DIFFUSION) {
                I.coef[k] = M::kFast ? I.dtc * I.coef[k] : Strict<C>::mul(I.dtc, I.coef[k]);
                const bool interior = (j >= 1) && (j < npts - 1);
                I.fk[k] = interior ? I.Fs : S(0);
                I.dk[k] = C(0);
                if (VAR == kVarFast && !interior) I.coef[k] = C(0);
            } else {
                const bool updated = j < npts - 1;          // hyperbolic.py:146: u[0 : Nx-1]
                I.dk[k] = updated ? I.dtc : C(0);
                I.fk[k] = S(0);
            }
        }
    }

    const int t0 = a.buf.time_index[env];
    int nsub = cfg.nt - 1 - t0;
    nsub = nsub < 0 ? 0 : (nsub > cfg.substeps ? cfg.substeps : nsub);
    const int t_end = t0 + nsub;

    // Half-warp pairs: when both envs exist and sit at the same time index, everything that steers the substep
    // loop below is identical for the two segments (the configuration is uniform over the batch and the fused-form
    // spans are then evaluated over the whole warp), so the loop runs in lock step and may use full-warp masks.
    I.wide = false;
    if constexpr (Integrator<ENV, M, K, VAR, LPE>::kCanWide) {
        if ((env | 1) < a.buf.n_envs)                 // warp-uniform: both envs of this warp exist -> 32 lanes here
            I.wide = __all_sync(kFull, t0 == __shfl_xor_sync(kFull, t0, 16));
    }

    double* carry = a.buf.carry + env * a.carry_len;
    double bsum = carry[0];

    I.g = actuate<S>(a, env, S(0));               // constant over the step unless kVarDyn
    I.fuse2 = false;
    // |n_j| <= (1 + 4F + max|c|) * max(|p|, |g|) =: rho * m, so after s substeps |u| <= m * rho^s.  A span of substeps
    // may use the fused form while m * rho^s stays below 2^kFuseLimit (2*u_j can then never overflow).
    constexpr double kFuseLimit = (sizeof(S) == 8) ? 1000.0 : 100.0;
    double fuse_log2_rho = 1.0;
    int fuse_left = 0;                             // substeps still covered by the current bound
    if constexpr (Integrator<ENV, M, K, VAR, LPE>::kCanFuse) {
        double cm = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) cm = fmax(cm, fabs(static_cast<double>(I.coef[k])));
        // lock-stepped half-warp pairs (`wide`) must take identical spans / fused-or-strict decisions: one (conservative)
        // bound for both envs, also when their per-env coefficient rows put rho in different binades
        cm = I.wide ? seg_max<32>(cm, kFull) : seg_max<LPE>(cm, mask);
        // upper bound of log2(rho) from the exponent field (no log2() call): rho in [2^e, 2^(e+1)) -> e + 1
        const double rho = 1.0 + 4.0 * fabs(static_cast<double>(I.Fs)) + cm;
        fuse_log2_rho = static_cast<double>(static_cast<int>((static_cast<unsigned>(__double2hiint(rho)) >> 20) & 0x7ffu) - 1022);
    }
    auto fuse_span = [&]() -> int {                // substeps for which the fused form is provably exact from here
        double m = fabs(static_cast<double>(I.g));
        bool finite = true;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double v = fabs(static_cast<double>(I.u[k]));
            m = fmax(m, v);                        // fmax drops NaN -> caught by `finite`
            finite = finite & (v < 1.0e300);
        }
        if (I.wide) {                               // one (conservative) bound for both envs keeps them in lock step
            m = seg_max<32>(m, kFull);
            finite = __all_sync(kFull, finite) && (m < 1.0e300);
        } else {
            m = seg_max<LPE>(m, mask);
            finite = __all_sync(mask, finite) && (m < 1.0e300);
        }
        if (!finite) return 0;
        // log2(max(m, 1)) < (exponent of max(m, 1)) + 1, again from the exponent field
        const int e_m = static_cast<int>((static_cast<unsigned>(__double2hiint(fmax(m, 1.0))) >> 20) & 0x7ffu) - 1022;
        const double span = (kFuseLimit - 1.0 - static_cast<double>(e_m)) / fuse_log2_rho;
        return span > 1.0e6 ? 1000000 : (span < 0.0 ? 0 : static_cast<int>(span));
    };

    // reward bookkeeping that needs mid-step rows (uniform over the warp)
    const bool tuned = cfg.reward_kind == PDECG_REWARD_TUNED1D;
    const bool ring_every = tuned && a.tuned_general;
    const int snap_at = (tuned && !ring_every && cfg.substeps > kLookback) ? nsub - kLookback : -1;   // substeps done when row t-100 is live
    const bool want_diff = !tuned && cfg.horizon == PDECG_HORIZON_DIFFERENTIAL && nsub > 0;
    const bool want_hor = !tuned && cfg.horizon == PDECG_HORIZON_T;
    const int hor_m = t_end > cfg.t_horizon_length ? cfg.t_horizon_length : t_end;   // norm_reward.py:65-72
    const int hor_from = nsub - hor_m + 1;          // rows after `done` substeps, done in [hor_from, nsub]
    double lb_norm = 0.0;                          // ||u_{t-100}||_2
    double diff_norm = 0.0;                        // ||u_t - u_{t-1}||_p   (NormReward differential)
    double hor_acc = 0.0;                          // sum of the last L row norms (NormReward t-horizon)
    S prev[K];
#pragma unroll
    for (int k = 0; k < K; ++k) prev[k] = S(0);

    int done = 0;
    while (true) {
        // ---- events at the current row (after `done` substeps) ----
        if (VAR == kVarFast && done == 1) I.fix_boundaries();
        if (done == snap_at) lb_norm = sqrt(chunk_sumsq<S, K, LPE>(I.u, j0, npts, mask));
        if (ring_every && done >= 1) {             // general look-back: norm of every row, ring indexed by t mod 101
            const double nrm = static_cast<double>(static_cast<S>(sqrt(chunk_sumsq<S, K, LPE>(I.u, j0, npts, mask))));
            if (lane == 0) carry[1 + (t0 + done) % (kLookback + 1)] = nrm;
        }
        if (want_hor && done >= hor_from && done >= 1)
            hor_acc += static_cast<double>(static_cast<S>(chunk_norm<S, K, LPE>(I.u, j0, npts, cfg.norm_ord, mask)));
        if (want_diff && done == nsub) {
            S d[K];
#pragma unroll
            for (int k = 0; k < K; ++k) d[k] = Strict<S>::sub(I.u[k], prev[k]);
            diff_norm = chunk_norm<S, K, LPE>(d, j0, npts, cfg.norm_ord, mask);
        }
        if (done >= nsub) break;
        if (want_diff && done == nsub - 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) prev[k] = I.u[k];
        }
        // ---- next stop ----
        int next = nsub;
        if (VAR == kVarFast && done == 0) next = 1;
        if (snap_at > done && snap_at < next) next = snap_at;
        if (ring_every) next = done + 1;
        if (want_diff && nsub - 1 > done && nsub - 1 < next) next = nsub - 1;
        if (want_hor) {
            int h = hor_from > done + 1 ? hor_from : done + 1;
            if (h < next) next = h;
        }
        if constexpr (Integrator<ENV, M, K, VAR, LPE>::kCanFuse) {
            if (VAR != kVarDyn) {
                if (fuse_left <= 0) {
                    const int span = fuse_span();
                    I.fuse2 = span >= 8;
                    fuse_left = I.fuse2 ? span : 16;      // too close to the limit: 16 strict substeps, then re-check
                }
                if (done + fuse_left < next) next = done + fuse_left;
                fuse_left -= next - done;
            }
        }
        I.advance(a, env, next - done);
        done = next;
    }
    if constexpr (VAR == kVarDyn) bsum += I.bsum_dyn;
    else bsum += static_cast<double>(nsub) * fabs(static_cast<double>(I.g));

    // ---- write back state, bookkeeping ---------------------------------------------------------
    if (nsub > 0) {
        if constexpr (VAR == kVarFast) I.fix_boundaries();   // frozen slots may have absorbed a NaN/inf neighbour
        store_row_chunk<S, K, LPE>(I.u, urow, npts, sm, lane, mask);
    }
    auto store_obs_row = [&](void* base) {                   // [n_envs, npts] in the state dtype, or float32 on request
        if (sizeof(S) == 8 && a.buf.obs_copy_f32)
            store_row_chunk_as<float, S, K, LPE>(I.u, static_cast<float*>(base) + env * npts, npts, sm, lane, mask);
        else
            store_row_chunk<S, K, LPE>(I.u, static_cast<S*>(base) + env * npts, npts, sm, lane, mask);
    };
    if (a.buf.obs_copy != nullptr) store_obs_row(a.buf.obs_copy);   // private observation for the caller (no separate device copy)
    S (&u)[K] = I.u;

    const double sumsq = chunk_sumsq<S, K, LPE>(u, j0, npts, mask);
    const S norm2 = static_cast<S>(sqrt(sumsq));               // np.linalg.norm(u[t]) in the state dtype
    const bool terminated = t_end >= cfg.nt - 1;                // hyperbolic.py:171-180
    // per-env blow-up flag (no counterpart in the reference, where inf / NaN simply propagate): the row holds an inf / NaN, or
    // values so large that their squares overflow float64; optionally folded into `truncated`
    const bool nonfinite = !(sumsq <= 1.7976931348623157e308);
    const bool truncated = (cfg.limit_pde_state_size && (norm2 >= static_cast<S>(cfg.max_state_value))) || (cfg.truncate_nonfinite && nonfinite);
    const S dxs = I.dxs;

    double reward;
    if (tuned) {
        // look-back norm from the ring when the row t-100 sits on an earlier step boundary
        if (ring_every) {
            __syncwarp(mask);
            lb_norm = (t_end - kLookback >= 0) ? carry[1 + (t_end - kLookback) % (kLookback + 1)] : 0.0;
        } else if (cfg.substeps <= kLookback && cfg.substeps >= 1) {
            const int slot = 1 + ((t_end / cfg.substeps) % a.ring_len);
            lb_norm = carry[slot];
            __syncwarp(mask);
            if (lane == 0 && nsub > 0) carry[slot] = static_cast<double>(norm2);
        }
        if (terminated && norm2 < S(20)) {
            // tuned_reward_1d.py:36-37, evaluated in the state dtype like numpy (python floats are weak)
            S acc = static_cast<S>(bsum);
            S r = Strict<S>::sub(Strict<S>::sub(static_cast<S>(cfg.terminate_reward), Strict<S>::div(acc, S(1000))), norm2);
            reward = static_cast<double>(r);
        } else if (truncated) {
            reward = cfg.truncate_penalty * static_cast<double>(cfg.reward_nt - t_end);
        } else {
            reward = static_cast<double>(Strict<S>::sub(static_cast<S>(lb_norm), norm2));
        }
    } else {
        const S coeff = static_cast<S>(cfg.norm_coeff);
        if (terminated) {
            reward = cfg.terminate_reward;
        } else if (truncated) {
            reward = cfg.truncate_penalty * static_cast<double>(cfg.reward_nt - t_end);
        } else if (cfg.horizon == PDECG_HORIZON_TEMPORAL || nsub == 0) {
            S nn = (cfg.norm_ord == kNormL2) ? norm2 : static_cast<S>(chunk_norm<S, K, LPE>(u, j0, npts, cfg.norm_ord, mask));
            reward = static_cast<double>(Strict<S>::div(-nn, coeff));
        } else if (cfg.horizon == PDECG_HORIZON_DIFFERENTIAL) {
            reward = static_cast<double>(Strict<S>::div(static_cast<S>(diff_norm), coeff));
        } else {
            // norm_reward.py:63-73: python-level accumulation of S-typed norms
            S mean = Strict<S>::div(static_cast<S>(hor_acc), static_cast<S>(hor_m));
            reward = static_cast<double>(Strict<S>::div(-mean, coeff));
        }
    }

    // ---- sensing (hyperbolic.py:69-120, parabolic.py:69-118) ------------------------------------
    S o = S(0);
    if (cfg.sensing_loc != PDECG_SENSE_FULL) {
        if (cfg.sensing_loc == PDECG_SENSE_COLLOCATED) {
            const S ub = row_value<S, K, LPE>(u, npts - 1, mask);
            if (cfg.control_type == PDECG_CONTROL_NEUMANN) o = ub;
            else o = Strict<S>::div(Strict<S>::sub(ub, row_value<S, K, LPE>(u, npts - 2, mask)), dxs);
        } else {
            const S u0 = row_value<S, K, LPE>(u, 0, mask);
            if (cfg.sensing_type == PDECG_SENSETYPE_NEUMANN) o = Strict<S>::div(Strict<S>::sub(row_value<S, K, LPE>(u, 1, mask), u0), dxs);
            else o = u0;
        }
        if (lane == 0) static_cast<S*>(a.buf.obs)[env] = o;
    }

    if (lane == 0) {
        carry[0] = bsum;
        a.buf.time_index[env] = t_end;
        a.buf.reward[env] = reward;
        a.buf.terminated[env] = terminated ? 1 : 0;
        a.buf.truncated[env] = truncated ? 1 : 0;
        if (a.buf.nonfinite != nullptr) a.buf.nonfinite[env] = nonfinite ? 1 : 0;
    }
    if (a.ar.enabled && (terminated || truncated)) {
        // the terminal observation survives the restart in the caller's final_obs buffer (what SB3 / gymnasium call
        // `terminal_observation` / `final_observation`; transport1Dppo.py:88-90 bootstraps from it on truncation)
        if (a.ar.final_obs != nullptr) {
            if (cfg.sensing_loc == PDECG_SENSE_FULL) store_obs_row(a.ar.final_obs);
            else if (lane == 0) put_obs<S>(a.ar.final_obs, a.buf.obs_copy_f32 != 0, env, o);
        }
        autoreset_env<S, LPE>(a, env, lane, mask);
    }
}

// Lanes per env, points per lane and kernel variant for this configuration (uniform over the batch).
// Two envs per warp (LPE = 16) are used for the reference's default grids when that wastes fewer lane-slots:
// 201 points (RD) -> 16 x 13 = 208 slots instead of 32 x 7 = 224; 100 points (Transport) -> 16 x 7 = 112
// instead of 32 x 4 = 128.
// Small batches are latency-bound (one control step is a chain of `substeps` dependent stencil updates per warp), so
// they use one env per warp with the fewest points per lane: TransportPDE1D x 1 env 0.089 -> 0.059 ms per step, x 256
// envs 0.078 -> 0.060 ms; from ~1,500 envs on the two-envs-per-warp mapping wins (2,048 envs: 0.108 vs 0.130 ms).  The
// choice depends only on the size of the whole batch (`map_envs`), never on how a host-buffer step is chunked, so
// step() and step_host() reduce the reward norms in the same order.
inline int64_t halfwarp_min_envs() {
    const char* e = getenv("PDECG_BC1D_HALFWARP_MIN");
    return e ? atoll(e) : 1536;
}
inline void pick_variant(const Bc1dArgs& a, int* LPE, int* K, int* var) {
    const bool dyn = a.cfg.env == PDECG_ENV_REACTION_DIFFUSION && a.cfg.control_type == PDECG_CONTROL_NEUMANN;
    const int k32 = (a.npts + 31) / 32, k16 = (a.npts + 15) / 16;
    const bool half = !dyn && (k16 == 7 || k16 == 13) && (k16 < 2 * k32) && a.map_envs >= halfwarp_min_envs() && !getenv("PDECG_BC1D_NO_HALFWARP");
    *LPE = half ? 16 : 32;
    *K = half ? k16 : (k32 <= 8 ? k32 : ((k32 + 1) / 2) * 2);     // 257..512 points: even K in {10, 12, 14, 16}
    // the frozen-slot variant zeroes the halos of the segment's edge lanes, so it is valid for every npts <= LPE*K
    *var = dyn ? kVarDyn : (getenv("PDECG_BC1D_SAFE") ? kVarSafe : kVarFast);
}

template <int ENV, class M, int VAR>
inline int launch_k(const Bc1dArgs& a, int LPE, int K, cudaStream_t st) {
    const int64_t per_block = (LPE == 16) ? 2 * kStepWarps : kStepWarps;
    const unsigned grid = static_cast<unsigned>((a.buf.n_envs + per_block - 1) / per_block);
    if (LPE == 16) {
        if constexpr (VAR == kVarFast) {
            if (K == 7) bc1d_step_kernel<ENV, M, 7, VAR, 16><<<grid, kStepThreads, 0, st>>>(a);
            else if (K == 13) bc1d_step_kernel<ENV, M, 13, VAR, 16><<<grid, kStepThreads, 0, st>>>(a);
            else return fail(PDECG_ERR_UNSUPPORTED, "unsupported half-warp points-per-lane %d", K);
        } else {
            return fail(PDECG_ERR_UNSUPPORTED, "half-warp mapping only exists for the fast variant");
        }
    } else {
        switch (K) {
#define PDECG_CASE(KK) case KK: bc1d_step_kernel<ENV, M, KK, VAR, 32><<<grid, kStepThreads, 0, st>>>(a); break;
            PDECG_CASE(1) PDECG_CASE(2) PDECG_CASE(3) PDECG_CASE(4) PDECG_CASE(5) PDECG_CASE(6) PDECG_CASE(7) PDECG_CASE(8)
            PDECG_CASE(10) PDECG_CASE(12) PDECG_CASE(14) PDECG_CASE(16)
#undef PDECG_CASE
            default: return fail(PDECG_ERR_UNSUPPORTED, "unsupported points-per-lane %d", K);
        }
    }
    count_launch();
    return check_launch("bc1d_step_kernel");
}

template <int ENV, class M>
inline int launch_cta(const Bc1dArgs& a, cudaStream_t st);       // bc1d_large.cuh

template <int ENV, class M>
inline int launch_var(const Bc1dArgs& a, cudaStream_t st) {
    if (a.npts > kWarpMaxPoints) return launch_cta<ENV, M>(a, st);   // beyond the register-resident limit: CTA per env
    int LPE, K, var;
    pick_variant(a, &LPE, &K, &var);
    if (var == kVarFast) return launch_k<ENV, M, kVarFast>(a, LPE, K, st);
    if constexpr (ENV == PDECG_ENV_REACTION_DIFFUSION) {
        if (var == kVarDyn) return launch_k<ENV, M, kVarDyn>(a, LPE, K, st);
    }
    return launch_k<ENV, M, kVarSafe>(a, LPE, K, st);
}

// one function per precision mode, defined in bc1d_inst_<mode>.cu
int launch_bc1d_refb32(const Bc1dArgs& a, cudaStream_t st);
int launch_bc1d_refb64(const Bc1dArgs& a, cudaStream_t st);
int launch_bc1d_fp64(const Bc1dArgs& a, cudaStream_t st);
int launch_bc1d_fp32(const Bc1dArgs& a, cudaStream_t st);

}  // namespace pdecg

#include "bc1d_large.cuh"
