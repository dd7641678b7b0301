// traffic_arz.cu -- TrafficPDE1D (Aw-Rascle-Zhang) step / reset kernels + C ABI (sm_100a; float64, opt-in float32).
//
// Reference behaviour restated here (no code shared): traffic_arz_env.py:23-280 and
// rewards/traffic_arz_reward.py:12-22 of lukebhan/PDEControlGym.
//
// One warp = one freeway.  (rho, y) rows live in registers; per substep the Richtmyer two-step
// Lax-Wendroff update needs the right neighbour's node state/fluxes (shuffle down) and the left
// neighbour's midpoint fluxes (shuffle up).  Node and midpoint fluxes are computed once and shared
// between the two cells that use them -- bit-identical to the reference, which evaluates them twice.
#include "pdecg_internal.h"

#ifndef PDECG_ARZ_MINBLOCKS
#define PDECG_ARZ_MINBLOCKS 8    // 64 registers, 32 warps/SM (measured at 262k envs: 10 -> 0.187, 9 -> 0.185, 8 -> 0.178, 7 -> 0.180, 6 -> 0.189 ms)
#endif
#ifndef PDECG_ARZ_MINBLOCKS_H
#define PDECG_ARZ_MINBLOCKS_H 4  // two envs per warp, up to 4 points per lane
#endif
namespace pdecg {

struct ArzArgs {
    pdecg_arz_cfg cfg;
    pdecg_arz_buffers buf;
    const double* ar_profile;   // pdecg_arz_step_autoreset: finished envs restart inside the step launch (nullptr: off)
};

// T = double: the reference's arithmetic (numpy float64, operation order kept, never contracted) -- bit-identical states.
// T = float : the opt-in fast mode (PDECG_PREC_FP32): float32 state, observation and arithmetic, same operation order.
template <typename T>
struct ArzCtx {      // vm and the hoisted exact division by the constant rm
    T vm;
    ConstDiv<T> by_rm;
    __device__ __forceinline__ explicit ArzCtx(const pdecg_arz_cfg& c) : vm(static_cast<T>(c.vm)), by_rm(static_cast<T>(c.rm)) {}
};
template <typename T>
__device__ __forceinline__ T veq(const ArzCtx<T>& c, T rho) {          // :271-272
    return Strict<T>::mul(c.vm, Strict<T>::sub(T(1), c.by_rm(rho)));
}
template <typename T>
__device__ __forceinline__ T flux_r(const ArzCtx<T>& c, T rho, T y) {   // :275-276
    return Strict<T>::add(y, Strict<T>::mul(rho, veq(c, rho)));
}
template <typename T>
__device__ __forceinline__ T flux_y(const ArzCtx<T>& c, T rho, T y) {   // :279-280
    return Strict<T>::mul(y, Strict<T>::add(Strict<T>::div(y, rho), veq(c, rho)));
}
__device__ __forceinline__ double clipd(double v, double lo, double hi) {            // np.clip keeps NaN
    v = (v < lo) ? lo : v;
    return (v > hi) ? hi : v;
}

// Row <-> register-chunk transfers (`lane` = the lane's position in its env's segment).  With K <= 2 points per lane,
// and in the two-envs-per-warp mapping, the chunk is read / written directly (lane L touches
// the 8K bytes at offset 8KL: every 32-byte sector is consumed by at most two adjacent lanes, and the second access
// of a line hits L1), which drops the shared-memory staging and its two warp barriers per row; wider chunks go through
// the staged, fully coalesced path.
template <int K, typename T, int LPE = 32>
__device__ __forceinline__ void arz_load_issue(T (&tmp)[K], const T* __restrict__ g, int npts, int lane) {
    if constexpr (K <= 2 || LPE != 32) {
#pragma unroll
        for (int k = 0; k < K; ++k) tmp[k] = (lane * K + k < npts) ? g[lane * K + k] : T(0);
    } else {
        load_row_issue<T, T, K>(tmp, g, npts, lane);
    }
}
template <int K, typename T, int LPE = 32>
__device__ __forceinline__ void arz_load_finish(T (&r)[K], const T (&tmp)[K], T* sm, int lane) {
    if constexpr (K <= 2 || LPE != 32) {
#pragma unroll
        for (int k = 0; k < K; ++k) r[k] = tmp[k];
    } else {
        load_row_finish<T, K>(r, tmp, sm, lane);
    }
}
template <int K, typename T, int LPE = 32>
__device__ __forceinline__ void arz_store(const T (&r)[K], T* __restrict__ g, int npts, T* sm, int lane) {
    if constexpr (K <= 2 || LPE != 32) {
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (lane * K + k < npts) g[lane * K + k] = r[k];
    } else {
        store_row_chunk<T, K>(r, g, npts, sm, lane);
    }
}

template <typename T, int LPE>
__device__ __noinline__ void arz_autoreset_env(const ArzArgs& a, int64_t env, int l, unsigned segm);   // defined after the reset kernel

template <typename T>
struct ArzCoef { T c1, c2, c3, c4, q_in, q_out; };

// One Richtmyer / Lax-Wendroff substep (:175-222) on the warp's register-resident (rho, y) row.
//
// FAST = false: every division is the IEEE one (ConstDiv's per-call window test for /rm, the compiler's div.rn with its
// own range test and slow path for y/rho) -- valid for every input.
// FAST = true : the same operations in the same order with the range tests taken out of the dependency chains: /rm is
// ConstDiv's 3-instruction tail, y/rho is fast_div's straight-line sequence, Veq(rho) of a node is evaluated once and
// shared by the boundary row and the node fluxes, and the boundary rows are selects instead of divergent branches.
// Every division operand is folded into ONE predicate (DivWindow); the function returns it, and the caller redoes the
// substep with FAST = false if any lane of the warp saw an operand outside the window -- so the states stay
// bit-identical to the reference's for every input, at ~2/3 of the instructions on the ordinary path.
template <int K, typename T, bool FAST, int LPE>
__device__ __forceinline__ bool arz_substep(T (&r)[K], T (&y)[K], const ArzCtx<T>& ctx, const ArzCoef<T>& co, int j0, int M) {
    using D = Strict<T>;
    using W = DivWindow<T>;
    const T c1 = co.c1, c2 = co.c2, c3 = co.c3, c4 = co.c4;
    bool ok = true;
    T fr[K], fy[K];
    if constexpr (FAST) {
        // ---- boundary densities (:175-190): rho[0] = rho[1], rho[M-1] = rho[M-2] ----
        const T r_left = __shfl_up_sync(kFull, r[K - 1], 1, LPE);                     // node j0-1
        if constexpr (K == 1) {
            const T r_right = __shfl_down_sync(kFull, r[0], 1, LPE);
            r[0] = (j0 == 0) ? r_right : r[0];
        } else {
            r[0] = (j0 == 0) ? r[1] : r[0];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const T rp = (k > 0) ? r[k > 0 ? k - 1 : 0] : r_left;                // r[j-1] (M >= 4: never the node just rewritten)
            r[k] = (j0 + k == M - 1) ? rp : r[k];
        }
        // ---- Veq at the nodes, boundary momenta, node fluxes ----
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = j0 + k;
            const T ve = D::mul(ctx.vm, D::sub(T(1), ctx.by_rm.fast(r[k])));
            const bool first = (k == 0) && (j == 0), last = (j == M - 1);
            const T yb = D::sub(first ? co.q_in : co.q_out, D::mul(r[k], ve));
            y[k] = (first || last) ? yb : y[k];
            ok = ok & W::b(r[k]) & W::a(y[k]);
            fr[k] = D::add(y[k], D::mul(r[k], ve));
            fy[k] = D::mul(y[k], D::add(fast_div(y[k], r[k]), ve));
        }
    } else {
        const T r_right = __shfl_down_sync(kFull, r[0], 1, LPE);                      // node j0+K
        const T r_left = __shfl_up_sync(kFull, r[K - 1], 1, LPE);                     // node j0-1
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = j0 + k;
            const T rn = (k < K - 1) ? r[k < K - 1 ? k + 1 : 0] : r_right;       // r[j+1]
            const T rp = (k > 0) ? r[k > 0 ? k - 1 : 0] : r_left;                // r[j-1]
            if (j == 0) {
                r[k] = rn;
                y[k] = D::sub(co.q_in, D::mul(r[k], veq(ctx, r[k])));
            }
            if (j == M - 1) {
                r[k] = rp;
                y[k] = D::sub(co.q_out, D::mul(r[k], veq(ctx, r[k])));
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            fr[k] = flux_r(ctx, r[k], y[k]);
            fy[k] = flux_y(ctx, r[k], y[k]);
        }
    }
    const T r_nx = __shfl_down_sync(kFull, r[0], 1, LPE), y_nx = __shfl_down_sync(kFull, y[0], 1, LPE);
    const T fr_nx = __shfl_down_sync(kFull, fr[0], 1, LPE), fy_nx = __shfl_down_sync(kFull, fy[0], 1, LPE);
    // ---- midpoints j+1/2 (:202-215) and their fluxes ----
    T ym[K], frm[K], fym[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const bool lastk = (k == K - 1);
        const T rb = lastk ? r_nx : r[lastk ? 0 : k + 1], yb = lastk ? y_nx : y[lastk ? 0 : k + 1];
        const T frb = lastk ? fr_nx : fr[lastk ? 0 : k + 1], fyb = lastk ? fy_nx : fy[lastk ? 0 : k + 1];
        const T rmid = D::sub(D::mul(T(0.5), D::add(rb, r[k])), D::mul(c1, D::sub(frb, fr[k])));
        const T ysum = D::add(yb, y[k]);
        const T ymid = D::sub(D::sub(D::mul(T(0.5), ysum), D::mul(c1, D::sub(fyb, fy[k]))), D::mul(c2, ysum));
        ym[k] = ymid;
        if constexpr (FAST) {
            const T ve = D::mul(ctx.vm, D::sub(T(1), ctx.by_rm.fast(rmid)));
            ok = ok & W::b(rmid) & W::a(ymid);
            frm[k] = D::add(ymid, D::mul(rmid, ve));
            fym[k] = D::mul(ymid, D::add(fast_div(ymid, rmid), ve));
        } else {
            frm[k] = flux_r(ctx, rmid, ymid);
            fym[k] = flux_y(ctx, rmid, ymid);
        }
    }
    const T ym_pv = __shfl_up_sync(kFull, ym[K - 1], 1, LPE);
    const T frm_pv = __shfl_up_sync(kFull, frm[K - 1], 1, LPE);
    const T fym_pv = __shfl_up_sync(kFull, fym[K - 1], 1, LPE);
    // ---- interior update (:218-222) ----
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = j0 + k;
        const T yl = (k == 0) ? ym_pv : ym[k > 0 ? k - 1 : 0];
        const T frl = (k == 0) ? frm_pv : frm[k > 0 ? k - 1 : 0];
        const T fyl = (k == 0) ? fym_pv : fym[k > 0 ? k - 1 : 0];
        const T rnew = D::sub(r[k], D::mul(c3, D::sub(frm[k], frl)));
        const T ynew = D::sub(y[k], D::add(D::mul(c3, D::sub(fym[k], fyl)), D::mul(c4, D::add(ym[k], yl))));
        const bool interior = (j >= 1) && (j <= M - 2);
        r[k] = interior ? rnew : r[k];
        y[k] = interior ? ynew : y[k];
    }
    return ok;
}

// One env.step() for the env(s) of one warp.
//
// LPE lanes per env: 32 = one freeway per warp; 16 = two per warp (grids up to 64 points; opt-in, see arz_launch_step).
// Cross-lane traffic is full-mask with the segment's width, so both segments run in lock-step.
//
// FAST = true is the path the kernel runs: deferred-check divisions throughout; the operand-window predicate of every
// substep and of the final velocity is AND-ed into one flag, voted ONCE per step, and if any lane of the warp saw an
// exceptional operand the function returns false BEFORE storing anything -- the kernel then calls the out-of-line
// FAST = false instance, which redoes the whole step from the (untouched) global state with IEEE divisions.
template <int K, typename T, int LPE, bool FAST>
__device__ __forceinline__ bool arz_step_env(const ArzArgs& a, T* sm) {
    using D = Strict<T>;
    constexpr int EPW = 32 / LPE;                          // envs per warp
    const int lane = threadIdx.x & 31;
    const int l = lane & (LPE - 1);                        // lane within the env's segment
    const int64_t env_raw = (static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5)) * EPW + lane / LPE;
    const bool have = env_raw < a.buf.n_envs;              // odd batch: the last warp's second segment shadows the last env,
    const int64_t env = have ? env_raw : a.buf.n_envs - 1; // computes along (the shuffles are full-mask) and stores nothing
    const pdecg_arz_cfg& cfg = a.cfg;
    const int M = cfg.M;
    const int j0 = l * K;
    const ArzCtx<T> ctx(cfg);
    ArzCoef<T> co;
    co.c1 = static_cast<T>(cfg.c1); co.c2 = static_cast<T>(cfg.c2); co.c3 = static_cast<T>(cfg.c3); co.c4 = static_cast<T>(cfg.c4);

    // start the row loads first: the per-env scalars (one 48-byte row, read as three broadcast 16-byte loads) overlap them
    T* rrow = reinterpret_cast<T*>(a.buf.ry) + env * 2 * M;
    T* yrow = rrow + M;
    T r[K], y[K], r_raw[K], y_raw[K];
    arz_load_issue<K, T, LPE>(r_raw, rrow, M, l);
    arz_load_issue<K, T, LPE>(y_raw, yrow, M, l);
    double* frow = a.buf.fenv + env * PDECG_ARZ_NF64;
    static_assert(PDECG_ARZ_F_TIME == 0 && PDECG_ARZ_F_RS == 1 && PDECG_ARZ_F_VS == 2 && PDECG_ARZ_F_QS == 3 &&
                  PDECG_ARZ_F_ACT_LO == 4 && PDECG_ARZ_F_ACT_HI == 5 && PDECG_ARZ_NF64 == 6, "fenv row layout");
    const double2 f01 = reinterpret_cast<const double2*>(frow)[0], f23 = reinterpret_cast<const double2*>(frow)[1];
    const double2 f45 = reinterpret_cast<const double2*>(frow)[2];
    double time = f01.x;
    const T rs = static_cast<T>(f01.y), vs = static_cast<T>(f23.x), qs = static_cast<T>(f23.y);
    // action_space.low / .high as built in __init__ (:97-100): NOT refreshed when 'outlet-train' re-draws qs in reset()
    const double lo = f45.x, hi = f45.y;

    arz_load_finish<K, T, LPE>(r, r_raw, sm, l);
    arz_load_finish<K, T, LPE>(y, y_raw, sm, l);
    // padding slots (never stored, never read by valid points) hold (rho, y) = (1, 1): inside every division window
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (j0 + k >= M) { r[k] = T(1); y[k] = T(1); }
    }

    // :148-170 -- float time advances once per step; actions clipped to the action space
    time = __dadd_rn(time, cfg.dt);
    const double* act = a.buf.action + env * cfg.action_dim;
    const T a0 = static_cast<T>(clipd(act[0], lo, hi));
    const T a1 = (cfg.action_dim > 1) ? static_cast<T>(clipd(act[1], lo, hi)) : a0;
    const int st = cfg.simulation_type;
    co.q_in = (st == PDECG_ARZ_INLET || st == PDECG_ARZ_BOTH) ? a0 : qs;
    co.q_out = (st == PDECG_ARZ_OUTLET || st == PDECG_ARZ_OUTLET_TRAIN) ? a0 : ((st == PDECG_ARZ_BOTH) ? a1 : qs);

    // :173 -- the env's clock only moves once per step, so `time < T` is one decision per env and step
    const bool live = time < cfg.T;
    const bool any_live = (EPW == 1) ? live : __any_sync(kFull, live);
    bool ok = true;
    for (int count = 0; count < cfg.control_freq && any_live; ++count)
        ok = ok & arz_substep<K, T, FAST, LPE>(r, y, ctx, co, j0, M);
    if (EPW > 1 && !live) {                                // the other env of the warp integrated; this one is past T: its
        arz_load_issue<K, T, LPE>(r_raw, rrow, M, l);      // lanes ran along on a state that is now discarded
        arz_load_issue<K, T, LPE>(y_raw, yrow, M, l);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            r[k] = (j0 + k < M) ? r_raw[k] : T(1);
            y[k] = (j0 + k < M) ? y_raw[k] : T(1);
        }
        ok = true;
    }

    // ---- velocity, reward, flags, observation (:226-233) ----
    T v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if constexpr (FAST) {
            v[k] = D::add(fast_div(y[k], r[k]), D::mul(ctx.vm, D::sub(T(1), ctx.by_rm.fast(r[k]))));
            ok = ok & DivWindow<T>::b(r[k]) & DivWindow<T>::a(y[k]);
        } else {
            v[k] = D::add(D::div(y[k], r[k]), veq(ctx, r[k]));
        }
    }
    if constexpr (FAST) {
        if (!__all_sync(kFull, ok)) return false;          // nothing stored yet: the exact instance redoes the step
    }
    double sv = 0.0, sr = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double dv = static_cast<double>(D::sub(v[k], vs)), dr = static_cast<double>(D::sub(r[k], rs));
        if (j0 + k < M) {
            sv += dv * dv;
            sr += dr * dr;
        }
    }
    // traffic_arz_reward.py:22: -(||v - vs|| / vs + ||rho - rs|| / rs)   (float64 mode: numpy's operations on float64
    // norms; float32 mode: the norms are float32 values)
    double reward;
    if constexpr (LPE == 32) {
        // both sums in ONE butterfly: the lower half-warp ends up with sum_v, the upper with sum_r; each half takes its own
        // square root and quotient, and one more exchange adds the two terms
        const bool up = (lane & 16) != 0;
        double mine = up ? sr : sv;
        mine += __shfl_xor_sync(kFull, up ? sv : sr, 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mine += __shfl_xor_sync(kFull, mine, o);
        const double den = static_cast<double>(up ? rs : vs);
        if constexpr (sizeof(T) == 8) {
            const double term = __ddiv_rn(sqrt(mine), den);
            reward = -__dadd_rn(term, __shfl_xor_sync(kFull, term, 16));
            sv = mine;                                     // (only their being zero is looked at below)
        } else {
            const float term = __fdiv_rn(static_cast<float>(sqrt(mine)), static_cast<float>(den));
            reward = -static_cast<double>(__fadd_rn(term, __shfl_xor_sync(kFull, term, 16)));
        }
        sv = mine + __shfl_xor_sync(kFull, mine, 16);      // sum_v + sum_r, for the steady-state test
        sr = 0.0;
    } else {
        sv = seg_sum<LPE>(sv);
        sr = seg_sum<LPE>(sr);
        reward = (sizeof(T) == 8)
            ? -__dadd_rn(__ddiv_rn(sqrt(sv), static_cast<double>(vs)), __ddiv_rn(sqrt(sr), static_cast<double>(rs)))
            : -static_cast<double>(__fadd_rn(__fdiv_rn(static_cast<float>(sqrt(sv)), static_cast<float>(vs)), __fdiv_rn(static_cast<float>(sqrt(sr)), static_cast<float>(rs))));
    }
    const unsigned segm = seg_mask<LPE>(lane);
    // truncate() (:116-127).  "Steady" = every deviation exactly zero: impossible unless both sums of squares are zero
    // (the converse can fail through underflow), so the per-point test only runs in that case.
    bool steady = false;
    if (__any_sync(kFull, sv + sr == 0.0)) {
        bool st_l = true;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (j0 + k < M) st_l = st_l && (D::sub(v[k], vs) == T(0)) && (D::sub(r[k], rs) == T(0));
        steady = (__ballot_sync(kFull, !st_l) & segm) == 0u;
    }
    bool over = false;
    if (cfg.limit_pde_state_size) {
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (j0 + k < M) over = over || (v[k] > static_cast<T>(cfg.vm)) || (r[k] > static_cast<T>(cfg.rm));
        over = (__ballot_sync(kFull, over) & segm) != 0u;
    }
    bool term = time >= cfg.t_over_dt;                                                // :109-111
    if (term) time = 0.0;
    const bool trunc = over || steady;
    if (st != PDECG_ARZ_OUTLET_TRAIN) term = term || (reward > -0.00023);             // :233
    if (!have) return true;                                                           // shadow segment: no stores (no warp-wide op follows)

#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = (j0 + k < M) ? v[k] : T(0);
    arz_store<K, T, LPE>(r, rrow, M, sm, l);
    arz_store<K, T, LPE>(y, yrow, M, sm, l);
    const bool restart = a.ar_profile != nullptr && (term || trunc);
    if (st == PDECG_ARZ_OUTLET_TRAIN) {                                               // :231
#pragma unroll
        for (int k = 0; k < K; ++k) {
            r[k] = D::div(D::sub(r[k], rs), rs);                                       // (the state rows are stored already)
            v[k] = D::div(D::sub(v[k], vs), vs);
        }
    }
    T* orow = reinterpret_cast<T*>(a.buf.obs) + env * 2 * M;
    arz_store<K, T, LPE>(r, orow, M, sm, l);
    arz_store<K, T, LPE>(v, orow + M, M, sm, l);
    if (restart && a.buf.final_obs != nullptr) {                                      // terminal observation for the caller
        T* frow_obs = reinterpret_cast<T*>(a.buf.final_obs) + env * 2 * M;
        arz_store<K, T, LPE>(r, frow_obs, M, sm, l);
        arz_store<K, T, LPE>(v, frow_obs + M, M, sm, l);
    }
    if (l == 0) {
        frow[PDECG_ARZ_F_TIME] = time;
        a.buf.reward[env] = reward;
        a.buf.terminated[env] = term ? 1 : 0;
        a.buf.truncated[env] = trunc ? 1 : 0;
    }
    if (restart) arz_autoreset_env<T, LPE>(a, env, l, segm);
    return true;
}

template <int K, typename T, int LPE>
__device__ __noinline__ void arz_step_env_exact(const ArzArgs& a, T* sm) {
    arz_step_env<K, T, LPE, false>(a, sm);
}

template <int K, typename T, int LPE>
__global__ void __launch_bounds__(kThreads, LPE == 16 ? PDECG_ARZ_MINBLOCKS_H : (K <= 8 ? PDECG_ARZ_MINBLOCKS : 1))
arz_step_kernel(const __grid_constant__ ArzArgs a) {
    __shared__ T sm_all[kWarpsPerBlock][(K <= 2 || LPE != 32) ? 1 : 32 * K];
    const int wib = threadIdx.x >> 5;
    if ((static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + wib) * (32 / LPE) >= a.buf.n_envs) return;      // whole warp
    const ArzCtx<T> ctx(a.cfg);
    bool done = false;
    if (ctx.by_rm.fast_ok()) done = arz_step_env<K, T, LPE, true>(a, sm_all[wib]);     // uniform
    if (!done) arz_step_env_exact<K, T, LPE>(a, sm_all[wib]);                         // rare: exceptional operands somewhere in the warp
}

// TrafficPDE1D.reset (:237-267) for one env, executed by its warp.  `keep_outputs`: leave reward / flags alone
// (fused auto-reset: they keep the step's values).
template <typename T>
__device__ __forceinline__ void arz_reset_env(const ArzArgs& a, const double* __restrict__ profile, int64_t env, int lane,
                                              bool keep_outputs, int nlanes = 32) {
    using D = Strict<T>;
    const pdecg_arz_cfg& cfg = a.cfg;
    const ArzCtx<T> ctx(cfg);
    const int M = cfg.M;
    const double* frow = a.buf.fenv + env * PDECG_ARZ_NF64;
    const T rs = static_cast<T>(frow[PDECG_ARZ_F_RS]), qs = static_cast<T>(frow[PDECG_ARZ_F_QS]);
    T* rrow = reinterpret_cast<T*>(a.buf.ry) + env * 2 * M;
    T* orow = reinterpret_cast<T*>(a.buf.obs) + env * 2 * M;
    for (int j = lane; j < M; j += nlanes) {
        // :259-261  r = rs*profile ; y = qs*1 - vm*r + (vm/rm)*r**2 ; v = y/r + Veq(r)
        const T r = D::mul(rs, static_cast<T>(profile[j]));
        const T y = D::add(D::sub(D::mul(qs, T(1)), D::mul(static_cast<T>(cfg.vm), r)), D::mul(static_cast<T>(cfg.vm_over_rm), D::mul(r, r)));
        const T v = D::add(D::div(y, r), veq(ctx, r));
        rrow[j] = r;
        rrow[M + j] = y;
        orow[j] = r;                               // reset() returns concat(r, v) un-normalised in every mode,
        orow[M + j] = v;                           // 'outlet-train' included (:263)
    }
    if (lane == 0 && !keep_outputs) {
        a.buf.reward[env] = 0.0;
        a.buf.terminated[env] = 0;
        a.buf.truncated[env] = 0;
    }
}

// fused same-step auto-reset: out of line so the (rare) restart does not weigh on the step kernel's registers
template <typename T, int LPE>
__device__ __noinline__ void arz_autoreset_env(const ArzArgs& a, int64_t env, int l, unsigned segm) {
    __syncwarp(segm);                              // the step's own row / observation stores come first
    arz_reset_env<T>(a, a.ar_profile, env, l, true, LPE);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) arz_reset_kernel(const ArzArgs a, const double* __restrict__ profile,
                                                             const uint8_t* __restrict__ mask) {
    const int lane = threadIdx.x & 31;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    if (env >= a.buf.n_envs) return;
    if (mask != nullptr && mask[env] == 0) return;
    arz_reset_env<T>(a, profile, env, lane, false);
}

static int arz_validate(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf) {
    if (!cfg || !buf) return fail(PDECG_ERR_INVALID, "null cfg/buffers");
    if (cfg->abi_version != PDECG_ABI_VERSION) return fail(PDECG_ERR_INVALID, "abi_version mismatch");
    if (cfg->simulation_type < PDECG_ARZ_INLET || cfg->simulation_type > PDECG_ARZ_OUTLET_TRAIN) return fail(PDECG_ERR_INVALID, "Invalid simulation type");
    if (cfg->M < 4) return fail(PDECG_ERR_INVALID, "need at least 4 grid points");
    if (cfg->M > 512) return fail(PDECG_ERR_UNSUPPORTED, "grid of %d points exceeds the register-resident limit of 512 points per env", cfg->M);
    if (cfg->control_freq < 1) return fail(PDECG_ERR_INVALID, "control_freq must be a positive integer (got %d)", cfg->control_freq);
    if (cfg->action_dim != ((cfg->simulation_type == PDECG_ARZ_BOTH) ? 2 : 1)) return fail(PDECG_ERR_INVALID, "action_dim must be 2 for 'both', else 1");
    if (cfg->precision < PDECG_PREC_REF || cfg->precision > PDECG_PREC_FP32) return fail(PDECG_ERR_INVALID, "bad precision %d", cfg->precision);
    if (buf->n_envs < 0) return fail(PDECG_ERR_INVALID, "n_envs < 0");
    if (buf->n_envs > 0 && (!buf->ry || !buf->obs || !buf->fenv || !buf->reward || !buf->terminated || !buf->truncated))
        return fail(PDECG_ERR_INVALID, "null device buffer");
    if (reinterpret_cast<uintptr_t>(buf->fenv) % 16 != 0) return fail(PDECG_ERR_INVALID, "fenv must be 16-byte aligned");
    return PDECG_OK;
}

}  // namespace pdecg

using namespace pdecg;

extern "C" int pdecg_arz_reset(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, const double* profile,
                               const uint8_t* mask, pdecg_stream_t stream) {
    int rc = arz_validate(cfg, buf);
    if (rc != PDECG_OK) return rc;
    if (buf->n_envs == 0) return PDECG_OK;
    if (!profile) return fail(PDECG_ERR_INVALID, "null profile buffer");
    ArzArgs a{*cfg, *buf, nullptr};
    const unsigned grid = static_cast<unsigned>((buf->n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (cfg->precision == PDECG_PREC_FP32) arz_reset_kernel<float><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, profile, mask);
    else arz_reset_kernel<double><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, profile, mask);
    count_launch();
    return check_launch("arz_reset_kernel");
}

static int arz_launch_step(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, const double* ar_profile, pdecg_stream_t stream) {
    int rc = arz_validate(cfg, buf);
    if (rc != PDECG_OK) return rc;
    if (buf->n_envs == 0) return PDECG_OK;
    if (!buf->action) return fail(PDECG_ERR_INVALID, "null action buffer");
    ArzArgs a{*cfg, *buf, ar_profile};
    const unsigned grid = static_cast<unsigned>((buf->n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Two envs per warp (16 lanes each): 29 % fewer instructions per env-step, but 118 registers at 4 points per lane
    // leave 16 warps per SM and the kernel turns latency-bound (0.230 vs 0.199 ms per 262k-env step) -- opt-in only.
    if (cfg->M <= 64 && getenv("PDECG_ARZ_LPE16")) {
        const int64_t warps = (buf->n_envs + 1) / 2;
        const unsigned grid2 = static_cast<unsigned>((warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
        switch ((cfg->M + 15) / 16) {
#define PDECG_CASE(KK)                                                                              \
    case KK:                                                                                        \
        if (cfg->precision == PDECG_PREC_FP32) arz_step_kernel<KK, float, 16><<<grid2, kThreads, 0, st>>>(a); \
        else arz_step_kernel<KK, double, 16><<<grid2, kThreads, 0, st>>>(a);                         \
        break;
            PDECG_CASE(1) PDECG_CASE(2) PDECG_CASE(3) PDECG_CASE(4)
#undef PDECG_CASE
        }
        count_launch();
        return check_launch("arz_step_kernel");
    }
    int K = (cfg->M + 31) / 32;
    if (K > 8) K = (K + 1) / 2 * 2;                    // 257..512 points: even K in {10, 12, 14, 16}
    switch (K) {
#define PDECG_CASE(KK)                                                                          \
    case KK:                                                                                    \
        if (cfg->precision == PDECG_PREC_FP32) arz_step_kernel<KK, float, 32><<<grid, kThreads, 0, st>>>(a); \
        else arz_step_kernel<KK, double, 32><<<grid, kThreads, 0, st>>>(a);                          \
        break;
        PDECG_CASE(1) PDECG_CASE(2) PDECG_CASE(3) PDECG_CASE(4) PDECG_CASE(5) PDECG_CASE(6) PDECG_CASE(7) PDECG_CASE(8)
        PDECG_CASE(10) PDECG_CASE(12) PDECG_CASE(14) PDECG_CASE(16)
#undef PDECG_CASE
        default: return fail(PDECG_ERR_UNSUPPORTED, "unsupported points-per-lane %d", K);
    }
    count_launch();
    return check_launch("arz_step_kernel");
}

extern "C" int pdecg_arz_step(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, pdecg_stream_t stream) {
    return arz_launch_step(cfg, buf, nullptr, stream);
}

extern "C" int pdecg_arz_step_autoreset(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, const double* profile,
                                        pdecg_stream_t stream) {
    if (!profile) return fail(PDECG_ERR_INVALID, "null profile buffer");
    return arz_launch_step(cfg, buf, profile, stream);
}

// Host-buffer step: H2D(actions) -> kernel -> D2H(observation, reward, flags) on `stream`, then stream-synchronised.
extern "C" int pdecg_arz_step_host(const pdecg_arz_cfg* cfg, const pdecg_arz_buffers* buf, const double* action_host,
                                   double* obs_host, double* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                                   const double* autoreset_profile, pdecg_stream_t stream) {
    int rc = arz_validate(cfg, buf);
    if (rc != PDECG_OK) return rc;
    if (buf->n_envs == 0) return PDECG_OK;
    if (!action_host || !buf->action) return fail(PDECG_ERR_INVALID, "null action buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n = static_cast<size_t>(buf->n_envs);
    cudaError_t e = cudaMemcpyAsync(const_cast<double*>(buf->action), action_host, n * cfg->action_dim * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "H2D actions: %s", cudaGetErrorString(e));
    if ((rc = arz_launch_step(cfg, buf, autoreset_profile, stream)) != PDECG_OK) return rc;
    const size_t esz = cfg->precision == PDECG_PREC_FP32 ? sizeof(float) : sizeof(double);
    if (obs_host && (e = cudaMemcpyAsync(obs_host, buf->obs, n * 2 * cfg->M * esz, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H obs: %s", cudaGetErrorString(e));
    if (reward_host && (e = cudaMemcpyAsync(reward_host, buf->reward, n * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H reward: %s", cudaGetErrorString(e));
    if (terminated_host && (e = cudaMemcpyAsync(terminated_host, buf->terminated, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H terminated: %s", cudaGetErrorString(e));
    if (truncated_host && (e = cudaMemcpyAsync(truncated_host, buf->truncated, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H truncated: %s", cudaGetErrorString(e));
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
    return PDECG_OK;
}
