// brain_tumor.cu -- BrainTumor1D (+ TherapyWrapper semantics): kernels + C ABI (sm_100a, float64).
//
// Reference behaviour restated here (no code shared): brain_tumor_env.py:33-505 and
// rewards/brain_tumor_reward.py:31-73 of lukebhan/PDEControlGym.
//
// One warp = one patient.  The density row c(x) stays in registers; the wrapper-level loops of the
// reference (whole Growth stage inside reset(), whole Post-Therapy stage inside one step(), forced
// weekend days, the benchmark episode) run INSIDE one launch with a per-env (warp-uniform) trip count,
// so a multi-day advance costs one HBM round trip of the row instead of one per day.
#include <math_constants.h>

#include "pdecg_internal.h"

#ifndef PDECG_BT_MINBLOCKS
#define PDECG_BT_MINBLOCKS 6   // 80 registers, 24 warps/SM: the single-day launch is latency/HBM-bound (measured: 8 -> 0.264, 6 -> 0.252, 5 -> 0.268, 4 -> 0.285 ms at 262k envs)
#endif
#ifndef PDECG_BT_MINBLOCKS_MULTI
#define PDECG_BT_MINBLOCKS_MULTI 4   // 128 registers: the multi-day (TherapyWrapper) launches are instruction-bound
#endif
namespace pdecg {

struct BtArgs {
    pdecg_bt_cfg cfg;
    pdecg_bt_buffers buf;
    int mode;
    const double* ar_init;      // pdecg_bt_step_autoreset: finished envs restart inside the RAW_STEP launch (nullptr: off)
    int64_t ar_stride;
};


struct BtEnv {   // the reference's per-env attributes (brain_tumor_env.py:89-101)
    int t, stage, growthDays, therapyDays, postTherapyDays, simulationDays;
    int firstTherapyDay, firstPostTherapyDay, cDeathDay, t_benchmark;
    int consecutive, treatment_calls, violations;
    double remaining, D, rho, last_dose;
};

// right-most index j with c_j >= thr (getTumorRadius, :104-121); -1 when the tumour is invisible
template <int K, typename T>
__device__ __forceinline__ int rightmost_ge(const T (&c)[K], int j0, int nx, T thr) {
    const int kmax = nx - j0;                         // slots k >= kmax of this lane are padding
    int best = -1;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (k < kmax && c[k] >= thr) best = k;        // ascending k -> keeps the largest
    }
    return warp_max_i(best >= 0 ? j0 + best : -1);
}

// _update_fd (:221-245).  HASR == false encodes "R is None"; treat_to = number of leading points with
// xScale <= treatmentRadius (they receive R = rin, the others R = 1 - exp(-alpha*0) = 0).
struct BtDivs {      // the two loop-invariant divisors of _update_fd (:236-237)
    ConstDiv<double> by_dx2, by_k;
    unsigned tiny_hi;   // high word of the largest power of two <= k*2^-60
    unsigned om_limit;  // |u_c| with a high word below this takes the 3-instruction quotient (0: never)
    unsigned k_hi;      // high word of k: 0 <= v < k for sure when hiword(v) < k_hi (as unsigned)
    bool dx2_one;       // dx**2 == 1 (the reference's grid): the division by it is the identity
    __device__ __forceinline__ explicit BtDivs(const pdecg_bt_cfg& cfg) : by_k(cfg.k) {
        const double dx2 = cfg.dx * cfg.dx;                     // self.dx ** 2 (one rounded multiply)
        dx2_one = (dx2 == 1.0);                                 // the reference's grid: no reciprocal to set up (uniform branch)
        if (!dx2_one) by_dx2 = ConstDiv<double>(dx2);
        const unsigned khi = static_cast<unsigned>(__double2hiint(cfg.k)) & 0x7ff00000u;    // 2^floor(log2 k)
        tiny_hi = (khi > (61u << 20)) ? khi - (60u << 20) : 0u;
        // ConstDiv<double>::in_range is 0x07b00000 <= hi < 0x78300000; everything below tiny_hi is fine too (see
        // bt_fd), so one unsigned compare covers both whenever the two windows overlap
        om_limit = (by_k.d_ok && tiny_hi >= 0x07b00000u) ? 0x78300000u : 0u;
        k_hi = static_cast<unsigned>(__double2hiint(cfg.k));
    }
};

// Exact form of one _update_fd for any operand (per-point divisions and clip); the fast form below falls back to
// it for the whole chunk when a window test fails.  xs = this lane's xScale chunk, R_j = rin where xs_j <= treat_r.
// xScale of slot k of this lane: from the register chunk (multi-day launches) or straight from global memory
template <int K, bool XSREG>
__device__ __forceinline__ double bt_xs(const double (&xs)[K], const double* __restrict__ xg, int j0, int k, int nx) {
    if constexpr (XSREG) return xs[k];
    else return (j0 + k < nx) ? xg[j0 + k] : CUDART_INF;
}

template <int K, bool HASR, bool XSREG>
__device__ __forceinline__ void bt_fd_exact(const pdecg_bt_cfg& cfg, const BtDivs& dv, const double (&c)[K], double (&n)[K], int lane, int j0,
                                             double hl, double hr, double D, double rho, double rin, double treat_r, const double (&xs)[K],
                                             const double* __restrict__ xg) {
    const int nx = cfg.nx;
    double t[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double cl = (k == 0) ? hl : c[k > 0 ? k - 1 : 0];
        const double cr = (k == K - 1) ? hr : c[k < K - 1 ? k + 1 : K - 1];
        const double cc = c[k];
        const double lap = dv.by_dx2(__dadd_rn(__dsub_rn(cr, __dmul_rn(2.0, cc)), cl));
        const double diff = __dmul_rn(D, lap);
        const unsigned cc_hi = static_cast<unsigned>(__double2hiint(cc)) & 0x7fffffffu;
        const double om = (cc_hi < dv.tiny_hi) ? 1.0 : __dsub_rn(1.0, dv.by_k(cc));
        const double pro = __dmul_rn(__dmul_rn(rho, cc), om);
        double acc = __dadd_rn(diff, pro);
        if constexpr (HASR) {
            const double rj = (bt_xs<K, XSREG>(xs, xg, j0, k, nx) <= treat_r) ? rin : 0.0;
            acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(rj, cc), om));
        }
        t[k] = __dadd_rn(cc, __dmul_rn(cfg.dt, acc));
    }
    const double from_left = __shfl_up_sync(kFull, t[K - 1], 1);
    const double from_right = __shfl_down_sync(kFull, t[0], 1);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = t[k];
        if (k == 0) v = (lane == 0) ? ((K > 1) ? t[K > 1 ? 1 : 0] : from_right) : v;
        if (j0 + k == nx - 1) v = (k > 0) ? t[k > 0 ? k - 1 : 0] : from_left;
        n[k] = (v < 0.0) ? 0.0 : ((v > cfg.k) ? cfg.k : v);   // np.clip keeps NaN
    }
}

template <int K, bool DX2ONE, bool HASR, bool XSREG>
__device__ __forceinline__ void bt_fd(const pdecg_bt_cfg& cfg, const BtDivs& dv, const double (&c)[K], double (&n)[K], int lane, int j0,
                                       double D, double rho, double rin, double treat_r, const double (&xs)[K], const double* __restrict__ xg) {
    const int nx = cfg.nx;
    const double hl = __shfl_up_sync(kFull, c[K - 1], 1);
    const double hr = __shfl_down_sync(kFull, c[0], 1);
    // (1.0 - u_c/k): the quotient comes from the 3-instruction exact tail for the whole chunk with ONE window test
    // per lane.  For |u_c| < k*2^-60 the quotient is below 2^-60 whatever its low bits are and the difference
    // rounds to exactly 1.0, so the far field (zeros, denormals) needs neither a test nor div.rn's slow path.
    double t[K];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double cl = (k == 0) ? hl : c[k > 0 ? k - 1 : 0];
        const double cr = (k == K - 1) ? hr : c[k < K - 1 ? k + 1 : K - 1];
        const double cc = c[k];
        // diffusion = D * (((u_r - 2.0*u_c) + u_l) / dx**2)
        const double s = __dadd_rn(__dsub_rn(cr, __dmul_rn(2.0, cc)), cl);
        const double diff = __dmul_rn(D, DX2ONE ? s : dv.by_dx2(s));
        const unsigned cc_hi = static_cast<unsigned>(__double2hiint(cc)) & 0x7fffffffu;
        ok = ok & (cc_hi < dv.om_limit);
        // proliferation = (rho*u_c) * (1.0 - u_c/k)
        const double om = __dsub_rn(1.0, dv.by_k.fast(cc));
        const double pro = __dmul_rn(__dmul_rn(rho, cc), om);
        double acc = __dadd_rn(diff, pro);
        if constexpr (HASR) {
            const double rj = (bt_xs<K, XSREG>(xs, xg, j0, k, nx) <= treat_r) ? rin : 0.0;
            acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(rj, cc), om));
        }                                                   // else `- 0.0`: the identity for every double
        t[k] = __dadd_rn(cc, __dmul_rn(cfg.dt, acc));
    }
    // nextU[0] = nextU[1]; nextU[-1] = nextU[-2]  (mirrored ends), then clip to [0, k].  Padding slots (j >= nx) are
    // left to evolve: valid points never read them (u[nx-1] is the mirror of u[nx-2]), the radius scans mask them,
    // the clip keeps them finite and the store drops them.
    const double from_left = __shfl_up_sync(kFull, t[K - 1], 1);
    const double from_right = __shfl_down_sync(kFull, t[0], 1);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = t[k];
        if (k == 0) v = (lane == 0) ? ((K > 1) ? t[K > 1 ? 1 : 0] : from_right) : v;
        if (j0 + k == nx - 1) v = (k > 0) ? t[k > 0 ? k - 1 : 0] : from_left;
        ok = ok & (static_cast<unsigned>(__double2hiint(v)) < dv.k_hi);   // 0 <= v < k for sure: the clip is the identity
        n[k] = v;
    }
    // huge / negative / NaN densities, values within 2^-20 of k, or an exotic k: redo this lane's chunk exactly
    if (!ok) bt_fd_exact<K, HASR, XSREG>(cfg, dv, c, n, lane, j0, hl, hr, D, rho, rin, treat_r, xs, xg);
}

// Single-day launches (80 registers): the same arithmetic in two passes with partial fall-backs, which keeps fewer
// values live than the fused form above (measured 0.252 vs 0.281 ms per 262,144-env raw step).
template <int K, bool DX2ONE, bool HASR>
__device__ __forceinline__ void bt_fd_lean(const pdecg_bt_cfg& cfg, const BtDivs& dv, const double (&c)[K], double (&n)[K], int lane, int j0,
                                            double D, double rho, double rin, double treat_r, const double* __restrict__ xg) {
    const int nx = cfg.nx;
    const double hl = __shfl_up_sync(kFull, c[K - 1], 1);
    const double hr = __shfl_down_sync(kFull, c[0], 1);
    // (1.0 - u_c/k): the quotient comes from the 3-instruction exact tail for the whole chunk with ONE window test
    // per lane.  For |u_c| < k*2^-60 the quotient is below 2^-60 whatever its low bits are and the difference
    // rounds to exactly 1.0, so the far field (zeros, denormals) needs neither a test nor div.rn's slow path.
    double diff[K], q[K];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double cl = (k == 0) ? hl : c[k > 0 ? k - 1 : 0];
        const double cr = (k == K - 1) ? hr : c[k < K - 1 ? k + 1 : K - 1];
        const double cc = c[k];
        // diffusion = D * (((u_r - 2.0*u_c) + u_l) / dx**2)
        const double s = __dadd_rn(__dsub_rn(cr, __dmul_rn(2.0, cc)), cl);
        diff[k] = __dmul_rn(D, DX2ONE ? s : dv.by_dx2(s));
        const unsigned cc_hi = static_cast<unsigned>(__double2hiint(cc)) & 0x7fffffffu;
        ok = ok & (cc_hi < dv.om_limit);
        q[k] = dv.by_k.fast(cc);
    }
    if (!ok) {                                             // huge / NaN densities or an exotic k: exact division
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned cc_hi = static_cast<unsigned>(__double2hiint(c[k])) & 0x7fffffffu;
            q[k] = (cc_hi < dv.tiny_hi) ? 0.0 : dv.by_k(c[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double cc = c[k];
        // proliferation = (rho*u_c) * (1.0 - u_c/k)
        const double om = __dsub_rn(1.0, q[k]);
        const double pro = __dmul_rn(__dmul_rn(rho, cc), om);
        double acc = __dadd_rn(diff[k], pro);
        if constexpr (HASR) {
            const double rj = (bt_xs<K, false>(c, xg, j0, k, nx) <= treat_r) ? rin : 0.0;
            acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(rj, cc), om));
        }                                                   // else `- 0.0`: the identity for every double
        n[k] = __dadd_rn(cc, __dmul_rn(cfg.dt, acc));
    }
    // nextU[0] = nextU[1]; nextU[-1] = nextU[-2]  (mirrored ends), then clip to [0, k].  Padding slots (j >= nx) are
    // left to evolve: valid points never read them (u[nx-1] is the mirror of u[nx-2]), the radius scans mask them,
    // the clip keeps them finite and the store drops them.
    const double from_left = __shfl_up_sync(kFull, n[K - 1], 1);
    const double from_right = __shfl_down_sync(kFull, n[0], 1);
    bool inside = true;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = n[k];
        if (k == 0) v = (lane == 0) ? ((K > 1) ? n[K > 1 ? 1 : 0] : from_right) : v;
        if (j0 + k == nx - 1) v = (k > 0) ? n[k > 0 ? k - 1 : 0] : from_left;
        inside = inside & (static_cast<unsigned>(__double2hiint(v)) < dv.k_hi);   // 0 <= v < k for sure
        n[k] = v;
    }
    if (!inside) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double v = n[k];
            n[k] = (v < 0.0) ? 0.0 : ((v > cfg.k) ? cfg.k : v);   // np.clip keeps NaN
        }
    }
}

// float32 fast mode (PDECG_PREC_FP32): the same update in float32, same operation order, plain IEEE divisions and clip
template <int K, bool HASR>
__device__ __forceinline__ void bt_fd_f32(const pdecg_bt_cfg& cfg, const float (&c)[K], float (&n)[K], int lane, int j0,
                                           float D, float rho, float rin, double treat_r, const double* __restrict__ xg) {
    using F = Strict<float>;
    const int nx = cfg.nx;
    const float dx2 = F::mul(static_cast<float>(cfg.dx), static_cast<float>(cfg.dx)), kk = static_cast<float>(cfg.k), dt = static_cast<float>(cfg.dt);
    const float hl = __shfl_up_sync(kFull, c[K - 1], 1);
    const float hr = __shfl_down_sync(kFull, c[0], 1);
    float t[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float cl = (k == 0) ? hl : c[k > 0 ? k - 1 : 0];
        const float cr = (k == K - 1) ? hr : c[k < K - 1 ? k + 1 : K - 1];
        const float cc = c[k];
        const float diff = F::mul(D, F::div(F::add(F::sub(cr, F::mul(2.0f, cc)), cl), dx2));
        const float om = F::sub(1.0f, F::div(cc, kk));
        float acc = F::add(diff, F::mul(F::mul(rho, cc), om));
        if constexpr (HASR) {
            const float rj = ((j0 + k < nx ? xg[j0 + k] : CUDART_INF) <= treat_r) ? rin : 0.0f;
            acc = F::sub(acc, F::mul(F::mul(rj, cc), om));
        }
        t[k] = F::add(cc, F::mul(dt, acc));
    }
    const float from_left = __shfl_up_sync(kFull, t[K - 1], 1);
    const float from_right = __shfl_down_sync(kFull, t[0], 1);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float v = t[k];
        if (k == 0) v = (lane == 0) ? ((K > 1) ? t[K > 1 ? 1 : 0] : from_right) : v;
        if (j0 + k == nx - 1) v = (k > 0) ? t[k > 0 ? k - 1 : 0] : from_left;
        n[k] = (v < 0.0f) ? 0.0f : ((v > kk) ? kk : v);
    }
}

__device__ __forceinline__ void bt_book(BtEnv& e) {
    if (e.stage == PDECG_BT_THERAPY) {
        e.therapyDays = e.t - e.growthDays;
        e.simulationDays = e.growthDays + e.therapyDays;
    } else if (e.stage == PDECG_BT_POST) {
        e.postTherapyDays = e.t - e.therapyDays - e.growthDays;
        e.simulationDays = e.growthDays + e.therapyDays + e.postTherapyDays;
    }
}

// One call of BrainTumor1D.step (:123-218) for an env with t < nt-1.
template <int K, bool XSREG, typename T>
__device__ __forceinline__ void bt_day(const BtArgs& a, const BtDivs& dv, const double (&xs)[K], BtEnv& e, T (&u)[K], int lane, int j0,
                                        double control, double& reward, bool& term, bool& trunc) {
    const pdecg_bt_cfg& cfg = a.cfg;
    const int nx = cfg.nx;
    e.t += 1;
    T n[K];
    double treat_r = 0.0, applied = 0.0;
    const int stage0 = e.stage;
    if (stage0 == PDECG_BT_THERAPY) {
        // :158-166
        const double want = __dmul_rn(control, cfg.total_dosage);
        applied = (e.remaining < want) ? e.remaining : want;          // python min(a, b)
        e.remaining = __dsub_rn(e.remaining, applied);
        e.last_dose = applied;
        // _compute_radiation_field (:247-265) uses the PREVIOUS day's T2 radius
        const int t2 = rightmost_ge<K, T>(u, j0, nx, static_cast<T>(cfg.t2_threshold_k));
        treat_r = (t2 < 0) ? 0.0 : __dadd_rn(__dmul_rn(static_cast<double>(t2), cfg.dx), 25.0);
        // R = 1 - exp(-alpha*BED) where xScale <= treatmentRadius, 1 - exp(-alpha*0) = 0 elsewhere (:259-265).
        // A zero dose gives BED = 0 and R = 1 - exp(-0) = +0 everywhere: skip the exp().
        double rin = 0.0;
        if (applied != 0.0) {
            const double bed = __dadd_rn(applied, __ddiv_rn(__dmul_rn(applied, applied), cfg.alpha_beta_ratio));
            rin = __dsub_rn(1.0, exp_cr(__dmul_rn(-cfg.alpha, bed)));     // correctly rounded (pdecg_common.cuh)
        }
        if constexpr (sizeof(T) == 4) {
            bt_fd_f32<K, true>(cfg, u, n, lane, j0, static_cast<float>(e.D), static_cast<float>(e.rho), static_cast<float>(rin), treat_r, a.buf.x_scale);
        } else if constexpr (XSREG) {
            if (dv.dx2_one) bt_fd<K, true, true, true>(cfg, dv, u, n, lane, j0, e.D, e.rho, rin, treat_r, xs, a.buf.x_scale);
            else bt_fd<K, false, true, true>(cfg, dv, u, n, lane, j0, e.D, e.rho, rin, treat_r, xs, a.buf.x_scale);
        } else {
            if (dv.dx2_one) bt_fd_lean<K, true, true>(cfg, dv, u, n, lane, j0, e.D, e.rho, rin, treat_r, a.buf.x_scale);
            else bt_fd_lean<K, false, true>(cfg, dv, u, n, lane, j0, e.D, e.rho, rin, treat_r, a.buf.x_scale);
        }
    } else {
        e.last_dose = 0.0;                                           // dosage_vs_time[t] stays 0 outside Therapy (:56,160)
        if constexpr (sizeof(T) == 4) {
            bt_fd_f32<K, false>(cfg, u, n, lane, j0, static_cast<float>(e.D), static_cast<float>(e.rho), 0.0f, 0.0, a.buf.x_scale);
        } else if constexpr (XSREG) {
            if (dv.dx2_one) bt_fd<K, true, false, true>(cfg, dv, u, n, lane, j0, e.D, e.rho, 0.0, 0.0, xs, a.buf.x_scale);
            else bt_fd<K, false, false, true>(cfg, dv, u, n, lane, j0, e.D, e.rho, 0.0, 0.0, xs, a.buf.x_scale);
        } else {
            if (dv.dx2_one) bt_fd_lean<K, true, false>(cfg, dv, u, n, lane, j0, e.D, e.rho, 0.0, 0.0, a.buf.x_scale);
            else bt_fd_lean<K, false, false>(cfg, dv, u, n, lane, j0, e.D, e.rho, 0.0, 0.0, a.buf.x_scale);
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) u[k] = n[k];

    const int t1 = rightmost_ge<K, T>(u, j0, nx, static_cast<T>(cfg.t1_threshold_k));
    const double t1_radius = __dmul_rn(static_cast<double>(t1), cfg.dx);
    if (stage0 == PDECG_BT_GROWTH) {
        e.growthDays = e.t;
        if (t1 >= 0 && t1_radius >= cfg.t1_detection_radius) {       // :151-155
            e.firstTherapyDay = e.t + 1;
            e.stage = PDECG_BT_THERAPY;
        }
    } else if (stage0 == PDECG_BT_THERAPY) {
        if (e.remaining < cfg.dosage_termination_threshold) {        // :171-176
            e.therapyDays = e.t - e.growthDays;
            e.firstPostTherapyDay = e.t + 1;
            e.stage = PDECG_BT_POST;
        }
    }
    // terminate() (:285-311) then truncate() (:313-345)
    term = e.t >= cfg.nt - 1;
    if (term) bt_book(e);
    trunc = (t1 >= 0) && (t1_radius >= cfg.t1_death_radius);
    if (trunc && e.cDeathDay < 0) {
        e.cDeathDay = e.t;
        bt_book(e);
    }
    // rewards (brain_tumor_reward.py:46-73)
    const bool ended = term || trunc;
    if (stage0 == PDECG_BT_THERAPY) {
        if (e.t_benchmark < 0) reward = 0.0;
        else if (ended) reward = static_cast<double>(e.t - e.t_benchmark);
        else {
            const double maxsafe = __dmul_rn(116.0, pow(treat_r, -0.685));
            const double ratio = __ddiv_rn(__dsub_rn(applied, maxsafe), __dsub_rn(cfg.total_dosage, maxsafe));
            double cl = (0.0 > ratio) ? 0.0 : ratio;       // python max(ratio, 0.0)
            cl = (1.0 < cl) ? 1.0 : cl;                    // python min(., 1.0)
            reward = __dmul_rn(-50.0, pow(cl, 1.0 / 3.0));
        }
    } else if (e.stage == PDECG_BT_POST && ended) {
        reward = (e.t_benchmark < 0) ? 0.0 : static_cast<double>(e.t - e.t_benchmark);
    } else {
        reward = 0.0;
    }
}

// MULTI = false: PDECG_BT_RAW_STEP only (one day per launch); MULTI = true: the TherapyWrapper modes.
template <typename T>
__device__ __noinline__ void bt_autoreset_env_t(const BtArgs& a, int64_t env, int lane);

template <int K, bool MULTI, typename T>
__global__ void __launch_bounds__(kThreads, K > 8 ? 1 : (MULTI ? PDECG_BT_MINBLOCKS_MULTI : PDECG_BT_MINBLOCKS)) bt_run_kernel(const __grid_constant__ BtArgs a) {
    __shared__ T sm_all[kWarpsPerBlock][32 * K];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + wib;
    if (env >= a.buf.n_envs) return;
    T* sm = sm_all[wib];
    const pdecg_bt_cfg& cfg = a.cfg;
    const int nx = cfg.nx;
    const int j0 = lane * K;
    T* const obs_all = reinterpret_cast<T*>(a.buf.obs);

    // start the row load first: the scalar loads and their broadcast then overlap its latency
    T* urow = reinterpret_cast<T*>(a.buf.u) + env * nx;
    T u[K], u_raw[K];
    constexpr int SURE = (K <= 8) ? K - 1 : K - 2;      // rows of a coalesced pass (j = lane + 32 i) that lie inside the row for every lane
    load_row_issue<T, T, K, 32, SURE>(u_raw, urow, nx, lane);
    // per-env scalars: the 64-byte integer row and the 32-byte float row, read by every lane as broadcast 16-byte loads
    int* irow = a.buf.ienv + env * PDECG_BT_NINT;
    double* frow = a.buf.fenv + env * PDECG_BT_NF64;
    static_assert(PDECG_BT_I_T == 0 && PDECG_BT_I_STAGE == 1 && PDECG_BT_I_GROWTH_DAYS == 2 && PDECG_BT_I_THERAPY_DAYS == 3 &&
                  PDECG_BT_I_POST_DAYS == 4 && PDECG_BT_I_SIM_DAYS == 5 && PDECG_BT_I_FIRST_THERAPY == 6 && PDECG_BT_I_FIRST_POST == 7 &&
                  PDECG_BT_I_DEATH_DAY == 8 && PDECG_BT_I_T_BENCHMARK == 9 && PDECG_BT_I_CONSECUTIVE == 10 &&
                  PDECG_BT_I_TREATMENT_CALLS == 11 && PDECG_BT_I_VIOLATIONS == 12 && PDECG_BT_NINT == 16, "ienv row layout");
    static_assert(PDECG_BT_F_REMAINING == 0 && PDECG_BT_F_D == 1 && PDECG_BT_F_RHO == 2 && PDECG_BT_F_LAST_DOSE == 3 && PDECG_BT_NF64 == 4, "fenv row layout");
    const int4 i0 = reinterpret_cast<const int4*>(irow)[0], i1 = reinterpret_cast<const int4*>(irow)[1];
    const int4 i2 = reinterpret_cast<const int4*>(irow)[2], i3 = reinterpret_cast<const int4*>(irow)[3];
    const double2 f0 = reinterpret_cast<const double2*>(frow)[0], f1 = reinterpret_cast<const double2*>(frow)[1];
    BtEnv e;
    e.t = i0.x; e.stage = i0.y; e.growthDays = i0.z; e.therapyDays = i0.w;
    e.postTherapyDays = i1.x; e.simulationDays = i1.y; e.firstTherapyDay = i1.z; e.firstPostTherapyDay = i1.w;
    e.cDeathDay = i2.x; e.t_benchmark = i2.y; e.consecutive = i2.z; e.treatment_calls = i2.w;
    e.violations = i3.x;
    e.remaining = f0.x; e.D = f0.y; e.rho = f1.x; e.last_dose = f1.y;

    load_row_finish<T, K>(u, u_raw, sm, lane);

    const BtDivs dv(cfg);
    // this lane's chunk of xScale (radiation-field mask): register-resident for the multi-day launches; the
    // single-day launch reads it from global memory on a therapy day
    double xs[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xs[k] = (MULTI && j0 + k < nx) ? a.buf.x_scale[j0 + k] : CUDART_INF;
    double reward = 0.0;
    bool term = false, trunc = false, valid = true, wrote_obs = false;
    const int last = cfg.nt - 1;

    if constexpr (!MULTI) {
        if (e.t < last) {
            bt_day<K, MULTI, T>(a, dv, xs, e, u, lane, j0, a.buf.action[env], reward, term, trunc);
        } else {
            valid = false;                       // the reference returns None here (:135)
            term = true;
        }
    } else if (a.mode == PDECG_BT_WRAP_RESET) {
        // TherapyWrapper.reset (:408-427); with a mask only the selected envs take part (the others keep their
        // consecutive-treatment-day counter and everything else: they are mid-therapy patients of the same batch)
        if (a.buf.mask != nullptr && a.buf.mask[env] == 0) {
            if (obs_all != nullptr) store_row_chunk<T, K, 32, SURE>(u, obs_all + env * nx, nx, sm, lane);
            return;
        }
        e.consecutive = 0;
        while (e.stage == PDECG_BT_GROWTH && e.t < last) {
            bt_day<K, MULTI, T>(a, dv, xs, e, u, lane, j0, 0.0, reward, term, trunc);
            if (term || trunc) break;
        }
    } else if (a.mode == PDECG_BT_RUN_TO_END) {
        // TherapyWrapper.benchmark (:480-505): zero-dose episode, then t_benchmark <- simulationDays
        while (!(term || trunc) && e.t < last) bt_day<K, MULTI, T>(a, dv, xs, e, u, lane, j0, 0.0, reward, term, trunc);
        e.t_benchmark = e.simulationDays;
    } else {
        // TherapyWrapper.step (:429-478)
        if (e.stage == PDECG_BT_POST) {
            while (!(term || trunc) && e.t < last) bt_day<K, MULTI, T>(a, dv, xs, e, u, lane, j0, 0.0, reward, term, trunc);
            if (e.t >= last && !(term || trunc)) { valid = false; term = true; }
        } else if (e.t < last) {
            const double control = a.buf.action[env];
            bt_day<K, MULTI, T>(a, dv, xs, e, u, lane, j0, control, reward, term, trunc);
            e.treatment_calls += 1;
            if (reward < 0.0) e.violations += 1;
            if (cfg.weekends) {
                e.consecutive = (control > 0.0) ? e.consecutive + 1 : 0;
                if (e.consecutive >= 5) {
                    e.consecutive = 0;
                    // the observation returned is the therapy day's row; the two rest days' results are
                    // discarded and the loop tests the STALE flags (:470-473)
                    if (obs_all != nullptr) {
                        store_row_chunk<T, K, 32, SURE>(u, obs_all + env * nx, nx, sm, lane);
                        wrote_obs = true;
                    }
                    for (int i = 0; i < 2; ++i) {
                        if (e.t < last) {
                            double r2;
                            bool te2, tr2;
                            bt_day<K, MULTI, T>(a, dv, xs, e, u, lane, j0, 0.0, r2, te2, tr2);
                        }
                        if (term || trunc) break;
                    }
                }
            }
        } else {
            valid = false;
            term = true;
        }
    }

    store_row_chunk<T, K, 32, SURE>(u, urow, nx, sm, lane);
    if (obs_all != nullptr && !wrote_obs) store_row_chunk<T, K, 32, SURE>(u, obs_all + env * nx, nx, sm, lane);

    // write back the scalars (lane 0: four + two 16-byte stores; the unused tail of the integer row keeps its contents)
    if (lane == 0) {
        int4* io = reinterpret_cast<int4*>(irow);
        io[0] = make_int4(e.t, e.stage, e.growthDays, e.therapyDays);
        io[1] = make_int4(e.postTherapyDays, e.simulationDays, e.firstTherapyDay, e.firstPostTherapyDay);
        io[2] = make_int4(e.cDeathDay, e.t_benchmark, e.consecutive, e.treatment_calls);
        io[3] = make_int4(e.violations, i3.y, i3.z, i3.w);
        double2* fo = reinterpret_cast<double2*>(frow);
        fo[0] = make_double2(e.remaining, f0.y);
        fo[1] = make_double2(f1.x, e.last_dose);
    }
    if (lane == 0) {
        a.buf.reward[env] = reward;
        a.buf.terminated[env] = term ? 1 : 0;
        a.buf.truncated[env] = trunc ? 1 : 0;
        if (a.buf.valid) a.buf.valid[env] = valid ? 1 : 0;
        if (a.buf.stage_out) a.buf.stage_out[env] = (!MULTI && a.ar_init != nullptr && (term || trunc)) ? PDECG_BT_GROWTH : e.stage;
    }
    if constexpr (!MULTI) {
        if (a.ar_init != nullptr && (term || trunc)) {
            if (a.buf.final_obs != nullptr)          // terminal observation for the caller, before the row is re-initialised
                store_row_chunk<T, K, 32, SURE>(u, reinterpret_cast<T*>(a.buf.final_obs) + env * nx, nx, sm, lane);
            bt_autoreset_env_t<T>(a, env, lane);
        }
    }
}

// BrainTumor1D.reset (:347-384) for one env, executed by its warp.  `keep_outputs`: leave reward / flags / valid alone
// (fused auto-reset: they keep the step's values).
template <typename T>
__device__ __forceinline__ void bt_reset_env(const BtArgs& a, const double* __restrict__ init, int64_t init_stride, int64_t env,
                                             int lane, bool keep_outputs) {
    const int nx = a.cfg.nx;
    const double* irow = init + env * init_stride;
    T* urow = reinterpret_cast<T*>(a.buf.u) + env * nx;
    for (int j = lane; j < nx; j += 32) {
        const T v = static_cast<T>(irow[j]);
        urow[j] = v;
        if (a.buf.obs) reinterpret_cast<T*>(a.buf.obs)[env * nx + j] = v;
    }
    // brain_tumor_env.py:363-379; t_benchmark, D, rho survive a reset
    int* ir = a.buf.ienv + env * PDECG_BT_NINT;
    if (lane < PDECG_BT_NINT) {
        int v = 0;
        if (lane == PDECG_BT_I_FIRST_THERAPY || lane == PDECG_BT_I_FIRST_POST || lane == PDECG_BT_I_DEATH_DAY) v = -1;
        if (lane == PDECG_BT_I_T_BENCHMARK || lane == PDECG_BT_I_TREATMENT_CALLS || lane == PDECG_BT_I_VIOLATIONS) v = ir[lane];
        ir[lane] = v;
    }
    if (lane == 0) {
        double* fr = a.buf.fenv + env * PDECG_BT_NF64;
        fr[PDECG_BT_F_REMAINING] = a.cfg.total_dosage;
        fr[PDECG_BT_F_LAST_DOSE] = 0.0;
        if (!keep_outputs) {
            a.buf.reward[env] = 0.0;
            a.buf.terminated[env] = 0;
            a.buf.truncated[env] = 0;
            if (a.buf.valid) a.buf.valid[env] = 1;
        }
    }
}

// fused same-step auto-reset: out of line so the (rare) restart does not weigh on the step kernel's registers
template <typename T>
__device__ __noinline__ void bt_autoreset_env_t(const BtArgs& a, int64_t env, int lane) {
    __syncwarp();                                  // the step's own row / scalar stores come first
    bt_reset_env<T>(a, a.ar_init, a.ar_stride, env, lane, true);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) bt_reset_kernel(const BtArgs a, const double* __restrict__ init, int64_t init_stride,
                                                            const uint8_t* __restrict__ mask) {
    const int lane = threadIdx.x & 31;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    if (env >= a.buf.n_envs) return;
    if (mask != nullptr && mask[env] == 0) return;
    bt_reset_env<T>(a, init, init_stride, env, lane, false);
}

static int bt_validate(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf) {
    if (!cfg || !buf) return fail(PDECG_ERR_INVALID, "null cfg/buffers");
    if (cfg->abi_version != PDECG_ABI_VERSION) return fail(PDECG_ERR_INVALID, "abi_version mismatch");
    if (cfg->nx < 4 || cfg->nt < 2) return fail(PDECG_ERR_INVALID, "need nx >= 4 and nt >= 2");
    if (cfg->nx > 512) return fail(PDECG_ERR_UNSUPPORTED, "grid of %d points exceeds the register-resident limit of 512 points per env", cfg->nx);
    if (!(cfg->dt > 0) || !(cfg->dx > 0) || !(cfg->k > 0)) return fail(PDECG_ERR_INVALID, "dt, dx, k must be positive");
    if (cfg->precision < PDECG_PREC_REF || cfg->precision > PDECG_PREC_FP32) return fail(PDECG_ERR_INVALID, "bad precision %d", cfg->precision);
    if (buf->n_envs < 0) return fail(PDECG_ERR_INVALID, "n_envs < 0");
    if (buf->n_envs > 0 && (!buf->u || !buf->ienv || !buf->fenv || !buf->reward || !buf->terminated || !buf->truncated || !buf->x_scale))
        return fail(PDECG_ERR_INVALID, "null device buffer");
    if (reinterpret_cast<uintptr_t>(buf->ienv) % 16 != 0 || reinterpret_cast<uintptr_t>(buf->fenv) % 16 != 0)
        return fail(PDECG_ERR_INVALID, "ienv / fenv must be 16-byte aligned");
    return PDECG_OK;
}

}  // namespace pdecg

using namespace pdecg;

extern "C" int pdecg_bt_reset(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, const double* init,
                              int64_t init_env_stride, const uint8_t* mask, pdecg_stream_t stream) {
    int rc = bt_validate(cfg, buf);
    if (rc != PDECG_OK) return rc;
    if (buf->n_envs == 0) return PDECG_OK;
    if (!init) return fail(PDECG_ERR_INVALID, "null init buffer");
    BtArgs a{*cfg, *buf, 0, nullptr, 0};
    const unsigned grid = static_cast<unsigned>((buf->n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (cfg->precision == PDECG_PREC_FP32) bt_reset_kernel<float><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, init, init_env_stride, mask);
    else bt_reset_kernel<double><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, init, init_env_stride, mask);
    count_launch();
    return check_launch("bt_reset_kernel");
}

static int bt_launch(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, int32_t mode, const double* ar_init, int64_t ar_stride,
                     pdecg_stream_t stream) {
    int rc = bt_validate(cfg, buf);
    if (rc != PDECG_OK) return rc;
    if (mode < PDECG_BT_RAW_STEP || mode > PDECG_BT_RUN_TO_END) return fail(PDECG_ERR_INVALID, "bad mode %d", mode);
    if (buf->n_envs == 0) return PDECG_OK;
    if ((mode == PDECG_BT_RAW_STEP || mode == PDECG_BT_WRAP_STEP) && !buf->action) return fail(PDECG_ERR_INVALID, "null action buffer");
    BtArgs a{*cfg, *buf, mode, ar_init, ar_stride};
    const unsigned grid = static_cast<unsigned>((buf->n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int K = (cfg->nx + 31) / 32;
    if (K > 8) K = (K + 1) / 2 * 2;                    // 257..512 points: even K in {10, 12, 14, 16}
    switch (K) {
#define PDECG_CASE(KK)                                                                   \
    case KK:                                                                             \
        if (cfg->precision == PDECG_PREC_FP32) {                                          \
            if (mode == PDECG_BT_RAW_STEP) bt_run_kernel<KK, false, float><<<grid, kThreads, 0, st>>>(a); \
            else bt_run_kernel<KK, true, float><<<grid, kThreads, 0, st>>>(a);            \
        } else if (mode == PDECG_BT_RAW_STEP) bt_run_kernel<KK, false, double><<<grid, kThreads, 0, st>>>(a); \
        else bt_run_kernel<KK, true, double><<<grid, kThreads, 0, st>>>(a);               \
        break;
        PDECG_CASE(1) PDECG_CASE(2) PDECG_CASE(3) PDECG_CASE(4) PDECG_CASE(5) PDECG_CASE(6) PDECG_CASE(7) PDECG_CASE(8)
        PDECG_CASE(10) PDECG_CASE(12) PDECG_CASE(14) PDECG_CASE(16)
#undef PDECG_CASE
        default: return fail(PDECG_ERR_UNSUPPORTED, "unsupported points-per-lane %d", K);
    }
    count_launch();
    return check_launch("bt_run_kernel");
}

extern "C" int pdecg_bt_run(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, int32_t mode, pdecg_stream_t stream) {
    return bt_launch(cfg, buf, mode, nullptr, 0, stream);
}

extern "C" int pdecg_bt_step_autoreset(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, const double* init,
                                       int64_t init_env_stride, pdecg_stream_t stream) {
    if (!init) return fail(PDECG_ERR_INVALID, "null init buffer");
    return bt_launch(cfg, buf, PDECG_BT_RAW_STEP, init, init_env_stride, stream);
}

// Host-buffer raw step: H2D(actions) -> kernel -> D2H(state = observation, reward, flags) on `stream`, then synchronised.
extern "C" int pdecg_bt_step_host(const pdecg_bt_cfg* cfg, const pdecg_bt_buffers* buf, const double* action_host,
                                  double* obs_host, double* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                                  uint8_t* valid_host, const double* autoreset_init, int64_t autoreset_init_env_stride,
                                  pdecg_stream_t stream) {
    int rc = bt_validate(cfg, buf);
    if (rc != PDECG_OK) return rc;
    if (buf->n_envs == 0) return PDECG_OK;
    if (!action_host || !buf->action) return fail(PDECG_ERR_INVALID, "null action buffer");
    if (valid_host && !buf->valid) return fail(PDECG_ERR_INVALID, "valid_host needs a device valid buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n = static_cast<size_t>(buf->n_envs);
    cudaError_t e = cudaMemcpyAsync(const_cast<double*>(buf->action), action_host, n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "H2D actions: %s", cudaGetErrorString(e));
    if ((rc = bt_launch(cfg, buf, PDECG_BT_RAW_STEP, autoreset_init, autoreset_init_env_stride, stream)) != PDECG_OK) return rc;
    const size_t esz = cfg->precision == PDECG_PREC_FP32 ? sizeof(float) : sizeof(double);
    if (obs_host && (e = cudaMemcpyAsync(obs_host, buf->u, n * cfg->nx * esz, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H obs: %s", cudaGetErrorString(e));
    if (reward_host && (e = cudaMemcpyAsync(reward_host, buf->reward, n * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H reward: %s", cudaGetErrorString(e));
    if (terminated_host && (e = cudaMemcpyAsync(terminated_host, buf->terminated, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H terminated: %s", cudaGetErrorString(e));
    if (truncated_host && (e = cudaMemcpyAsync(truncated_host, buf->truncated, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H truncated: %s", cudaGetErrorString(e));
    if (valid_host && (e = cudaMemcpyAsync(valid_host, buf->valid, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H valid: %s", cudaGetErrorString(e));
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
    return PDECG_OK;
}
