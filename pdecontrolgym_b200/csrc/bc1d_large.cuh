// bc1d_large.cuh -- TransportPDE1D / ReactionDiffusionPDE1D step kernels for grids beyond the register-resident
// limit of the warp kernels (more than 512 points per env): ONE CTA = ONE ENV.
//
// Reference behaviour restated here (no code shared): hyperbolic.py:126-227, parabolic.py:126-221,
// base_env_1d.py:20-45, tuned_reward_1d.py:25-40, norm_reward.py:29-73 of lukebhan/PDEControlGym.
//
// Two kernels, same arithmetic, operation order and reward / flag semantics as the warp kernels (bc1d_kernels.cuh):
// states are bit-identical to numpy, and HBM still carries one row in and out per control step.
//   bc1d_step_cta_reg_kernel : each thread keeps K contiguous points in registers and updates them in place; one
//                              (first, last) pair per thread crosses threads per substep through shared memory
//                              (default; 1.8-2.1e12 grid-point updates/s on a B200).
//   bc1d_step_cta_kernel     : the row lives in shared memory (two ping-pong copies + the coefficient row), thread t
//                              updates the points t, t + THREADS, ...; used for ReactionDiffusion with Neumann
//                              control, where the actuated value changes every substep.
// One block barrier per substep in both.  Included at the end of bc1d_kernels.cuh.
#pragma once
#include <type_traits>

namespace pdecg {

constexpr int kCtaMaxThreads = 1024;
constexpr int kCtaMaxPoints = 8192;      // 2 x 8 B state copies + 8 B coefficient per point = 192 KB of the 227 KB per CTA

// Sum / NaN-propagating max over the CTA in a fixed order (warp xor tree, then the warps' partial results in
// ascending order): deterministic from launch to launch.  `red` holds 32 doubles.  Every thread gets the result.
__device__ __forceinline__ double cta_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();                                   // protects `red` against the previous reduction's readers
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) tot += red[w];
    return tot;
}
__device__ __forceinline__ double cta_nanmax(double v, double* red) {
    v = seg_nanmax<32>(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) tot = nanmax(tot, red[w]);
    return tot;
}

template <typename S>
__device__ __forceinline__ double cta_sumsq(const S* row, int npts, double* red) {
    double s = 0.0;
    for (int j = threadIdx.x; j < npts; j += blockDim.x) {
        const double v = static_cast<double>(row[j]);
        s += v * v;
    }
    return cta_sum(s, red);
}

// ||row||_ord (or ||row - other||_ord when `other` is given; the difference is formed in the state dtype like numpy)
template <typename S>
__device__ __forceinline__ double cta_norm(const S* row, const S* other, int npts, int ord, double* red) {
    double acc = 0.0;
    for (int j = threadIdx.x; j < npts; j += blockDim.x) {
        const S d = other ? Strict<S>::sub(row[j], other[j]) : row[j];
        const double v = static_cast<double>(d);
        if (ord == kNormL2) acc += v * v;
        else if (ord == kNormL1) acc += fabs(v);
        else acc = nanmax(acc, fabs(v));
    }
    if (ord == kNormL2) return sqrt(cta_sum(acc, red));
    if (ord == kNormL1) return cta_sum(acc, red);
    return cta_nanmax(acc, red);
}

template <int ENV, class M>
__global__ void __launch_bounds__(kCtaMaxThreads, 1) bc1d_step_cta_kernel(const __grid_constant__ Bc1dArgs a) {
    using S = typename M::S;
    using C = typename M::C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    const pdecg_bc1d_cfg& cfg = a.cfg;
    const int npts = a.npts;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int64_t env = blockIdx.x;
    C* coef = reinterpret_cast<C*>(smem_raw);                         // RD: dt*beta_j ; Transport: beta_j
    S* cur = reinterpret_cast<S*>(coef + npts);
    S* nxt = cur + npts;

    const S Fs = static_cast<S>(cfg.dt / (cfg.dx * cfg.dx));          // parabolic.py:138 (python float -> S)
    const S dxs = static_cast<S>(cfg.dx);
    const C dtc = static_cast<C>(cfg.dt);
    const ConstDiv<S> by_dx(dxs);
    const int jb = npts - 1;                                          // actuated point (x = X; ghost point for RD)

    S* urow = static_cast<S*>(a.buf.u) + env * npts;
    const C* brow = static_cast<const C*>(a.buf.beta) + env * a.buf.beta_env_stride;
    for (int j = tid; j < npts; j += nth) {
        cur[j] = urow[j];
        const C b = brow[j];
        if constexpr (ENV == PDECG_ENV_REACTION_DIFFUSION) coef[j] = M::kFast ? dtc * b : Strict<C>::mul(dtc, b);   // parabolic.py:144
        else coef[j] = b;
    }
    __syncthreads();

    const int t0 = a.buf.time_index[env];
    int nsub = cfg.nt - 1 - t0;
    nsub = nsub < 0 ? 0 : (nsub > cfg.substeps ? cfg.substeps : nsub);
    const int t_end = t0 + nsub;
    double* carry = a.buf.carry + env * a.carry_len;

    const bool dyn = (ENV == PDECG_ENV_REACTION_DIFFUSION) && cfg.control_type == PDECG_CONTROL_NEUMANN;   // parabolic.py:148-150
    S g = actuate<S>(a, env, S(0));                                   // constant over the step unless `dyn`
    double bsum_dyn = 0.0;

    // reward bookkeeping that needs mid-step rows (block-uniform; same definitions as the warp kernel)
    const bool tuned = cfg.reward_kind == PDECG_REWARD_TUNED1D;
    const bool ring_every = tuned && a.tuned_general;
    const int snap_at = (tuned && !ring_every && cfg.substeps > kLookback) ? nsub - kLookback : -1;
    const bool want_diff = !tuned && cfg.horizon == PDECG_HORIZON_DIFFERENTIAL && nsub > 0;
    const bool want_hor = !tuned && cfg.horizon == PDECG_HORIZON_T;
    const int hor_m = t_end > cfg.t_horizon_length ? cfg.t_horizon_length : t_end;   // norm_reward.py:65-72
    const int hor_from = nsub - hor_m + 1;
    double lb_norm = 0.0, diff_norm = 0.0, hor_acc = 0.0;

    if (snap_at == 0) lb_norm = sqrt(cta_sumsq<S>(cur, npts, red));
    for (int done = 0; done < nsub;) {
        if (dyn) {
            g = actuate<S>(a, env, cur[jb - 1]);
            bsum_dyn += fabs(static_cast<double>(g));
        }
        if constexpr (ENV == PDECG_ENV_REACTION_DIFFUSION) {
            for (int j = tid; j < npts; j += nth) {
                S n;
                if (j == 0) n = S(0);                                 // parabolic.py:146
                else if (j == jb) n = g;
                else n = rd_point<M, false>(cur[j - 1], cur[j], cur[j + 1], Fs, coef[j]);
                nxt[j] = n;
            }
        } else {
            const S p0 = cur[0];
            for (int j = tid; j < npts; j += nth) {
                S n;
                if (j == jb) {
                    n = g;                                            // hyperbolic.py:143-145
                } else {
                    const S p = cur[j];
                    S dq;
                    if constexpr (M::kFast) dq = (cur[j + 1] - p) * by_dx.y;
                    else dq = by_dx(Strict<S>::sub(cur[j + 1], p));
                    n = tr_point<M>(p, dq, p0, dtc, coef[j]);
                }
                nxt[j] = n;
            }
        }
        __syncthreads();
        S* t = cur; cur = nxt; nxt = t;
        ++done;
        // ---- events at the current row (after `done` substeps) ----
        if (done == snap_at) lb_norm = sqrt(cta_sumsq<S>(cur, npts, red));
        if (ring_every) {                          // general look-back: norm of every row, ring indexed by t mod 101
            const double nrm = static_cast<double>(static_cast<S>(sqrt(cta_sumsq<S>(cur, npts, red))));
            if (tid == 0) carry[1 + (t0 + done) % (kLookback + 1)] = nrm;
        }
        if (want_hor && done >= hor_from)
            hor_acc += static_cast<double>(static_cast<S>(cta_norm<S>(cur, nullptr, npts, cfg.norm_ord, red)));
    }
    if (want_diff) diff_norm = cta_norm<S>(cur, nxt, npts, cfg.norm_ord, red);   // `nxt` still holds row t-1

    // ---- write back state ------------------------------------------------------------------------
    const bool c32 = a.buf.obs_copy_f32 != 0;
    for (int j = tid; j < npts; j += nth) {
        const S v = cur[j];
        if (nsub > 0) urow[j] = v;
        if (a.buf.obs_copy) put_obs<S>(a.buf.obs_copy, c32, env * npts + j, v);
    }
    const double sumsq = cta_sumsq<S>(cur, npts, red);
    const S norm2 = static_cast<S>(sqrt(sumsq));               // np.linalg.norm(u[t]) in the state dtype
    const bool terminated = t_end >= cfg.nt - 1;                // hyperbolic.py:171-180
    // per-env blow-up flag (no counterpart in the reference, where inf / NaN simply propagate): the row holds an inf / NaN, or
    // values so large that their squares overflow float64; optionally folded into `truncated`
    const bool nonfinite = !(sumsq <= 1.7976931348623157e308);
    const bool truncated = (cfg.limit_pde_state_size && (norm2 >= static_cast<S>(cfg.max_state_value))) || (cfg.truncate_nonfinite && nonfinite);
    double other_norm = 0.0;                                    // NormReward temporal with p != 2
    if (!tuned && !terminated && !truncated && (cfg.horizon == PDECG_HORIZON_TEMPORAL || nsub == 0) && cfg.norm_ord != kNormL2)
        other_norm = cta_norm<S>(cur, nullptr, npts, cfg.norm_ord, red);

    if (tid == 0) {
        double bsum = carry[0];
        if (dyn) bsum += bsum_dyn;
        else bsum += static_cast<double>(nsub) * fabs(static_cast<double>(g));
        double reward;
        if (tuned) {
            if (ring_every) {
                lb_norm = (t_end - kLookback >= 0) ? carry[1 + (t_end - kLookback) % (kLookback + 1)] : 0.0;
            } else if (cfg.substeps <= kLookback && cfg.substeps >= 1) {
                const int slot = 1 + ((t_end / cfg.substeps) % a.ring_len);
                lb_norm = carry[slot];
                if (nsub > 0) carry[slot] = static_cast<double>(norm2);
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
                S nn = (cfg.norm_ord == kNormL2) ? norm2 : static_cast<S>(other_norm);
                reward = static_cast<double>(Strict<S>::div(-nn, coeff));
            } else if (cfg.horizon == PDECG_HORIZON_DIFFERENTIAL) {
                reward = static_cast<double>(Strict<S>::div(static_cast<S>(diff_norm), coeff));
            } else {
                S mean = Strict<S>::div(static_cast<S>(hor_acc), static_cast<S>(hor_m));
                reward = static_cast<double>(Strict<S>::div(-mean, coeff));
            }
        }
        // ---- sensing (hyperbolic.py:69-120, parabolic.py:69-118) ----
        if (cfg.sensing_loc != PDECG_SENSE_FULL) {
            S o;
            if (cfg.sensing_loc == PDECG_SENSE_COLLOCATED) {
                if (cfg.control_type == PDECG_CONTROL_NEUMANN) o = cur[npts - 1];
                else o = Strict<S>::div(Strict<S>::sub(cur[npts - 1], cur[npts - 2]), dxs);
            } else {
                if (cfg.sensing_type == PDECG_SENSETYPE_NEUMANN) o = Strict<S>::div(Strict<S>::sub(cur[1], cur[0]), dxs);
                else o = cur[0];
            }
            static_cast<S*>(a.buf.obs)[env] = o;
            if (a.ar.enabled && a.ar.final_obs != nullptr && (terminated || truncated)) put_obs<S>(a.ar.final_obs, c32, env, o);
        }
        carry[0] = bsum;
        a.buf.time_index[env] = t_end;
        a.buf.reward[env] = reward;
        a.buf.terminated[env] = terminated ? 1 : 0;
        a.buf.truncated[env] = truncated ? 1 : 0;
        if (a.buf.nonfinite != nullptr) a.buf.nonfinite[env] = nonfinite ? 1 : 0;
    }
    if (a.ar.enabled && (terminated || truncated)) {
        if (a.ar.final_obs != nullptr && cfg.sensing_loc == PDECG_SENSE_FULL)     // terminal observation for the caller
            for (int j = tid; j < npts; j += nth) put_obs<S>(a.ar.final_obs, c32, env * npts + j, cur[j]);
        __syncthreads();                               // the step's own row / carry stores come first
        if (tid < 32) autoreset_env<S, 32>(a, env, tid, kFull);   // warp 0 restarts the env exactly like the reset kernel
    }
}

// ---- register-tiled variant ------------------------------------------------------------------------------------
// Thread t owns the K contiguous points [K*t, K*t + K) in registers and updates them IN PLACE (ascending order with
// the old left value carried along), so a substep costs one 2-value halo exchange through shared memory (double
// buffered: one block barrier per substep) instead of five shared-memory accesses and the index arithmetic per point.
// Used for every large-grid configuration except ReactionDiffusion with Neumann control (actuation re-evaluated
// from u[nx-1] every substep), which stays on the shared-memory kernel above.
template <typename S> struct Pair { S first, last; };

template <int ENV, class M, int K>
__global__ void __launch_bounds__(K <= 4 ? 512 : 1024, K <= 4 ? 2 : 1) bc1d_step_cta_reg_kernel(const __grid_constant__ Bc1dArgs a) {
    using S = typename M::S;
    using C = typename M::C;
    constexpr int kMaxThreads = K <= 4 ? 512 : 1024;
    __shared__ Pair<S> edge[2][kMaxThreads + 2];       // [parity][1 + thread]: (first, last) value of each thread's chunk
    __shared__ double red[32];
    __shared__ S sens[4];                              // u[0], u[1], u[npts-2], u[npts-1] of the final row
    const pdecg_bc1d_cfg& cfg = a.cfg;
    const int npts = a.npts;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int64_t env = blockIdx.x;
    const int j0 = tid * K;
    const int jb = npts - 1;                           // actuated point (x = X; ghost point for RD)

    const S Fs = static_cast<S>(cfg.dt / (cfg.dx * cfg.dx));          // parabolic.py:138 (python float -> S)
    const S dxs = static_cast<S>(cfg.dx);
    const C dtc = static_cast<C>(cfg.dt);
    const ConstDiv<S> by_dx(dxs);

    S* urow = static_cast<S*>(a.buf.u) + env * npts;
    const C* brow = static_cast<const C*>(a.buf.beta) + env * a.buf.beta_env_stride;
    S u[K];
    C coef[K];                                         // RD: dt*beta_j ; Transport: beta_j
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = j0 + k;
        // Transport padding slots (never stored, never read by valid points) hold a ramp instead of zeros so that the
        // differences formed there are ordinary numbers and stay on the division fast path
        u[k] = (j < npts) ? urow[j] : ((ENV == PDECG_ENV_TRANSPORT) ? static_cast<S>(j - npts + 1) : S(0));
        const C b = (j < npts) ? brow[j] : C(0);
        if constexpr (ENV == PDECG_ENV_REACTION_DIFFUSION) coef[k] = M::kFast ? dtc * b : Strict<C>::mul(dtc, b);   // parabolic.py:144
        else coef[k] = b;
    }
    if (tid == 0) {
        edge[0][0] = edge[1][0] = Pair<S>{S(0), S(0)};
        edge[0][nth + 1] = edge[1][nth + 1] = Pair<S>{S(0), S(0)};
    }

    const int t0 = a.buf.time_index[env];
    int nsub = cfg.nt - 1 - t0;
    nsub = nsub < 0 ? 0 : (nsub > cfg.substeps ? cfg.substeps : nsub);
    const int t_end = t0 + nsub;
    double* carry = a.buf.carry + env * a.carry_len;
    const S g = actuate<S>(a, env, S(0));              // constant over the step
    const int owner = jb / K;                          // thread holding the actuated point
    const bool owner_warp = (tid >> 5) == (owner >> 5);

    // reward bookkeeping that needs mid-step rows (block-uniform; same definitions as the warp kernel)
    const bool tuned = cfg.reward_kind == PDECG_REWARD_TUNED1D;
    const bool ring_every = tuned && a.tuned_general;
    const int snap_at = (tuned && !ring_every && cfg.substeps > kLookback) ? nsub - kLookback : -1;
    const bool want_diff = !tuned && cfg.horizon == PDECG_HORIZON_DIFFERENTIAL && nsub > 0;
    const bool want_hor = !tuned && cfg.horizon == PDECG_HORIZON_T;
    const int hor_m = t_end > cfg.t_horizon_length ? cfg.t_horizon_length : t_end;   // norm_reward.py:65-72
    const int hor_from = nsub - hor_m + 1;
    double lb_norm = 0.0, diff_norm = 0.0, hor_acc = 0.0;

    auto row_sumsq = [&]() -> double {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double v = (j0 + k < npts) ? static_cast<double>(u[k]) : 0.0;
            s += v * v;
        }
        return cta_sum(s, red);
    };
    auto row_norm = [&](int ord) -> double {
        if (ord == kNormL2) return sqrt(row_sumsq());
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (j0 + k < npts) {
                const double v = fabs(static_cast<double>(u[k]));
                acc = (ord == kNormL1) ? acc + v : nanmax(acc, v);
            }
        }
        return (ord == kNormL1) ? cta_sum(acc, red) : cta_nanmax(acc, red);
    };
    // one substep in place; DIFF: also accumulate ||new - old||_ord over the valid points (NormReward differential)
    auto substep = [&](int par, auto diff_tag, double& dacc) {
        constexpr bool DIFF = decltype(diff_tag)::value;
        edge[par][1 + tid] = Pair<S>{u[0], u[K - 1]};
        __syncthreads();
        const S hr = edge[par][2 + tid].first;
        if constexpr (ENV == PDECG_ENV_REACTION_DIFFUSION) {
            S left = edge[par][tid].last;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const S old = u[k];
                const S pr = (k == K - 1) ? hr : u[k < K - 1 ? k + 1 : K - 1];
                S n = rd_point<M, false>(left, old, pr, Fs, coef[k]);
                if (k == 0) n = (tid == 0) ? S(0) : n;                 // parabolic.py:146
                if (owner_warp) n = (j0 + k == jb) ? g : n;
                if constexpr (DIFF) {
                    if (j0 + k < npts) {
                        const double dv = fabs(static_cast<double>(Strict<S>::sub(n, old)));
                        dacc = (cfg.norm_ord == kNormL2) ? dacc + dv * dv : ((cfg.norm_ord == kNormL1) ? dacc + dv : nanmax(dacc, dv));
                    }
                }
                u[k] = n;
                left = old;
            }
        } else {
            const S p0 = edge[par][1].first;
            // (p_{j+1} - p_j) / dx for the whole chunk from the OLD values: 3-instruction exact tail per point and ONE
            // range test per chunk; operands outside the window (zero, denormal, inf, ...) redo the chunk exactly
            S x[K], dq[K];
            bool ok = true;
            [[maybe_unused]] float xmin = 0x1p100f, xmax = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const S pr = (k == K - 1) ? hr : u[k < K - 1 ? k + 1 : K - 1];
                if constexpr (M::kFast) {
                    dq[k] = (pr - u[k]) * by_dx.y;
                } else {
                    x[k] = Strict<S>::sub(pr, u[k]);
                    dq[k] = by_dx.fast(x[k]);
                    if constexpr (sizeof(S) == 4) {
                        xmin = fminf(xmin, fabsf(x[k]));
                        xmax = fmaxf(xmax, fabsf(x[k]));
                    } else {
                        ok = ok & by_dx.in_range(x[k]);
                    }
                }
            }
            if constexpr (!M::kFast) {
                if constexpr (sizeof(S) == 4) ok = by_dx.in_range_minmax(xmin, xmax);
                if (!ok) {
#pragma unroll
                    for (int k = 0; k < K; ++k) dq[k] = by_dx.exact(x[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const S old = u[k];
                S n = tr_point<M>(old, dq[k], p0, dtc, coef[k]);
                if (owner_warp) n = (j0 + k == jb) ? g : n;            // hyperbolic.py:143-145
                if constexpr (DIFF) {
                    if (j0 + k < npts) {
                        const double dv = fabs(static_cast<double>(Strict<S>::sub(n, old)));
                        dacc = (cfg.norm_ord == kNormL2) ? dacc + dv * dv : ((cfg.norm_ord == kNormL1) ? dacc + dv : nanmax(dacc, dv));
                    }
                }
                u[k] = n;
            }
        }
    };

    double dacc = 0.0;
    int done = 0;
    while (true) {
        // ---- events at the current row (after `done` substeps) ----
        if (done == snap_at) lb_norm = sqrt(row_sumsq());
        if (ring_every && done >= 1) {             // general look-back: norm of every row, ring indexed by t mod 101
            const double nrm = static_cast<double>(static_cast<S>(sqrt(row_sumsq())));
            if (tid == 0) carry[1 + (t0 + done) % (kLookback + 1)] = nrm;
        }
        if (want_hor && done >= hor_from && done >= 1) hor_acc += static_cast<double>(static_cast<S>(row_norm(cfg.norm_ord)));
        if (done >= nsub) break;
        if (want_diff && done == nsub - 1) {       // the last substep also accumulates ||u_t - u_{t-1}||
            substep(done & 1, std::true_type{}, dacc);
            ++done;
            continue;
        }
        // ---- next stop ----
        int next = nsub;
        if (snap_at > done && snap_at < next) next = snap_at;
        if (ring_every) next = done + 1;
        if (want_diff && nsub - 1 < next) next = nsub - 1;
        if (want_hor) {
            const int h = hor_from > done + 1 ? hor_from : done + 1;
            if (h < next) next = h;
        }
        // plain substeps up to the next stop (tight loop: nothing but the stencil and the halo exchange)
        for (int s = done; s < next; ++s) substep(s & 1, std::false_type{}, dacc);
        done = next;
    }
    if (want_diff) {
        if (cfg.norm_ord == kNormL2) diff_norm = sqrt(cta_sum(dacc, red));
        else if (cfg.norm_ord == kNormL1) diff_norm = cta_sum(dacc, red);
        else diff_norm = cta_nanmax(dacc, red);
    }

    // ---- write back state ------------------------------------------------------------------------
    const bool c32 = a.buf.obs_copy_f32 != 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = j0 + k;
        if (j < npts) {
            if (nsub > 0) urow[j] = u[k];
            if (a.buf.obs_copy) put_obs<S>(a.buf.obs_copy, c32, env * npts + j, u[k]);
            if (j == 0) sens[0] = u[k];
            if (j == 1) sens[1] = u[k];
            if (j == npts - 2) sens[2] = u[k];
            if (j == npts - 1) sens[3] = u[k];
        }
    }
    const double sumsq = row_sumsq();                          // (its barriers also publish `sens`)
    const S norm2 = static_cast<S>(sqrt(sumsq));               // np.linalg.norm(u[t]) in the state dtype
    const bool terminated = t_end >= cfg.nt - 1;                // hyperbolic.py:171-180
    // per-env blow-up flag (no counterpart in the reference, where inf / NaN simply propagate): the row holds an inf / NaN, or
    // values so large that their squares overflow float64; optionally folded into `truncated`
    const bool nonfinite = !(sumsq <= 1.7976931348623157e308);
    const bool truncated = (cfg.limit_pde_state_size && (norm2 >= static_cast<S>(cfg.max_state_value))) || (cfg.truncate_nonfinite && nonfinite);
    double other_norm = 0.0;                                    // NormReward temporal with p != 2
    if (!tuned && !terminated && !truncated && (cfg.horizon == PDECG_HORIZON_TEMPORAL || nsub == 0) && cfg.norm_ord != kNormL2)
        other_norm = row_norm(cfg.norm_ord);

    if (tid == 0) {
        double bsum = carry[0] + static_cast<double>(nsub) * fabs(static_cast<double>(g));
        double reward;
        if (tuned) {
            if (ring_every) {
                lb_norm = (t_end - kLookback >= 0) ? carry[1 + (t_end - kLookback) % (kLookback + 1)] : 0.0;
            } else if (cfg.substeps <= kLookback && cfg.substeps >= 1) {
                const int slot = 1 + ((t_end / cfg.substeps) % a.ring_len);
                lb_norm = carry[slot];
                if (nsub > 0) carry[slot] = static_cast<double>(norm2);
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
                S nn = (cfg.norm_ord == kNormL2) ? norm2 : static_cast<S>(other_norm);
                reward = static_cast<double>(Strict<S>::div(-nn, coeff));
            } else if (cfg.horizon == PDECG_HORIZON_DIFFERENTIAL) {
                reward = static_cast<double>(Strict<S>::div(static_cast<S>(diff_norm), coeff));
            } else {
                S mean = Strict<S>::div(static_cast<S>(hor_acc), static_cast<S>(hor_m));
                reward = static_cast<double>(Strict<S>::div(-mean, coeff));
            }
        }
        // ---- sensing (hyperbolic.py:69-120, parabolic.py:69-118) ----
        if (cfg.sensing_loc != PDECG_SENSE_FULL) {
            S o;
            if (cfg.sensing_loc == PDECG_SENSE_COLLOCATED) {
                if (cfg.control_type == PDECG_CONTROL_NEUMANN) o = sens[3];
                else o = Strict<S>::div(Strict<S>::sub(sens[3], sens[2]), dxs);
            } else {
                if (cfg.sensing_type == PDECG_SENSETYPE_NEUMANN) o = Strict<S>::div(Strict<S>::sub(sens[1], sens[0]), dxs);
                else o = sens[0];
            }
            static_cast<S*>(a.buf.obs)[env] = o;
            if (a.ar.enabled && a.ar.final_obs != nullptr && (terminated || truncated)) put_obs<S>(a.ar.final_obs, c32, env, o);
        }
        carry[0] = bsum;
        a.buf.time_index[env] = t_end;
        a.buf.reward[env] = reward;
        a.buf.terminated[env] = terminated ? 1 : 0;
        a.buf.truncated[env] = truncated ? 1 : 0;
        if (a.buf.nonfinite != nullptr) a.buf.nonfinite[env] = nonfinite ? 1 : 0;
    }
    if (a.ar.enabled && (terminated || truncated)) {
        if (a.ar.final_obs != nullptr && cfg.sensing_loc == PDECG_SENSE_FULL) {   // terminal observation for the caller
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (j0 + k < npts) put_obs<S>(a.ar.final_obs, c32, env * npts + j0 + k, u[k]);
        }
        __syncthreads();                               // the step's own row / carry stores come first
        if (tid < 32) autoreset_env<S, 32>(a, env, tid, kFull);   // warp 0 restarts the env exactly like the reset kernel
    }
}

template <int ENV, class M>
inline int launch_cta(const Bc1dArgs& a, cudaStream_t st) {
    using S = typename M::S;
    using C = typename M::C;
    if (a.npts > kCtaMaxPoints)
        return fail(PDECG_ERR_UNSUPPORTED, "grid of %d points exceeds the shared-memory-resident limit of %d points per env", a.npts, kCtaMaxPoints);
    const bool dyn = ENV == PDECG_ENV_REACTION_DIFFUSION && a.cfg.control_type == PDECG_CONTROL_NEUMANN;
    if (!dyn && !getenv("PDECG_BC1D_LARGE_SMEM")) {
        // register-tiled: 4 points per thread up to 2,048 points (<= 512 threads), 8 beyond (<= 1,024 threads)
        const unsigned grid = static_cast<unsigned>(a.buf.n_envs);
        if (a.npts <= 2048) {
            const int threads = ((a.npts + 3) / 4 + 31) / 32 * 32;
            bc1d_step_cta_reg_kernel<ENV, M, 4><<<grid, threads, 0, st>>>(a);
        } else {
            const int threads = ((a.npts + 7) / 8 + 31) / 32 * 32;
            bc1d_step_cta_reg_kernel<ENV, M, 8><<<grid, threads, 0, st>>>(a);
        }
        count_launch();
        return check_launch("bc1d_step_cta_reg_kernel");
    }
    const size_t smem = static_cast<size_t>(a.npts) * (sizeof(C) + 2 * sizeof(S));
    auto kern = bc1d_step_cta_kernel<ENV, M>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const int threads = a.npts <= 1024 ? 256 : (a.npts <= 4096 ? 512 : 1024);
    kern<<<static_cast<unsigned>(a.buf.n_envs), threads, smem, st>>>(a);
    count_launch();
    return check_launch("bc1d_step_cta_kernel");
}

}  // namespace pdecg
