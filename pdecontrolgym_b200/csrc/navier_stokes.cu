// navier_stokes.cu -- NavierStokes2D step / reset / pressure-solve kernels + C ABI (sm_100a, float64).
//
// Reference behaviour restated here (no code shared): environments2d/navier_stokes2D.py:9-194,
// base_env_2d.py:24-63 and rewards/ns_reward.py:18-28 of lukebhan/PDEControlGym.
//
// Generic kernel: one CTA = one env; u, v, u*, v*, p (ping-pong) live in shared memory for the whole
// control step (predictor -> boundary conditions -> N Jacobi sweeps -> corrector -> boundary
// conditions -> reward); HBM sees each field once in and once out.  No tensor cores: nothing here is
// a dense contraction.
//
// Jacobi sweep identity used below (derived from navier_stokes2D.py:105-113, order of the four
// Neumann copies included): after a sweep every boundary value equals the NEW value of the nearest
// interior point (index clamped to [1, n-2] in both directions).  Each thread therefore evaluates the
// Jacobi formula at the clamped index of its own point -- boundary points recompute their neighbour's
// value (same inputs, same operation order => same bits) instead of waiting for a second barrier.
#include <cstdlib>
#include <type_traits>

#include "pdecg_internal.h"

namespace pdecg {

struct NsArgs {
    pdecg_ns_cfg cfg;
    pdecg_ns_buffers buf;
    // solve-only entry point
    const double* su;
    const double* sv;
    double* sp;
    int solve_only;
    // pdecg_ns_step_autoreset: finished envs restart inside the step launch (ar_u == nullptr: off)
    const double* ar_u;
    const double* ar_v;
    const double* ar_p;
    int64_t ar_stride;
};

// NavierStokes2D.reset (:171-194) for one env, executed by `nth` cooperating threads (a CTA or a warp).
// `keep_outputs`: leave reward / flags / valid alone (fused auto-reset: they keep the step's values).
// T = double: the reference's arithmetic (bit-identical fields).  T = float: the opt-in fast mode (PDECG_PREC_FP32):
// float32 fields, observation and arithmetic in the same operation order; actions, the reference trajectory, the reward
// and the initial conditions handed to reset stay float64.
template <typename T>
__device__ __forceinline__ void ns_reset_env(const NsArgs& a, const double* __restrict__ iu, const double* __restrict__ iv,
                                             const double* __restrict__ ip, int64_t stride, int64_t env, int tid, int nth,
                                             bool keep_outputs) {
    const int npt = a.cfg.nx * a.cfg.ny;
    T* const bu = reinterpret_cast<T*>(a.buf.u);
    T* const bv = reinterpret_cast<T*>(a.buf.v);
    T* const bp = reinterpret_cast<T*>(a.buf.p);
    T* const bo = reinterpret_cast<T*>(a.buf.obs);
    for (int q = tid; q < npt; q += nth) {
        const T un = static_cast<T>(iu[env * stride + q]), vn = static_cast<T>(iv[env * stride + q]);
        bu[env * npt + q] = un;
        bv[env * npt + q] = vn;
        bp[env * npt + q] = static_cast<T>(ip[env * stride + q]);
        bo[(env * npt + q) * 2] = un;
        bo[(env * npt + q) * 2 + 1] = vn;
    }
    if (tid == 0) {
        a.buf.time_index[env] = 0;
        if (!keep_outputs) {
            a.buf.reward[env] = 0.0;
            a.buf.terminated[env] = 0;
            a.buf.truncated[env] = 0;
            if (a.buf.valid) a.buf.valid[env] = 1;
        }
    }
}

// navier_stokes_large.cu: the HBM-resident path for grids beyond kNsMaxPoints
size_t ns_large_workspace_bytes(const pdecg_ns_cfg* cfg, int64_t n_envs);
int ns_large_step(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* ar_u, const double* ar_v, const double* ar_p,
                  int64_t ar_stride, cudaStream_t st);
int ns_large_solve_pressure(const pdecg_ns_cfg* cfg, int64_t n_envs, const double* u, const double* v, double* p, void* work, cudaStream_t st);

using D = Strict<double>;
constexpr int kNsThreads = 256;
constexpr int kNsMaxPoints = 4096;   // 7 fields x 4096 x 8 B = 224 KB of the 227 KB an sm_100a CTA may use

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// central_difference (:9-15) and laplace (:17-22) at an INTERIOR point of a smem field with row pitch nx
template <typename T>
__device__ __forceinline__ T cd_x(const T* f, int i, int j, int nx, T two_dx) {
    return Strict<T>::div(Strict<T>::sub(f[i * nx + j + 1], f[i * nx + j - 1]), two_dx);
}
template <typename T>
__device__ __forceinline__ T cd_y(const T* f, int i, int j, int nx, T two_dy) {
    return Strict<T>::div(Strict<T>::sub(f[(i + 1) * nx + j], f[(i - 1) * nx + j]), two_dy);
}
template <typename T>
__device__ __forceinline__ T lap5(const T* f, int i, int j, int nx, T dxdy) {
    using S = Strict<T>;
    const T c = f[i * nx + j];
    T s = S::add(f[i * nx + j - 1], f[(i - 1) * nx + j]);
    s = S::sub(s, S::mul(T(4), c));
    s = S::add(s, f[i * nx + j + 1]);
    s = S::add(s, f[(i + 1) * nx + j]);
    return S::div(s, dxdy);
}

// apply_boundary (:68-91): sides in the reference's order, u before v; a barrier between sides keeps the
// corner values order-dependent exactly like the sequential numpy assignments.
template <typename T>
__device__ void apply_boundary(const pdecg_ns_cfg& cfg, T* fu, T* fv, T action, int tid, int nthreads) {
    const int nx = cfg.nx, ny = cfg.ny;
    for (int side = 0; side < 4; ++side) {
        const int len = (side < 2) ? nx : ny;
        for (int comp = 0; comp < 2; ++comp) {
            T* f = comp == 0 ? fu : fv;
            const int cond = cfg.bc[side][comp];
            for (int q = tid; q < len; q += nthreads) {
                int dst, src;
                if (side == PDECG_NS_LOWER) { dst = q; src = nx + q; }
                else if (side == PDECG_NS_UPPER) { dst = (ny - 1) * nx + q; src = (ny - 2) * nx + q; }
                else if (side == PDECG_NS_LEFT) { dst = q * nx; src = q * nx + 1; }
                else { dst = q * nx + nx - 1; src = q * nx + nx - 2; }
                if (cond == PDECG_NS_NEUMANN) f[dst] = f[src];
                else if (cond == PDECG_NS_DIRICHLET) f[dst] = T(0);
                else f[dst] = action;
            }
        }
        __syncthreads();
    }
}

// N Jacobi sweeps (:104-114) on smem ping-pong buffers pa (current) / pb; returns the buffer holding the result.
// `b` = (dx*dy) * rhs, per point (only interior entries are read).
template <int PT, typename T>
__device__ __forceinline__ T* jacobi(const pdecg_ns_cfg& cfg, T* pa, T* pb, const T* b, int tid, int nthreads) {
    using S = Strict<T>;
    const int nx = cfg.nx, ny = cfg.ny, npt = nx * ny;
    // each thread owns up to PT points (npt <= PT*nthreads is enforced on the host)
    int idx[PT], ci[PT];
    T bb[PT];
#pragma unroll
    for (int s = 0; s < PT; ++s) {
        const int q = tid + s * nthreads;
        idx[s] = q < npt ? q : -1;
        const int i = clampi((q < npt ? q : 0) / nx, 1, ny - 2), j = clampi((q < npt ? q : 0) % nx, 1, nx - 2);
        ci[s] = i * nx + j;
        bb[s] = b[ci[s]];
    }
    for (int it = 0; it < cfg.n_pressure_iters; ++it) {
#pragma unroll
        for (int s = 0; s < PT; ++s) {
            if (idx[s] >= 0) {
                const int c = ci[s];
                T acc = S::add(pa[c - 1], pa[c - nx]);     // ((left + up) + right) + down
                acc = S::add(acc, pa[c + 1]);
                acc = S::add(acc, pa[c + nx]);
                pb[idx[s]] = S::mul(T(0.25), S::sub(acc, bb[s]));
            }
        }
        __syncthreads();
        T* t = pa; pa = pb; pb = t;
    }
    return pa;
}

// 0.25 * (acc - b) in the reference's two roundings (float64); the float32 fast mode folds it into one FMA on b' = -b/4
__device__ __forceinline__ double jacobi_update(double acc, double b) { return D::mul(0.25, D::sub(acc, b)); }
__device__ __forceinline__ float jacobi_update(float acc, float bq) { return __fmaf_rn(acc, 0.25f, bq); }
__device__ __forceinline__ double jacobi_rhs(double b) { return b; }
__device__ __forceinline__ float jacobi_rhs(float b) { return __fmul_rn(-0.25f, b); }

// The same N sweeps with the pressure in REGISTERS (grids up to 82 x 82): the interior is tiled over the CTA as 16 thread
// columns x (2 per warp) thread rows of 5 x BC points.  By the sweep identity above only the interior is iterated, with
// clamped neighbours at the domain edge; inside a block neighbours are registers, left / right come from 16-lane
// shuffles, up / down from the partner half-warp by shuffle and from the adjacent warps through a double-buffered edge
// array in shared memory -- one block barrier and no other shared-memory traffic per sweep (the generic form above reads
// four neighbours and writes one value per point per sweep through shared memory: 2.5 Jacobi points per clock per SM at
// 41 x 41, against the FP64 pipe's 6.4).  The first sweep reads the true boundary values of the warm-start pressure from
// shared memory; the result goes back to `sp` (interior, then the boundary copies).  blockDim.x = 32 * ceil(ceil((ny-2)/5)/2).
constexpr int kBlkRows = 5;
template <int BC, typename T>
__device__ __forceinline__ void jacobi_blk(const pdecg_ns_cfg& cfg, T* sp, const T* sb, T* edges, int tid, int nthreads) {
    using S = Strict<T>;
    constexpr int BR = kBlkRows, EW = 16 * BC;
    const int nx = cfg.nx, ny = cfg.ny, npt = nx * ny, NXI = nx - 2, NYI = ny - 2;
    if (cfg.n_pressure_iters < 1) return;
    const int warp = tid >> 5, nwarps = nthreads >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const bool odd = (ty & 1) != 0;
    const int i0 = ty * BR, j0 = tx * BC;
    T p[BR][BC], n[BR][BC], b[BR][BC];
#pragma unroll
    for (int rr = 0; rr < BR; ++rr)
#pragma unroll
        for (int cc = 0; cc < BC; ++cc) {
            const int i = i0 + rr, j = j0 + cc;
            const bool ok = i < NYI && j < NXI;
            const int q = ok ? (i + 1) * nx + (j + 1) : nx + 1;
            b[rr][cc] = ok ? jacobi_rhs(sb[q]) : T(0);
            T acc = S::add(sp[q - 1], sp[q - nx]);          // sweep 1: true neighbours, incl. the stored boundary values
            acc = S::add(acc, sp[q + 1]);
            acc = S::add(acc, sp[q + nx]);
            p[rr][cc] = ok ? jacobi_update(acc, b[rr][cc]) : T(0);
        }
    // clamps at the domain edge: the neighbour beyond it is the point itself.  Top / left sit on a block boundary (select
    // on the halo); the last interior row / column may sit at any slot of one thread row / column.
    const int lb = NYI - 1, lr = NXI - 1;
    const bool is_top = ty == 0, is_left = tx == 0;
    const int sb_slot = (lb / BR == ty) ? lb - ty * BR : -1, sr_slot = (lr / BC == tx) ? lr - tx * BC : -1;
    auto sweep = [&](const T (&src)[BR][BC], T (&dst)[BR][BC], int par) {
        T up_h[BC], dn_h[BC], lf_h[BR], rt_h[BR];
        T* e_mine = edges + ((par * nwarps + warp) * 2) * EW;                  // [top | bottom] rows of this warp
#pragma unroll
        for (int cc = 0; cc < BC; ++cc) {
            const T pub = odd ? src[0][cc] : src[BR - 1][cc];                 // what the partner thread row needs from me
            const T got = __shfl_xor_sync(kFull, pub, 16);
            if (odd) e_mine[EW + tx * BC + cc] = src[BR - 1][cc];              // my bottom row -> warp + 1's up halo
            else e_mine[tx * BC + cc] = src[0][cc];                            // my top row    -> warp - 1's down halo
            up_h[cc] = got;
            dn_h[cc] = got;
        }
        __syncthreads();
        {
            const T* e_up = edges + ((par * nwarps + (warp > 0 ? warp - 1 : 0)) * 2 + 1) * EW + tx * BC;              // bottom row of the warp above
            const T* e_dn = edges + ((par * nwarps + (warp < nwarps - 1 ? warp + 1 : nwarps - 1)) * 2) * EW + tx * BC;   // top row of the warp below
#pragma unroll
            for (int cc = 0; cc < BC; ++cc) {
                const T far = odd ? e_dn[cc] : e_up[cc];
                if (odd) dn_h[cc] = far; else up_h[cc] = far;
                if (is_top) up_h[cc] = src[0][cc];
            }
        }
#pragma unroll
        for (int rr = 0; rr < BR; ++rr) {
            lf_h[rr] = __shfl_up_sync(kFull, src[rr][BC - 1], 1, 16);
            rt_h[rr] = __shfl_down_sync(kFull, src[rr][0], 1, 16);
            if (is_left) lf_h[rr] = src[rr][0];
        }
#pragma unroll
        for (int rr = 0; rr < BR; ++rr)
#pragma unroll
            for (int cc = 0; cc < BC; ++cc) {
                const T self = src[rr][cc];
                const T up = (rr > 0) ? src[rr > 0 ? rr - 1 : 0][cc] : up_h[cc];
                T dn = (rr < BR - 1) ? src[rr < BR - 1 ? rr + 1 : 0][cc] : dn_h[cc];
                const T lf = (cc > 0) ? src[rr][cc > 0 ? cc - 1 : 0] : lf_h[rr];
                T rt = (cc < BC - 1) ? src[rr][cc < BC - 1 ? cc + 1 : 0] : rt_h[rr];
                dn = (sb_slot == rr) ? self : dn;
                rt = (sr_slot == cc) ? self : rt;
                T acc = S::add(lf, up);                      // ((left + up) + right) + down
                acc = S::add(acc, rt);
                acc = S::add(acc, dn);
                dst[rr][cc] = jacobi_update(acc, b[rr][cc]);
            }
    };
    int it = 1;
    for (; it + 2 <= cfg.n_pressure_iters; it += 2) {       // register ping-pong
        sweep(p, n, 0);
        sweep(n, p, 1);
    }
    if (it < cfg.n_pressure_iters) {
        sweep(p, n, 0);
#pragma unroll
        for (int rr = 0; rr < BR; ++rr)
#pragma unroll
            for (int cc = 0; cc < BC; ++cc) p[rr][cc] = n[rr][cc];
    }
    __syncthreads();                                         // every thread has read what it needs of the old sp
#pragma unroll
    for (int rr = 0; rr < BR; ++rr)
#pragma unroll
        for (int cc = 0; cc < BC; ++cc) {
            const int i = i0 + rr, j = j0 + cc;
            if (i < NYI && j < NXI) sp[(i + 1) * nx + (j + 1)] = p[rr][cc];
        }
    __syncthreads();
    // boundary values = nearest interior value (the four Neumann copies of :110-113, in their order)
    for (int q = tid; q < npt; q += nthreads) {
        const int i = q / nx, j = q - i * nx;
        if (i < 1 || i > ny - 2 || j < 1 || j > nx - 2) sp[q] = sp[clampi(i, 1, ny - 2) * nx + clampi(j, 1, nx - 2)];
    }
    __syncthreads();
}

// THREADS x PT >= nx*ny: (256, 2) up to 512 points, (512, 4) up to 2,048, (1024, 4) up to 4,096 (64 x 64; the seven
// float64 fields of one env then fill the SM's 227 KB of shared memory)
// BC = 0: the shared-memory Jacobi above (any thread count).  BC = 1..5: the register-blocked Jacobi (jacobi_blk) with BC
// points per thread per row; the launch then uses 32 * ceil(ceil((ny-2)/5)/2) <= THREADS threads and no pong buffer.
template <int THREADS, int PT, typename T, int BC = 0>
__global__ void __launch_bounds__(THREADS) ns_step_kernel(const NsArgs a) {
    using S = Strict<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const pdecg_ns_cfg& cfg = a.cfg;
    const int nx = cfg.nx, ny = cfg.ny, npt = nx * ny;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int64_t env = blockIdx.x;
    // BC = 0: seven fields.  BC > 0: four -- u / v are dead after the predictor, so the right-hand side and the pressure
    // take their place, and the corrector updates u* / v* in place (more CTAs per SM: 54 KB instead of 94 KB at 41 x 41).
    T* su = smem;                 // u  -> u_next                      | BC > 0: u, then (dx*dy) * rhs
    T* sv = su + npt;             // v  -> v_next                      | BC > 0: v, then the pressure
    T* sus = sv + npt;            // u* (predictor)                    | BC > 0: ... -> u_next
    T* svs = sus + npt;           // v*                                | BC > 0: ... -> v_next
    T* sp0 = (BC > 0) ? sv : svs + npt;           // pressure (ping)
    T* sp1 = sp0 + npt;                           // pressure pong (BC = 0 only)
    T* sb = (BC > 0) ? su : sp1 + npt;            // (dx*dy) * rhs
    T* edges = svs + npt;         // BC > 0: [2][warps][2][16 * BC] rows exchanged between adjacent warps
    T* fu = (BC > 0) ? sus : su;  // where u_next / v_next end up
    T* fv = (BC > 0) ? svs : sv;
    (void)sp1; (void)edges;
    __shared__ double red[THREADS / 32];

    const T dx = static_cast<T>(cfg.dx), dy = static_cast<T>(cfg.dy), dt = static_cast<T>(cfg.dt);
    const T nu = static_cast<T>(cfg.viscosity), rho = static_cast<T>(cfg.density);
    const T two_dx = S::mul(T(2), dx), two_dy = S::mul(T(2), dy);
    const T dxdy = S::mul(dx, dy);

    if (a.solve_only) {
        if constexpr (sizeof(T) == 8) {          // NavierStokes2D.solve_pressure alone (float64 callers only)
            const double* gu = a.su + env * npt;
            const double* gv = a.sv + env * npt;
            double* gp = a.sp + env * npt;
            for (int q = tid; q < npt; q += nth) { sus[q] = gu[q]; svs[q] = gv[q]; sp0[q] = gp[q]; sb[q] = 0.0; }
            __syncthreads();
            const double rho_dt = D::div(cfg.density, cfg.dt);
            for (int q = tid; q < npt; q += nth) {
                const int i = q / nx, j = q % nx;
                if (i >= 1 && i <= ny - 2 && j >= 1 && j <= nx - 2)
                    sb[q] = D::mul(dxdy, D::mul(rho_dt, D::add(cd_x(sus, i, j, nx, two_dx), cd_y(svs, i, j, nx, two_dy))));
            }
            __syncthreads();
            double* pf;
            if constexpr (BC > 0) { jacobi_blk<BC, double>(cfg, sp0, sb, edges, tid, nth); pf = sp0; }
            else pf = jacobi<PT, double>(cfg, sp0, sp1, sb, tid, nth);
            for (int q = tid; q < npt; q += nth) gp[q] = pf[q];
        }
        return;
    }

    const int t0 = a.buf.time_index[env];
    T* const obs_all = reinterpret_cast<T*>(a.buf.obs);
    T* const fobs_all = reinterpret_cast<T*>(a.buf.final_obs);
    if (t0 + 1 >= cfg.nt || t0 + 1 >= cfg.n_uref || t0 + 1 >= cfg.n_action_ref) {
        // the reference raises IndexError here (U[time_index] / U_ref / action_ref out of range)
        if (tid == 0) {
            if (a.buf.valid) a.buf.valid[env] = 0;
            a.buf.terminated[env] = 1;
            a.buf.truncated[env] = 0;
            a.buf.reward[env] = 0.0;
        }
        if (a.ar_u != nullptr) {
            if (fobs_all != nullptr) {        // terminal observation = the (unchanged) current U[t]
                for (int q = tid; q < 2 * npt; q += nth) fobs_all[env * npt * 2 + q] = obs_all[env * npt * 2 + q];
                __syncthreads();
            }
            ns_reset_env<T>(a, a.ar_u, a.ar_v, a.ar_p, a.ar_stride, env, tid, nth, true);
        }
        return;
    }
    T* gu = reinterpret_cast<T*>(a.buf.u) + env * npt;
    T* gv = reinterpret_cast<T*>(a.buf.v) + env * npt;
    T* gp = reinterpret_cast<T*>(a.buf.p) + env * npt;
    const double action_d = a.buf.action[env];
    const T action = static_cast<T>(action_d);
    for (int q = tid; q < npt; q += nth) {
        su[q] = gu[q];
        sv[q] = gv[q];
        if constexpr (BC == 0) sp0[q] = gp[q];
    }
    __syncthreads();

    // ---- predictor (:130-138) ----
    for (int q = tid; q < npt; q += nth) {
        const int i = q / nx, j = q % nx;
        T us = su[q], vs = sv[q];
        if (i >= 1 && i <= ny - 2 && j >= 1 && j <= nx - 2) {
            const T u0 = su[q], v0 = sv[q];
            // u + dt * (((-u)*dudx - v*dudy) + nu*lap(u))
            T adv = S::sub(S::mul(-u0, cd_x(su, i, j, nx, two_dx)), S::mul(v0, cd_y(su, i, j, nx, two_dy)));
            us = S::add(u0, S::mul(dt, S::add(adv, S::mul(nu, lap5(su, i, j, nx, dxdy)))));
            adv = S::sub(S::mul(-u0, cd_x(sv, i, j, nx, two_dx)), S::mul(v0, cd_y(sv, i, j, nx, two_dy)));
            vs = S::add(v0, S::mul(dt, S::add(adv, S::mul(nu, lap5(sv, i, j, nx, dxdy)))));
        }
        sus[q] = us;
        svs[q] = vs;
    }
    __syncthreads();
    apply_boundary<T>(cfg, sus, svs, action, tid, nth);

    // ---- pressure Poisson (:94-116) ----
    const T rho_dt = S::div(rho, dt);
    for (int q = tid; q < npt; q += nth) {
        const int i = q / nx, j = q % nx;
        T bv = T(0);
        if (i >= 1 && i <= ny - 2 && j >= 1 && j <= nx - 2)
            bv = S::mul(dxdy, S::mul(rho_dt, S::add(cd_x(sus, i, j, nx, two_dx), cd_y(svs, i, j, nx, two_dy))));
        sb[q] = bv;
        if constexpr (BC > 0) sp0[q] = gp[q];            // the warm-start pressure takes v's place
    }
    __syncthreads();
    T* pf;
    if constexpr (BC > 0) { jacobi_blk<BC, T>(cfg, sp0, sb, edges, tid, nth); pf = sp0; }
    else pf = jacobi<PT, T>(cfg, sp0, sp1, sb, tid, nth);

    // ---- corrector (:143-146) ----
    const T dt_rho = S::div(dt, rho);
    for (int q = tid; q < npt; q += nth) {
        const int i = q / nx, j = q % nx;
        T un = sus[q], vn = svs[q];
        if (i >= 1 && i <= ny - 2 && j >= 1 && j <= nx - 2) {
            un = S::sub(un, S::mul(dt_rho, cd_x(pf, i, j, nx, two_dx)));
            vn = S::sub(vn, S::mul(dt_rho, cd_y(pf, i, j, nx, two_dy)));
        }
        fu[q] = un;
        fv[q] = vn;
    }
    __syncthreads();
    apply_boundary<T>(cfg, fu, fv, action, tid, nth);

    // ---- store, observation U[t], reward (:147-157, ns_reward.py:28) ----
    const int t1 = t0 + 1;
    const double* uref = a.buf.u_ref + static_cast<int64_t>(t1) * npt * 2;
    T* gobs = obs_all + env * npt * 2;
    double ss = 0.0;
    for (int q = tid; q < npt; q += nth) {
        const T un = fu[q], vn = fv[q];
        gu[q] = un;
        gv[q] = vn;
        gp[q] = pf[q];
        gobs[2 * q] = un;
        gobs[2 * q + 1] = vn;
        const double du = D::sub(static_cast<double>(un), uref[2 * q]), dv = D::sub(static_cast<double>(vn), uref[2 * q + 1]);
        ss += du * du;
        ss += dv * dv;
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < (nth >> 5); ++w) tot += red[w];
        const double nrm = sqrt(tot);
        // -1/2 * norm**2 / shape[1] / shape[2] - gamma/2 * norm(a - a_ref[t])**2
        double r1 = D::div(D::div(D::mul(-0.5, D::mul(nrm, nrm)), static_cast<double>(nx)), static_cast<double>(ny));   // / shape[1] / shape[2]
        const double da = fabs(D::sub(action_d, a.buf.action_ref[t1]));
        const double r2 = D::mul(D::div(cfg.gamma, 2.0), D::mul(da, da));
        a.buf.reward[env] = D::sub(r1, r2);
        a.buf.time_index[env] = t1;
        a.buf.terminated[env] = (t1 >= cfg.nt - 1) ? 1 : 0;      // :159-168
        a.buf.truncated[env] = 0;
        if (a.buf.valid) a.buf.valid[env] = 1;
    }
    if (a.ar_u != nullptr && t1 >= cfg.nt - 1) {          // same-step auto-reset (block-uniform condition)
        if (fobs_all != nullptr) {                        // terminal observation U[nt-1] for the caller
            T* fobs = fobs_all + env * npt * 2;
            for (int q = tid; q < npt; q += nth) { fobs[2 * q] = fu[q]; fobs[2 * q + 1] = fv[q]; }
        }
        __syncthreads();                                  // the step's own field / observation stores come first
        ns_reset_env<T>(a, a.ar_u, a.ar_v, a.ar_p, a.ar_stride, env, tid, nth, true);
    }
}

template <typename T>
__global__ void __launch_bounds__(kNsThreads) ns_reset_kernel(const NsArgs a, const double* __restrict__ iu, const double* __restrict__ iv,
                                                              const double* __restrict__ ip, int64_t stride, const uint8_t* __restrict__ mask) {
    const int64_t env = blockIdx.x;
    if (mask != nullptr && mask[env] == 0) return;
    ns_reset_env<T>(a, iu, iv, ip, stride, env, threadIdx.x, blockDim.x, false);
}

// ------------------------------------------------------------------------------------------------
// Specialised kernel for the reference grid (nx = ny = N+2): a (half-)WARP = ONE ENV, Jacobi in registers.
//
// The N x N interior is tiled over the LPE lanes of an env as an RG x CG grid of BR x BC blocks:
//   LPE = 16 (two envs per warp), N = 19: 4 x 4 lanes, 5 x 5 points each -- 400 lane-slots for 361 points, 20 halo
//            values per env-sweep.  One warp-sweep (two envs) is 125 DADD/DMUL + 40 SHFL + 40 FSEL.
//   LPE = 32 (one env per warp),   N = 19: 10 x 3 lanes, 2 x 7 points each -- 448 lane-slots (30 of 32 lanes), 18 halo
//            values: 70 DADD/DMUL + 36 SHFL + 36 FSEL per env-sweep (the round-1 mapping, kept for A/B runs).
// The FP64 pipe issues 16 lanes per clock per SM sub-partition, so the half-warp mapping needs 125 x 2 = 250 pipe cycles
// per TWO envs against 70 x 2 = 140 per ONE: 1.12x fewer FP64 slots per env, 1.8x fewer shuffles -- and at the BASELINE
// batch (1,024 envs = 512 warps on 592 sub-partitions) no sub-partition hosts more than one warp, where the one-env
// mapping put two warps on most of them (7 warps per SM).
// By the sweep identity above, after the first sweep every boundary value is a copy of the nearest interior value, so
// sweeps 2..n only need the interior with CLAMPED neighbours: inside a block neighbours are registers, across blocks
// they come from 2(BR+BC) f64 shuffles per sweep (segment-wide: width = LPE), and the clamps are selects on the halo
// values under four lane-uniform predicates.  No barrier and no shared-memory traffic inside the 2000-sweep loop:
// 5 DADD/DMUL per point per sweep.  The first sweep reads the true boundary values of the warm-start pressure from
// shared memory, so an arbitrary initial pressure field is handled exactly.
// All 32 lanes run the whole kernel in lock step whatever their env's state: a segment whose env does not exist or
// is past its last step computes on zeros and stores nothing, so every shuffle / barrier names the full warp.
template <int N, int BR, int BC, int RG, int CG, int WPB, int LPE, typename T>
__global__ void __launch_bounds__(WPB * 32) ns_step_warp_kernel(const NsArgs a) {
    static_assert(LPE == 32 || LPE == 16, "32 or 16 lanes per env");
    static_assert(RG * CG <= LPE && RG * BR >= N && CG * BC >= N, "bad tiling");
    constexpr int NX = N + 2, NPT = NX * NX, EPW = 32 / LPE;
    using S = Strict<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const pdecg_ns_cfg& cfg = a.cfg;
    const int lane = threadIdx.x & (LPE - 1), wib = threadIdx.x >> 5;
    const int seg = (threadIdx.x & 31) / LPE;
    const unsigned segmask = seg_mask<LPE>(threadIdx.x & 31);
    const int64_t env0 = (static_cast<int64_t>(blockIdx.x) * WPB + wib) * EPW;
    if (env0 >= a.buf.n_envs) return;                          // warp-uniform
    const int64_t env = env0 + seg;
    const bool exists = env < a.buf.n_envs;
    T* f0 = smem + static_cast<size_t>(wib * EPW + seg) * 4 * NPT;   // u   -> (dx*dy)*rhs
    T* f1 = f0 + NPT;                                     // v   -> pressure
    T* f2 = f1 + NPT;                                     // u*  -> u_next
    T* f3 = f2 + NPT;                                     // v*  -> v_next

    const int t0 = exists ? a.buf.time_index[env] : 0;
    const bool stale = exists && (t0 + 1 >= cfg.nt || t0 + 1 >= cfg.n_uref || t0 + 1 >= cfg.n_action_ref);
    const bool live = exists && !stale;
    if (stale) {                                               // the reference raises IndexError here (segment-uniform)
        if (lane == 0) {
            if (a.buf.valid) a.buf.valid[env] = 0;
            a.buf.terminated[env] = 1;
            a.buf.truncated[env] = 0;
            a.buf.reward[env] = 0.0;
        }
        if (a.ar_u != nullptr) {
            if (a.buf.final_obs != nullptr) {        // terminal observation = the (unchanged) current U[t]
                for (int q = lane; q < 2 * NPT; q += LPE)
                    reinterpret_cast<T*>(a.buf.final_obs)[env * NPT * 2 + q] = reinterpret_cast<T*>(a.buf.obs)[env * NPT * 2 + q];
                __syncwarp(segmask);
            }
            ns_reset_env<T>(a, a.ar_u, a.ar_v, a.ar_p, a.ar_stride, env, lane, LPE, true);
        }
    }
    __syncwarp();
    if (!__any_sync(kFull, live)) return;                      // warp-uniform
    const T dx = static_cast<T>(cfg.dx), dy = static_cast<T>(cfg.dy), dt = static_cast<T>(cfg.dt);
    const T nu = static_cast<T>(cfg.viscosity), rho = static_cast<T>(cfg.density);
    const T two_dx = S::mul(T(2), dx), two_dy = S::mul(T(2), dy);
    const T dxdy = S::mul(dx, dy);
    const ConstDiv<T> by_2dx(two_dx), by_2dy(two_dy), by_dxdy(dxdy);
    const int64_t envc = live ? env : 0;                       // pointers of a non-live segment are never dereferenced
    T* gu = reinterpret_cast<T*>(a.buf.u) + envc * NPT;
    T* gv = reinterpret_cast<T*>(a.buf.v) + envc * NPT;
    T* gp = reinterpret_cast<T*>(a.buf.p) + envc * NPT;
    const double action_d = live ? a.buf.action[env] : 0.0;
    const T action = static_cast<T>(action_d);

    for (int q = lane; q < NPT; q += LPE) { f0[q] = live ? gu[q] : T(0); f1[q] = live ? gv[q] : T(0); }
    __syncwarp();
    auto interior = [](int q) { const int i = q / NX, j = q - i * NX; return i >= 1 && i <= N && j >= 1 && j <= N; };
    auto cdx = [&](const T* f, int q) { return by_2dx(S::sub(f[q + 1], f[q - 1])); };
    auto cdy = [&](const T* f, int q) { return by_2dy(S::sub(f[q + NX], f[q - NX])); };
    auto lap = [&](const T* f, int q) {
        T s = S::add(f[q - 1], f[q - NX]);
        s = S::sub(s, S::mul(T(4), f[q]));
        s = S::add(s, f[q + 1]);
        s = S::add(s, f[q + NX]);
        return by_dxdy(s);
    };
    // ---- predictor (:130-138) ----
    for (int q = lane; q < NPT; q += LPE) {
        T us = f0[q], vs = f1[q];
        if (interior(q)) {
            const T u0 = us, v0 = vs;
            T adv = S::sub(S::mul(-u0, cdx(f0, q)), S::mul(v0, cdy(f0, q)));
            us = S::add(u0, S::mul(dt, S::add(adv, S::mul(nu, lap(f0, q)))));
            adv = S::sub(S::mul(-u0, cdx(f1, q)), S::mul(v0, cdy(f1, q)));
            vs = S::add(v0, S::mul(dt, S::add(adv, S::mul(nu, lap(f1, q)))));
        }
        f2[q] = us;
        f3[q] = vs;
    }
    __syncwarp();
    // apply_boundary (:68-91): sides in the reference's order, u before v
    auto apply_bc = [&](T* fu, T* fv) {
        for (int side = 0; side < 4; ++side) {
            for (int comp = 0; comp < 2; ++comp) {
                T* f = comp == 0 ? fu : fv;
                const int cond = cfg.bc[side][comp];
                for (int q = lane; q < NX; q += LPE) {
                    int dst, src;
                    if (side == PDECG_NS_LOWER) { dst = q; src = NX + q; }
                    else if (side == PDECG_NS_UPPER) { dst = (NX - 1) * NX + q; src = (NX - 2) * NX + q; }
                    else if (side == PDECG_NS_LEFT) { dst = q * NX; src = q * NX + 1; }
                    else { dst = q * NX + NX - 1; src = q * NX + NX - 2; }
                    if (cond == PDECG_NS_NEUMANN) f[dst] = f[src];
                    else if (cond == PDECG_NS_DIRICHLET) f[dst] = T(0);
                    else f[dst] = action;
                }
            }
            __syncwarp();
        }
    };
    apply_bc(f2, f3);
    // ---- (dx*dy)*rhs into f0, warm-start pressure into f1 ----
    const T rho_dt = S::div(rho, dt);
    for (int q = lane; q < NPT; q += LPE) {
        f0[q] = interior(q) ? S::mul(dxdy, S::mul(rho_dt, S::add(cdx(f2, q), cdy(f3, q)))) : T(0);
        f1[q] = live ? gp[q] : T(0);
    }
    __syncwarp();

    // ---- Jacobi sweeps (:104-114) on registers ----
    const int rg = lane / CG, cg = lane - rg * CG;
    const int i0 = rg * BR, j0 = cg * BC;                    // interior coordinates of this lane's block
    if (cfg.n_pressure_iters >= 1) {
        T p[BR][BC], b[BR][BC];
#pragma unroll
        for (int rr = 0; rr < BR; ++rr)
#pragma unroll
            for (int cc = 0; cc < BC; ++cc) {
                const int i = i0 + rr, j = j0 + cc;
                const bool ok = (lane < RG * CG) && i < N && j < N;
                const int q = ok ? (i + 1) * NX + (j + 1) : NX + 1;
                // float32 mode: b holds -(dx*dy)*rhs/4 so that the update is one FMA (4 instead of 5 operations per point-sweep)
                b[rr][cc] = ok ? (sizeof(T) == 4 ? S::mul(T(-0.25), f0[q]) : f0[q]) : T(0);
                // sweep 1: true neighbours (incl. stored boundary values) from shared memory
                T acc = S::add(f1[q - 1], f1[q - NX]);
                acc = S::add(acc, f1[q + 1]);
                acc = S::add(acc, f1[q + NX]);
                p[rr][cc] = ok ? jacobi_update(acc, b[rr][cc]) : T(0);
            }
        constexpr int RLAST = (N - 1) / BR, RSLOT = (N - 1) % BR;     // block / slot holding interior row N-1
        constexpr int CLAST = (N - 1) / BC, CSLOT = (N - 1) % BC;
        const bool top = rg == 0, bottom = rg == RLAST, leftmost = cg == 0, rightmost = cg == CLAST;
        // Halo values of a block (the neighbouring blocks' edge rows / columns); the clamp at the domain edge is a select
        // on the halo: the neighbour beyond the edge is the point itself.
        struct Halo { T up[BC], dn[BC], lf[BR], rt[BR]; };
        auto exchange_rows = [&](const T (&src)[BR][BC], Halo& h) {
#pragma unroll
            for (int cc = 0; cc < BC; ++cc) {
                h.up[cc] = __shfl_up_sync(kFull, src[BR - 1][cc], CG, LPE);
                h.dn[cc] = __shfl_down_sync(kFull, src[0][cc], CG, LPE);
                h.up[cc] = top ? src[0][cc] : h.up[cc];
                if constexpr (RSLOT == BR - 1) h.dn[cc] = bottom ? src[BR - 1][cc] : h.dn[cc];
            }
        };
        auto exchange_cols = [&](const T (&src)[BR][BC], Halo& h) {
#pragma unroll
            for (int rr = 0; rr < BR; ++rr) {
                h.lf[rr] = __shfl_up_sync(kFull, src[rr][BC - 1], 1, LPE);
                h.rt[rr] = __shfl_down_sync(kFull, src[rr][0], 1, LPE);
                h.lf[rr] = leftmost ? src[rr][0] : h.lf[rr];
                if constexpr (CSLOT == BC - 1) h.rt[rr] = rightmost ? src[rr][BC - 1] : h.rt[rr];
            }
        };
        auto exchange = [&](const T (&src)[BR][BC], Halo& h) { exchange_rows(src, h); exchange_cols(src, h); };
        auto point = [&](const T (&src)[BR][BC], const Halo& h, int rr, int cc) {
            const T self = src[rr][cc];
            const T up = (rr > 0) ? src[rr > 0 ? rr - 1 : 0][cc] : h.up[cc];
            T dn = (rr < BR - 1) ? src[rr < BR - 1 ? rr + 1 : 0][cc] : h.dn[cc];
            if constexpr (RSLOT < BR - 1) { if (rr == RSLOT) dn = bottom ? self : dn; }
            const T lf = (cc > 0) ? src[rr][cc > 0 ? cc - 1 : 0] : h.lf[rr];
            T rt = (cc < BC - 1) ? src[rr][cc < BC - 1 ? cc + 1 : 0] : h.rt[rr];
            if constexpr (CSLOT < BC - 1) { if (cc == CSLOT) rt = rightmost ? self : rt; }
            T acc = S::add(lf, up);             // ((left + up) + right) + down
            acc = S::add(acc, rt);
            acc = S::add(acc, dn);
            return jacobi_update(acc, b[rr][cc]);
        };
        // One sweep `src` (with its halo `hs`) -> `dst` (and its halo `hd`).  The block's EDGE points come first and
        // their shuffles are issued at once; the interior points, which need no halo, are computed while the shuffles
        // are in flight: the sweep-to-sweep recurrence (edge values -> shuffle -> edge values) overlaps with them.
        auto sweep = [&](const T (&src)[BR][BC], const Halo& hs, T (&dst)[BR][BC], Halo& hd) {
#pragma unroll
            for (int rr = 0; rr < BR; ++rr)
#pragma unroll
                for (int cc = 0; cc < BC; ++cc)
                    if (rr == 0 || rr == BR - 1 || cc == 0 || cc == BC - 1) dst[rr][cc] = point(src, hs, rr, cc);
            exchange(dst, hd);
#pragma unroll
            for (int rr = 1; rr < BR - 1; ++rr)
#pragma unroll
                for (int cc = 1; cc < BC - 1; ++cc) dst[rr][cc] = point(src, hs, rr, cc);
        };
        T n[BR][BC];
        Halo hp, hn;
        if (cfg.n_pressure_iters > 1) exchange(p, hp);
        int it = 1;
        for (; it + 2 <= cfg.n_pressure_iters; it += 2) {   // register ping-pong: no copies in the steady state
            sweep(p, hp, n, hn);
            sweep(n, hn, p, hp);
        }
        if (it < cfg.n_pressure_iters) {
            sweep(p, hp, n, hn);
#pragma unroll
            for (int rr = 0; rr < BR; ++rr)
#pragma unroll
                for (int cc = 0; cc < BC; ++cc) p[rr][cc] = n[rr][cc];
        }
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < BR; ++rr)
#pragma unroll
            for (int cc = 0; cc < BC; ++cc) {
                const int i = i0 + rr, j = j0 + cc;
                if ((lane < RG * CG) && i < N && j < N) f1[(i + 1) * NX + (j + 1)] = p[rr][cc];
            }
        __syncwarp();
        // boundary values = nearest interior value (the four Neumann copies of :110-113, in their order)
        for (int q = lane; q < NPT; q += LPE) {
            if (!interior(q)) {
                const int i = q / NX, j = q - i * NX;
                f1[q] = f1[clampi(i, 1, N) * NX + clampi(j, 1, N)];
            }
        }
        __syncwarp();
    }

    // ---- corrector (:143-146) ----
    const T dt_rho = S::div(dt, rho);
    for (int q = lane; q < NPT; q += LPE) {
        if (interior(q)) {
            f2[q] = S::sub(f2[q], S::mul(dt_rho, cdx(f1, q)));
            f3[q] = S::sub(f3[q], S::mul(dt_rho, cdy(f1, q)));
        }
    }
    __syncwarp();
    apply_bc(f2, f3);

    // ---- store, observation U[t], reward ----
    const int t1 = t0 + 1;
    const double* uref = a.buf.u_ref + static_cast<int64_t>(live ? t1 : 0) * NPT * 2;
    T* gobs = reinterpret_cast<T*>(a.buf.obs) + envc * NPT * 2;
    using V2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
    double ss = 0.0;
    if (live) {
        for (int q = lane; q < NPT; q += LPE) {
            const T un = f2[q], vn = f3[q];
            gu[q] = un;
            gv[q] = vn;
            gp[q] = f1[q];
            const double du = D::sub(static_cast<double>(un), uref[2 * q]), dv = D::sub(static_cast<double>(vn), uref[2 * q + 1]);
            ss += du * du;
            ss += dv * dv;
        }
        // interleaved (u, v) observation: coalesced pairs
        for (int q = lane; q < NPT; q += LPE) {
            V2 pr;
            pr.x = f2[q];
            pr.y = f3[q];
            reinterpret_cast<V2*>(gobs)[q] = pr;
        }
    }
    ss = seg_sum<LPE>(ss);                                     // full-warp mask: every lane is here
    if (live && lane == 0) {
        const double nrm = sqrt(ss);
        double r1 = D::div(D::div(D::mul(-0.5, D::mul(nrm, nrm)), static_cast<double>(NX)), static_cast<double>(NX));
        const double da = fabs(D::sub(action_d, a.buf.action_ref[t1]));
        const double r2 = D::mul(D::div(cfg.gamma, 2.0), D::mul(da, da));
        a.buf.reward[env] = D::sub(r1, r2);
        a.buf.time_index[env] = t1;
        a.buf.terminated[env] = (t1 >= cfg.nt - 1) ? 1 : 0;
        a.buf.truncated[env] = 0;
        if (a.buf.valid) a.buf.valid[env] = 1;
    }
    if (live && a.ar_u != nullptr && t1 >= cfg.nt - 1) {  // same-step auto-reset (segment-uniform condition)
        if (a.buf.final_obs != nullptr) {                 // terminal observation U[nt-1] for the caller
            V2* fobs = reinterpret_cast<V2*>(reinterpret_cast<T*>(a.buf.final_obs) + env * NPT * 2);
            for (int q = lane; q < NPT; q += LPE) {
                V2 pr;
                pr.x = f2[q];
                pr.y = f3[q];
                fobs[q] = pr;
            }
        }
        __syncwarp(segmask);                              // the step's own field / observation stores come first
        ns_reset_env<T>(a, a.ar_u, a.ar_v, a.ar_p, a.ar_stride, env, lane, LPE, true);
    }
}

static int ns_validate_cfg(const pdecg_ns_cfg* cfg) {
    if (!cfg) return fail(PDECG_ERR_INVALID, "null cfg");
    if (cfg->abi_version != PDECG_ABI_VERSION) return fail(PDECG_ERR_INVALID, "abi_version mismatch");
    if (cfg->nx < 3 || cfg->ny < 3) return fail(PDECG_ERR_INVALID, "need nx, ny >= 3");
    if (cfg->nx > 8192 || cfg->ny > 8192) return fail(PDECG_ERR_UNSUPPORTED, "grid %dx%d exceeds the limit of 8192 points per direction", cfg->nx, cfg->ny);
    if (cfg->n_pressure_iters < 0) return fail(PDECG_ERR_INVALID, "maximum_pressure_iteration < 0");
    if (!(cfg->dt > 0) || !(cfg->dx > 0) || !(cfg->dy > 0)) return fail(PDECG_ERR_INVALID, "dt, dx, dy must be positive");
    for (int s = 0; s < 4; ++s)
        for (int c = 0; c < 2; ++c)
            if (cfg->bc[s][c] < PDECG_NS_DIRICHLET || cfg->bc[s][c] > PDECG_NS_CONTROLLABLE) return fail(PDECG_ERR_INVALID, "bad boundary condition code");
    if (cfg->precision < PDECG_PREC_REF || cfg->precision > PDECG_PREC_FP32) return fail(PDECG_ERR_INVALID, "bad precision %d", cfg->precision);
    if (cfg->precision == PDECG_PREC_FP32 && (cfg->nx * cfg->ny > kNsMaxPoints || getenv("PDECG_NS_TILED") != nullptr))
        return fail(PDECG_ERR_UNSUPPORTED, "the float32 mode covers the on-chip grids only (nx * ny <= %d)", kNsMaxPoints);
    return PDECG_OK;
}

// Grids of more than kNsMaxPoints points per env (or any grid with PDECG_NS_TILED set: A/B runs, tests) take the
// HBM-resident path of navier_stokes_large.cu; everything else stays on chip for the whole step.
static bool ns_use_large(const pdecg_ns_cfg* cfg) {
    return cfg->nx * cfg->ny > kNsMaxPoints || getenv("PDECG_NS_TILED") != nullptr;
}

static bool ns_f32(const pdecg_ns_cfg* cfg) { return cfg->precision == PDECG_PREC_FP32; }
static size_t ns_smem_bytes(const pdecg_ns_cfg* cfg) {
    return static_cast<size_t>(7) * cfg->nx * cfg->ny * (ns_f32(cfg) ? sizeof(float) : sizeof(double));
}

// generic (CTA-per-env) kernel: pick the instantiation for this grid, opt in to its shared memory, launch
template <int THREADS, int PT>
static int ns_launch_generic_t(const NsArgs& a, int64_t n_envs, size_t smem, cudaStream_t st) {
    void (*kern)(const NsArgs) = ns_f32(&a.cfg) ? ns_step_kernel<THREADS, PT, float> : ns_step_kernel<THREADS, PT, double>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<static_cast<unsigned>(n_envs), THREADS, smem, st>>>(a);
    count_launch();
    return check_launch("ns_step_kernel");
}
// register-blocked Jacobi: 16 thread columns x BC points cover the interior width, 5-point thread rows the height
template <int BC>
static int ns_launch_blk_t(const NsArgs& a, int64_t n_envs, cudaStream_t st) {
    const int rows = (a.cfg.ny - 2 + kBlkRows - 1) / kBlkRows, warps = (rows + 1) / 2;
    const size_t esz = ns_f32(&a.cfg) ? sizeof(float) : sizeof(double);
    const size_t smem = (static_cast<size_t>(4) * a.cfg.nx * a.cfg.ny + static_cast<size_t>(2) * warps * 2 * 16 * BC) * esz;
    void (*kern)(const NsArgs) = ns_f32(&a.cfg) ? ns_step_kernel<256, 1, float, BC> : ns_step_kernel<256, 1, double, BC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<static_cast<unsigned>(n_envs), warps * 32, smem, st>>>(a);
    count_launch();
    return check_launch("ns_step_kernel (register-blocked Jacobi)");
}
static int ns_launch_generic(const NsArgs& a, int64_t n_envs, cudaStream_t st) {
    const int npt = a.cfg.nx * a.cfg.ny;
    // Register-blocked Jacobi up to 50 x 50 (1..3 points per thread per row): 3.29 vs 4.32 ms per 1,024-env step at
    // 41 x 41.  Wider blocks lose to the shared-memory form (7.02 vs 5.86 ms per 592-env step at 64 x 64): with a runtime
    // grid size the clamp at the last interior row / column is a select on every point, ~19 instructions per point-sweep
    // of which 5 are FP64, and the launch turns issue-bound; PDECG_NS_BLK_JACOBI forces it for any on-chip grid (tests).
    const int bc = (a.cfg.nx - 2 + 15) / 16;
    const bool force = getenv("PDECG_NS_BLK_JACOBI") != nullptr;
    if (a.cfg.nx - 2 <= 16 * 5 && a.cfg.ny - 2 <= 16 * kBlkRows && (bc <= 3 || force) && !getenv("PDECG_NS_SMEM_JACOBI")) {
        switch (bc) {
            case 0: case 1: return ns_launch_blk_t<1>(a, n_envs, st);
            case 2: return ns_launch_blk_t<2>(a, n_envs, st);
            case 3: return ns_launch_blk_t<3>(a, n_envs, st);
            case 4: return ns_launch_blk_t<4>(a, n_envs, st);
            default: return ns_launch_blk_t<5>(a, n_envs, st);
        }
    }
    const size_t smem = ns_smem_bytes(&a.cfg);
    if (npt <= 2 * 256) return ns_launch_generic_t<256, 2>(a, n_envs, smem, st);
    if (npt <= 4 * 512) return ns_launch_generic_t<512, 4>(a, n_envs, smem, st);
    return ns_launch_generic_t<1024, 4>(a, n_envs, smem, st);
}

}  // namespace pdecg

using namespace pdecg;

extern "C" int pdecg_ns_reset(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* init_u, const double* init_v,
                              const double* init_p, int64_t init_env_stride, const uint8_t* mask, pdecg_stream_t stream) {
    int rc = ns_validate_cfg(cfg);
    if (rc != PDECG_OK) return rc;
    if (!buf || buf->n_envs < 0) return fail(PDECG_ERR_INVALID, "bad buffers");
    if (buf->n_envs == 0) return PDECG_OK;
    if (!init_u || !init_v || !init_p || !buf->u || !buf->v || !buf->p || !buf->obs || !buf->time_index || !buf->reward || !buf->terminated || !buf->truncated)
        return fail(PDECG_ERR_INVALID, "null device buffer");
    NsArgs a{*cfg, *buf, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0};
    if (ns_f32(cfg)) ns_reset_kernel<float><<<static_cast<unsigned>(buf->n_envs), kNsThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, init_u, init_v, init_p, init_env_stride, mask);
    else ns_reset_kernel<double><<<static_cast<unsigned>(buf->n_envs), kNsThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, init_u, init_v, init_p, init_env_stride, mask);
    count_launch();
    return check_launch("ns_reset_kernel");
}

static int ns_launch_step(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* ar_u, const double* ar_v,
                          const double* ar_p, int64_t ar_stride, pdecg_stream_t stream) {
    int rc = ns_validate_cfg(cfg);
    if (rc != PDECG_OK) return rc;
    if (!buf || buf->n_envs < 0) return fail(PDECG_ERR_INVALID, "bad buffers");
    if (buf->n_envs == 0) return PDECG_OK;
    if (!buf->u || !buf->v || !buf->p || !buf->obs || !buf->action || !buf->u_ref || !buf->action_ref || !buf->time_index || !buf->reward || !buf->terminated || !buf->truncated)
        return fail(PDECG_ERR_INVALID, "null device buffer");
    if (ns_use_large(cfg)) return ns_large_step(cfg, buf, ar_u, ar_v, ar_p, ar_stride, static_cast<cudaStream_t>(stream));
    NsArgs a{*cfg, *buf, nullptr, nullptr, nullptr, 0, ar_u, ar_v, ar_p, ar_stride};
    if (cfg->nx == 21 && cfg->ny == 21 && !getenv("PDECG_NS_GENERIC")) {
        // the reference grid (examples/NavierStokes/NS2Dppo.py:29-32): register-resident kernel, two envs per warp
        // (4 x 4 lanes of 5 x 5 points each); PDECG_NS_LPE32 selects the one-env-per-warp mapping (10 x 3 lanes of 2 x 7).
        // Warps never cooperate, so the CTA size only sets the scheduling granularity: small batches (fewer warps than
        // the GPU can hold at full rate) use one-warp CTAs so that the block scheduler spreads them evenly over the SMs.
        constexpr int kN = 19;
        const bool lpe32 = getenv("PDECG_NS_LPE32") != nullptr;
        const int epw = lpe32 ? 1 : 2;
        const int64_t warps = (buf->n_envs + epw - 1) / epw;
        const bool small = (warps < 148 * 16 || getenv("PDECG_NS_WPB1")) && !getenv("PDECG_NS_WPB4");
        const int wpb = small ? 1 : kWarpsPerBlock;
        const bool f32 = ns_f32(cfg);
        const size_t smem_w = static_cast<size_t>(wpb) * epw * 4 * (kN + 2) * (kN + 2) * (f32 ? sizeof(float) : sizeof(double));
        void (*kern)(const NsArgs);
        if (f32) {
            if (lpe32) kern = small ? ns_step_warp_kernel<kN, 2, 7, 10, 3, 1, 32, float> : ns_step_warp_kernel<kN, 2, 7, 10, 3, kWarpsPerBlock, 32, float>;
            else kern = small ? ns_step_warp_kernel<kN, 5, 5, 4, 4, 1, 16, float> : ns_step_warp_kernel<kN, 5, 5, 4, 4, kWarpsPerBlock, 16, float>;
        } else if (lpe32) kern = small ? ns_step_warp_kernel<kN, 2, 7, 10, 3, 1, 32, double> : ns_step_warp_kernel<kN, 2, 7, 10, 3, kWarpsPerBlock, 32, double>;
        else kern = small ? ns_step_warp_kernel<kN, 5, 5, 4, 4, 1, 16, double> : ns_step_warp_kernel<kN, 5, 5, 4, 4, kWarpsPerBlock, 16, double>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_w));
        if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        const unsigned grid = static_cast<unsigned>((warps + wpb - 1) / wpb);
        kern<<<grid, wpb * 32, smem_w, static_cast<cudaStream_t>(stream)>>>(a);
        count_launch();
        return check_launch("ns_step_warp_kernel");
    }
    return ns_launch_generic(a, buf->n_envs, static_cast<cudaStream_t>(stream));
}

extern "C" int pdecg_ns_step(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, pdecg_stream_t stream) {
    return ns_launch_step(cfg, buf, nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int pdecg_ns_step_autoreset(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* init_u, const double* init_v,
                                       const double* init_p, int64_t init_env_stride, pdecg_stream_t stream) {
    if (!init_u || !init_v || !init_p) return fail(PDECG_ERR_INVALID, "null init buffer");
    return ns_launch_step(cfg, buf, init_u, init_v, init_p, init_env_stride, stream);
}

extern "C" size_t pdecg_ns_workspace_bytes(const pdecg_ns_cfg* cfg, int64_t n_envs) {
    if (!cfg || n_envs <= 0 || cfg->nx < 3 || cfg->ny < 3 || !ns_use_large(cfg)) return 0;
    return ns_large_workspace_bytes(cfg, n_envs);
}

extern "C" int pdecg_ns_solve_pressure(const pdecg_ns_cfg* cfg, int64_t n_envs, const double* u, const double* v, double* p,
                                       void* work, pdecg_stream_t stream) {
    int rc = ns_validate_cfg(cfg);
    if (rc != PDECG_OK) return rc;
    if (n_envs < 0) return fail(PDECG_ERR_INVALID, "n_envs < 0");
    if (n_envs == 0) return PDECG_OK;
    if (!u || !v || !p) return fail(PDECG_ERR_INVALID, "null device buffer");
    if (ns_f32(cfg)) return fail(PDECG_ERR_UNSUPPORTED, "pdecg_ns_solve_pressure is float64 only");
    if (ns_use_large(cfg)) return ns_large_solve_pressure(cfg, n_envs, u, v, p, work, static_cast<cudaStream_t>(stream));
    NsArgs a{};
    a.cfg = *cfg;
    a.buf.n_envs = n_envs;
    a.su = u; a.sv = v; a.sp = p; a.solve_only = 1;
    return ns_launch_generic(a, n_envs, static_cast<cudaStream_t>(stream));
}

// Host-buffer step: H2D(actions) -> kernel -> D2H(observation U[t], reward, flags) on `stream`, then synchronised.
extern "C" int pdecg_ns_step_host(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* action_host,
                                  double* obs_host, double* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                                  uint8_t* valid_host, const double* autoreset_u, const double* autoreset_v,
                                  const double* autoreset_p, int64_t autoreset_env_stride, pdecg_stream_t stream) {
    int rc = ns_validate_cfg(cfg);
    if (rc != PDECG_OK) return rc;
    if (!buf || buf->n_envs < 0) return fail(PDECG_ERR_INVALID, "bad buffers");
    if (buf->n_envs == 0) return PDECG_OK;
    if (!action_host || !buf->action) return fail(PDECG_ERR_INVALID, "null action buffer");
    if (valid_host && !buf->valid) return fail(PDECG_ERR_INVALID, "valid_host needs a device valid buffer");
    if ((autoreset_u || autoreset_v || autoreset_p) && !(autoreset_u && autoreset_v && autoreset_p)) return fail(PDECG_ERR_INVALID, "auto-reset needs all three initial fields");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n = static_cast<size_t>(buf->n_envs), npt = static_cast<size_t>(cfg->nx) * cfg->ny;
    cudaError_t e = cudaMemcpyAsync(const_cast<double*>(buf->action), action_host, n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "H2D actions: %s", cudaGetErrorString(e));
    if ((rc = ns_launch_step(cfg, buf, autoreset_u, autoreset_v, autoreset_p, autoreset_env_stride, stream)) != PDECG_OK) return rc;
    const size_t esz = ns_f32(cfg) ? sizeof(float) : sizeof(double);
    if (obs_host && (e = cudaMemcpyAsync(obs_host, buf->obs, n * npt * 2 * esz, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H obs: %s", cudaGetErrorString(e));
    if (reward_host && (e = cudaMemcpyAsync(reward_host, buf->reward, n * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H reward: %s", cudaGetErrorString(e));
    if (terminated_host && (e = cudaMemcpyAsync(terminated_host, buf->terminated, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H terminated: %s", cudaGetErrorString(e));
    if (truncated_host && (e = cudaMemcpyAsync(truncated_host, buf->truncated, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H truncated: %s", cudaGetErrorString(e));
    if (valid_host && (e = cudaMemcpyAsync(valid_host, buf->valid, n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H valid: %s", cudaGetErrorString(e));
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
    return PDECG_OK;
}
