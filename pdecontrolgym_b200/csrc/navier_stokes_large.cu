// navier_stokes_large.cu -- NavierStokes2D on grids that do not fit one SM's shared memory (> 64 x 64): the
// HBM-resident path (sm_100a, float64).
//
// Reference behaviour restated here (no code shared): environments2d/navier_stokes2D.py:9-22 (central_difference,
// laplace), :68-91 (apply_boundary), :94-116 (solve_pressure), :118-157 (step) and rewards/ns_reward.py:18-28.
//
// One control step = a sequence of launches over ALL envs of the batch:
//   predictor (vectorised stencil) -> boundary conditions -> (dx*dy)*rhs -> ONE plain Jacobi sweep that reads the true
//   boundary ring of the warm-start pressure -> ceil((n-1)/KT) launches of the TMA-tiled, temporally blocked Jacobi
//   kernel (KT sweeps per HBM round trip) -> pressure ring -> corrector -> boundary conditions -> observation / reward.
//
// The tiled Jacobi kernel is the hot one (>= 99 % of a 2000-sweep step):
//   * p and b = (dx*dy)*rhs live in HBM in a padded layout -- a margin of kPad rows / columns above / left of the field and
//     an even row pitch -- described by 3-D tensor maps [env][y][x]: the halo of a tile on the top / left edge then has
//     non-negative, 16-byte aligned box origins (see kPad);
//   * a persistent CTA per SM walks over 80 x 80 tiles (core (80-2KT)^2 + halo KT); `cp.async.bulk.tensor` (TMA) brings
//     the next tile's p and b into the other shared-memory stage while the current one is being swept (mbarrier
//     complete_tx), out-of-range coordinates are zero-filled by the TMA unit;
//   * the tile is swept in REGISTERS: 16 x 16 threads own 5 x 5 points each, left / right halos travel by 16-lane
//     shuffles, up / down halos by shuffle inside a warp (two thread rows per warp) and through a double-buffered
//     shared-memory edge array between warps: one block barrier per sweep, 5 DADD/DMUL per point per sweep;
//   * after KT sweeps the tile's core is exact (the halo absorbs the missing neighbours), goes back to the stage's
//     shared-memory buffer as a dense box and leaves through a TMA store.
// Sweep identity used throughout (see navier_stokes.cu): after a sweep every boundary value equals the new value of
// the nearest interior point, so sweeps 2..n never read the ring -- an interior point next to it uses ITSELF as that
// neighbour (a select on the halo / slot), and the ring is materialised once after the last sweep.
#include <cuda.h>

#include <cstdlib>

#include "pdecg_internal.h"

namespace pdecg {

using D = Strict<double>;

struct NsLargeArgs {
    pdecg_ns_cfg cfg;
    pdecg_ns_buffers buf;
    double* us;          // [n_envs, ny, nx]   predictor u*
    double* vs;          // [n_envs, ny, nx]   predictor v*
    double* b;           // [n_envs, ny + kPad, nxp]  (dx*dy)*rhs, padded layout: field point (i, j) at (i + kPad, j + kPad)
    double* pa;          // [n_envs, ny + kPad, nxp]  pressure ping
    double* pb;          // [n_envs, ny + kPad, nxp]  pressure pong
    int nxp;             // row pitch of b / pa / pb in elements (nx + kPad rounded up to even: TMA needs 16-byte strides)
    const double* ar_u;  // fused same-step auto-reset (nullptr: off)
    const double* ar_v;
    const double* ar_p;
    int64_t ar_stride;
};

constexpr int kLgThreads = 256;
// Margin above / left of the padded fields.  The TMA unit needs every box to START on a 16-byte boundary of the innermost
// dimension (an odd float64 column faults with "illegal instruction"; coordinates beyond the extent are fine and zero-
// filled): tile origins sit at field column 1 - KT + c * CORE, so with KT = 10 and an even core the margin must be odd.
constexpr int kPad = 11;
__host__ __device__ inline int padded_pitch(int nx) { return (nx + kPad + 1) & ~1; }

// ---- simple whole-grid kernels (one thread per point, env-major; a handful of launches per step) ----------------
// valid[env] / staleness: an env past its last step must not be advanced (the reference raises IndexError): every
// kernel below skips it through `stale[env]` computed by the first kernel of the step.
__global__ void __launch_bounds__(kLgThreads) ns_lg_predictor_kernel(const NsLargeArgs a, uint8_t* __restrict__ stale) {
    const pdecg_ns_cfg& cfg = a.cfg;
    const int nx = cfg.nx, ny = cfg.ny;
    const int64_t npt = static_cast<int64_t>(nx) * ny;
    const int64_t env = blockIdx.y;
    const int t0 = a.buf.time_index[env];
    const bool st = (t0 + 1 >= cfg.nt || t0 + 1 >= cfg.n_uref || t0 + 1 >= cfg.n_action_ref);
    if (blockIdx.x == 0 && threadIdx.x == 0) stale[env] = st ? 1 : 0;
    if (st) return;
    const double* __restrict__ u = a.buf.u + env * npt;
    const double* __restrict__ v = a.buf.v + env * npt;
    double* __restrict__ us = a.us + env * npt;
    double* __restrict__ vs = a.vs + env * npt;
    const ConstDiv<double> by_2dx(D::mul(2.0, cfg.dx)), by_2dy(D::mul(2.0, cfg.dy)), by_dxdy(D::mul(cfg.dx, cfg.dy));
    for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < npt; q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(q / nx), j = static_cast<int>(q - static_cast<int64_t>(i) * nx);
        double un = u[q], vn = v[q];
        if (i >= 1 && i <= ny - 2 && j >= 1 && j <= nx - 2) {
            const double u0 = un, v0 = vn;
            auto cdx = [&](const double* f) { return by_2dx(D::sub(f[q + 1], f[q - 1])); };
            auto cdy = [&](const double* f) { return by_2dy(D::sub(f[q + nx], f[q - nx])); };
            auto lap = [&](const double* f) {
                double s = D::add(f[q - 1], f[q - nx]);
                s = D::sub(s, D::mul(4.0, f[q]));
                s = D::add(s, f[q + 1]);
                s = D::add(s, f[q + nx]);
                return by_dxdy(s);
            };
            // u + dt * (((-u)*dudx - v*dudy) + nu*lap(u))     (:130-138)
            double adv = D::sub(D::mul(-u0, cdx(u)), D::mul(v0, cdy(u)));
            un = D::add(u0, D::mul(cfg.dt, D::add(adv, D::mul(cfg.viscosity, lap(u)))));
            adv = D::sub(D::mul(-u0, cdx(v)), D::mul(v0, cdy(v)));
            vn = D::add(v0, D::mul(cfg.dt, D::add(adv, D::mul(cfg.viscosity, lap(v)))));
        }
        us[q] = un;
        vs[q] = vn;
    }
}

// apply_boundary (:68-91) on fields in global memory: one CTA per env, sides in the reference's order, u before v; the
// block barrier between sides keeps the corner values order-dependent exactly like the sequential numpy assignments.
__global__ void __launch_bounds__(kLgThreads) ns_lg_bc_kernel(const NsLargeArgs a, double* __restrict__ fu_all, double* __restrict__ fv_all,
                                                              const uint8_t* __restrict__ stale) {
    const pdecg_ns_cfg& cfg = a.cfg;
    const int nx = cfg.nx, ny = cfg.ny;
    const int64_t env = blockIdx.x;
    if (stale[env]) return;
    double* fu = fu_all + env * static_cast<int64_t>(nx) * ny;
    double* fv = fv_all + env * static_cast<int64_t>(nx) * ny;
    const double action = a.buf.action[env];
    for (int side = 0; side < 4; ++side) {
        const int len = (side < 2) ? nx : ny;
        for (int comp = 0; comp < 2; ++comp) {
            double* f = comp == 0 ? fu : fv;
            const int cond = cfg.bc[side][comp];
            for (int q = threadIdx.x; q < len; q += blockDim.x) {
                int64_t dst, src;
                if (side == PDECG_NS_LOWER) { dst = q; src = nx + q; }
                else if (side == PDECG_NS_UPPER) { dst = static_cast<int64_t>(ny - 1) * nx + q; src = static_cast<int64_t>(ny - 2) * nx + q; }
                else if (side == PDECG_NS_LEFT) { dst = static_cast<int64_t>(q) * nx; src = static_cast<int64_t>(q) * nx + 1; }
                else { dst = static_cast<int64_t>(q) * nx + nx - 1; src = static_cast<int64_t>(q) * nx + nx - 2; }
                if (cond == PDECG_NS_NEUMANN) f[dst] = f[src];
                else if (cond == PDECG_NS_DIRICHLET) f[dst] = 0.0;
                else f[dst] = action;
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}

// b = (dx*dy) * (rho/dt * (d u*/dx + d v*/dy)) on interior points (0 elsewhere) and ONE Jacobi sweep with the TRUE
// neighbours of the warm-start pressure (its boundary ring included): p1 = 1/4 (((l + u) + r) + d - b)   (:100-108)
__global__ void __launch_bounds__(kLgThreads) ns_lg_rhs_first_sweep_kernel(const NsLargeArgs a, const double* __restrict__ p_src_all, int64_t p_src_env,
                                                                           int p_src_pitch, const double* __restrict__ su_all, const double* __restrict__ sv_all,
                                                                           int64_t s_env, int do_sweep, const uint8_t* __restrict__ stale) {
    const pdecg_ns_cfg& cfg = a.cfg;
    const int nx = cfg.nx, ny = cfg.ny, nxp = a.nxp;
    const int64_t env = blockIdx.y;
    if (stale != nullptr && stale[env]) return;
    const int64_t npt = static_cast<int64_t>(nx) * ny, npp = static_cast<int64_t>(nxp) * (ny + kPad);
    const double* __restrict__ us = su_all + env * s_env;
    const double* __restrict__ vs = sv_all + env * s_env;
    const double* __restrict__ p = p_src_all + env * p_src_env;
    double* __restrict__ b = a.b + env * npp;
    double* __restrict__ p1 = a.pa + env * npp;
    const ConstDiv<double> by_2dx(D::mul(2.0, cfg.dx)), by_2dy(D::mul(2.0, cfg.dy));
    const double dxdy = D::mul(cfg.dx, cfg.dy), rho_dt = D::div(cfg.density, cfg.dt);
    for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < npt; q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(q / nx), j = static_cast<int>(q - static_cast<int64_t>(i) * nx);
        const int64_t qp = static_cast<int64_t>(i + kPad) * nxp + (j + kPad);
        const int64_t qs = static_cast<int64_t>(i) * p_src_pitch + j;
        double bv = 0.0, pv = p[qs];
        if (i >= 1 && i <= ny - 2 && j >= 1 && j <= nx - 2) {
            bv = D::mul(dxdy, D::mul(rho_dt, D::add(by_2dx(D::sub(us[q + 1], us[q - 1])), by_2dy(D::sub(vs[q + nx], vs[q - nx])))));
            if (do_sweep) {
                double acc = D::add(p[qs - 1], p[qs - p_src_pitch]);
                acc = D::add(acc, p[qs + 1]);
                acc = D::add(acc, p[qs + p_src_pitch]);
                pv = D::mul(0.25, D::sub(acc, bv));
            }
        }
        b[qp] = bv;
        p1[qp] = pv;
    }
}

// pressure ring = nearest interior value (the four Neumann copies of :110-113 leave exactly that, corners included), and
// the dense copy-out of the padded field into the caller's p
__global__ void __launch_bounds__(kLgThreads) ns_lg_ring_copy_kernel(const NsLargeArgs a, const double* __restrict__ p_fin_all, double* __restrict__ dst_all,
                                                                     int64_t dst_env, int fill_ring, const uint8_t* __restrict__ stale) {
    const pdecg_ns_cfg& cfg = a.cfg;
    const int nx = cfg.nx, ny = cfg.ny, nxp = a.nxp;
    const int64_t env = blockIdx.y;
    if (stale != nullptr && stale[env]) return;
    const int64_t npt = static_cast<int64_t>(nx) * ny;
    const double* __restrict__ p = p_fin_all + env * static_cast<int64_t>(nxp) * (ny + kPad);
    double* __restrict__ dst = dst_all + env * dst_env;
    for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < npt; q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(q / nx), j = static_cast<int>(q - static_cast<int64_t>(i) * nx);
        int ci = i, cj = j;
        if (fill_ring) {
            ci = i < 1 ? 1 : (i > ny - 2 ? ny - 2 : i);
            cj = j < 1 ? 1 : (j > nx - 2 ? nx - 2 : j);
        }
        dst[q] = p[static_cast<int64_t>(ci + kPad) * nxp + (cj + kPad)];
    }
}

// corrector (:143-146): u = u* - dt/rho * dp/dx, v = v* - dt/rho * dp/dy on interior points
__global__ void __launch_bounds__(kLgThreads) ns_lg_corrector_kernel(const NsLargeArgs a, const uint8_t* __restrict__ stale) {
    const pdecg_ns_cfg& cfg = a.cfg;
    const int nx = cfg.nx, ny = cfg.ny;
    const int64_t env = blockIdx.y;
    if (stale[env]) return;
    const int64_t npt = static_cast<int64_t>(nx) * ny;
    const double* __restrict__ p = a.buf.p + env * npt;
    const double* __restrict__ us = a.us + env * npt;
    const double* __restrict__ vs = a.vs + env * npt;
    double* __restrict__ u = a.buf.u + env * npt;
    double* __restrict__ v = a.buf.v + env * npt;
    const ConstDiv<double> by_2dx(D::mul(2.0, cfg.dx)), by_2dy(D::mul(2.0, cfg.dy));
    const double dt_rho = D::div(cfg.dt, cfg.density);
    for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < npt; q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(q / nx), j = static_cast<int>(q - static_cast<int64_t>(i) * nx);
        double un = us[q], vn = vs[q];
        if (i >= 1 && i <= ny - 2 && j >= 1 && j <= nx - 2) {
            un = D::sub(un, D::mul(dt_rho, by_2dx(D::sub(p[q + 1], p[q - 1]))));
            vn = D::sub(vn, D::mul(dt_rho, by_2dy(D::sub(p[q + nx], p[q - nx]))));
        }
        u[q] = un;
        v[q] = vn;
    }
}

// observation U[t] (interleaved u, v), NSReward, bookkeeping, fused auto-reset: one CTA per env
__global__ void __launch_bounds__(kLgThreads) ns_lg_finish_kernel(const NsLargeArgs a, const uint8_t* __restrict__ stale) {
    const pdecg_ns_cfg& cfg = a.cfg;
    const int nx = cfg.nx, ny = cfg.ny;
    const int64_t env = blockIdx.x;
    const int64_t npt = static_cast<int64_t>(nx) * ny;
    const int tid = threadIdx.x, nth = blockDim.x;
    __shared__ double red[kLgThreads / 32];
    double* gu = a.buf.u + env * npt;
    double* gv = a.buf.v + env * npt;
    double* gp = a.buf.p + env * npt;
    double* gobs = a.buf.obs + env * npt * 2;
    auto restart = [&]() {
        const double* iu = a.ar_u + env * a.ar_stride;
        const double* iv = a.ar_v + env * a.ar_stride;
        const double* ip = a.ar_p + env * a.ar_stride;
        for (int64_t q = tid; q < npt; q += nth) {
            const double un = iu[q], vn = iv[q];
            gu[q] = un;
            gv[q] = vn;
            gp[q] = ip[q];
            gobs[2 * q] = un;
            gobs[2 * q + 1] = vn;
        }
        if (tid == 0) a.buf.time_index[env] = 0;
    };
    if (stale[env]) {
        if (tid == 0) {
            if (a.buf.valid) a.buf.valid[env] = 0;
            a.buf.terminated[env] = 1;
            a.buf.truncated[env] = 0;
            a.buf.reward[env] = 0.0;
        }
        if (a.ar_u != nullptr) {
            if (a.buf.final_obs != nullptr) {
                for (int64_t q = tid; q < 2 * npt; q += nth) a.buf.final_obs[env * npt * 2 + q] = gobs[q];
                __syncthreads();
            }
            restart();
        }
        return;
    }
    const int t1 = a.buf.time_index[env] + 1;
    const double* uref = a.buf.u_ref + static_cast<int64_t>(t1) * npt * 2;
    const bool fin = a.ar_u != nullptr && t1 >= cfg.nt - 1 && a.buf.final_obs != nullptr;
    double ss = 0.0;
    for (int64_t q = tid; q < npt; q += nth) {
        const double un = gu[q], vn = gv[q];
        reinterpret_cast<double2*>(gobs)[q] = make_double2(un, vn);
        if (fin) reinterpret_cast<double2*>(a.buf.final_obs + env * npt * 2)[q] = make_double2(un, vn);
        const double du = D::sub(un, uref[2 * q]), dv = D::sub(vn, uref[2 * q + 1]);
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
        const double action = a.buf.action[env];
        double r1 = D::div(D::div(D::mul(-0.5, D::mul(nrm, nrm)), static_cast<double>(nx)), static_cast<double>(ny));
        const double da = fabs(D::sub(action, a.buf.action_ref[t1]));
        const double r2 = D::mul(D::div(cfg.gamma, 2.0), D::mul(da, da));
        a.buf.reward[env] = D::sub(r1, r2);
        a.buf.time_index[env] = t1;
        a.buf.terminated[env] = (t1 >= cfg.nt - 1) ? 1 : 0;
        a.buf.truncated[env] = 0;
        if (a.buf.valid) a.buf.valid[env] = 1;
    }
    if (a.ar_u != nullptr && t1 >= cfg.nt - 1) {
        __syncthreads();
        restart();
    }
}

// ---- TMA / mbarrier plumbing (PTX) ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the tiled, temporally blocked Jacobi kernel -------------------------------------------------------------------
constexpr int kTile = 80;                 // tile edge incl. the halo: 16 threads x 5 points
constexpr int kBlk = 5;                   // points per thread per direction
constexpr int kTG = kTile / kBlk;         // 16 x 16 threads
constexpr int kTileElems = kTile * kTile;
constexpr size_t kStageBytes = 2 * static_cast<size_t>(kTileElems) * sizeof(double);        // p tile + b tile
constexpr size_t kEdgeBytes = 2 * 8 * 2 * static_cast<size_t>(kTile) * sizeof(double);     // [parity][warp][top|bottom][80]
constexpr size_t kJacobiSmem = 2 * kStageBytes + kEdgeBytes + 64;

struct JacobiArgs {
    int nx, ny, n_envs;
    int sweeps;            // sweeps this launch (<= KT)
    int tiles_x, tiles_y;  // core tiles per env
    int64_t total_tiles;
};

template <int KT>
__global__ void __launch_bounds__(kLgThreads, 1) ns_lg_jacobi_tiled_kernel(const __grid_constant__ CUtensorMap map_p_in, const __grid_constant__ CUtensorMap map_b,
                                                                           const __grid_constant__ CUtensorMap map_p_out, const JacobiArgs ja) {
    static_assert(KT % kBlk == 0 && 2 * KT < kTile && KT - 1 <= kPad, "halo depth: a multiple of the thread block edge, within the margin");
    static_assert((1 - KT + kPad) % 2 == 0 && KT % 2 == 0 && (kTile - 2 * KT) % 2 == 0, "load and store boxes must start on even float64 columns");
    constexpr int CORE = kTile - 2 * KT;               // core edge in points
    constexpr int T0 = KT / kBlk;                      // first thread row / column of the core
    extern __shared__ __align__(128) unsigned char smem_raw[];
    auto stage_p = [&](int s) { return reinterpret_cast<double*>(smem_raw + static_cast<size_t>(s) * kStageBytes); };
    auto stage_b = [&](int s) { return stage_p(s) + kTileElems; };
    double* edges = reinterpret_cast<double*>(smem_raw + 2 * kStageBytes);            // [2][8][2][80]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * kStageBytes + kEdgeBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid / kTG, tx = tid - ty * kTG;     // thread row = 2*warp + (lane >> 4), column = lane & 15
    const bool odd = (ty & 1) != 0;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_coords = [&](int64_t t, int& env, int& gy0, int& gx0, int& cy, int& cx) {
        const int per_env = ja.tiles_x * ja.tiles_y;
        env = static_cast<int>(t / per_env);
        const int r = static_cast<int>(t - static_cast<int64_t>(env) * per_env);
        cy = r / ja.tiles_x;
        cx = r - cy * ja.tiles_x;
        gy0 = 1 + cy * CORE - KT;                    // field row of tile-local row 0 (interior index 1 sits at local KT)
        gx0 = 1 + cx * CORE - KT;                    // (>= 1 - KT: non-negative once shifted by kPad into the padded layout)
    };
    auto issue_load = [&](int64_t t, int stage) {
        int env, gy0, gx0, cy, cx;
        tile_coords(t, env, gy0, gx0, cy, cx);
        mbar_expect_tx(&bars[stage], static_cast<uint32_t>(kStageBytes));
        tma_load_3d(stage_p(stage), &map_p_in, gx0 + kPad, gy0 + kPad, env, &bars[stage]);
        tma_load_3d(stage_b(stage), &map_b, gx0 + kPad, gy0 + kPad, env, &bars[stage]);
    };

    const int64_t first = blockIdx.x, stride = gridDim.x;
    if (tid == 0 && first < ja.total_tiles) issue_load(first, 0);

    int it = 0;
    for (int64_t t = first; t < ja.total_tiles; t += stride, ++it) {
        const int stage = it & 1;
        if (tid == 0 && t + stride < ja.total_tiles) {
            tma_wait_read_all();                       // the store that staged its result in the other stage has read it
            issue_load(t + stride, stage ^ 1);
        }
        int env, gy0, gx0, cy, cx;
        tile_coords(t, env, gy0, gx0, cy, cx);
        mbar_wait(&bars[stage], (it >> 1) & 1);

        double p[kBlk][kBlk], n[kBlk][kBlk], b[kBlk][kBlk];
        {
            const double* sp = stage_p(stage) + (ty * kBlk) * kTile + tx * kBlk;
            const double* sb = stage_b(stage) + (ty * kBlk) * kTile + tx * kBlk;
#pragma unroll
            for (int rr = 0; rr < kBlk; ++rr)
#pragma unroll
                for (int cc = 0; cc < kBlk; ++cc) {
                    p[rr][cc] = sp[rr * kTile + cc];
                    b[rr][cc] = sb[rr * kTile + cc];
                }
        }
        // clamps at the domain edge (interior index 1 / n-2): the neighbour beyond it is the point itself.  Top / left
        // fall on a thread-block boundary by construction (select on the halo); bottom / right may sit at any slot.
        const int lb = ja.ny - 2 - gy0, lr = ja.nx - 2 - gx0;            // tile-local row / column of the last interior index
        const bool is_top = (cy == 0) && (ty == T0), is_left = (cx == 0) && (tx == T0);
        const bool has_b = lb >= 0 && lb < kTile && (lb / kBlk) == ty, has_r = lr >= 0 && lr < kTile && (lr / kBlk) == tx;
        const int sb_slot = has_b ? lb - ty * kBlk : -1, sr_slot = has_r ? lr - tx * kBlk : -1;
        const bool edge_tile = (lb < kTile) || (lr < kTile);             // block-uniform: bottom / right boundary inside this tile

        auto sweep = [&](auto edge_tag, const double (&src)[kBlk][kBlk], double (&dst)[kBlk][kBlk], int par) {
            constexpr bool EDGE = decltype(edge_tag)::value;
            double up_h[kBlk], dn_h[kBlk], lf_h[kBlk], rt_h[kBlk];
            // up / down: partner thread row inside the warp by shuffle, the other side through the edge array
            double* e_mine = edges + ((par * 8 + warp) * 2) * kTile;     // [top | bottom] of this warp
#pragma unroll
            for (int cc = 0; cc < kBlk; ++cc) {
                const double pub = odd ? src[0][cc] : src[kBlk - 1][cc];  // what the partner row needs from me
                const double got = __shfl_xor_sync(kFull, pub, 16);
                // rows shared with the neighbouring warps
                if (odd) e_mine[kTile + tx * kBlk + cc] = src[kBlk - 1][cc];   // my bottom row -> warp + 1's up halo
                else e_mine[tx * kBlk + cc] = src[0][cc];                       // my top row    -> warp - 1's down halo
                up_h[cc] = got;
                dn_h[cc] = got;
            }
            __syncthreads();
            {
                const double* e_up = edges + ((par * 8 + (warp > 0 ? warp - 1 : 0)) * 2 + 1) * kTile + tx * kBlk;   // bottom row of the warp above
                const double* e_dn = edges + ((par * 8 + (warp < 7 ? warp + 1 : 7)) * 2) * kTile + tx * kBlk;       // top row of the warp below
#pragma unroll
                for (int cc = 0; cc < kBlk; ++cc) {
                    const double far = odd ? e_dn[cc] : e_up[cc];
                    if (odd) dn_h[cc] = far; else up_h[cc] = far;
                    if (is_top) up_h[cc] = src[0][cc];
                }
            }
#pragma unroll
            for (int rr = 0; rr < kBlk; ++rr) {
                lf_h[rr] = __shfl_up_sync(kFull, src[rr][kBlk - 1], 1, 16);
                rt_h[rr] = __shfl_down_sync(kFull, src[rr][0], 1, 16);
                if (is_left) lf_h[rr] = src[rr][0];
            }
#pragma unroll
            for (int rr = 0; rr < kBlk; ++rr)
#pragma unroll
                for (int cc = 0; cc < kBlk; ++cc) {
                    const double self = src[rr][cc];
                    const double up = (rr > 0) ? src[rr > 0 ? rr - 1 : 0][cc] : up_h[cc];
                    double dn = (rr < kBlk - 1) ? src[rr < kBlk - 1 ? rr + 1 : 0][cc] : dn_h[cc];
                    const double lf = (cc > 0) ? src[rr][cc > 0 ? cc - 1 : 0] : lf_h[rr];
                    double rt = (cc < kBlk - 1) ? src[rr][cc < kBlk - 1 ? cc + 1 : 0] : rt_h[rr];
                    if constexpr (EDGE) {
                        dn = (sb_slot == rr) ? self : dn;
                        rt = (sr_slot == cc) ? self : rt;
                    }
                    double acc = D::add(lf, up);             // ((left + up) + right) + down
                    acc = D::add(acc, rt);
                    acc = D::add(acc, dn);
                    dst[rr][cc] = D::mul(0.25, D::sub(acc, b[rr][cc]));
                }
        };
        auto run = [&](auto edge_tag) {
            int s = 0;
            for (; s + 2 <= ja.sweeps; s += 2) {        // register ping-pong
                sweep(edge_tag, p, n, 0);
                sweep(edge_tag, n, p, 1);
            }
            if (s < ja.sweeps) {
                sweep(edge_tag, p, n, 0);
#pragma unroll
                for (int rr = 0; rr < kBlk; ++rr)
#pragma unroll
                    for (int cc = 0; cc < kBlk; ++cc) p[rr][cc] = n[rr][cc];
            }
        };
        if (edge_tile) run(std::true_type{});
        else run(std::false_type{});

        // core -> dense [CORE][CORE] box in this stage's p buffer -> TMA store (the TMA unit clips what lies outside the field)
        __syncthreads();                                // every thread has taken its p / b out of the stage and left the last sweep
        if (ty >= T0 && ty < T0 + CORE / kBlk && tx >= T0 && tx < T0 + CORE / kBlk) {
            double* so = stage_p(stage) + ((ty - T0) * kBlk) * CORE + (tx - T0) * kBlk;
#pragma unroll
            for (int rr = 0; rr < kBlk; ++rr)
#pragma unroll
                for (int cc = 0; cc < kBlk; ++cc) so[rr * CORE + cc] = p[rr][cc];
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tma_store_3d(&map_p_out, stage_p(stage), gx0 + KT + kPad, gy0 + KT + kPad, env);
            tma_commit();
        }
    }
    if (tid == 0) tma_wait_all();
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 3-D map [env][y][x] over a padded field (extent nx + kPad by ny + kPad, the field itself starts at (kPad, kPad)); box = bw x bh x 1
static int make_map(CUtensorMap* m, double* base, int nx, int ny, int nxp, int64_t n_envs, int bw, int bh) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(PDECG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(nx + kPad), static_cast<cuuint64_t>(ny + kPad), static_cast<cuuint64_t>(n_envs)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(nxp) * 8, static_cast<cuuint64_t>(nxp) * (ny + kPad) * 8};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PDECG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    return PDECG_OK;
}

constexpr int kKT = 10;      // sweeps per HBM round trip (halo depth): core 60 x 60 of an 80 x 80 tile

size_t ns_large_workspace_bytes(const pdecg_ns_cfg* cfg, int64_t n_envs) {
    const size_t nxp = static_cast<size_t>(padded_pitch(cfg->nx));
    // u*, v* dense; b, p ping, p pong padded; + one stale byte per env (rounded up to 256 B)
    return static_cast<size_t>(n_envs) * (2 * static_cast<size_t>(cfg->nx) * cfg->ny + 3 * nxp * (cfg->ny + kPad)) * sizeof(double) +
           ((static_cast<size_t>(n_envs) + 255) & ~static_cast<size_t>(255));
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// N Jacobi sweeps on the padded fields: a.pa holds sweep 1's result on entry; returns the buffer with the final one
static int jacobi_sweeps(const NsLargeArgs& a, int64_t n_envs, int remaining, cudaStream_t st, double** p_final) {
    const int kt = kKT;
    const int core = kTile - 2 * kt;
    const int nx = a.cfg.nx, ny = a.cfg.ny;
    double* cur = a.pa;
    double* oth = a.pb;
    if (remaining > 0) {
        CUtensorMap m_b, m_in[2], m_out[2];
        int rc;
        if ((rc = make_map(&m_b, a.b, nx, ny, a.nxp, n_envs, kTile, kTile)) != PDECG_OK) return rc;
        if ((rc = make_map(&m_in[0], a.pa, nx, ny, a.nxp, n_envs, kTile, kTile)) != PDECG_OK) return rc;
        if ((rc = make_map(&m_in[1], a.pb, nx, ny, a.nxp, n_envs, kTile, kTile)) != PDECG_OK) return rc;
        if ((rc = make_map(&m_out[0], a.pa, nx, ny, a.nxp, n_envs, core, core)) != PDECG_OK) return rc;
        if ((rc = make_map(&m_out[1], a.pb, nx, ny, a.nxp, n_envs, core, core)) != PDECG_OK) return rc;
        JacobiArgs ja;
        ja.nx = nx; ja.ny = ny; ja.n_envs = static_cast<int>(n_envs);
        ja.tiles_x = (nx - 2 + core - 1) / core;
        ja.tiles_y = (ny - 2 + core - 1) / core;
        ja.total_tiles = static_cast<int64_t>(ja.tiles_x) * ja.tiles_y * n_envs;
        const unsigned grid = static_cast<unsigned>(ja.total_tiles < sm_count() ? ja.total_tiles : sm_count());
        cudaFuncSetAttribute(ns_lg_jacobi_tiled_kernel<kKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kJacobiSmem));
        int which = 0;                                  // 0: cur == pa
        while (remaining > 0) {
            ja.sweeps = remaining < kt ? remaining : kt;
            ns_lg_jacobi_tiled_kernel<kKT><<<grid, kLgThreads, kJacobiSmem, st>>>(m_in[which], m_b, m_out[which ^ 1], ja);
            count_launch();
            if ((rc = check_launch("ns_lg_jacobi_tiled_kernel")) != PDECG_OK) return rc;
            remaining -= ja.sweeps;
            which ^= 1;
            double* t = cur; cur = oth; oth = t;
        }
    }
    *p_final = cur;
    return PDECG_OK;
}

static void carve(NsLargeArgs& a, void* work, int64_t n_envs, uint8_t** stale) {
    const int64_t npt = static_cast<int64_t>(a.cfg.nx) * a.cfg.ny, npp = static_cast<int64_t>(a.nxp) * (a.cfg.ny + kPad);
    double* w = static_cast<double*>(work);
    a.us = w; w += n_envs * npt;
    a.vs = w; w += n_envs * npt;
    a.b = w; w += n_envs * npp;
    a.pa = w; w += n_envs * npp;
    a.pb = w; w += n_envs * npp;
    *stale = reinterpret_cast<uint8_t*>(w);
}

static unsigned grid_x(int64_t npt) {
    const int64_t g = (npt + kLgThreads * 4 - 1) / (kLgThreads * 4);
    return static_cast<unsigned>(g < 1 ? 1 : (g > 1024 ? 1024 : g));
}

int ns_large_step(const pdecg_ns_cfg* cfg, const pdecg_ns_buffers* buf, const double* ar_u, const double* ar_v, const double* ar_p,
                  int64_t ar_stride, cudaStream_t st) {
    if (!buf->work) return fail(PDECG_ERR_INVALID, "grids beyond %d points per env need the scratch buffer pdecg_ns_buffers::work (pdecg_ns_workspace_bytes)", 4096);
    NsLargeArgs a{};
    a.cfg = *cfg;
    a.buf = *buf;
    a.nxp = padded_pitch(cfg->nx);
    a.ar_u = ar_u; a.ar_v = ar_v; a.ar_p = ar_p; a.ar_stride = ar_stride;
    uint8_t* stale = nullptr;
    const int64_t n = buf->n_envs, npt = static_cast<int64_t>(cfg->nx) * cfg->ny;
    carve(a, buf->work, n, &stale);
    const dim3 g2(grid_x(npt), static_cast<unsigned>(n));
    int rc;
    ns_lg_predictor_kernel<<<g2, kLgThreads, 0, st>>>(a, stale);
    count_launch();
    if ((rc = check_launch("ns_lg_predictor_kernel")) != PDECG_OK) return rc;
    ns_lg_bc_kernel<<<static_cast<unsigned>(n), kLgThreads, 0, st>>>(a, a.us, a.vs, stale);
    count_launch();
    const int iters = cfg->n_pressure_iters;
    // b, and sweep 1 (true ring values) into the padded ping buffer (iters == 0: a plain padded copy of p)
    ns_lg_rhs_first_sweep_kernel<<<g2, kLgThreads, 0, st>>>(a, buf->p, npt, cfg->nx, a.us, a.vs, npt, iters >= 1 ? 1 : 0, stale);
    count_launch();
    if ((rc = check_launch("ns_lg_rhs_first_sweep_kernel")) != PDECG_OK) return rc;
    double* p_final = a.pa;
    if (iters >= 2 && (rc = jacobi_sweeps(a, n, iters - 1, st, &p_final)) != PDECG_OK) return rc;
    ns_lg_ring_copy_kernel<<<g2, kLgThreads, 0, st>>>(a, p_final, buf->p, npt, iters >= 1 ? 1 : 0, stale);
    count_launch();
    ns_lg_corrector_kernel<<<g2, kLgThreads, 0, st>>>(a, stale);
    count_launch();
    ns_lg_bc_kernel<<<static_cast<unsigned>(n), kLgThreads, 0, st>>>(a, buf->u, buf->v, stale);
    count_launch();
    ns_lg_finish_kernel<<<static_cast<unsigned>(n), kLgThreads, 0, st>>>(a, stale);
    count_launch();
    return check_launch("ns_lg_finish_kernel");
}

// solve_pressure alone (:94-116) for external callers: p <- n sweeps on rhs(u, v), warm-started from p
int ns_large_solve_pressure(const pdecg_ns_cfg* cfg, int64_t n_envs, const double* u, const double* v, double* p, void* work, cudaStream_t st) {
    if (!work) return fail(PDECG_ERR_INVALID, "grids beyond %d points per env need a scratch buffer (pdecg_ns_workspace_bytes)", 4096);
    NsLargeArgs a{};
    a.cfg = *cfg;
    a.buf.n_envs = n_envs;
    a.nxp = padded_pitch(cfg->nx);
    uint8_t* stale = nullptr;
    const int64_t npt = static_cast<int64_t>(cfg->nx) * cfg->ny;
    carve(a, work, n_envs, &stale);
    const dim3 g2(grid_x(npt), static_cast<unsigned>(n_envs));
    const int iters = cfg->n_pressure_iters;
    int rc;
    ns_lg_rhs_first_sweep_kernel<<<g2, kLgThreads, 0, st>>>(a, p, npt, cfg->nx, u, v, npt, iters >= 1 ? 1 : 0, nullptr);
    count_launch();
    if ((rc = check_launch("ns_lg_rhs_first_sweep_kernel")) != PDECG_OK) return rc;
    double* p_final = a.pa;
    if (iters >= 2 && (rc = jacobi_sweeps(a, n_envs, iters - 1, st, &p_final)) != PDECG_OK) return rc;
    ns_lg_ring_copy_kernel<<<g2, kLgThreads, 0, st>>>(a, p_final, p, npt, iters >= 1 ? 1 : 0, nullptr);
    count_launch();
    return check_launch("ns_lg_ring_copy_kernel");
}

}  // namespace pdecg
