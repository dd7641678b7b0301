/*
 * pdecg_oracle.c -- plain-C restatement of the Transport / ReactionDiffusion, Traffic-ARZ and NavierStokes2D
 * integrators.  TEST INFRASTRUCTURE ONLY.
 *
 * Scalar loops with explicit types and one IEEE operation per statement, in the order numpy evaluates the
 * reference's expressions (hyperbolic.py:140-156, parabolic.py:141-151, base_env_1d.py:36-39,
 * traffic_arz_env.py:131-233,270-280, navier_stokes2D.py:9-22,68-157 of lukebhan/PDEControlGym).  It exists to (1) check the CUDA kernels at FULL batch sizes, where the numpy oracle
 * (oracle/np_oracle.py, which is pinned bit-for-bit to the live reference) is too slow, and (2) give bench.py a
 * compiled multi-threaded CPU number next to the numpy one.  tests/test_c_oracle.py pins this file to the numpy
 * oracle bit for bit.  Build: make -C oracle  (gcc -O3 -mavx2 -ffp-contract=off -fno-fast-math -fopenmp: vectorised but never re-associated or contracted).
 * Nothing under pdecontrolgym_b200/ links or loads it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int env;           /* 0 = Transport, 1 = ReactionDiffusion */
    int state_f64;     /* 0: float32 state ("ref"), 1: float64 state ("fp64") */
    int beta_f64;      /* dtype of beta (only matters for float32 state) */
    int action_f64;    /* dtype of the action */
    int nx, nt, substeps;
    int control_neumann, normalize;
    double dt, dx, max_control_value;
} pdecg_oracle_cfg;

/* base_env_1d.py:36-39, evaluated in the dtype of its argument */
static float normalize_f(float a, int on, float mx) { if (!on) return a; float t = a + 1.0f; t = t * mx; return t - mx; }
static double normalize_d(double a, int on, double mx) { if (!on) return a; double t = a + 1.0; t = t * mx; return t - mx; }

/* value written at the actuated point; `nb` = neighbour state value used by Neumann control */
static double actuate(const pdecg_oracle_cfg* c, const void* action, int64_t e, double nb) {
    if (!c->action_f64) {
        float a = ((const float*)action)[e];
        if (!c->control_neumann) return (double)normalize_f(a, c->normalize, (float)c->max_control_value);
        float adx = a * (float)c->dx;
        if (c->state_f64) { double w = (double)adx + nb; return normalize_d(w, c->normalize, c->max_control_value); }
        float w = adx + (float)nb;
        return (double)normalize_f(w, c->normalize, (float)c->max_control_value);
    } else {
        double a = ((const double*)action)[e];
        if (!c->control_neumann) return normalize_d(a, c->normalize, c->max_control_value);
        double w = a * c->dx;
        w = w + nb;
        return normalize_d(w, c->normalize, c->max_control_value);
    }
}

/* One control step (`substeps` explicit updates, stopping at time index nt-1) for envs [0, n).
 * u: [n, npts] state (float or double), beta: [n or 1, npts] (float or double), time_index: [n] int32.
 * scratch: [threads, 2*npts] doubles.  Returns 0. */
int pdecg_oracle_bc1d_step(const pdecg_oracle_cfg* c, int64_t n, void* u, const void* beta, int64_t beta_stride,
                           const void* action, int32_t* time_index) {
    const int npts = c->env == 1 ? c->nx + 1 : c->nx;
    const int nx = c->nx;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        int t = time_index[e];
        if (c->state_f64) {
            double* p = (double*)malloc(sizeof(double) * 2 * npts);
            double* q = p + npts;
            double* ue = (double*)u + e * npts;
            const double* be = (const double*)beta + e * beta_stride;
            memcpy(p, ue, sizeof(double) * npts);
            const double F = c->dt / (c->dx * c->dx);
            for (int i = 0; i < c->substeps && t < c->nt - 1; ++i, ++t) {
                if (c->env == 1) {
                    for (int j = 1; j < nx; ++j) {
                        double a = 2.0 * p[j]; a = p[j - 1] - a; a = a + p[j + 1]; a = F * a; a = p[j] + a;
                        double r = c->dt * be[j]; r = r * p[j];
                        q[j] = a + r;
                    }
                    q[0] = 0.0;
                    q[nx] = actuate(c, action, e, p[nx - 1]);
                } else {
                    q[nx - 1] = actuate(c, action, e, 0.0);
                    for (int j = 0; j < nx - 1; ++j) {
                        double d = p[j + 1] - p[j]; d = d / c->dx;
                        double b = p[0] * be[j]; d = d + b; d = c->dt * d;
                        q[j] = p[j] + d;
                    }
                }
                memcpy(p, q, sizeof(double) * npts);
            }
            memcpy(ue, p, sizeof(double) * npts);
            free(p);
        } else {
            float* p = (float*)malloc(sizeof(float) * 2 * npts);
            float* q = p + npts;
            float* ue = (float*)u + e * npts;
            memcpy(p, ue, sizeof(float) * npts);
            const float Ff = (float)(c->dt / (c->dx * c->dx));
            const float dtf = (float)c->dt, dxf = (float)c->dx;
            for (int i = 0; i < c->substeps && t < c->nt - 1; ++i, ++t) {
                if (c->env == 1) {
                    for (int j = 1; j < nx; ++j) {
                        float a = 2.0f * p[j]; a = p[j - 1] - a; a = a + p[j + 1]; a = Ff * a; a = p[j] + a;
                        if (c->beta_f64) {
                            double r = c->dt * ((const double*)beta)[e * beta_stride + j]; r = r * (double)p[j];
                            q[j] = (float)((double)a + r);
                        } else {
                            float r = dtf * ((const float*)beta)[e * beta_stride + j]; r = r * p[j];
                            q[j] = a + r;
                        }
                    }
                    q[0] = 0.0f;
                    q[nx] = (float)actuate(c, action, e, (double)p[nx - 1]);
                } else {
                    q[nx - 1] = (float)actuate(c, action, e, 0.0);
                    for (int j = 0; j < nx - 1; ++j) {
                        float d = p[j + 1] - p[j]; d = d / dxf;
                        if (c->beta_f64) {
                            double b = (double)p[0] * ((const double*)beta)[e * beta_stride + j];
                            double s = (double)d + b; s = c->dt * s;
                            q[j] = (float)((double)p[j] + s);
                        } else {
                            float b = p[0] * ((const float*)beta)[e * beta_stride + j];
                            float s = d + b; s = dtf * s;
                            q[j] = p[j] + s;
                        }
                    }
                }
                memcpy(p, q, sizeof(float) * npts);
            }
            memcpy(ue, p, sizeof(float) * npts);
            free(p);
        }
        time_index[e] = t;
    }
    return 0;
}

int pdecg_oracle_max_points(void) { return 1 << 20; }


/* ------------------------------------------------------------------------------------------------------------
 * TrafficPDE1D (Aw-Rascle-Zhang), traffic_arz_env.py:131-233 -- state integrator + velocity, all float64.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int M, control_freq;
    int sim_type;          /* 0 inlet, 1 outlet, 2 both, 3 outlet-train */
    int action_dim;
    double T, dt, dx, tau, vm, rm;
} pdecg_oracle_arz_cfg;

static double arz_veq(const pdecg_oracle_arz_cfg* c, double rho) { double t = rho / c->rm; t = 1.0 - t; return c->vm * t; }      /* :271-272 */
static double arz_fr(const pdecg_oracle_arz_cfg* c, double rho, double y) { double t = rho * arz_veq(c, rho); return y + t; }     /* :275-276 */
static double arz_fy(const pdecg_oracle_arz_cfg* c, double rho, double y) { double t = y / rho; t = t + arz_veq(c, rho); return y * t; } /* :279-280 */

/* r, y: [n, M] in/out; v: [n, M] out; action: [n, action_dim]; qs: [n]; time: [n] (float time, advanced by dt per step
 * and zeroed when it reaches T/dt, :109-111,148).  Returns 0. */
int pdecg_oracle_arz_step(const pdecg_oracle_arz_cfg* c, int64_t n, double* r, double* y, double* v, const double* action,
                          const double* qs, double* time) {
    const int M = c->M;
    const double c1 = c->dt / (2 * c->dx), c3 = c->dt / c->dx;            /* :202,218 (python floats) */
    const double c2 = 0.25 * c->dt / c->tau, c4 = 0.5 * c->dt / c->tau;   /* :208,221 */
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        double* re = r + e * M;
        double* ye = y + e * M;
        double* buf = (double*)malloc(sizeof(double) * 7 * M);
        double *fr = buf, *fy = buf + M, *rm_ = buf + 2 * M, *ym_ = buf + 3 * M, *frm = buf + 4 * M, *fym = buf + 5 * M, *tmp = buf + 6 * M;
        (void)tmp;
        double t = time[e] + c->dt;
        const double lo = qs[e] * 0.8, hi = 1.2 * qs[e];                  /* :98-100 */
        double a0 = action[e * c->action_dim];
        a0 = a0 < lo ? lo : a0; a0 = a0 > hi ? hi : a0;
        double a1 = a0;
        if (c->action_dim > 1) { a1 = action[e * c->action_dim + 1]; a1 = a1 < lo ? lo : a1; a1 = a1 > hi ? hi : a1; }
        const double q_in = (c->sim_type == 0 || c->sim_type == 2) ? a0 : qs[e];
        const double q_out = (c->sim_type == 1 || c->sim_type == 3) ? a0 : (c->sim_type == 2 ? a1 : qs[e]);
        for (int count = 0; count < c->control_freq && t < c->T; ++count) {   /* :173 */
            re[0] = re[1];                                                    /* :175-190 */
            { double w = re[0] * arz_veq(c, re[0]); ye[0] = q_in - w; }
            re[M - 1] = re[M - 2];
            { double w = re[M - 1] * arz_veq(c, re[M - 1]); ye[M - 1] = q_out - w; }
            for (int j = 0; j < M; ++j) { fr[j] = arz_fr(c, re[j], ye[j]); fy[j] = arz_fy(c, re[j], ye[j]); }
            for (int j = 0; j < M - 1; ++j) {                                 /* midpoint j+1/2 (:202-215) */
                double s = re[j + 1] + re[j]; s = 0.5 * s;
                double d = fr[j + 1] - fr[j]; d = c1 * d;
                rm_[j] = s - d;
                double ys = ye[j + 1] + ye[j];
                double h = 0.5 * ys;
                double dy = fy[j + 1] - fy[j]; dy = c1 * dy;
                h = h - dy;
                double rel = c2 * ys;
                ym_[j] = h - rel;
                frm[j] = arz_fr(c, rm_[j], ym_[j]);
                fym[j] = arz_fy(c, rm_[j], ym_[j]);
            }
            for (int j = 1; j < M - 1; ++j) {                                 /* :218-222 (uses the old r_j, y_j) */
                double d = frm[j] - frm[j - 1]; d = c3 * d;
                double dy = fym[j] - fym[j - 1]; dy = c3 * dy;
                double s = ym_[j] + ym_[j - 1]; s = c4 * s;
                dy = dy + s;
                fr[j] = re[j] - d;                                            /* fr / fy reused as the new rows */
                fy[j] = ye[j] - dy;
            }
            for (int j = 1; j < M - 1; ++j) { re[j] = fr[j]; ye[j] = fy[j]; }
        }
        double* ve = v + e * M;
        for (int j = 0; j < M; ++j) { double q = ye[j] / re[j]; ve[j] = q + arz_veq(c, re[j]); }   /* :226 */
        if (t >= c->T / c->dt) t = 0.0;                                       /* :109-111 */
        time[e] = t;
        free(buf);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * NavierStokes2D (Chorin projection, Jacobi pressure solve), navier_stokes2D.py:68-157 -- all float64.
 * Arrays are [n, ny, nx] (row = y).  bc[side][component]: 0 Dirichlet (0), 1 Neumann (copy), 2 Controllable (action);
 * sides in the order apply_boundary visits them: lower (row 0), upper (row -1), left (col 0), right (col -1).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int nx, ny, n_iter;
    int bc[4][2];
    double dt, dx, dy, nu, rho;
} pdecg_oracle_ns_cfg;

static void ns_apply_boundary(const pdecg_oracle_ns_cfg* c, double* u, double* v, double a) {   /* :68-91 */
    const int nx = c->nx, ny = c->ny;
    for (int side = 0; side < 4; ++side)
        for (int comp = 0; comp < 2; ++comp) {
            double* f = comp == 0 ? u : v;
            const int cond = c->bc[side][comp];
            const int len = side < 2 ? nx : ny;
            for (int q = 0; q < len; ++q) {
                int dst, src;
                if (side == 0) { dst = q; src = nx + q; }
                else if (side == 1) { dst = (ny - 1) * nx + q; src = (ny - 2) * nx + q; }
                else if (side == 2) { dst = q * nx; src = q * nx + 1; }
                else { dst = q * nx + nx - 1; src = q * nx + nx - 2; }
                if (cond == 1) f[dst] = f[src];
                else if (cond == 0) f[dst] = 0.0;
                else f[dst] = a;
            }
        }
}
static double ns_cdx(const double* f, int i, int j, int nx, double two_dx) { double d = f[i * nx + j + 1] - f[i * nx + j - 1]; return d / two_dx; }   /* :9-15 */
static double ns_cdy(const double* f, int i, int j, int nx, double two_dy) { double d = f[(i + 1) * nx + j] - f[(i - 1) * nx + j]; return d / two_dy; }
static double ns_lap(const double* f, int i, int j, int nx, double dxdy) {                                                                              /* :17-22 */
    double s = f[i * nx + j - 1] + f[(i - 1) * nx + j];
    double m = 4 * f[i * nx + j];
    s = s - m; s = s + f[i * nx + j + 1]; s = s + f[(i + 1) * nx + j];
    return s / dxdy;
}

/* u, v, p: [n, ny, nx] in/out; action: [n].  One call = one env step (predictor, BC, n_iter Jacobi sweeps, corrector, BC). */
int pdecg_oracle_ns_step(const pdecg_oracle_ns_cfg* c, int64_t n, double* u, double* v, double* p, const double* action) {
    const int nx = c->nx, ny = c->ny, npt = nx * ny;
    const double two_dx = 2 * c->dx, two_dy = 2 * c->dy, dxdy = c->dx * c->dy;
    const double rho_dt = c->rho / c->dt, dt_rho = c->dt / c->rho;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        double* ue = u + e * npt; double* ve = v + e * npt; double* pe = p + e * npt;
        double* buf = (double*)malloc(sizeof(double) * 4 * npt);
        double *us = buf, *vs = buf + npt, *rhs = buf + 2 * npt, *pn = buf + 3 * npt;
        memcpy(us, ue, sizeof(double) * npt);
        memcpy(vs, ve, sizeof(double) * npt);
        for (int i = 1; i < ny - 1; ++i)
            for (int j = 1; j < nx - 1; ++j) {                            /* :130-138 */
                const double u0 = ue[i * nx + j], v0 = ve[i * nx + j];
                double a = -u0 * ns_cdx(ue, i, j, nx, two_dx);
                double b = v0 * ns_cdy(ue, i, j, nx, two_dy);
                a = a - b;
                b = c->nu * ns_lap(ue, i, j, nx, dxdy);
                a = a + b; a = c->dt * a;
                us[i * nx + j] = u0 + a;
                a = -u0 * ns_cdx(ve, i, j, nx, two_dx);
                b = v0 * ns_cdy(ve, i, j, nx, two_dy);
                a = a - b;
                b = c->nu * ns_lap(ve, i, j, nx, dxdy);
                a = a + b; a = c->dt * a;
                vs[i * nx + j] = v0 + a;
            }
        ns_apply_boundary(c, us, vs, action[e]);
        memset(rhs, 0, sizeof(double) * npt);
        for (int i = 1; i < ny - 1; ++i)
            for (int j = 1; j < nx - 1; ++j) {                            /* :98-101 */
                double d = ns_cdx(us, i, j, nx, two_dx) + ns_cdy(vs, i, j, nx, two_dy);
                rhs[i * nx + j] = rho_dt * d;
            }
        for (int it = 0; it < c->n_iter; ++it) {                          /* :104-114 */
            memcpy(pn, pe, sizeof(double) * npt);
            for (int i = 1; i < ny - 1; ++i)
                for (int j = 1; j < nx - 1; ++j) {
                    double s = pe[i * nx + j - 1] + pe[(i - 1) * nx + j];
                    s = s + pe[i * nx + j + 1]; s = s + pe[(i + 1) * nx + j];
                    double b = dxdy * rhs[i * nx + j];
                    s = s - b;
                    pn[i * nx + j] = 0.25 * s;
                }
            for (int i = 0; i < ny; ++i) pn[i * nx + nx - 1] = pn[i * nx + nx - 2];
            for (int j = 0; j < nx; ++j) pn[j] = pn[nx + j];
            for (int i = 0; i < ny; ++i) pn[i * nx] = pn[i * nx + 1];
            for (int j = 0; j < nx; ++j) pn[(ny - 1) * nx + j] = pn[(ny - 2) * nx + j];
            memcpy(pe, pn, sizeof(double) * npt);
        }
        memcpy(ue, us, sizeof(double) * npt);
        memcpy(ve, vs, sizeof(double) * npt);
        for (int i = 1; i < ny - 1; ++i)
            for (int j = 1; j < nx - 1; ++j) {                            /* :143-146 */
                double g = dt_rho * ns_cdx(pe, i, j, nx, two_dx);
                ue[i * nx + j] = us[i * nx + j] - g;
                g = dt_rho * ns_cdy(pe, i, j, nx, two_dy);
                ve[i * nx + j] = vs[i * nx + j] - g;
            }
        ns_apply_boundary(c, ue, ve, action[e]);
        free(buf);
    }
    return 0;
}


/* ---- correctly rounded exp(): operation-for-operation restatement of exp_cr() in pdecontrolgym_b200/csrc/pdecg_common.cuh ----
 * (double-double arithmetic, fixed order; compiled with -ffp-contract=off, fma() from libm).  tests/test_exp_cr.py compares
 * it with mpmath's correctly rounded value; the GPU self-test compares the device function with this one bit for bit. */
typedef struct { double hi, lo; } pdecg_dd;
static pdecg_dd dd_quick(double a, double b) { pdecg_dd r; r.hi = a + b; r.lo = b + -(r.hi + -a); return r; }
static pdecg_dd dd_add(pdecg_dd a, pdecg_dd b) {
    const double s = a.hi + b.hi;
    const double bb = s + -a.hi;
    double e = (a.hi + -(s + -bb)) + (b.hi + -bb);
    e = e + (a.lo + b.lo);
    return dd_quick(s, e);
}
static pdecg_dd dd_mul(pdecg_dd a, pdecg_dd b) {
    const double p = a.hi * b.hi;
    double e = fma(a.hi, b.hi, -p);
    e = e + ((a.hi * b.lo) + (a.lo * b.hi));
    return dd_quick(p, e);
}
double pdecg_oracle_exp_cr(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) return HUGE_VAL;
    if (x < -745.1332191019412) return 0.0;
    const double kf = rint(x * 0x1.71547652b82fep+0);
    const pdecg_dd ln2 = {0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56};
    pdecg_dd mk = {-kf, 0.0}, xx = {x, 0.0}, one = {1.0, 0.0};
    pdecg_dd r = dd_add(xx, dd_mul(mk, ln2));
    r.hi = r.hi * 0x1p-8;
    r.lo = r.lo * 0x1p-8;
    const pdecg_dd c[10] = {{0x1.ae64567f544e4p-26, -0x1.c062e06d1f209p-80}, {0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76},
                            {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73}, {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
                            {0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73}, {0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65},
                            {0x1.1111111111111p-7, 0x1.1111111111111p-63}, {0x1.5555555555555p-5, 0x1.5555555555555p-59},
                            {0x1.5555555555555p-3, 0x1.5555555555555p-57}, {0x1.0000000000000p-1, 0.0}};
    pdecg_dd acc = c[0];
    for (int i = 1; i < 10; ++i) acc = dd_add(dd_mul(acc, r), c[i]);
    acc = dd_add(dd_mul(acc, r), one);
    acc = dd_add(dd_mul(acc, r), one);
    for (int i = 0; i < 8; ++i) acc = dd_mul(acc, acc);
    return scalbn(acc.hi + acc.lo, (int)kf);
}
void pdecg_oracle_exp_cr_array(const double* x, double* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = pdecg_oracle_exp_cr(x[i]);
}
