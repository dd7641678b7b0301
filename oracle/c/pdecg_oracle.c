/*
 * pdecg_oracle.c -- plain-C restatement of the Transport / ReactionDiffusion integrators.  TEST INFRASTRUCTURE ONLY.
 *
 * Scalar loops with explicit types and one IEEE operation per statement, in the order numpy evaluates the
 * reference's expressions (hyperbolic.py:140-156, parabolic.py:141-151, base_env_1d.py:36-39 of
 * lukebhan/PDEControlGym).  It exists to (1) check the CUDA kernels at FULL batch sizes, where the numpy oracle
 * (oracle/np_oracle.py, which is pinned bit-for-bit to the live reference) is too slow, and (2) give bench.py a
 * compiled multi-threaded CPU number next to the numpy one.  tests/test_c_oracle.py pins this file to the numpy
 * oracle bit for bit.  Build: make -C oracle  (gcc -O3 -mavx2 -ffp-contract=off -fno-fast-math -fopenmp: vectorised but never re-associated or contracted).
 * Nothing under pdecontrolgym_b200/ links or loads it.
 */
#include <math.h>
#include <stdint.h>
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
            double p[520], q[520];
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
        } else {
            float p[520], q[520];
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
        }
        time_index[e] = t;
    }
    return 0;
}

int pdecg_oracle_max_points(void) { return 520; }
