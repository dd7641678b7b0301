"""Shared builders for the parity tests: the same parameter dict drives the oracle and the CUDA env."""
import numpy as np

from oracle import np_oracle as O

TOL_F32_REDUCTION = 2e-6   # rewards / norms from float32 rows: numpy's sdot / pairwise order is not reproducible
TOL_F64_REDUCTION = 1e-12  # same for float64 rows


def bc1d_case(env, nx=None, steps=12):
    """Small full-episode configuration (terminates exactly at `steps` control steps)."""
    if env == "rd":
        nx = 200 if nx is None else nx
        dx = 1.0 / nx
        dt = 0.4 * dx * dx                       # F = dt/dx^2 = 0.4 like the reference defaults
        sub = 100
        T = dt * sub * steps
        npts = nx + 1
        beta = 50 * np.cos(8 * np.arccos(np.linspace(0, 1, npts)))
    else:
        nx = 100 if nx is None else nx
        dx = 1.0 / nx
        dt = dx / 100
        sub = 500
        T = dt * sub * steps
        npts = nx
        beta = 5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, npts)))
    kw = dict(T=T, dt=dt, X=1, dx=dx, control_sample_rate=dt * sub)
    nt = int(round(T / dt) + 1)
    assert nt - 1 == sub * steps, (nt, sub, steps)
    assert int(round(kw["control_sample_rate"] / dt)) == sub
    assert int(round(1 / dx)) == nx
    return kw, npts, beta, nt


def make_ics(rng, B, npts):
    return rng.uniform(1, 3, size=(B, 1)) * np.ones((B, npts)) + 0.1 * rng.standard_normal((B, npts))


def reward_pair(kind, nt_r):
    """(oracle spec, product reward object) for 'tuned' or ('horizon', ord)."""
    import pdecontrolgym_b200 as pcg
    if kind == "tuned":
        return O.TunedReward1DSpec(nt_r, -1e3, 3e2), pcg.TunedReward1D(nt_r, -1e3, 3e2)
    horizon, ordv = kind
    name = {1: "1", 2: "2", np.inf: "inf"}[ordv]
    return (O.NormRewardSpec(nt_r, norm=ordv, horizon=horizon),
            pcg.NormReward(nt_r, norm=name, horizon=horizon))


def assert_rewards_close(r_gpu, r_ref, scale, tol, what=""):
    r_gpu = np.asarray(r_gpu, dtype=np.float64)
    r_ref = np.asarray(r_ref, dtype=np.float64)
    both_nan = np.isnan(r_gpu) & np.isnan(r_ref)
    same_inf = np.isinf(r_gpu) & (r_gpu == r_ref)
    exact = r_gpu == r_ref
    ok = exact | both_nan | same_inf | (np.abs(r_gpu - r_ref) <= tol * np.maximum(scale, 1e-30))
    assert ok.all(), f"{what}: rewards differ gpu={r_gpu[~ok]} ref={r_ref[~ok]} scale={np.broadcast_to(scale, r_gpu.shape)[~ok]}"
