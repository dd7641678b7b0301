"""CPU: the plain-C restatement (oracle/c/pdecg_oracle.c) equals the numpy oracle bit for bit -- and through it
the live reference -- for every precision / beta dtype / action dtype / control type combination."""
import itertools

import numpy as np
import pytest

from oracle import np_oracle as O
from oracle.c_oracle import ArzCOracle, Bc1dCOracle, NsCOracle
from tests.helpers import bc1d_case, make_ics


@pytest.mark.parametrize("env", ["rd", "tr"])
def test_c_oracle_matches_numpy_oracle(env):
    for precision, bd, ad, ct, nz in itertools.product(("ref", "fp64"), (np.float32, np.float64), (np.float32, np.float64),
                                                      ("Dirchilet", "Neumann"), (False, True)):
        kw, npts, beta, nt = bc1d_case(env, None, 5)
        beta = beta.astype(bd)
        rng = np.random.default_rng(7)
        B = 6
        ics = make_ics(rng, B, npts)
        steps = 3 if (env == "rd" and ct == "Neumann" and nz) else 5
        common = dict(normalize=nz, control_type=ct, max_control_value=20)
        Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
        o = Orc(B, reward=O.NormRewardSpec(nt - 1, norm=2), precision=precision, **common, **kw)
        c = Bc1dCOracle(env, precision=precision, **common, **kw)
        o.reset(ics, beta)
        c.reset(ics, beta)
        for k in range(steps):
            a = (rng.uniform(-1, 1, B) if nz else rng.uniform(-3, 3, B)).astype(ad)
            o.step(a)
            c.step(a)
            assert c.u.dtype == o.u.dtype
            assert np.array_equal(c.u, o.u, equal_nan=True), (env, precision, bd, ad, ct, nz, k)
        assert (c.t == o.t).all()


def test_c_oracle_stops_at_episode_end_and_other_grids():
    for env, nx in (("rd", 37), ("tr", 255), ("rd", 255)):
        kw, npts, beta, nt = bc1d_case(env, nx, 2)
        ics = make_ics(np.random.default_rng(1), 3, npts)
        Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
        o = Orc(3, reward=O.NormRewardSpec(nt - 1, norm=2), precision="fp64", **kw)
        c = Bc1dCOracle(env, precision="fp64", **kw)
        o.reset(ics, beta)
        c.reset(ics, beta)
        for k in range(4):                 # two steps past the end of the episode
            a = np.full(3, 0.3, dtype=np.float32)
            o.step(a)
            c.step(a)
            assert np.array_equal(c.u, o.u)
        assert (c.t == nt - 1).all()


def test_c_oracle_large_grids():
    """No size limit in the C restatement: the grids served by the CTA-per-env CUDA kernels."""
    for env, nx in (("rd", 700), ("tr", 1500), ("rd", 4097)):
        kw, npts, beta, nt = bc1d_case(env, nx, 2)
        ics = make_ics(np.random.default_rng(2), 2, npts)
        Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
        for precision, bd in (("ref", np.float32), ("fp64", np.float64)):
            o = Orc(2, reward=O.NormRewardSpec(nt - 1, norm=2), precision=precision, normalize=True, **kw)
            c = Bc1dCOracle(env, precision=precision, normalize=True, **kw)
            o.reset(ics, beta.astype(bd))
            c.reset(ics, beta.astype(bd))
            for k in range(2):
                a = np.array([0.3, -0.7], dtype=np.float32)
                o.step(a)
                c.step(a)
                assert np.array_equal(c.u, o.u)


@pytest.mark.parametrize("st,cf,dx,dt", [("inlet", 1, 10, 0.25), ("outlet", 3, 10, 0.25), ("both", 2, 10, 0.25), ("both", 2, 4, 0.1), ("outlet", 1, 1, 0.025)])
def test_c_oracle_traffic_arz_matches_numpy_oracle(st, cf, dx, dt):
    """Plain-C Traffic-ARZ integrator == numpy oracle bit for bit (states, velocities, float time), across the
    time-based termination and on finer grids."""
    B = 5
    kw = dict(T=20 * dt / 0.25, dt=dt, X=500, dx=dx)
    o = O.TrafficARZOracle(B, simulation_type=st, control_freq=cf, **kw)
    o.reset()
    c = ArzCOracle(simulation_type=st, control_freq=cf, **kw)
    c.reset(o.r, o.y, o.qs)
    rng = np.random.default_rng(3)
    for k in range(400 if dx == 10 else 100):
        a = rng.uniform(0.7 * 1.2, 1.3 * 1.2, size=(B, 2 if st == "both" else 1))
        obs = o.step(a)[0]
        got = c.step(a)
        assert np.array_equal(got, obs), f"step {k}"
        assert np.array_equal(c.y, o.y) and np.array_equal(c.time, o.time_index), f"step {k}"


@pytest.mark.parametrize("n,iters", [(21, 37), (11, 5), (33, 12), (64, 3)])
def test_c_oracle_navier_stokes_matches_numpy_oracle(n, iters):
    """Plain-C NavierStokes2D step == numpy oracle bit for bit (u, v, p), mixed boundary conditions."""
    bc = {"upper": ["Controllable", "Neumann"], "lower": ["Neumann", "Dirchilet"], "left": ["Dirchilet", "Controllable"], "right": ["Neumann", "Neumann"]}
    h = 1.0 / (n - 1)
    P = dict(T=0.2 * (h / 0.05) ** 2, dt=1e-3 * (h / 0.05) ** 2, X=1, dx=h, Y=1, dy=h)
    B = 3
    rng = np.random.default_rng(n)
    o = O.NavierStokesOracle(B, boundary_condition=bc, U_ref=np.zeros((200, n, n, 2)), action_ref=2.0 * np.ones(1000), maximum_pressure_iteration=iters, **P)
    u0, v0, p0 = (rng.uniform(-1, 1, (B, n, n)) for _ in range(3))
    o.reset(u0, v0, p0)
    c = NsCOracle(P["dt"], h, h, n, n, bc, maximum_pressure_iteration=iters)
    c.reset(u0, v0, p0)
    for k in range(6):
        a = rng.uniform(2, 4, B)
        U = o.step(a)[0]
        got = c.step(a)
        assert np.array_equal(got, U), f"step {k}"
        assert np.array_equal(c.p, o.p), f"step {k}: pressure"


def test_c_oracle_property_random_arz_and_ns_configurations():
    """Hypothesis: random physical parameters, grids, control frequencies and boundary-condition tables -- the C and
    numpy restatements stay bit-identical (NaN / inf patterns of unstable draws included)."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
    @given(sim=st.sampled_from(["inlet", "outlet", "both"]), cf=st.integers(1, 4), M=st.integers(5, 140),
           vm=st.floats(10, 60), rm=st.floats(0.1, 0.3), tau=st.floats(5, 120), cfl=st.floats(0.1, 1.2), seed=st.integers(0, 10 ** 6))
    def arz(sim, cf, M, vm, rm, tau, cfl, seed):
        X = 500.0
        dx = X / (M - 1)
        if len(np.arange(0, X + dx, dx)) != M:
            return
        dt = cfl * dx / vm
        ro_s = 0.75 * rm
        kw = dict(T=30 * dt, dt=dt, X=X, dx=dx)
        with np.errstate(all="ignore"):
            o = O.TrafficARZOracle(3, simulation_type=sim, control_freq=cf, v_max=vm, ro_max=rm, tau=tau, ro_steady=ro_s,
                                   v_steady=vm * (1 - ro_s / rm), **kw)
            o.reset()
            c = ArzCOracle(simulation_type=sim, control_freq=cf, v_max=vm, ro_max=rm, tau=tau, **kw)
            c.reset(o.r, o.y, o.qs)
            rng = np.random.default_rng(seed)
            for k in range(12):
                a = o.qs[:, None] * rng.uniform(0.6, 1.4, size=(3, 2 if sim == "both" else 1))
                obs = o.step(a)[0]
                got = c.step(a)
                assert np.array_equal(got, obs, equal_nan=True) and np.array_equal(c.y, o.y, equal_nan=True)
                assert np.array_equal(c.time, o.time_index)

    @settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
    @given(n=st.integers(4, 40), iters=st.integers(0, 12), nu=st.floats(0.01, 0.5), rho=st.floats(0.5, 2.0),
           bcs=st.lists(st.sampled_from(["Dirchilet", "Neumann", "Controllable"]), min_size=8, max_size=8), seed=st.integers(0, 10 ** 6))
    def ns(n, iters, nu, rho, bcs, seed):
        bc = {side: [bcs[2 * i], bcs[2 * i + 1]] for i, side in enumerate(("upper", "lower", "left", "right"))}
        h = 1.0 / (n - 1)
        dt = 0.2 * (0.5 * h * h / nu)
        P = dict(T=50 * dt, dt=dt, X=1, dx=h, Y=1, dy=h)
        rng = np.random.default_rng(seed)
        o = O.NavierStokesOracle(2, boundary_condition=bc, U_ref=np.zeros((60, n, n, 2)), action_ref=np.ones(100),
                                 maximum_pressure_iteration=iters, viscosity=nu, density=rho, **P)
        u0, v0, p0 = (rng.uniform(-2, 2, (2, n, n)) for _ in range(3))
        o.reset(u0, v0, p0)
        c = NsCOracle(dt, h, h, n, n, bc, viscosity=nu, density=rho, maximum_pressure_iteration=iters)
        c.reset(u0, v0, p0)
        for k in range(4):
            a = rng.uniform(-3, 3, 2)
            with np.errstate(all="ignore"):
                U = o.step(a)[0]
            assert np.array_equal(c.step(a), U, equal_nan=True) and np.array_equal(c.p, o.p, equal_nan=True)

    arz()
    ns()
