"""GPU: vector-env protocol adapters and the ARZ backstepping driver."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_adapters_shapes_and_semantics(cuda_device):
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.adapters import GymnasiumVectorAdapter, SB3VecEnvAdapter
    kw = dict(T=0.002, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3, normalize=True, reward_class=pcg.TunedReward1D(200))
    env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=6, device=cuda_device, autoreset=True, **kw)
    opts = {"init_condition": pcg.UniformScaleIC(1, 10), "beta": np.zeros(201)}
    g = GymnasiumVectorAdapter(env)
    obs, info = g.reset(seed=3, options=opts)
    assert isinstance(obs, np.ndarray) and obs.shape == (6, 201) and obs.dtype == np.float32
    obs, r, te, tr, info = g.step(np.zeros((6, 1), dtype=np.float32))
    assert r.shape == (6,) and te.dtype == bool and not te.any()
    obs, r, te, tr, info = g.step(np.zeros((6, 1), dtype=np.float32))
    assert te.all()
    s = SB3VecEnvAdapter(env)
    obs = s.reset()
    s.step_async(np.zeros((6, 1), dtype=np.float32))
    obs, r, dones, infos = s.step_wait()
    assert obs.shape == (6, 201) and r.dtype == np.float32 and len(infos) == 6 and not dones.any()
    obs, r, dones, infos = s.step(np.zeros((6, 1), dtype=np.float32))
    assert dones.all() and not infos[0]["TimeLimit.truncated"]
    with pytest.raises(ValueError):
        SB3VecEnvAdapter(pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=2, device=cuda_device, **kw))


def test_arz_backstepping_drives_to_steady_state(cuda_device):
    """Closed loop with the tutorial's outlet law: the reward (minus distance to the steady state) improves
    markedly over the open-loop run, for every env of the batch (Trafficarz1d_tutorial.rst:153-184)."""
    import torch
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.controllers import TrafficARZBackstepping
    kw = dict(T=240, dt=0.25, X=500, dx=10, simulation_type="outlet", num_envs=4, device=cuda_device)
    closed, opened = pcg.make("PDEControlGym-TrafficPDE1D", **kw), pcg.make("PDEControlGym-TrafficPDE1D", **kw)
    ctl = TrafficARZBackstepping(closed)
    closed.reset()
    opened.reset()
    for k in range(959):
        oc, rc, *_ = closed.step(ctl())
        oo, ro, *_ = opened.step(torch.full((4, 1), 1.2, dtype=torch.float64))
    assert torch.isfinite(rc).all() and (rc > ro).all() and (rc > -0.2).all(), (rc, ro)


def test_navier_stokes_adjoint_optimiser_matches_numpy(cuda_device):
    """controllers.NavierStokesAdjoint vs a numpy restatement of examples/NavierStokes/NS2Doptimization.py:55-107 built on
    the oracle's solve_pressure: costates and actions to 1e-10 (same operation order; torch's scalar division is a
    reciprocal multiply)."""
    import torch
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.controllers import NavierStokesAdjoint
    from oracle import np_oracle as O
    N, T, n, iters = 3, 9, 21, 40
    BC = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    rng = np.random.default_rng(2)
    ut, vt = rng.uniform(-1, 1, (T + 4, n, n)), rng.uniform(-1, 1, (T + 4, n, n))
    kw = dict(boundary_condition=BC, U_ref=np.stack([ut, vt], axis=-1), action_ref=2.0 * np.ones(100), maximum_pressure_iteration=iters,
              T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=cuda_device, reward_class=pcg.NSReward(0.1), **kw)
    orc = O.NavierStokesOracle(N, gamma=0.1, **kw)
    init = tuple(rng.uniform(-1, 1, (N, n, n)) for _ in range(3))
    env.reset(options={"init": init})
    U, V = [], []
    for t in range(T):
        obs, *_ = env.step(torch.from_numpy(rng.uniform(2, 4, N)))
        U.append(obs[..., 0].clone())
        V.append(obs[..., 1].clone())
    U, V = torch.stack(U), torch.stack(V)
    adj = NavierStokesAdjoint(env)
    L1, L2 = adj.costates(U, V, ut, vt)
    acts = adj.actions(U, V, ut, vt, 2.0)

    # numpy restatement
    dx = dy = 0.05
    dt = 1e-3
    Un, Vn = U.cpu().numpy(), V.cpu().numpy()

    def walls(a1, a2):
        for a in (a1, a2):
            a[:, :, [-1, 0]] = 0.
            a[:, [-1, 0], :] = 0.
        return a1, a2
    lam1, lam2, pressure = (np.zeros((N, n, n)) for _ in range(3))
    Lam1, Lam2 = [lam1], [lam2]
    for t in range(T - 1):
        dl1dx, dl1dy = O._cd_x(lam1, dx), O._cd_y(lam1, dy)
        dl2dx, dl2dy = O._cd_x(lam2, dx), O._cd_y(lam2, dy)
        lap1, lap2 = O._lap(lam1, dx, dy), O._lap(lam2, dx, dy)
        dlam1dt = -2 * dl1dx * Un[-1 - t] - dl1dy * Vn[-1 - t] - dl2dx * Vn[-1 - t] - 0.1 * lap1 + (Un[-1 - t] - ut[-1 - t])
        dlam2dt = -2 * dl2dy * Vn[-1 - t] - dl1dy * Un[-1 - t] - dl2dx * Un[-1 - t] - 0.1 * lap2 + (Vn[-1 - t] - vt[-1 - t])
        lam1 = lam1 - dt * dlam1dt
        lam2 = lam2 - dt * dlam2dt
        lam1, lam2 = walls(lam1, lam2)
        pressure = orc.solve_pressure(lam1, lam2, pressure)
        lam1 = lam1 - dt * O._cd_x(pressure, dx)
        lam2 = lam2 - dt * O._cd_y(pressure, dy)
        lam1, lam2 = walls(lam1, lam2)
        Lam1.append(lam1)
        Lam2.append(lam2)
    Lam1, Lam2 = np.stack(Lam1[::-1]), np.stack(Lam2[::-1])
    # torch divides by a Python scalar through its reciprocal, so the costates agree to rounding, not bit for bit
    np.testing.assert_allclose(L1.cpu().numpy(), Lam1, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(L2.cpu().numpy(), Lam2, rtol=1e-10, atol=1e-13)
    want = np.stack([2.0 - 0.1 / 0.1 * O._cd_y(Lam1[t], dy)[:, -2, :].sum(axis=-1) * 5 * dx for t in range(T)])
    np.testing.assert_allclose(acts.cpu().numpy(), want, rtol=1e-10, atol=1e-12)
