"""GPU: vector-env protocol adapters and the ARZ backstepping driver."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_adapters_follow_the_vector_env_protocols(cuda_device):
    """The gymnasium (same-step and next-step auto-reset) and SB3 adapters on a CUDA env, driven by the vendored
    protocol checkers; terminal observations equal the oracle's last row, restarted rows equal the initial condition."""
    import pdecontrolgym_b200 as pcg
    from oracle import np_oracle as O
    from pdecontrolgym_b200.adapters import GymnasiumVectorAdapter, SB3VecEnvAdapter
    from tests.vecenv_protocol import drive_gymnasium, drive_sb3
    N = 6
    kw = dict(T=0.003, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3, normalize=True)       # 3 control steps per episode
    rng = np.random.default_rng(2)
    ic = rng.uniform(1, 2, (N, 1)) * np.ones((N, 201))
    beta = (50 * np.cos(8 * np.arccos(np.linspace(0, 1, 201)))).astype(np.float32)
    acts = rng.uniform(-1, 1, (8, N, 1)).astype(np.float32)
    orc = O.ReactionDiffusionOracle(N, reward=O.TunedReward1DSpec(300, -1e3, 3e2), **kw)
    orc.reset(ic, beta)
    rows = [orc.step(acts[k].reshape(-1))[0].copy() for k in range(3)]                         # one oracle episode
    opts = {"init_condition": ic, "beta": beta}

    env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=N, device=cuda_device, autoreset=True, final_obs=True,
                   reward_class=pcg.TunedReward1D(300), **kw)
    log = drive_gymnasium(GymnasiumVectorAdapter(env), lambda k: acts[k], 4, (201,), reset_kwargs={"options": opts})
    assert log[0][0].dtype == np.float32 and np.array_equal(log[0][0], rows[0]) and not log[0][2].any()
    obs, rew, term, trunc, infos = log[2]
    assert term.all() and np.array_equal(infos["final_obs"], rows[2]) and np.array_equal(obs, ic.astype(np.float32))
    assert np.array_equal(log[3][0], orc_first_step(O, kw, ic, beta, acts[3]))                 # the new episode continues from the IC

    log = drive_sb3(SB3VecEnvAdapter(env), lambda k: acts[k], 3, (201,))
    obs, rew, dones, infos = log[2]
    assert dones.all() and all(np.array_equal(infos[i]["terminal_observation"], rows[2][i]) for i in range(N))
    assert not infos[0]["TimeLimit.truncated"] and np.array_equal(obs, ic.astype(np.float32))
    with pytest.raises(ValueError):
        SB3VecEnvAdapter(pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=2, device=cuda_device, reward_class=pcg.TunedReward1D(300), **kw))

    env2 = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=N, device=cuda_device, autoreset=False,
                    reward_class=pcg.TunedReward1D(300), **kw)
    log = drive_gymnasium(GymnasiumVectorAdapter(env2, autoreset_mode="next_step"), lambda k: acts[k], 6, (201,), reset_kwargs={"options": opts})
    assert log[2][2].all() and np.array_equal(log[2][0], rows[2])                              # the finishing step returns the terminal row
    assert not log[3][2].any() and np.array_equal(log[3][0], ic.astype(np.float32)) and (log[3][1] == 0).all()
    assert np.array_equal(log[4][0], orc_first_step(O, kw, ic, beta, acts[4]))


def orc_first_step(O, kw, ic, beta, act):
    orc = O.ReactionDiffusionOracle(ic.shape[0], reward=O.TunedReward1DSpec(300, -1e3, 3e2), **kw)
    orc.reset(ic, beta)
    return orc.step(act.reshape(-1))[0]


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


def test_navier_stokes_adjoint_optimiser_matches_the_reference_run(cuda_device):
    """controllers.NavierStokesAdjoint vs tests/golden/ns_adjoint_reference_run.npz -- examples/NavierStokes/NS2Doptimization.py:62-107
    executed with the reference's own env / central_difference / laplace / solve_pressure (tests/golden/gen_golden.py::gen_ns_adjoint):
    the forward rollout is bit-identical, costates and the new action sequence agree to 1e-10 (torch's element-wise ops
    round a scalar division through a reciprocal multiply)."""
    import os
    import torch
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.controllers import NavierStokesAdjoint
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ns_adjoint_reference_run.npz"))
    N, T, n = 2, len(g["actions_in"]), 21
    BC = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=cuda_device, reward_class=pcg.NSReward(0.1), boundary_condition=BC,
                   U_ref=g["uref_head"], action_ref=2.0 * np.ones(100), maximum_pressure_iteration=60, T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)
    c = g["init"]
    env.reset(options={"init": tuple(np.full((n, n), ci) for ci in c)})
    U, V = [], []
    for t in range(T):
        obs, *_ = env.step(torch.full((N,), float(g["actions_in"][t]), dtype=torch.float64))
        U.append(obs[..., 0].clone())
        V.append(obs[..., 1].clone())
    U, V = torch.stack(U), torch.stack(V)
    for b in range(N):
        assert np.array_equal(U[:, b].cpu().numpy(), g["U"]) and np.array_equal(V[:, b].cpu().numpy(), g["V"]), "forward rollout"
    adj = NavierStokesAdjoint(env)
    L1, L2 = adj.costates(U, V, g["u_target_tail"], g["v_target_tail"])
    acts = adj.actions(U, V, g["u_target_tail"], g["v_target_tail"], 2.0)
    for b in range(N):
        np.testing.assert_allclose(L1[:, b].cpu().numpy(), g["Lam1"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(L2[:, b].cpu().numpy(), g["Lam2"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(acts[:, b].cpu().numpy(), g["actions_out"], rtol=1e-10)
