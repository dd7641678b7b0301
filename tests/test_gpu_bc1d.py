"""CUDA parity: TransportPDE1D / ReactionDiffusionPDE1D kernels vs the numpy oracle, through the
public vector-env API (which calls the C ABI).  States / observations must be BIT-IDENTICAL in the
`ref` and `fp64` modes; rewards (norm reductions) within a stated tolerance."""
import itertools

import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import (TOL_F32_REDUCTION, TOL_F64_REDUCTION, assert_rewards_close, bc1d_case,
                           make_ics, reward_pair)

pytestmark = pytest.mark.gpu


def run_pair(cuda_device, env, precision, control_type, sensing, normalize, beta_dtype, act_dtype,
             reward_kind, nx=None, B=5, steps=12, limit=False, seed=0):
    import torch
    import pdecontrolgym_b200 as pcg
    kw, npts, beta, nt = bc1d_case(env, nx, steps)
    sensing_loc, sensing_type = sensing
    nt_r = nt - 1
    ospec, prew = reward_pair(reward_kind, nt_r)
    common = dict(normalize=normalize, sensing_loc=sensing_loc, control_type=control_type,
                  sensing_type=sensing_type, limit_pde_state_size=limit, max_state_value=30.0,
                  max_control_value=20, **kw)
    rng = np.random.default_rng(seed)
    ics = make_ics(rng, B, npts)
    beta = beta.astype(beta_dtype)
    Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
    oracle_prec = "fp64" if precision == "fp64" else "ref"
    oenv = Orc(B, reward=ospec, precision=oracle_prec, **common)
    o_ref = oenv.reset(ics, beta)
    eid = "PDEControlGym-ReactionDiffusionPDE1D" if env == "rd" else "PDEControlGym-TransportPDE1D"
    genv = pcg.make(eid, num_envs=B, device=cuda_device, precision=precision, reward_class=prew,
                    sensing_noise_func=None, reset_init_condition_func=None,
                    reset_recirculation_func=None, **common)
    o_gpu, _ = genv.reset(options={"init_condition": ics, "beta": beta})
    assert np.array_equal(o_gpu.cpu().numpy(), o_ref), "reset observation"
    tol = TOL_F64_REDUCTION if precision == "fp64" else TOL_F32_REDUCTION
    for k in range(steps + 1):          # one extra call on finished envs
        a = (rng.uniform(-1, 1, B) if normalize else rng.uniform(-3, 3, B)).astype(act_dtype)
        o_ref, r_ref, te_ref, tr_ref = oenv.step(a)
        o_gpu, r_gpu, te_gpu, tr_gpu, _ = genv.step(torch.from_numpy(a))
        og = o_gpu.cpu().numpy()
        assert og.dtype == o_ref.dtype
        if precision == "fp32":
            np.testing.assert_allclose(og, o_ref, rtol=2e-3, atol=2e-3)
        else:
            assert np.array_equal(og, o_ref, equal_nan=True), f"step {k}: state/obs not bit-identical (max |d|={np.nanmax(np.abs(og - o_ref))})"
            assert np.array_equal(te_gpu.cpu().numpy(), te_ref) and np.array_equal(tr_gpu.cpu().numpy(), tr_ref), f"step {k}: flags"
        if k < steps:
            scale = 2 * np.sqrt((oenv.u.astype(np.float64) ** 2).sum(axis=1)) + np.abs(r_ref)
            if precision == "fp32":
                assert_rewards_close(r_gpu.cpu().numpy(), r_ref, scale + 1.0, 5e-3, f"step {k}")
            else:
                assert_rewards_close(r_gpu.cpu().numpy(), r_ref, scale, tol, f"step {k}")
    assert te_ref.all()


SENSINGS = [("full", None), ("collocated", None), ("opposite", "Neumann"), ("opposite", "Dirchilet")]


@pytest.mark.parametrize("env", ["rd", "tr"])
@pytest.mark.parametrize("precision,beta_dtype", [("ref", np.float32), ("ref", np.float64), ("fp64", np.float64)])
def test_default_grid_all_variants(cuda_device, env, precision, beta_dtype):
    for control_type, sensing, normalize, act_dtype, rk in itertools.product(
            ("Dirchilet", "Neumann"), SENSINGS, (False, True), (np.float32, np.float64),
            ("tuned", ("temporal", 2))):
        if env == "rd" and sensing == ("opposite", "Dirchilet"):
            continue
        if env == "rd" and control_type == "Neumann" and normalize:
            steps = 3   # reference quirk Q4: blows up to inf/nan within a few steps; still compared
        else:
            steps = 12
        run_pair(cuda_device, env, precision, control_type, sensing, normalize, beta_dtype, act_dtype, rk,
                 steps=steps)


@pytest.mark.parametrize("env", ["rd", "tr"])
@pytest.mark.parametrize("rk", ["tuned", ("temporal", 1), ("temporal", np.inf), ("differential", 2),
                                ("differential", 1), ("t-horizon", 2)])
def test_reward_kinds(cuda_device, env, rk):
    for precision, bd in (("ref", np.float32), ("fp64", np.float64)):
        run_pair(cuda_device, env, precision, "Dirchilet", ("full", None), True, bd, np.float32, rk)


@pytest.mark.parametrize("env", ["rd", "tr"])
@pytest.mark.parametrize("nx", [8, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 160, 191, 192, 200, 223, 224, 255, 256, 300, 384, 447, 511])
def test_grid_sizes(cuda_device, env, nx):
    """Every points-per-lane instantiation (K = 1..8, 10, 12, 14, 16), incl. sizes that are not multiples of 32."""
    for precision, bd in (("ref", np.float32), ("ref", np.float64), ("fp64", np.float64)):
        run_pair(cuda_device, env, precision, "Dirchilet", ("full", None), False, bd, np.float32, "tuned",
                 nx=nx, B=3, steps=4)
        run_pair(cuda_device, env, precision, "Neumann", ("collocated", None), False, bd, np.float64, ("temporal", 2),
                 nx=nx, B=3, steps=4)


@pytest.mark.parametrize("env", ["rd", "tr"])
def test_truncation(cuda_device, env):
    run_pair(cuda_device, env, "ref", "Dirchilet", ("full", None), False, np.float32, np.float32, "tuned", limit=True)
    run_pair(cuda_device, env, "fp64", "Dirchilet", ("full", None), False, np.float64, np.float32, ("temporal", 2), limit=True)


@pytest.mark.parametrize("env", ["rd", "tr"])
def test_fp32_fast_mode(cuda_device, env):
    run_pair(cuda_device, env, "fp32", "Dirchilet", ("full", None), True, np.float32, np.float32, "tuned", steps=6)


def test_batch_sizes_and_independence(cuda_device):
    """Ragged batch sizes (not multiples of the 4-env CTA) and batch-composition invariance."""
    import torch
    import pdecontrolgym_b200 as pcg
    kw, npts, beta, nt = bc1d_case("rd", None, 3)
    rng = np.random.default_rng(3)
    ics = make_ics(rng, 37, npts)
    acts = rng.uniform(-1, 1, (3, 37)).astype(np.float32)
    outs = {}
    for B in (1, 2, 3, 5, 37):
        env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=B, device=cuda_device, precision="fp64",
                       reward_class=pcg.NormReward(nt - 1), normalize=True, **kw)
        env.reset(options={"init_condition": ics[:B], "beta": beta})
        for k in range(3):
            o, r, te, tr, _ = env.step(torch.from_numpy(acts[k, :B]))
        outs[B] = (o.cpu().numpy(), r.cpu().numpy())
    for B in (1, 2, 3, 5):
        assert np.array_equal(outs[B][0], outs[37][0][:B])
        assert np.array_equal(outs[B][1], outs[37][1][:B])


def test_empty_batch_and_errors(cuda_device):
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200._lib import PdecgError
    kw, npts, beta, nt = bc1d_case("tr", None, 3)
    env = pcg.make("PDEControlGym-TransportPDE1D", num_envs=0, device=cuda_device,
                   reward_class=pcg.TunedReward1D(nt - 1), **kw)
    o, _ = env.reset(options={"init_condition": np.zeros(npts), "beta": beta})
    o, r, te, tr, _ = env.step(np.zeros(0, dtype=np.float32))
    assert o.shape == (0, npts) and r.shape == (0,)
    with pytest.raises(Exception, match="Invalid sensing_loc"):
        pcg.make("PDEControlGym-TransportPDE1D", num_envs=1, device=cuda_device, sensing_loc="middle",
                 reward_class=pcg.TunedReward1D(nt - 1), **kw)
    with pytest.raises(Exception, match="not viable"):
        pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=1, device=cuda_device, sensing_loc="opposite",
                 sensing_type="Dirchilet", reward_class=pcg.TunedReward1D(nt - 1), **kw)
    kw2 = dict(kw)
    kw2["dx"] = 1.0 / 9000
    big = pcg.make("PDEControlGym-TransportPDE1D", num_envs=1, device=cuda_device,
                   reward_class=pcg.NormReward(nt - 1), **kw2)
    with pytest.raises(PdecgError, match="shared-memory-resident"):
        big.reset(options={"init_condition": np.zeros(9000), "beta": np.zeros(9000)})


@pytest.mark.parametrize("autoreset", [False, True])
@pytest.mark.parametrize("N", [3, 4099, 20001, 40003])      # 1 / 2 / 5 / 9 chunks (a short first one) over four streams
def test_step_host_matches_device_step(cuda_device, N, autoreset):
    """pdecg_bc1d_step_host (one H2D -> chunked kernels pipelined over four streams -> D2H per chunk, reward / flags once)
    == pdecg_bc1d_step, with and without the fused same-step auto-reset (per-env initial rows: the chunk offsets matter)."""
    import torch
    import pdecontrolgym_b200 as pcg
    kw, npts, beta, nt = bc1d_case("rd", None, 3)
    rng = np.random.default_rng(N)
    ics = make_ics(rng, N, npts)
    mk = lambda: pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=N, device=cuda_device, precision="fp64",  # noqa: E731
                          reward_class=pcg.NormReward(nt - 1), normalize=True, autoreset=autoreset, **kw)
    e1, e2 = mk(), mk()
    for e in (e1, e2):
        e.reset(options={"init_condition": ics, "beta": np.tile(beta, (N, 1)) * rng.uniform(0.9, 1.1)})
    e2.beta.copy_(e1.beta)
    obs_h = torch.empty(N, npts, dtype=torch.float64).pin_memory()
    rew_h = torch.empty(N, dtype=torch.float64).pin_memory()
    te_h = torch.empty(N, dtype=torch.uint8).pin_memory()
    tr_h = torch.empty(N, dtype=torch.uint8).pin_memory()
    for k in range(7 if autoreset else 3):
        a = rng.uniform(-1, 1, N).astype(np.float32)
        o, r, te, tr, _ = e1.step(torch.from_numpy(a))
        e2.step_host(torch.from_numpy(a).pin_memory(), obs_h, rew_h, te_h, tr_h)
        assert np.array_equal(o.cpu().numpy(), obs_h.numpy())
        assert np.array_equal(r.cpu().numpy(), rew_h.numpy())
        assert np.array_equal(te.cpu().numpy(), te_h.numpy().astype(bool)) and np.array_equal(tr.cpu().numpy(), tr_h.numpy().astype(bool))
        assert bool(te_h.all()) == (k % 3 == 2)
        if autoreset and k % 3 == 2:
            assert np.array_equal(obs_h.numpy(), ics) and torch.equal(e1.time_index, e2.time_index) and int(e2.time_index.max()) == 0


@pytest.mark.parametrize("env", ["rd", "tr"])
@pytest.mark.parametrize("sub,steps,extra", [(30, 6, 0), (7, 20, 3), (100, 3, 40), (150, 2, 77), (64, 4, 0)])
def test_tuned_reward_any_substep_count(cuda_device, env, sub, steps, extra):
    """TunedReward1D look-back (||u[t-100]||) for substep counts that neither divide 100 nor exceed it, and for
    episodes whose last control step is cut short (nt-1 not a multiple of substeps): general 101-entry ring."""
    import torch
    import pdecontrolgym_b200 as pcg
    nx = 40
    dx = 1.0 / nx
    dt = 0.4 * dx * dx if env == "rd" else dx / 100
    npts = nx + 1 if env == "rd" else nx
    n_sub_total = sub * steps + extra
    T = dt * n_sub_total
    nt = int(round(T / dt) + 1)
    assert nt - 1 == n_sub_total and nt > 100
    kw = dict(T=T, dt=dt, X=1, dx=dx, control_sample_rate=dt * sub, normalize=True, max_control_value=20)
    rng = np.random.default_rng(sub)
    B = 4
    ics = make_ics(rng, B, npts)
    beta = np.cos(np.linspace(0, 3, npts)).astype(np.float32)
    Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
    oenv = Orc(B, reward=O.TunedReward1DSpec(nt - 1, -1e3, 3e2), **kw)
    eid = "PDEControlGym-ReactionDiffusionPDE1D" if env == "rd" else "PDEControlGym-TransportPDE1D"
    genv = pcg.make(eid, num_envs=B, device=cuda_device, reward_class=pcg.TunedReward1D(nt - 1, -1e3, 3e2), **kw)
    oenv.reset(ics, beta)
    genv.reset(options={"init_condition": ics, "beta": beta})
    n_calls = steps + (1 if extra else 0)
    for k in range(n_calls):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        o_r, r_r, te_r, tr_r = oenv.step(a)
        o_g, r_g, te_g, tr_g, _ = genv.step(torch.from_numpy(a))
        assert np.array_equal(o_g.cpu().numpy(), o_r) and np.array_equal(te_g.cpu().numpy(), te_r)
        scale = 2 * np.sqrt((oenv.u.astype(np.float64) ** 2).sum(axis=1)) + np.abs(r_r)
        assert_rewards_close(r_g.cpu().numpy(), r_r, scale, TOL_F32_REDUCTION, f"call {k}")
    assert te_r.all()


@pytest.mark.parametrize("env,precision", [("rd", "fp64"), ("rd", "ref"), ("tr", "ref"), ("tr", "fp64")])
@pytest.mark.parametrize("rk", ["tuned", ("temporal", 2)])
def test_staggered_time_indices_within_a_warp(cuda_device, env, precision, rk):
    """Two envs share a warp on the default grids (16 lanes each).  Envs of a pair that sit at different time
    indices (masked reset) or finish at different steps must not disturb each other; odd batch -> a lone env."""
    import torch
    import pdecontrolgym_b200 as pcg
    steps, B = 6, 5
    kw, npts, beta, nt = bc1d_case(env, None, steps)
    ospec, prew = reward_pair(rk, nt - 1)
    common = dict(normalize=True, sensing_loc="full", control_type="Dirchilet", sensing_type=None,
                  limit_pde_state_size=False, max_state_value=1e10, max_control_value=20, **kw)
    rng = np.random.default_rng(11)
    ics, ics2 = make_ics(rng, B, npts), make_ics(rng, B, npts)
    bdt = np.float64 if precision == "fp64" else np.float32
    beta = beta.astype(bdt)
    Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
    oprec = "fp64" if precision == "fp64" else "ref"
    old, new = Orc(B, reward=ospec, precision=oprec, **common), Orc(B, reward=reward_pair(rk, nt - 1)[0], precision=oprec, **common)
    old.reset(ics, beta)
    eid = "PDEControlGym-ReactionDiffusionPDE1D" if env == "rd" else "PDEControlGym-TransportPDE1D"
    genv = pcg.make(eid, num_envs=B, device=cuda_device, precision=precision, reward_class=prew, **common)
    genv.reset(options={"init_condition": ics, "beta": beta})
    for k in range(2):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        old.step(a)
        genv.step(torch.from_numpy(a))
    mask = np.array([False, True, True, False, False])
    genv.reset(options={"init_condition": ics2, "beta": beta, "mask": mask})
    new.reset(ics2, beta)
    tol = TOL_F64_REDUCTION if precision == "fp64" else TOL_F32_REDUCTION
    for k in range(steps + 1):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        o_old, r_old, te_old, _ = old.step(a)
        o_new, r_new, te_new, _ = new.step(a)
        o_gpu, r_gpu, te_gpu, _, _ = genv.step(torch.from_numpy(a))
        o_ref = np.where(mask[:, None], o_new, o_old)
        assert np.array_equal(o_gpu.cpu().numpy(), o_ref), f"step {k}: state not bit-identical"
        assert np.array_equal(te_gpu.cpu().numpy(), np.where(mask, te_new, te_old)), f"step {k}: terminated"
        live = np.where(mask, k < steps, k < steps - 2)         # rewards of finished envs are not compared
        r_ref = np.where(mask, r_new, r_old)
        scale = 2 * np.sqrt((o_ref.astype(np.float64) ** 2).sum(axis=1)) + np.abs(r_ref)
        assert_rewards_close(r_gpu.cpu().numpy()[live], r_ref[live], scale[live], tol, f"step {k}")


# ---- grids beyond 512 points per env: the CTA-per-env shared-memory kernel (bc1d_large.cuh) ----------------------
@pytest.mark.parametrize("env", ["rd", "tr"])
@pytest.mark.parametrize("nx", [512, 513, 600, 1024, 1025, 2500, 4097, 8191])
def test_large_grid_sizes(cuda_device, env, nx):
    """256-, 512- and 1024-thread launches, sizes that are not multiples of the block size, every precision mode."""
    steps = 3 if nx <= 1025 else 2
    for precision, bd in (("ref", np.float32), ("ref", np.float64), ("fp64", np.float64)):
        run_pair(cuda_device, env, precision, "Dirchilet", ("full", None), False, bd, np.float32, "tuned",
                 nx=nx, B=3, steps=steps)
        run_pair(cuda_device, env, precision, "Neumann", ("collocated", None), False, bd, np.float64, ("temporal", 2),
                 nx=nx, B=3, steps=steps)


@pytest.mark.parametrize("env", ["rd", "tr"])
def test_large_grid_all_variants(cuda_device, env):
    """Control / sensing / normalisation / action-dtype / reward variants and truncation on a 700-point grid."""
    for control_type, sensing, normalize, act_dtype in itertools.product(
            ("Dirchilet", "Neumann"), SENSINGS, (False, True), (np.float32, np.float64)):
        if env == "rd" and sensing == ("opposite", "Dirchilet"):
            continue
        steps = 3 if (env == "rd" and control_type == "Neumann" and normalize) else 4
        run_pair(cuda_device, env, "ref", control_type, sensing, normalize, np.float32, act_dtype, "tuned", nx=700, B=2, steps=steps)
        run_pair(cuda_device, env, "fp64", control_type, sensing, normalize, np.float64, act_dtype, ("temporal", 2), nx=700, B=2, steps=steps)
    for rk in [("temporal", 1), ("temporal", np.inf), ("differential", 2), ("differential", 1), ("differential", np.inf), ("t-horizon", 2)]:
        for precision, bd in (("ref", np.float32), ("fp64", np.float64)):
            run_pair(cuda_device, env, precision, "Dirchilet", ("full", None), True, bd, np.float32, rk, nx=700, B=2, steps=4)
    run_pair(cuda_device, env, "ref", "Dirchilet", ("full", None), False, np.float32, np.float32, "tuned", nx=700, B=3, steps=12, limit=True)
    run_pair(cuda_device, env, "fp64", "Dirchilet", ("full", None), False, np.float64, np.float32, ("temporal", 2), nx=700, B=3, steps=12, limit=True)
    run_pair(cuda_device, env, "fp32", "Dirchilet", ("full", None), True, np.float32, np.float32, "tuned", nx=700, B=2, steps=3)


@pytest.mark.parametrize("env", ["rd", "tr"])
@pytest.mark.parametrize("sub,steps,extra", [(30, 5, 0), (150, 2, 77)])
def test_large_grid_tuned_reward_general_ring(cuda_device, env, sub, steps, extra):
    """TunedReward1D look-back through the 101-entry per-substep ring on a 640-point grid."""
    import torch
    import pdecontrolgym_b200 as pcg
    nx = 640
    dx = 1.0 / nx
    dt = 0.4 * dx * dx if env == "rd" else dx / 100
    npts = nx + 1 if env == "rd" else nx
    n_sub_total = sub * steps + extra
    T = dt * n_sub_total
    nt = int(round(T / dt) + 1)
    assert nt - 1 == n_sub_total and nt > 100
    kw = dict(T=T, dt=dt, X=1, dx=dx, control_sample_rate=dt * sub, normalize=True, max_control_value=20)
    rng = np.random.default_rng(sub)
    B = 3
    ics = make_ics(rng, B, npts)
    beta = np.cos(np.linspace(0, 3, npts)).astype(np.float32)
    Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
    oenv = Orc(B, reward=O.TunedReward1DSpec(nt - 1, -1e3, 3e2), **kw)
    eid = "PDEControlGym-ReactionDiffusionPDE1D" if env == "rd" else "PDEControlGym-TransportPDE1D"
    genv = pcg.make(eid, num_envs=B, device=cuda_device, reward_class=pcg.TunedReward1D(nt - 1, -1e3, 3e2), **kw)
    oenv.reset(ics, beta)
    genv.reset(options={"init_condition": ics, "beta": beta})
    for k in range(steps + (1 if extra else 0)):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        o_r, r_r, te_r, tr_r = oenv.step(a)
        o_g, r_g, te_g, tr_g, _ = genv.step(torch.from_numpy(a))
        assert np.array_equal(o_g.cpu().numpy(), o_r) and np.array_equal(te_g.cpu().numpy(), te_r)
        scale = 2 * np.sqrt((oenv.u.astype(np.float64) ** 2).sum(axis=1)) + np.abs(r_r)
        assert_rewards_close(r_g.cpu().numpy(), r_r, scale, TOL_F32_REDUCTION, f"call {k}")
    assert te_r.all()


@pytest.mark.parametrize("env_id,precision", [("PDEControlGym-ReactionDiffusionPDE1D", "fp64"), ("PDEControlGym-TransportPDE1D", "ref")])
@pytest.mark.parametrize("random_ic", [False, True])
def test_large_grid_fused_autoreset_and_step_host(cuda_device, env_id, precision, random_ic):
    """On a 900-point grid: fused same-step auto-reset == step + masked reset, and step_host == step."""
    import torch
    import pdecontrolgym_b200 as pcg
    rd = "Reaction" in env_id
    N, steps, nx = 5, 3, 900
    dx = 1.0 / nx
    if rd:
        dt, sub, npts = 0.4 * dx * dx, 100, nx + 1
    else:
        dt, sub, npts = dx / 100, 300, nx
    kw = dict(T=steps * sub * dt, dt=dt, X=1, dx=dx, control_sample_rate=sub * dt, normalize=True, precision=precision,
              limit_pde_state_size=True, max_state_value=1e10)
    kw["reward_class"] = pcg.TunedReward1D(int(round(kw["T"] / dt)), -1e3, 3e2)
    rng = np.random.default_rng(9)
    beta = rng.uniform(-3, 3, (N, npts)).astype(np.float64 if precision == "fp64" else np.float32)
    ic = pcg.UniformScaleIC(1.0, 10.0) if random_ic else rng.uniform(1, 2, (N, npts))
    fused = pcg.make(env_id, num_envs=N, device=cuda_device, autoreset=True, **kw)
    manual = pcg.make(env_id, num_envs=N, device=cuda_device, autoreset=False, **kw)
    hosted = pcg.make(env_id, num_envs=N, device=cuda_device, autoreset=True, **kw)
    for e in (fused, manual, hosted):
        e.reset(seed=5, options={"init_condition": ic, "beta": beta})
    stag = np.array([True, False, False, True, True])
    a0 = torch.zeros(N, device=cuda_device)
    for e in (fused, manual, hosted):
        e.step(a0)
        e.reset(options={"mask": stag})
    sdt = torch.float64 if precision == "fp64" else torch.float32
    obs_h = torch.empty(N, npts, dtype=sdt).pin_memory()
    rew_h = torch.empty(N, dtype=torch.float64).pin_memory()
    te_h = torch.empty(N, dtype=torch.uint8).pin_memory()
    tr_h = torch.empty(N, dtype=torch.uint8).pin_memory()
    n_done = 0
    for k in range(2 * steps + 2):
        a = torch.from_numpy(rng.uniform(-1, 1, N).astype(np.float32))
        o1, r1, te1, tr1, _ = fused.step(a)
        o2, r2, te2, tr2, _ = manual.step(a)
        hosted.step_host(a.pin_memory(), obs_h, rew_h, te_h, tr_h)
        done = te2 | tr2
        if done.any():
            o2, _ = manual.reset(options={"mask": done})
            n_done += int(done.sum())
        assert torch.equal(te1, te2) and torch.equal(tr1, tr2) and torch.equal(r1, r2), f"step {k}: outputs"
        assert torch.equal(o1, o2), f"step {k}: observation"
        assert np.array_equal(o1.cpu().numpy(), obs_h.numpy()) and np.array_equal(r1.cpu().numpy(), rew_h.numpy())
        assert np.array_equal(te1.cpu().numpy(), te_h.numpy().astype(bool))
        for name in ("u", "time_index", "_carry", "_episode"):
            assert torch.equal(getattr(fused, name), getattr(manual, name)), f"step {k}: {name}"
    assert n_done >= N


@pytest.mark.parametrize("env", ["rd", "tr"])
def test_large_grid_shared_memory_variant(cuda_device, env, monkeypatch):
    """The shared-memory CTA kernel (kept for RD + Neumann control, where the actuation changes every substep) on the
    configurations the register-tiled kernel normally serves."""
    monkeypatch.setenv("PDECG_BC1D_LARGE_SMEM", "1")
    for nx in (513, 1500, 4097):
        for precision, bd in (("ref", np.float32), ("fp64", np.float64)):
            run_pair(cuda_device, env, precision, "Dirchilet", ("full", None), True, bd, np.float32, "tuned", nx=nx, B=2, steps=2)
    for rk in [("temporal", 1), ("differential", 2), ("differential", np.inf), ("t-horizon", 2)]:
        run_pair(cuda_device, env, "fp64", "Dirchilet", ("opposite", "Neumann"), True, np.float64, np.float64, rk, nx=700, B=2, steps=3)
    run_pair(cuda_device, env, "ref", "Dirchilet", ("full", None), False, np.float32, np.float32, "tuned", nx=700, B=3, steps=12, limit=True)


@pytest.mark.parametrize("env", ["rd", "tr"])
def test_small_batch_latency_mapping(cuda_device, env, monkeypatch):
    """Default lane mapping of batches below 1,536 envs (one env per warp, fewest points per lane) on the reference
    grids: same bits as the oracle, and step_host() of a chunked batch agrees with step() bit for bit (rewards included:
    both use the mapping of the whole batch)."""
    import torch
    import pdecontrolgym_b200 as pcg
    monkeypatch.delenv("PDECG_BC1D_HALFWARP_MIN")
    for precision, bd in (("ref", np.float32), ("ref", np.float64), ("fp64", np.float64)):
        run_pair(cuda_device, env, precision, "Dirchilet", ("full", None), True, bd, np.float32, "tuned", B=5, steps=4)
        run_pair(cuda_device, env, precision, "Neumann", ("collocated", None), False, bd, np.float64, ("differential", 2), B=3, steps=3)
    # a batch above the threshold whose host-step chunks fall below it
    monkeypatch.setenv("PDECG_BC1D_HALFWARP_MIN", "4096")
    kw, npts, beta, nt = bc1d_case(env, None, 2)
    N = 12290                      # 3 chunks of 4100, 4100 and 4090 envs
    eid = "PDEControlGym-ReactionDiffusionPDE1D" if env == "rd" else "PDEControlGym-TransportPDE1D"
    rng = np.random.default_rng(1)
    ics = make_ics(rng, N, npts)
    mk = lambda: pcg.make(eid, num_envs=N, device=cuda_device, precision="ref", reward_class=pcg.NormReward(nt - 1), normalize=True, **kw)  # noqa: E731
    e1, e2 = mk(), mk()
    for e in (e1, e2):
        e.reset(options={"init_condition": ics, "beta": beta.astype(np.float32)})
    obs_h = torch.empty(N, npts, dtype=torch.float32).pin_memory()
    rew_h = torch.empty(N, dtype=torch.float64).pin_memory()
    a = rng.uniform(-1, 1, N).astype(np.float32)
    o, r, te, tr, _ = e1.step(torch.from_numpy(a))
    e2.step_host(torch.from_numpy(a).pin_memory(), obs_h, rew_h, None, None)
    assert np.array_equal(o.cpu().numpy(), obs_h.numpy()) and np.array_equal(r.cpu().numpy(), rew_h.numpy())


def test_halfwarp_pair_with_per_env_beta_in_different_binades(cuda_device):
    """Two envs share a warp (16 lanes each) and run the substep loop in lock step.  Their growth bound
    rho = 1 + 4F + max|dt*beta| decides how many substeps may use the exact-FMA form; with PER-ENV beta rows that put
    rho on either side of a power of two (F just below 0.25: rho ~ 2) the two envs must still agree on that count.
    States stay bit-identical to the oracle for both."""
    import torch
    import pdecontrolgym_b200 as pcg
    nx, sub, steps, B = 200, 100, 6, 8
    dx = 1.0 / nx
    dt = 0.2499 * dx * dx
    kw = dict(T=dt * sub * steps, dt=dt, X=1, dx=dx, control_sample_rate=dt * sub, normalize=True, max_control_value=20)
    nt_r = sub * steps
    rng = np.random.default_rng(11)
    ics = make_ics(rng, B, nx + 1)
    shape = np.cos(8 * np.arccos(np.linspace(0, 1, nx + 1)))
    beta = np.empty((B, nx + 1))
    beta[1::2] = (1e-3 / dt) * shape[None, :]                            # odd envs: max|dt*beta| = 1e-3 -> rho = 2.0006 > 2
    beta[0::2] = (1e-5 / dt) * shape[None, :]                            # even envs: max|dt*beta| = 1e-5 -> rho = 1.99961 < 2
    assert 1 + 4 * 0.2499 + 1e-5 < 2 < 1 + 4 * 0.2499 + 1e-3
    oenv = O.ReactionDiffusionOracle(B, reward=O.NormRewardSpec(nt_r, norm=2), precision="fp64", **kw)
    genv = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=B, device=cuda_device, precision="fp64",
                    reward_class=pcg.NormReward(nt_r, norm="2"), **kw)
    o_ref = oenv.reset(ics, beta)
    o_gpu, _ = genv.reset(options={"init_condition": ics, "beta": beta})
    assert np.array_equal(o_gpu.cpu().numpy(), o_ref)
    for k in range(steps):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        o_ref, r_ref, te_ref, tr_ref = oenv.step(a)
        o_gpu, r_gpu, te_gpu, tr_gpu, _ = genv.step(torch.from_numpy(a))
        assert np.array_equal(o_gpu.cpu().numpy(), o_ref), f"step {k}"
        np.testing.assert_allclose(r_gpu.cpu().numpy(), r_ref, rtol=1e-12)
    assert te_ref.all() and te_gpu.all()


@pytest.mark.parametrize("env,nx", [("rd", None), ("tr", None), ("rd", 700), ("tr", 2500)])
def test_float32_observations_and_final_obs(cuda_device, env, nx):
    """obs_dtype='float32' on a float64 batch: every observation the kernels hand out (step, reset, fused restart,
    final_obs, step_host) is the oracle's float64 row rounded once to float32 (`astype(np.float32)`), while the resident
    state stays float64 and bit-identical.  final_obs holds the terminal row of the envs that restarted; the returned
    observation of those envs is the new episode's initial row.  Warp kernels and both CTA-per-env kernels."""
    import torch
    import pdecontrolgym_b200 as pcg
    steps, B = 4, 6
    kw, npts, beta, nt = bc1d_case(env, nx, steps)
    nt_r = nt - 1
    common = dict(normalize=True, max_control_value=20, **kw)
    rng = np.random.default_rng(5)
    ics = make_ics(rng, B, npts)
    Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
    oenv = Orc(B, reward=O.NormRewardSpec(nt_r, norm=2), precision="fp64", **common)
    eid = "PDEControlGym-ReactionDiffusionPDE1D" if env == "rd" else "PDEControlGym-TransportPDE1D"
    mk = lambda **extra: pcg.make(eid, num_envs=B, device=cuda_device, precision="fp64", reward_class=pcg.NormReward(nt_r, norm="2"),  # noqa: E731
                                  autoreset=True, obs_dtype="float32", final_obs=True, **extra, **common)
    genv, henv = mk(), mk()
    o_ref = oenv.reset(ics, beta)
    o_gpu, _ = genv.reset(options={"init_condition": ics, "beta": beta})
    henv.reset(options={"init_condition": ics, "beta": beta})
    assert o_gpu.dtype == torch.float32 and np.array_equal(o_gpu.cpu().numpy(), o_ref.astype(np.float32))
    obs_h = torch.empty(B, npts, dtype=torch.float32).pin_memory()
    rew_h, te_h, tr_h = torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.uint8).pin_memory(), torch.empty(B, dtype=torch.uint8).pin_memory()
    for k in range(steps):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        o_ref, r_ref, te_ref, tr_ref = oenv.step(a)
        o_gpu, r_gpu, te_gpu, tr_gpu, info = genv.step(torch.from_numpy(a))
        henv.step_host(torch.from_numpy(a), obs_h, rew_h, te_h, tr_h)
        assert o_gpu.dtype == torch.float32
        if k < steps - 1:
            assert np.array_equal(o_gpu.cpu().numpy(), o_ref.astype(np.float32)), f"step {k}"
            assert np.array_equal(genv.u.cpu().numpy(), o_ref), f"step {k}: resident float64 state"
            assert np.array_equal(obs_h.numpy(), o_ref.astype(np.float32)), f"step {k}: step_host"
        else:
            assert te_gpu.all() and te_h.numpy().all()
            fin = info["final_obs"]
            assert fin.dtype == torch.float32 and np.array_equal(fin.cpu().numpy(), o_ref.astype(np.float32)), "terminal observation"
            assert np.array_equal(henv._final_obs.cpu().numpy(), o_ref.astype(np.float32)), "terminal observation (step_host)"
            assert np.array_equal(o_gpu.cpu().numpy(), ics.astype(np.float32)), "post-reset observation is the initial row"
            assert np.array_equal(obs_h.numpy(), ics.astype(np.float32))
            assert np.array_equal(genv.u.cpu().numpy(), ics)
        np.testing.assert_allclose(r_gpu.cpu().numpy(), r_ref, rtol=1e-12)
        np.testing.assert_allclose(rew_h.numpy(), r_ref, rtol=1e-12)


@pytest.mark.parametrize("sensing", [("collocated", None), ("opposite", "Neumann")])
def test_final_obs_scalar_sensing(cuda_device, sensing):
    """Scalar sensing: final_obs is [num_envs] and holds the terminal sensor value of the restarted envs."""
    import torch
    import pdecontrolgym_b200 as pcg
    steps, B = 3, 5
    kw, npts, beta, nt = bc1d_case("tr", None, steps)
    common = dict(normalize=True, max_control_value=20, sensing_loc=sensing[0], sensing_type=sensing[1], **kw)
    rng = np.random.default_rng(6)
    ics = make_ics(rng, B, npts)
    oenv = O.TransportOracle(B, reward=O.TunedReward1DSpec(nt - 1, -1e3, 3e2), precision="ref", **common)
    genv = pcg.make("PDEControlGym-TransportPDE1D", num_envs=B, device=cuda_device, precision="ref",
                    reward_class=pcg.TunedReward1D(nt - 1, -1e3, 3e2), autoreset=True, final_obs=True, **common)
    oenv.reset(ics, beta.astype(np.float32))
    o0, _ = genv.reset(options={"init_condition": ics, "beta": beta.astype(np.float32)})
    for k in range(steps):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        o_ref, r_ref, te_ref, tr_ref = oenv.step(a)
        o_gpu, r_gpu, te_gpu, tr_gpu, info = genv.step(torch.from_numpy(a))
    assert te_gpu.all()
    assert np.array_equal(info["final_obs"].cpu().numpy(), np.asarray(o_ref).reshape(-1))
    assert np.array_equal(o_gpu.cpu().numpy(), o0.cpu().numpy())
