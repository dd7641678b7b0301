"""CUDA parity: NeuronPDE1D kernel vs the numpy oracle (itself pinned to the shimmed reference class in
tests/test_oracle_vs_reference.py::test_neuron).  The control law's trapezoid integral is a warp reduction tree on the GPU
and a sequential Python sum in the reference, so states are compared to 1e-10 relative (north-star bar), the moving
boundary index exactly."""
import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
P = dict(T=180, dt=5e-6, X=12e-6, dx=0.01)


@pytest.mark.parametrize("n_steps", [1, 7])
def test_vs_oracle(cuda_device, n_steps):
    import pdecontrolgym_b200 as pcg
    B = 3
    env = pcg.make("PDEControlGym-NeuronPDE1D", num_envs=B, device=cuda_device, **P)
    orc = O.NeuronOracle(B, **P)
    o_g, info = env.reset()
    assert np.array_equal(o_g.cpu().numpy(), orc.u) and np.array_equal(info["L"].cpu().numpy(), orc.L)
    for k in range(600 // n_steps):
        for _ in range(n_steps):
            u_r, L_r, te_r = orc.step()
        o_g, r, te, tr, info = env.step(n_steps=n_steps)
        np.testing.assert_allclose(o_g.cpu().numpy(), u_r, rtol=1e-10, atol=0, err_msg=f"call {k}")
        np.testing.assert_allclose(info["Z"].cpu().numpy(), orc.Z, rtol=1e-12)
        np.testing.assert_allclose(info["control"].cpu().numpy(), orc.control, rtol=1e-9)
        assert np.array_equal(info["L"].cpu().numpy(), L_r) and not te.any()


def test_long_run_growth_and_batch_of_initial_lengths(cuda_device):
    """400,000 steps (2 s of simulated time) in launches of 20,000: the axon grows (L advances), every env stays finite, envs
    with different initial lengths evolve differently, and a reward descriptor yields -||u[:L]||_2."""
    import torch
    import pdecontrolgym_b200 as pcg
    B = 4
    env = pcg.make("PDEControlGym-NeuronPDE1D", num_envs=B, device=cuda_device, reward_class=pcg.NormReward(1, norm="2"), **P)
    _, info0 = env.reset(options={"LSubZero": np.array([1e-6, 1e-6, 2e-6, 3e-6])})
    for _ in range(20):
        obs, r, te, tr, info = env.step(n_steps=20000)
    assert torch.isfinite(obs).all() and (info["L"] >= info0["L"]).all() and (info["L"] > info0["L"]).any()
    assert torch.equal(obs[0], obs[1]) and not torch.equal(obs[0], obs[2])
    L = info["L"].cpu().numpy()
    want = np.array([-np.linalg.norm(obs[b, :L[b]].cpu().numpy()) for b in range(B)])
    np.testing.assert_allclose(r.cpu().numpy(), want, rtol=1e-12)
    assert abs(float(env.time_index[0]) - 400000 * 5e-6) < 1e-9
