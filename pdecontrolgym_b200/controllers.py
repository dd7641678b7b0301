"""Batched model-based (backstepping) boundary controllers -- the baselines the reference's examples use to
validate each env (examples/transportPDE/transport1Dbackstepping.py:22-36,
examples/reactionDiffusionPDE/reactionDiffusion1DBackstepping.py:22-39,
docs/source/tutorials/Trafficarz1d_tutorial.rst:153-184).

The gain kernels are computed once on the host (numpy, float64, the examples' recurrences); applying a
controller is one [num_envs, npts] x [npts] product on the device, so closed-loop episodes of the whole batch
run without host round trips.  These are drivers around the hot path ("next" row of SURVEY.md 8(f)), written in
plain PyTorch on purpose: the product is a library GEMV, not a kernel of this package."""
from __future__ import annotations

import numpy as np
import torch


class TransportBackstepping:
    """u(1,t) = sum_i k_i u_i * 1e-2 with the Volterra kernel of transport1Dbackstepping.py:22-36."""

    def __init__(self, beta_on_kernel_grid, dx, device="cuda"):
        theta = np.asarray(beta_on_kernel_grid)
        kappa = np.zeros(len(theta))
        for i in range(len(theta)):
            acc = 0
            for j in range(i):
                acc += (kappa[i - j] * theta[j]) * dx
            kappa[i] = acc - theta[i]
        self.kernel = torch.from_numpy(np.flip(kappa).copy()).to(device)

    def __call__(self, obs):
        return (obs.to(torch.float64) @ self.kernel) * 1e-2


class ReactionDiffusionBackstepping:
    """u(1,t) = dx * sum_{i<n-1} k[-1, i] u_i with the kernel PDE solve of reactionDiffusion1DBackstepping.py:22-39."""

    def __init__(self, beta_on_kernel_grid, dx, device="cuda"):
        a = np.asarray(beta_on_kernel_grid)
        n = len(a)
        k = np.zeros((n, n))
        k[1][1] = -(a[1] + a[0]) * dx / 4
        for i in range(1, n - 1):
            k[i + 1][0] = 0
            k[i + 1][i + 1] = k[i][i] - dx / 4.0 * (a[i - 1] + a[i])
            k[i + 1][i] = k[i][i] - dx / 2 * a[i]
            for j in range(1, i):
                k[i + 1][j] = -k[i - 1][j] + k[i][j + 1] + k[i][j - 1] + a[j] * (dx ** 2) * (k[i][j + 1] + k[i][j - 1]) / 2
        self.dx = dx
        self.gain = torch.from_numpy(k[-1].copy()).to(device)

    def __call__(self, obs):
        m = obs.shape[1] - 1
        g = self.gain[:m] if self.gain.numel() >= m else torch.nn.functional.pad(self.gain, (0, m - self.gain.numel()))
        return (obs[:, :m].to(torch.float64) @ g) * self.dx


class TrafficARZBackstepping:
    """Outlet (or both-boundary) flow-rate law of Trafficarz1d_tutorial.rst:153-184 for a batched `TrafficPDE1D`."""

    def __init__(self, env):
        self.env = env
        x = torch.from_numpy(np.arange(0, env.L + env.dx, env.dx)).to(env.device)
        self.x = x

    def __call__(self, obs=None):
        e = self.env
        if e.simulation_type == "inlet":
            return e.qs.clone().unsqueeze(1)
        vs, rs, qs, ps = e.vs.unsqueeze(1), e.rs.unsqueeze(1), e.qs.unsqueeze(1), e.ps.unsqueeze(1)
        x = self.x.unsqueeze(0)
        lam1 = vs
        lam2 = vs + rs * (-e.vm / e.rm)
        Kk = -(1 / (1 * ps)) * (-1 / e.tau) * torch.exp(-x / (e.tau * vs))
        Mk = -Kk
        cv = Mk + (lam2 / lam1) * Kk * torch.exp(x / (vs * e.tau))
        cq = ((lam1 - lam2) / lam1) * Kk * torch.exp(x / (vs * e.tau))
        v, r = e.v, e.r
        q_out = qs.squeeze(1) + rs.squeeze(1) * torch.trapz(cv * (v - vs), dx=e.dx, dim=1) + torch.trapz(cq * (r * v - qs), dx=e.dx, dim=1)
        if e.simulation_type == "both":
            return torch.stack((e.qs, q_out), dim=1)
        return q_out.unsqueeze(1)
