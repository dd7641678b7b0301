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


class NavierStokesAdjoint:
    """Adjoint-based open-loop optimiser of examples/NavierStokes/NS2Doptimization.py:55-107 for a whole batch of
    `NavierStokes2D` envs: a backward costate sweep over a recorded trajectory (each sweep step projects the costate
    with the env's Jacobi pressure solve, `pdecg_ns_solve_pressure`, one launch for all envs), then the boundary
    gradient gives the new action sequence.  Element-wise arithmetic follows the example's operation order."""

    def __init__(self, env, viscosity: float = 0.1, gamma: float = 0.1):
        self.env = env
        self.nu, self.gamma = viscosity, gamma

    @staticmethod
    def _zero_walls(a1, a2):                             # apply_boundary of the example (:55-60)
        for a in (a1, a2):
            a[..., :, -1] = 0.
            a[..., :, 0] = 0.
            a[..., -1, :] = 0.
            a[..., 0, :] = 0.
        return a1, a2

    def costates(self, U, V, u_target, v_target):
        """U, V: [T, N, ny, nx] recorded velocities (after each of T steps); u_target, v_target: [NT, ny, nx] with
        NT >= T (indexed from the end like the example).  Returns Lam1, Lam2: [T, N, ny, nx] in forward time order."""
        from .envs_navier_stokes import central_difference as cd, laplace
        e = self.env
        dx, dy, dt = e.dx, e.dy, e.dt
        dev = e.device
        U, V = (torch.as_tensor(x, device=dev, dtype=torch.float64) for x in (U, V))
        ut, vt = (torch.as_tensor(x, device=dev, dtype=torch.float64) for x in (u_target, v_target))
        T = U.shape[0]
        lam1, lam2 = torch.zeros_like(U[0]), torch.zeros_like(U[0])
        pressure = torch.zeros_like(U[0])
        L1, L2 = [lam1], [lam2]
        for t in range(T - 1):                           # :85-101
            u, v = U[-1 - t], V[-1 - t]
            dl1dx, dl1dy = cd(lam1, "x", dx), cd(lam1, "y", dy)
            dl2dx, dl2dy = cd(lam2, "x", dx), cd(lam2, "y", dy)
            lap1, lap2 = laplace(lam1, dx, dy), laplace(lam2, dx, dy)
            dlam1dt = -2 * dl1dx * u - dl1dy * v - dl2dx * v - self.nu * lap1 + (u - ut[-1 - t])
            dlam2dt = -2 * dl2dy * v - dl1dy * u - dl2dx * u - self.nu * lap2 + (v - vt[-1 - t])
            lam1 = lam1 - dt * dlam1dt
            lam2 = lam2 - dt * dlam2dt
            lam1, lam2 = self._zero_walls(lam1, lam2)
            pressure = e.solve_pressure(lam1, lam2, pressure)
            lam1 = lam1 - dt * cd(pressure, "x", dx)
            lam2 = lam2 - dt * cd(pressure, "y", dy)
            lam1, lam2 = self._zero_walls(lam1, lam2)
            L1.append(lam1)
            L2.append(lam2)
        return torch.stack(L1[::-1]), torch.stack(L2[::-1])

    def actions(self, U, V, u_target, v_target, action_ref=2.0):
        """New action sequence [T, N] (:103-106): a_t = a_ref,t - nu/gamma * sum_x(d lam1 / dy at the row below the
        controlled wall) * 5 * dx."""
        from .envs_navier_stokes import central_difference as cd
        L1, _ = self.costates(U, V, u_target, v_target)
        T = L1.shape[0]
        aref = torch.as_tensor(action_ref, device=L1.device, dtype=torch.float64)
        aref = aref.expand(T) if aref.dim() == 0 else aref[:T]
        g = cd(L1, "y", self.env.dy)[..., -2, :].sum(dim=-1)          # [T, N]
        return aref[:, None] - self.nu / self.gamma * g * 5 * self.env.dx
