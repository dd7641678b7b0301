"""Model-based backstepping controllers used by the reference's examples to produce its published
known-answer values K1 / K2 (examples/transportPDE/transport1Dbackstepping.py:15-36,88-99 and
examples/reactionDiffusionPDE/reactionDiffusion1DBackstepping.py:15-39,90-100).  Restated here as test
drivers (host-side numpy; one scalar action per control step)."""
import math

import numpy as np


def chebyshev_beta(x, gamma, amplitude):
    """beta(x) = amplitude*cos(gamma*acos(x)) stored as float32, like solveBetaFunction of the examples."""
    out = np.zeros(len(x), dtype=np.float32)
    for i, v in enumerate(x):
        out[i] = amplitude * math.cos(gamma * math.acos(v))
    return out


def transport_kernel(theta, dx):
    """Successive-approximation solution of the Volterra kernel equation, flipped (transport1Dbackstepping.py:22-29)."""
    kappa = np.zeros(len(theta))
    for i in range(len(theta)):
        acc = 0
        for j in range(i):
            acc += (kappa[i - j] * theta[j]) * dx
        kappa[i] = acc - theta[i]
    return np.flip(kappa)


def transport_control(kernel, u):
    res = 0
    for i in range(len(u)):
        res += kernel[i] * u[i]
    return res * 1e-2


def rd_kernel(beta, dx):
    """Finite-difference solution of the kernel PDE on the triangle (reactionDiffusion1DBackstepping.py:22-35)."""
    n = len(beta)
    k = np.zeros((n, n))
    a = beta
    k[1][1] = -(a[1] + a[0]) * dx / 4
    for i in range(1, n - 1):
        k[i + 1][0] = 0
        k[i + 1][i + 1] = k[i][i] - dx / 4.0 * (a[i - 1] + a[i])
        k[i + 1][i] = k[i][i] - dx / 2 * a[i]
        for j in range(1, i):
            k[i + 1][j] = -k[i - 1][j] + k[i][j + 1] + k[i][j - 1] + a[j] * (dx ** 2) * (k[i][j + 1] + k[i][j - 1]) / 2
    return k


def rd_control(kernel, u, dx):
    return sum(kernel[-1][0:len(u) - 1] * u[0:len(u) - 1]) * dx


TRANSPORT_KW = dict(T=5, dt=1e-4, X=1, dx=1e-2, normalize=False, sensing_loc="full", control_type="Dirchilet",
                    sensing_type=None, limit_pde_state_size=True, max_state_value=1e10, max_control_value=20,
                    control_sample_rate=0.1)
RD_KW = dict(T=1, dt=1e-5, X=1, dx=5e-3, normalize=False, sensing_loc="full", control_type="Dirchilet",
             sensing_type=None, limit_pde_state_size=True, max_state_value=1e10, max_control_value=20,
             control_sample_rate=0.001)

# published: HyperbolicPDEExample.ipynb:820,823 ; ParabolicPDEExample.ipynb:33152,33155   (reward, sum of ||obs||_2)
K1 = {1: (289.8379892610982, 106.08644131330854), 10: (198.37980383622647, 1060.8641808569605)}
K2 = {1: (299.82330386477986, 1275.4394707666174), 10: (298.23300542427603, 12754.398135204445)}


def run_episode(reset_fn, step_fn, control_fn):
    """reset_fn() -> obs row (numpy 1-D); step_fn(a) -> (obs row, reward, terminated, truncated)."""
    obs = reset_fn()
    total, norm_sum, n = 0.0, 0.0, 0
    term = trunc = False
    while not (term or trunc):
        a = control_fn(obs)
        obs, r, term, trunc = step_fn(a)
        total += float(r)
        norm_sum += float(np.linalg.norm(obs))
        n += 1
    return total, norm_sum, n
