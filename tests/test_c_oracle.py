"""CPU: the plain-C restatement (oracle/c/pdecg_oracle.c) equals the numpy oracle bit for bit -- and through it
the live reference -- for every precision / beta dtype / action dtype / control type combination."""
import itertools

import numpy as np
import pytest

from oracle import np_oracle as O
from oracle.c_oracle import Bc1dCOracle
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
