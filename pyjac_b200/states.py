"""State batches: the bundled PaSR trajectories and the synthetic benchmark states.

A *state row* is what the reference's scalar API takes (docs/faqs.rst:82-87):
``y = [T, Y_0 .. Y_{NSP-2}]`` in pyJac's internal (moved-last) species order, plus a
pressure.  Batches here are row-major ``y[n, NSP]``, ``P[n]``.
"""
from __future__ import annotations

import numpy as np


def pasr_states(npy_path: str, mech):
    """Rows ``[t, T, P, Y_0..Y_{NSP-1}]`` (original species order) from a PaSR ``.npy``
    (partially_stirred_reactor.py:715-742), renormalised exactly as the reference's
    functional tester does before evaluating (functional_tester/test.py:1254-1258)."""
    d = np.load(npy_path)
    d = d.reshape(-1, d.shape[-1]).copy()
    if d.shape[1] != mech.NSP + 3:
        raise ValueError('PaSR rows have %d columns, mechanism needs %d' % (d.shape[1], mech.NSP + 3))
    ls = 3 + mech.last_spec_original
    for i in range(len(d)):
        d[i, 3:] /= np.sum(d[i, 3:])
        d[i, ls] = 1. - np.sum(d[i, 3:ls]) - np.sum(d[i, ls + 1:])
    Y = d[:, 3:][:, mech.fwd_spec_map]
    y = np.concatenate([d[:, 1:2], Y[:, :-1]], axis=1)
    return d[:, 2].copy(), np.ascontiguousarray(y)


def synthetic_states(nsp: int, n: int, seed: int = 0):
    """T ~ U[800, 2400] K, P ~ logU[0.5, 25] atm, Y ~ Dirichlet(0.5) over
    all species with the last one recomputed as 1 - sum(others)."""
    rng = np.random.default_rng(seed)
    T = rng.uniform(800.0, 2400.0, n)
    P = np.exp(rng.uniform(np.log(0.5), np.log(25.0), n)) * 101325.0
    Y = rng.dirichlet(np.full(nsp, 0.5), n)
    Y[:, -1] = 1.0 - Y[:, :-1].sum(axis=1)
    y = np.concatenate([T[:, None], Y[:, :-1]], axis=1)
    return P, np.ascontiguousarray(y)
