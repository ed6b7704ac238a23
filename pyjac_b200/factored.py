"""Host-side view of the factored Jacobian record (include/pyjac_b200.h).

Per state ``NF = NSP + 3 (NSP - 1) + NNZ`` doubles::

    fac[0 : NSP]                     J[0, j]            energy-equation row
    fac[NSP + k]                     J[k+1, 0]          temperature column
    fac[NSP + (NSP-1) + k]           WA_k
    fac[NSP + 2 (NSP-1) + k]         WB_k
    fac[NSP + 3 (NSP-1) + p]         S_p at (rows[p], cols[p])

    J[i, j] = ca[j] WA_{i-1} + cb[j] WB_{i-1} + S(i, j)       i, j >= 1

The reference only has the dense form (its ``sparse_multiplier`` emitter, create_jacobian.py:3301-3404, is
broken at :3322); these helpers describe the record and convert it to the dense layout for callers that want the reference's format
(a format conversion of results the GPU computed -- there is no CPU evaluation of anything in this package).
"""
import numpy as np


def pattern_from_tables(T):
    """(rows, cols, ca, cb) from a table dict (pyjac_b200.tables.build) -- the same values the library
    returns through pyjac_factored_pattern."""
    nsp = int(T['dims'][0])
    nnz = int(T['p5_cfg'][15])
    cf = np.asarray(T['p5_colfac']).reshape(nsp, 2)
    ca, cb = cf[:, 0].copy(), cf[:, 1].copy()
    ca[0] = cb[0] = 0.0
    return np.asarray(T['fac_rows'])[:nnz], np.asarray(T['fac_cols'])[:nnz], ca, cb


def expand(fac, nsp, rows, cols, ca, cb):
    """Records (n, NF) -> dense Jacobians (n, NSP*NSP), column-major per state (jac[i + NSP*j])."""
    fac = np.asarray(fac)
    n, last = fac.shape[0], nsp - 1
    assert fac.shape[1] == nsp + 3 * last + len(rows)
    J = np.zeros((n, nsp, nsp))                     # J[s, col, row]
    J[:, :, 0] = fac[:, :nsp]
    J[:, 0, 1:] = fac[:, nsp:nsp + last]
    wa, wb = fac[:, nsp + last:nsp + 2 * last], fac[:, nsp + 2 * last:nsp + 3 * last]
    J[:, 1:, 1:] = ca[None, 1:, None] * wa[:, None, :] + cb[None, 1:, None] * wb[:, None, :]
    J[:, cols, rows] += fac[:, nsp + 3 * last:]
    return J.reshape(n, nsp * nsp)
