"""Static work schedule ("plan") of the sm_100a Jacobian kernel (csrc/jacobian.cuh).

The kernel evaluates GS states per thread block at a time.  A warp's 32 lanes are split into
NSUB = 64 / GS *sub-groups* of GS / 2 lanes; every lane carries two neighbouring states (one
16-byte shared-memory access serves both), and the NSUB sub-groups of a warp work on NSUB
different table items (reactions, species, Jacobian rows) at the same time.  All mechanism
indexing is therefore decoded once per item for GS states, and what a warp does is fully
determined by the tables below, which are built here, once per mechanism, for a given
(GS, warps per block):

  phase B   reactions, in *rounds* of NSUB reactions of one kind, rounds dealt to warps by
            estimated cost (longest processing time first)
  phase C   per species: sum over its reactions of nu * (net rate, T-column term, X1, X2);
            the NSUB sub-groups of a warp split one species' list
  phase DE  per Jacobian element: all NSP*(NSP-1) species-row elements sorted by the length of
            their sparse contribution list and cut into *steps* of NSUB elements of (nearly)
            the same length, padded with null contributions; steps dealt to warps by cost.
            The energy-equation row (one element per column, enthalpy-weighted lists) forms a
            second class of steps run after the first.

This replaces the per-reaction / per-species statement unrolling of the reference's
generators (pyjac/core/create_jacobian.py:2650-2976, 3095-3254; rate_subs.py:1425-1527).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

SMEM_LIMIT = 232448          # bytes of dynamic shared memory one block may opt in to (sm_100)
NSCAL = 16                   # per-state scalar rows kept in shared memory
NPART = 7                    # per-warp partial sums
GS_CHOICES = (32, 16, 8, 4, 2)
SP_SLOTS, RX_SLOTS = 8, 5    # C B dB hW WA(Y) WB WT cp  /  net tT X1 X2 dH
SLOT_WA, SLOT_WT, SLOT_CP = 4, 6, 7
NULL_E = 0x3FFFFF            # element index of a padding element
MAX_L2 = 1023

# cost estimates (warp instructions) used only for load balancing
COST_PLAIN, COST_IRREV, COST_THREE = 215.0, 150.0, 40.0
COST_PM = {'thd': 300.0, 'lind': 450.0, 'troe': 900.0, 'sri': 1200.0}
COST_EFF = 8.0
COST_C_ITEM, COST_C_IT = 90.0, 18.0
COST_J_STEP, COST_J_SPARSE, COST_J_IT = 22.0, 8.0, 16.0
COST_T_STEP, COST_T_IT = 40.0, 18.0
COST_DOTS = 150.0


def layout(nsp: int, nr: int, nraw: int, gs: int, nw: int) -> Dict[str, int]:
    """Shared-memory carve-up in doubles.  A *row* is GS doubles (one value per state);
    species and reactions own SP_SLOTS / RX_SLOTS consecutive rows each."""
    off = 0
    L: Dict[str, int] = {}

    def take(name, rows):
        nonlocal off
        L[name] = off
        off += rows * gs

    take('SP', (nsp + 1) * SP_SLOTS)     # species nsp: the empty reaction slot (C = 1, others 0)
    take('RX', (nr + 1) * RX_SLOTS)      # reaction nr: zeros (padding of the phase C lists)
    take('RAW', nraw + 1)                # row nraw: zero (null contributions)
    take('SC', NSCAL)
    take('PA', nw * NPART)
    L['total'] = off
    return L


def choose_gs(nsp: int, nr: int, nraw: int, nw: int) -> int:
    for gs in GS_CHOICES:
        if layout(nsp, nr, nraw, gs, nw)['total'] * 8 <= SMEM_LIMIT:
            return gs
    return 0


def _lpt(costs: Sequence[float], nw: int, init: Sequence[float] = None) -> List[List[int]]:
    """Longest-processing-time-first assignment of items to nw bins; bins keep LPT order."""
    load = list(init) if init is not None else [0.0] * nw
    bins: List[List[int]] = [[] for _ in range(nw)]
    for ix in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        w = min(range(nw), key=lambda b: (load[b], b))
        bins[w].append(ix)
        load[w] += costs[ix]
    return bins, load


def hi16(c: float) -> int:
    bits = int(np.float64(c).view(np.uint64))
    if bits & ((1 << 48) - 1):
        raise ValueError('coefficient %r not representable in 16 bits' % c)
    return bits >> 48


def _f64_words(x: float) -> List[int]:
    bits = int(np.float64(x).view(np.uint64))
    return [bits & 0xFFFFFFFF, bits >> 32]


def build_plan(nsp: int, nr: int, nraw: int, first_pm: int, kinds: List[str], is_rev: List[bool],
               has3: List[bool], n_eff: List[int], red: List[List[Tuple[int, float]]],
               contrib: Dict[Tuple[int, int], List[Tuple[int, float]]],
               tcontrib: Dict[int, List[Tuple[int, int]]], sp_w: Sequence[float],
               sp_iw: Sequence[float], sp_mwf: Sequence[float], gs: int, nt: int
               ) -> Dict[str, np.ndarray]:
    """kinds[p] in {'plain','thd','lind','troe','sri'} per kernel-order reaction p;
    contrib[(k, j)] = [(raw row, nu)], tcontrib[j] = [(raw row, reaction)]."""
    assert gs in GS_CHOICES and nt % 32 == 0 and 64 <= nt <= 512
    if nsp > 2000:
        raise ValueError('too many species for the 22-bit element index')
    nw = nt // 32
    nsub = 64 // gs
    last = nsp - 1
    i32 = lambda x: np.asarray(x, dtype=np.int32)
    u32 = lambda x: np.asarray(x, dtype=np.uint32).view(np.int32)
    P: Dict[str, np.ndarray] = {}

    # ---------------------------------------------------------------- phase B
    pm = list(range(first_pm, nr))
    plain = sorted(range(first_pm), key=lambda p: (has3[p], not is_rev[p], p))
    rounds: List[Tuple[bool, List[int], float]] = []
    for c0 in range(0, len(pm), nsub):
        grp = pm[c0:c0 + nsub]
        cost = max(COST_PM[kinds[p]] for p in grp) + COST_EFF * max(n_eff[p] for p in grp)
        if any(has3[p] for p in grp):
            cost += COST_THREE
        rounds.append((True, grp, cost))
    for c0 in range(0, len(plain), nsub):
        grp = plain[c0:c0 + nsub]
        cost = COST_PLAIN if any(is_rev[p] for p in grp) else COST_IRREV
        if any(has3[p] for p in grp):
            cost += COST_THREE
        rounds.append((False, grp, cost))
    bins, _ = _lpt([r[2] for r in rounds], nw)
    b_off, b_npm, b_item = [0], [], []
    for w in range(nw):
        mine = sorted(bins[w], key=lambda ix: (not rounds[ix][0], ix))     # pm rounds first
        b_npm.append(sum(1 for ix in mine if rounds[ix][0]))
        for ix in mine:
            grp = rounds[ix][1]
            b_item += grp + [-1] * (nsub - len(grp))
        b_off.append(len(b_item) // nsub)
    P['p5_b_off'] = i32(b_off)
    P['p5_b_npm'] = i32(b_npm)
    P['p5_b_item'] = i32(b_item or [-1] * nsub)

    # ---------------------------------------------------------------- phase C
    c_cost = [COST_C_ITEM + COST_C_IT * ((len(red[k]) + nsub - 1) // nsub) for k in range(nsp)]
    bins, _ = _lpt(c_cost, nw)
    c_off, c_item, c_con = [0], [], []
    for w in range(nw):
        for k in bins[w]:
            words = [p | (hi16(nu) << 16) for p, nu in red[k]]
            nit = (len(words) + nsub - 1) // nsub
            words += [nr] * (nit * nsub - len(words))          # reaction row nr is all zero
            c_item += [k, len(c_con) // nsub, nit]
            c_con += words
        c_off.append(len(c_item) // 3)
    P['p5_c_off'] = i32(c_off)
    P['p5_c_item'] = i32(c_item)
    P['p5_c_con'] = u32(c_con + [nr] * nsub)

    # ---------------------------------------------------------------- phase DE, species rows
    # element (col, k): output row k + 1 of column col; col 0 is the temperature column
    elems = []
    for col in range(nsp):
        for k in range(last):
            lst = contrib.get((k, col - 1), []) if col else []
            elems.append((len(lst), col, k, lst))
    elems.sort(key=lambda e: (-e[0], e[1], e[2]))
    steps = []
    for c0 in range(0, len(elems), nsub):
        grp = elems[c0:c0 + nsub]
        L2 = (grp[0][0] + 1) // 2
        if L2 > MAX_L2:
            raise ValueError('sparse Jacobian element with too many contributions')
        units: List[List[int]] = []            # units[u][sub] = [x, y]
        row, pw = [], []
        for s in range(nsub):
            if s < len(grp):
                _, col, k, lst = grp[s]
                slot = SLOT_WT if col == 0 else SLOT_WA
                row.append([(col * nsp + k + 1) | (L2 << 22), (k * SP_SLOTS + slot) | (col << 16)])
                pw.append(_f64_words(sp_iw[col - 1] * sp_w[k]) if col else [0, 0])
            else:
                row.append([NULL_E | (L2 << 22), SLOT_WA])
                pw.append([0, 0])
        units.append(row)
        if L2:
            units.append(pw)
            for i2 in range(L2):
                u = []
                for s in range(nsub):
                    lst = grp[s][3] if s < len(grp) else []
                    w = [src | (hi16(c) << 16) for src, c in lst[2 * i2:2 * i2 + 2]]
                    u.append(w + [nraw] * (2 - len(w)))        # null: zero raw row, coefficient +0
                units.append(u)
        steps.append((COST_J_STEP + (COST_J_SPARSE + COST_J_IT * L2 if L2 else 0.0), units))
    init = [0.0] * nw
    init[0] = COST_DOTS
    bins, load = _lpt([st[0] for st in steps], nw, init)
    e_off, e_nst, e_str = [0], [], []
    for w in range(nw):
        e_nst.append(len(bins[w]))
        for ix in bins[w]:
            for u in steps[ix][1]:
                for xy in u:
                    e_str += xy
        e_off.append(len(e_str) // (2 * nsub))
    P['p5_e_off'] = i32(e_off)
    P['p5_e_nst'] = i32(e_nst)
    P['p5_e_str'] = u32(e_str + [0] * (2 * nsub))

    # ---------------------------------------------------------------- phase DE, energy row
    telems = sorted(((len(tcontrib.get(j, [])), j) for j in range(last)), key=lambda e: (-e[0], e[1]))
    tsteps = []
    for c0 in range(0, len(telems), nsub):
        grp = telems[c0:c0 + nsub]
        L2 = (grp[0][0] + 1) // 2
        if L2 > MAX_L2:
            raise ValueError('energy-row element with too many contributions')
        units = [[[((j + 1) * nsp) | (L2 << 22), j + 1] for _, j in grp] +
                 [[NULL_E | (L2 << 22), 1]] * (nsub - len(grp))]
        for i2 in range(L2):
            u = []
            for s in range(nsub):
                lst = tcontrib.get(grp[s][1], []) if s < len(grp) else []
                w = [src | (p << 16) for src, p in lst[2 * i2:2 * i2 + 2]]
                u.append(w + [nraw | (nr << 16)] * (2 - len(w)))
            units.append(u)
        tsteps.append((COST_T_STEP + COST_T_IT * L2, units))
    bins, _ = _lpt([st[0] for st in tsteps], nw, load)
    t_off, t_nst, t_str = [0], [], []
    for w in range(nw):
        t_nst.append(len(bins[w]))
        for ix in bins[w]:
            for u in tsteps[ix][1]:
                for xy in u:
                    t_str += xy
        t_off.append(len(t_str) // (2 * nsub))
    P['p5_t_off'] = i32(t_off)
    P['p5_t_nst'] = i32(t_nst)
    P['p5_t_str'] = u32(t_str + [0] * (2 * nsub))

    # per column: (1 / W_j, (1 / W_j) (W_j / W_N)); the temperature column takes W_k * T-term as is
    colfac = [1.0, 0.0]
    for j in range(last):
        colfac += [sp_iw[j], sp_iw[j] * sp_mwf[j]]
    P['p5_colfac'] = np.asarray(colfac, dtype=np.float64)

    L = layout(nsp, nr, nraw, gs, nw)
    # threads meeting at the named barrier before the energy-row steps: warp 0 (arrives after the
    # dot products) and every other warp that owns such steps (waits)
    waiters = sum(1 for w in range(1, nw) if t_nst[w])
    t_sync = 32 * (waiters + 1) if waiters else 0
    P['p5_cfg'] = i32([gs, nt, nw, nsub, L['SP'], L['RX'], L['RAW'], L['SC'], L['PA'], L['total'], t_sync]
                      + [0] * 5)
    return P
