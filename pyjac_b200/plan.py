"""Static work schedule ("plan") of the sm_100a kernel (csrc/eval.cuh).

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

PLAN_VERSION = 3             # of the p5_* schedule tables (checked by the library against its own)
SMEM_LIMIT = 232448          # bytes of dynamic shared memory one block may opt in to (sm_100)
NSCAL = 24                   # per-state scalar rows: 2 x 8 from phase A0, 5 derived in DE, 2 x P (PLOG)
NPART = 7                    # per-warp partial sums
GS_CHOICES = (32, 16, 8, 4, 2)
DEFAULT_THREADS = 384         # 12 warps with 168 registers each measured faster than 16 x 128
SP_SLOTS, RX_SLOTS = 8, 5    # C B dB hW WA(Y) WB WT cp  /  net tT X1 X2 dH
SLOT_WA, SLOT_WT, SLOT_CP = 4, 6, 7
NULL_E = 0x3FFFFF            # element index of a padding element
NONE32 = 0xFFFFFFFF
MAX_L2 = 1023

# cost estimates (warp instructions) used only for load balancing
COST_PLAIN, COST_IRREV, COST_THREE = 215.0, 150.0, 40.0
COST_PM = {'thd': 300.0, 'lind': 450.0, 'troe': 900.0, 'sri': 1200.0}
COST_EFF = 8.0
COST_PLOG = 120.0
COST_C_ITEM, COST_C_IT = 90.0, 28.0
# phase DE costs are in units of 22 cycles, fitted to per-warp clock measurements on a B200
# (tools/phase_clocks.py) and re-tuned by throughput for the 12-warp plan (tools/cost_sweep.sh:
# warp 0's share and the dense rows were under-estimated, +5.6 % once corrected): the classes are
# bound by shared-memory / L2 latency, not by issue
COST_S_STEP, COST_S_OVF = 40.0, 11.0
COST_D_ITEM, COST_D_COL = 63.0, 17.0
COST_T_ITEM, COST_T_IT = 130.0, 11.0
D_MAX_COLS = 24              # a row with more dense-only columns is cut into several items
COST_DOTS = 640.0            # warp 0: energy-equation scalars + the next group's phase A0


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
    take('RX', (nr + 2) * RX_SLOTS)      # reactions nr, nr + 1: zeros (padding of the gather lists)
    take('RAW', nraw + 2)                # rows nraw, nraw + 1: zeros (null contributions)
    take('SC', NSCAL)
    take('PA', nw * NPART)
    take('CF', -(-2 * nsp // gs))        # per column (1/W_j, (1/W_j)(W_j/W_N)) as plain doubles
    L['total'] = off
    return L


def choose_gs(nsp: int, nr: int, nraw: int, nw: int) -> int:
    """Largest number of states per block whose working set fits in shared memory; 0 if not even
    two states fit."""
    for gs in GS_CHOICES:
        if layout(nsp, nr, nraw, gs, nw)['total'] * 8 <= SMEM_LIMIT:
            return gs
    return 0


SMEM_MIN_GS = 8              # below this many states per block in shared memory the plan moves
                             # the working set to global memory instead (measured on a B200: USC-II-shaped
                             # mechanism 5.3e6 states/s with 2 states in shared memory, 1.08e7 with 8 in global)


def wsg_shape(nsp: int) -> Tuple[int, int]:
    """(states per block, threads per block) of a plan whose working set lives in global memory.
    Measured on a B200: n-heptane-sized mechanisms run best with 16 states (128-byte Jacobian
    stores) and two or more 128-register blocks per SM, USC-II-sized ones with 8 states and one
    384-thread block."""
    return (16, 256) if nsp >= 256 else (8, DEFAULT_THREADS)


def _lpt(costs: Sequence[float], nw: int, init: Sequence[float] = None) -> List[List[int]]:
    """Longest-processing-time-first assignment of items to nw bins; bins keep LPT order."""
    load = list(init) if init is not None else [0.0] * nw
    bins: List[List[int]] = [[] for _ in range(nw)]
    for ix in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        w = min(range(nw), key=lambda b: (load[b], b))
        bins[w].append(ix)
        load[w] += costs[ix]
    return bins, load


def balance_banks(lists: List[List[int]], half) -> None:
    """Reorder each of the lists (one per sub-group, the order of its entries is free) so that at
    every position about half of the lists hold an entry on either half of a 128-byte bank line
    (``half(entry)`` in {0, 1}): the shared-memory reads of one instruction then need 4 rather than
    up to 8 wavefronts.  Greedy: list s prefers half (s + position) & 1."""
    pools = [[[e for e in lst if half(e) == h] for h in (0, 1)] for lst in lists]
    out: List[List[int]] = [[] for _ in lists]
    for pos in range(max((len(lst) for lst in lists), default=0)):
        for s_, pool in enumerate(pools):
            if len(out[s_]) >= len(lists[s_]):
                continue
            want = (s_ + pos) & 1
            pick = want if pool[want] else 1 - want
            out[s_].append(pool[pick].pop())
    for lst, new in zip(lists, out):
        lst[:] = new


def balance_pairs(lists: List[List[int]], half, wild=None) -> None:
    """Reorder the lists (one per sub-group, all of one length, the order of their entries is free) for
    rows of 64 bytes: a 16-byte shared-memory access is served a quarter warp at a time, i.e. sub-groups
    2 i and 2 i + 1 together, and takes one wavefront if their two rows lie on different halves of a
    128-byte bank line, two otherwise (measured on a B200 with ncu's per-instruction wavefront counts:
    6.0 wavefronts per ld.shared.v2.f64 for rows of random parity, 4.0 when every pair is split).  So
    the entries of lists 2 i and 2 i + 1 are matched position by position to opposite halves, as many
    as there are; ``wild(v)`` -> (the entry on half 0, on half 1) for padding entries that may sit on
    either half."""
    for i in range(0, len(lists) - 1, 2):
        A, B = lists[i], lists[i + 1]
        assert len(A) == len(B)
        pools = []
        for lst in (A, B):
            h = [[], []]
            w = []
            for v in lst:
                if wild is not None and wild(v) is not None:
                    w.append(wild(v))
                else:
                    h[half(v)].append(v)
            pools.append((h, w))
        (ha, wa_), (hb, wb_) = pools
        outa, outb = [], []
        for x, y in ((0, 1), (1, 0)):                       # real entries on opposite halves
            while ha[x] and hb[y]:
                outa.append(ha[x].pop()); outb.append(hb[y].pop())
        for x in (0, 1):                                    # a real entry and a padding entry opposite to it
            while ha[x] and wb_:
                outa.append(ha[x].pop()); outb.append(wb_.pop()[1 - x])
            while hb[x] and wa_:
                outb.append(hb[x].pop()); outa.append(wa_.pop()[1 - x])
        while wa_ and wb_:
            outa.append(wa_.pop()[0]); outb.append(wb_.pop()[1])
        rest_a = ha[0] + ha[1] + [w_[0] for w_ in wa_]       # what is left meets on one half
        rest_b = hb[0] + hb[1] + [w_[1] for w_ in wb_]
        assert len(rest_a) == len(rest_b)
        A[:] = outa + rest_a
        B[:] = outb + rest_b


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
               sp_iw: Sequence[float], sp_mwf: Sequence[float], gs: int, nt: int, wsg: int = 0
               ) -> Dict[str, np.ndarray]:
    """kinds[p] in {'plain','plog','thd','lind','troe','sri'} per kernel-order reaction p;
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
    plain = sorted(range(first_pm), key=lambda p: (has3[p], kinds[p] == 'plog', not is_rev[p], p))
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
        if any(kinds[p] == 'plog' for p in grp):
            cost += COST_PLOG
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
    P['p5_b_item'] = i32(b_item + [-1] * (2 * nsub))         # ends with two null rounds (look-ahead)

    # byte offsets inside the shared-memory regions (a row is GS doubles)
    RB = gs * 8
    SPB, RXB = SP_SLOTS * RB, RX_SLOTS * RB
    PF = 16                                    # null units after each stream (prefetch runs ahead)

    def expand(lst):
        """[(source, nu)] -> (sources with weight +1, sources with weight -1), nu an integer."""
        plus, minus = [], []
        for src, c in lst:
            if not float(c).is_integer():
                raise ValueError('non-integer coefficient %r' % c)
            (plus if c > 0 else minus).extend([src] * int(abs(c)))
        return plus, minus

    # ---------------------------------------------------------------- phase C
    # Per species k: sum over its reactions of nu_ki * (net, tT, X1, X2), nu split into +1 and -1
    # entries (byte offsets of reaction rows).  D sub-groups share one species; a warp round
    # handles NSUB / D species of similar list length.  Round r of a warp: header unit
    # {species row offset or NONE32, 1 if this sub-group stores the result}, then nP units of
    # two +1 entries and nM units of two -1 entries per sub-group, all padded with the zero row.
    c_lists = [expand(red[k]) for k in range(nsp)]

    def c_schedule(coop):
        """Rounds of NSUB / coop species of similar list length, dealt to the warps longest first;
        returns (rounds per warp, largest per-warp cost)."""
        def n_units(n_entries):
            return (-(-n_entries // coop) + 1) // 2
        per_round = nsub // coop
        order_c = sorted(range(nsp), key=lambda k: (-(n_units(len(c_lists[k][0])) + n_units(len(c_lists[k][1]))), k))
        chunks = [order_c[c0:c0 + per_round] for c0 in range(0, nsp, per_round)]
        cost = [COST_C_ITEM + COST_C_IT * (max(n_units(len(c_lists[k][0])) for k in ch) +
                                           max(n_units(len(c_lists[k][1])) for k in ch)) for ch in chunks]
        bins, load = _lpt(cost, nw)
        return [[chunks[ix] for ix in sorted(b)] for b in bins], max(load), n_units

    # D sub-groups share one species: the largest power of two that still lets every species be
    # summed in one round of the block (splitting further to even out the warps measured slower:
    # the phase is bound by shared-memory throughput, not by its longest warp)
    coop = nsub
    while coop > 1 and nw * (nsub // coop) < nsp:
        coop //= 2
    per_warp, _, n_units = c_schedule(coop)
    c_off, c_item, c_str = [0], [], []
    for w in range(nw):
        for ch in per_warp[w]:
            n_p = max(n_units(len(c_lists[k][0])) for k in ch)
            n_m = max(n_units(len(c_lists[k][1])) for k in ch)
            c_item += [len(c_str) // (2 * nsub), n_p, n_m, 0]
            # header unit, then the units; sub-group sb works for species ch[sb // coop], part sb % coop
            subs_k = [ch[sb // coop] if sb // coop < len(ch) else None for sb in range(nsub)]
            for sb in range(nsub):
                k = subs_k[sb]
                c_str += [NONE32 if k is None else k * SPB + (k & 1) * RB, 1 if (k is not None and sb % coop == 0) else 0]
            for which, n in ((0, n_p), (1, n_m)):
                per_sub = []
                for sb in range(nsub):
                    k = subs_k[sb]
                    lst = [] if k is None else c_lists[k][which][sb % coop::coop]
                    per_sub.append([q_ * RXB for q_ in lst] + [(nr + ((sb + i) & 1)) * RXB for i in range(2 * n - len(lst))])
                if RB == 64:
                    balance_pairs(per_sub, lambda v: (v // RB) & 1,
                                  lambda v: ((nr + (nr & 1)) * RXB, (nr + 1 - (nr & 1)) * RXB) if v // RXB >= nr else None)
                else:
                    balance_banks(per_sub, lambda v: (v // RB) & 1 if (RB & 127) else 0)
                for u in range(n):
                    for sb in range(nsub):
                        c_str += per_sub[sb][2 * u:2 * u + 2]
        c_off.append(len(c_item) // 4)
    P['p5_c_off'] = i32(c_off)
    P['p5_c_item'] = i32(c_item + [len(c_str) // (2 * nsub), 0, 0, 0])      # + null round (look-ahead)
    P['p5_c_str'] = u32(c_str + [nr * RXB] * (2 * nsub * PF))

    # ---------------------------------------------------------------- phase DE
    # element (col, k): output row k + 1 of column col; col 0 is the temperature column.
    # Three classes of work, each with its own stream of uint2 / uint4 units per sub-group:
    #   S  elements with a sparse part (NSUB per step, sorted by padded list length L >= 1): two
    #      uint4 per element, {e | L << 22, species-row offset of WA | col << 20, the double
    #      (1/W_j) W_k} and the first two units {+1 raw row, -1 raw row, +1, -1} (byte offsets;
    #      padding = zero row); units 3..L, padded to a multiple of four, go to an overflow
    #      stream of uint2; a warp takes two steps at a time
    #   D  dense-only elements by row: a sub-group keeps W_k a_k, W_k b_k of one species row in
    #      registers and walks that row's dense-only columns: header {species-row offset or NONE,
    #      element index of the row's temperature-column entry}, then n units {e, col}
    #   T  the energy-equation row: pl.tcoop sub-groups share one column's enthalpy-weighted
    #      list: header {col | store << 16, offset of cp_j}, then n units {raw row, reaction row}
    def sp_even(k):
        """Byte offset of species k's even-slot base; odd slots sit at (base ^ RB): the slot pair
        is swapped for odd k so that rows of different species fall on different banks."""
        return k * SPB + (k & 1) * RB
    zr = nraw * RB
    null_sp = sp_even(0) + SLOT_WA * RB

    # signed entry = byte offset of a raw row | 1 if its weight is -1 (offsets are multiples of 16)
    elems = []
    for col in range(1, nsp):
        for k in range(last):
            plus, minus = expand(contrib.get((k, col - 1), []))
            ent = [x * RB for x in plus] + [(x * RB) | 1 for x in minus]
            ent.sort(key=lambda v: v & ~1)
            elems.append(((len(ent) + 1) // 2, col, k, ent))
    sparse = sorted((e for e in elems if e[0] > 0), key=lambda e: (-e[0], e[1], e[2]))
    if RB == 64:
        # Neighbouring sub-groups (2 i, 2 i + 1) share a shared-memory phase (balance_pairs): among the
        # elements of one list length, pair those whose raw rows lean to one half of the bank lines with
        # those leaning to the other
        # (the same holds for the species rows W_k a_k / W_k b_k of the element's row k, whose half is k & 1)
        lean = lambda e: sum(1 if (v // RB) & 1 else -1 for v in e[3])
        paired = []
        for L in sorted({e[0] for e in sparse}, reverse=True):
            pools = [sorted((e for e in sparse if e[0] == L and (e[2] & 1) == par), key=lambda e: (lean(e), e[1], e[2]))
                     for par in (0, 1)]
            while pools[0] and pools[1]:                  # rows of opposite parity, raw rows leaning opposite ways
                paired += [pools[0].pop(0), pools[1].pop()]
            run = pools[0] + pools[1]
            run.sort(key=lambda e: (lean(e), e[1], e[2]))
            while len(run) > 1:
                paired += [run.pop(0), run.pop()]
            paired += run
        sparse = paired

    s_steps = []                               # (A words, B words, overflow units, L)
    for c0 in range(0, len(sparse), nsub):
        grp = sparse[c0:c0 + nsub]
        L = grp[0][0]                          # units of two entries
        if L > MAX_L2:
            raise ValueError('sparse Jacobian element with too many contributions')
        n_ovf = (max(L - 2, 0) + 3) // 4 * 4              # overflow units come in batches of four
        A, B, ovf = [], [], [[] for _ in range(n_ovf)]
        # pad with the two all-zero raw rows (nraw, nraw + 1: one on either half of a bank line),
        # then order every list so that each position is spread over both halves
        ents = [list(grp[sb][3]) if sb < len(grp) else [] for sb in range(nsub)]
        for sb in range(nsub):
            ents[sb] += [(nraw + ((sb + i) & 1)) * RB for i in range(2 * L - len(ents[sb]))]
        if RB == 64:
            zw = lambda v: ((nraw + (nraw & 1)) * RB | (v & 1), (nraw + 1 - (nraw & 1)) * RB | (v & 1)) if (v & ~1) // RB >= nraw else None
            balance_pairs(ents, lambda v: (v // RB) & 1, zw)
        else:
            balance_banks(ents, lambda v: (v // RB) & 1 if (RB & 127) else 0)
        for sb in range(nsub):                 # positions past 2 L: skipped (L = 1) or padding of a batch
            ents[sb] += [(nraw + ((sb + i) & 1)) * RB for i in range(2 * (n_ovf + 2) - 2 * L)]
        for sb in range(nsub):
            if sb < len(grp):
                _, col, k, _ = grp[sb]
                A.append([(col * nsp + k + 1) | (L << 22), (sp_even(k) + SLOT_WA * RB) | (col << 20)]
                         + _f64_words(sp_iw[col - 1] * sp_w[k]))
            else:
                A.append([NULL_E | (L << 22), null_sp, 0, 0])
            pe = ents[sb]
            B.append(pe[:4])
            for i in range(n_ovf):
                ovf[i].append(pe[4 + 2 * i:6 + 2 * i])
        s_steps.append((A, B, ovf, L))
    null_s = ([[NULL_E, null_sp, 0, 0]] * nsub, [[zr] * 4] * nsub, [], 0)
    s_pairs = []
    for c0 in range(0, len(s_steps), 2):
        pr_ = s_steps[c0:c0 + 2]
        s_pairs.append(pr_ + [null_s] * (2 - len(pr_)))

    # D: rows sorted by their number of dense-only columns, NSUB rows per item
    sparse_set = {(e[1], e[2]) for e in sparse}
    # Factored output: [energy row: nsp][T column: nsp - 1][W_k a_k: nsp - 1][W_k b_k: nsp - 1]
    # [sparse block (W_k / W_j) S_kj in column-major pattern order].  fac_map: dense element -> slot of the
    # sparse block (-1: the element has no sparse part)
    fac_pat = sorted(sparse_set)
    fac_map = [-1] * (nsp * nsp)
    for s_, (col, k) in enumerate(fac_pat):
        fac_map[col * nsp + k + 1] = nsp + 3 * last + s_
    P['p5_fac_map'] = i32(fac_map)
    P['fac_rows'] = i32([k + 1 for col, k in fac_pat] or [0])
    P['fac_cols'] = i32([col for col, k in fac_pat] or [0])
    fac_nnz = len(fac_pat)
    d_rows = []                                # (row, its dense-only columns, owns the temperature column)
    for k in range(last):
        cols = [col for col in range(1, nsp) if (col, k) not in sparse_set]
        parts = [cols[c0:c0 + D_MAX_COLS] for c0 in range(0, len(cols), D_MAX_COLS)] or [[]]
        d_rows += [(k, part, i == 0) for i, part in enumerate(parts)]
    d_rows.sort(key=lambda r: (-len(r[1]), r[0]))
    d_items = []
    for c0 in range(0, len(d_rows), nsub):
        grp = d_rows[c0:c0 + nsub]
        n = (len(grp[0][1]) + 3) // 4 * 4                  # columns come in batches of four
        units = [[[sp_even(k), k + 1 if own else NULL_E] for k, _, own in grp] + [[NONE32, NULL_E]] * (nsub - len(grp))]
        for u in range(n):
            units.append([[cols[u] * nsp + k + 1, cols[u]] if u < len(cols) else [NULL_E, 0] for k, cols, _ in grp]
                         + [[NULL_E, 0]] * (nsub - len(grp)))
        d_items.append((n, units))

    # T: columns sorted by list length; tcoop sub-groups per column, one round per warp if possible
    tcoop = nsub
    while tcoop > 1 and nw * (nsub // tcoop) < last:
        tcoop //= 2
    t_per = nsub // tcoop
    t_cols = sorted(range(last), key=lambda j: (-len(tcontrib.get(j, [])), j))
    t_items = []
    for c0 in range(0, last, t_per):
        grp = t_cols[c0:c0 + t_per]
        n = max(-(-len(tcontrib.get(j, [])) // tcoop) for j in grp)
        n = (n + 3) // 4 * 4                               # units come in batches of four
        if n > 0xFFFF:
            raise ValueError('energy-row element with too many contributions')
        hdr, per_sub = [], []
        for sb in range(nsub):
            j = grp[sb // tcoop] if sb // tcoop < len(grp) else None
            if j is None:
                hdr.append([0, null_sp])
                per_sub.append([[(nraw + ((sb + i) & 1)) * RB, nr * RXB] for i in range(n)])
            else:
                lst = tcontrib.get(j, [])[sb % tcoop::tcoop]
                hdr.append([(j + 1) | ((1 if sb % tcoop == 0 else 0) << 16), (sp_even(j) ^ RB) + (SLOT_CP - 1) * RB])
                per_sub.append([[src * RB, rx_ * RXB] for src, rx_ in lst]
                               + [[(nraw + ((sb + i) & 1)) * RB, nr * RXB] for i in range(n - len(lst))])
        keyed = [[(pair[0] << 32) | pair[1] for pair in per_sub[sb]] for sb in range(nsub)]
        balance_banks(keyed, lambda v: ((v >> 32) // RB) & 1 if (RB & 127) else 0)
        per_sub = [[[v >> 32, v & 0xFFFFFFFF] for v in keyed[sb]] for sb in range(nsub)]
        units = [hdr] + [[per_sub[sb][u] for sb in range(nsub)] for u in range(n)]
        t_items.append((n, units))

    # one longest-first assignment over all three classes
    costs = ([COST_T_ITEM + COST_T_IT * n for n, _ in t_items] +
             [sum(COST_S_STEP + COST_S_OVF * len(st[2]) for st in pr_) for pr_ in s_pairs] +
             [COST_D_ITEM + COST_D_COL * n for n, _ in d_items])
    init = [0.0] * nw
    init[0] = COST_DOTS
    allbins, _ = _lpt(costs, nw, init)
    nT, nS = len(t_items), len(s_pairs)
    s_off, s_str, o_off, o_str = [0], [], [0], []
    d_off, d_item, d_str, t_off, t_item, t_str = [0], [], [], [0], [], []
    for w in range(nw):
        mine = sorted(allbins[w])
        for ix in mine:
            if ix < nT:
                t_item += [len(t_str) // (2 * nsub), t_items[ix][0]]
                for u in t_items[ix][1]:
                    for xy in u:
                        t_str += xy
            elif ix < nT + nS:
                for A, B, ovf, _ in s_pairs[ix - nT]:
                    for ws in A:
                        s_str += ws
                    for ws in B:
                        s_str += ws
                    for u in ovf:
                        for xy in u:
                            o_str += xy
            else:
                d_item += [len(d_str) // (2 * nsub), d_items[ix - nT - nS][0]]
                for u in d_items[ix - nT - nS][1]:
                    for xy in u:
                        d_str += xy
        t_off.append(len(t_item) // 2)
        d_off.append(len(d_item) // 2)
        s_off.append(len(s_str) // (8 * nsub))          # in steps
        o_off.append(len(o_str) // (2 * nsub))
    P['p5_d_off'] = i32(d_off)
    P['p5_d_item'] = i32(d_item + [len(d_str) // (2 * nsub), 0])
    P['p5_d_str'] = u32(d_str + [NULL_E, 0] * (nsub * PF))
    P['p5_s_off'] = i32(s_off)
    P['p5_s_str'] = u32(s_str + ([NULL_E, null_sp, 0, 0] * nsub + [zr] * (4 * nsub)) * 4)
    P['p5_o_off'] = i32(o_off)
    P['p5_o_str'] = u32(o_str + [zr] * (2 * nsub * PF))
    P['p5_t_off'] = i32(t_off)
    P['p5_t_item'] = i32(t_item + [len(t_str) // (2 * nsub), 0])
    P['p5_t_str'] = u32(t_str + [zr, nr * RXB] * (nsub * PF))
    t_nst = [t_off[w + 1] - t_off[w] for w in range(nw)]

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
    P['p5_cfg'] = i32([gs, nt, nw, nsub, L['SP'], L['RX'], L['RAW'], L['SC'], L['PA'], L['total'], t_sync, coop, tcoop, L['CF'],
                       1 if wsg else 0, fac_nnz])
    return P
