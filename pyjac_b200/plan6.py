"""Record streams of the Jacobian kernel ``pj6::k_jac6`` (csrc/jac6.cuh).

Same algebra and the same state-pair / sub-group mapping as :mod:`pyjac_b200.plan` (a warp's
lanes are NSUB = 64 / GS sub-groups of GS / 2 lanes, a lane carries two states, shared memory
holds rows of GS doubles), but *everything a warp needs to know about the mechanism arrives as
one private stream of fixed-format records*: a record is NSUB x 16 bytes, sub-group s reads word
s.  The streams live in global memory (L2-resident) and are copied ahead of their use into a
small per-warp ring in shared memory by bulk-asynchronous copies (``cp.async.bulk`` completing
on an mbarrier), so that no table word is fetched through the load/store pipe from global
memory and a table read costs one shared-memory wavefront per record.

Per group of GS states a warp consumes, in this order (block barriers between the phases):

  B   rounds of NSUB reactions: four records = the reaction's 64-byte record, transposed
  C   species rounds: header {species rows, #records, W_k} + records of eight 16-bit signed
      reaction rows (sums of net rate, T-column term, X1) + records of signed correction rows
      (X1 + X2 of the few reactions where that is not zero);
      energy-row rounds: header {column} + records of four (raw row, reaction row) pairs
  DE  segments: header {species rows of one Jacobian row per sub-group, W_k} + L element
      records {element, column, flags, six signed 16-bit raw rows}: a sub-group keeps
      W_k a_k, W_k b_k of its row in registers and walks a piece of that row; the records of a
      piece are sorted by their number of entries and every record carries the step's
      (warp-uniform) entry count, so that the kernel issues only as many gathers as the
      longest list of the step needs; then records of the energy-equation row.

Replaces, like plan.py, the statement unrolling of pyjac/core/create_jacobian.py:2650-2976,
3095-3254 and rate_subs.py:1425-1527.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .plan import (COST_EFF, COST_IRREV, COST_PLAIN, COST_PLOG, COST_PM, COST_THREE, NONE32, NPART,
                   NSCAL, SMEM_LIMIT, _f64_words, _lpt, balance_banks)

PLAN_VERSION = 4             # of the p6_* record streams (checked by the library against its own)
CHB = 512                    # bytes per stream chunk (one bulk copy)
NSLOT = 3                    # ring slots per warp (two measured equally fast: profiles/README.md)
SP_SLOTS, RX_SLOTS = 6, 4    # C B|WB dB|WA hW WT cp  /  net tT X1 dH
F_NULL = 1 << 28             # reaction record of a padding sub-group
F_CORR = 1 << 29             # the reaction writes X1 + X2 to its correction row
ENT_NULL = 0x7FFF
D_CIN, D_COUT, D_VALID, D_HDR = 1 << 28, 1 << 29, 1 << 30, 1 << 31      # carry in / out, element exists, header record
MAX_NSP = 255                # 16-bit element index, 8-bit column
GS6 = (4, 8, 16, 32)         # states per block k_jac6 is built for

# cost model (warp instructions) for load balancing only
C6_B_PM, C6_C_HDR, C6_C_REC, C6_X_REC, C6_T_HDR, C6_T_REC = 1.0, 70.0, 120.0, 60.0, 40.0, 45.0
C6_D_REC, C6_D_ENT, C6_D_HDR, C6_E_REC = 27.0, 8.0, 30.0, 40.0
C6_DOTS = 500.0              # warp 0: energy-equation scalars + the next group's phase A0


def layout6(nsp: int, nr: int, ncorr: int, nraw: int, gs: int, nw: int) -> Dict[str, int]:
    """Shared-memory carve-up: region offsets in doubles (multiples of 16 doubles = 128 bytes),
    ``ring`` / ``mbar`` / ``bytes`` in bytes."""
    off = 0
    L: Dict[str, int] = {}

    al = max(16, 2 * gs)                        # regions start on a multiple of two rows (slot swizzle)

    def take(name, doubles):
        nonlocal off
        L[name] = off
        off += -(-doubles // al) * al

    take('SP', (nsp + 1) * SP_SLOTS * gs)       # species nsp: the empty reaction slot
    take('RX', (nr + 2) * RX_SLOTS * gs)        # reactions nr, nr + 1: zeros
    take('XC', (ncorr + 2) * gs)                # rows ncorr, ncorr + 1: zeros
    take('RAW', (nraw + 2) * gs)                # rows nraw, nraw + 1: zeros
    take('ET', max(nsp - 1, 1) * gs)            # energy-row gathers, one row per column
    take('SC', NSCAL * gs)
    take('PA', nw * NPART * gs)
    take('CF', 2 * nsp)
    L['ring'] = off * 8
    L['mbar'] = L['ring'] + nw * NSLOT * CHB
    L['bytes'] = L['mbar'] + nw * NSLOT * 8
    return L


def fits(nsp: int, nr: int, ncorr: int, nraw: int, gs: int, nw: int) -> bool:
    return nsp <= MAX_NSP and nraw + 2 < ENT_NULL and (nr + 2) * RX_SLOTS + 1 < ENT_NULL and \
        layout6(nsp, nr, ncorr, nraw, gs, nw)['bytes'] <= SMEM_LIMIT


def build_plan6(nsp: int, nr: int, nraw: int, first_pm: int, p_c0: int, kinds: List[str],
                is_rev: List[bool], has3: List[bool], n_eff: List[int], rec6: np.ndarray,
                red: List[List[Tuple[int, float]]], corr_rx: Sequence[bool],
                contrib: Dict[Tuple[int, int], List[Tuple[int, float]]],
                tcontrib: Dict[int, List[Tuple[int, int]]], sp_w: Sequence[float],
                sp_iw: Sequence[float], sp_mwf: Sequence[float], gs: int, nt: int) -> Dict[str, np.ndarray]:
    """rec6[p] = the 16-int record of kernel-order reaction p; corr_rx[p]: reaction p >= p_c0 may
    have X1 + X2 != 0; the other arguments as :func:`pyjac_b200.plan.build_plan`."""
    assert nt % 32 == 0 and 64 <= nt <= 512 and 64 % gs == 0
    nw, nsub, last = nt // 32, 64 // gs, nsp - 1
    ncorr = nr - p_c0
    RB = gs * 8
    SPB, RXB = SP_SLOTS * RB, RX_SLOTS * RB
    rec_words = nsub * 4                       # uint32 words per record
    chr_ = CHB // (nsub * 16)                  # records per chunk
    assert chr_ >= 1
    halves = (RB & 127) != 0                   # rows narrower than a bank line: two per line

    def sp_even(k):
        return k * SPB + (k & 1) * RB

    def rxv(p):
        """Row index (in units of rows) of reaction p's even-slot base: like the species rows, the slot
        pairs of odd reactions are swapped so that equal slots of different reactions spread over both
        halves of a bank line."""
        return p * RX_SLOTS + (p & 1)

    def expand(lst):
        """[(source, nu)] -> signed unit entries [(source, sign)], nu an integer."""
        out = []
        for src, c in lst:
            if not float(c).is_integer():
                raise ValueError('non-integer coefficient %r' % c)
            out += [(src, 0 if c > 0 else 1)] * int(abs(c))
        return out

    streams: List[List[int]] = [[] for _ in range(nw)]     # uint32 words per warp
    hdr = np.zeros((nw, 8), dtype=np.int64)

    def emit(w, rec):
        """rec: NSUB lists of four uint32 words."""
        assert len(rec) == nsub and all(len(r) == 4 for r in rec)
        for r in rec:
            streams[w] += [int(v) & 0xFFFFFFFF for v in r]

    def align(w, fill):
        """Pads warp w's stream to a chunk boundary with records of `fill` words per sub-group."""
        while len(streams[w]) % (CHB // 4):
            emit(w, [list(fill)] * nsub)

    # ---------------------------------------------------------------- phase B
    pm = list(range(first_pm, nr))
    plain = sorted(range(first_pm), key=lambda p: (has3[p], kinds[p] == 'plog', not is_rev[p], p))
    rounds: List[Tuple[bool, List[int], float]] = []
    for c0 in range(0, len(pm), nsub):
        grp = pm[c0:c0 + nsub]
        cost = max(COST_PM[kinds[p]] for p in grp) + COST_EFF * max(n_eff[p] for p in grp)
        if any(has3[p] for p in grp):
            cost += COST_THREE
        rounds.append((True, grp, cost * C6_B_PM))
    for c0 in range(0, len(plain), nsub):
        grp = plain[c0:c0 + nsub]
        cost = COST_PLAIN if any(is_rev[p] for p in grp) else COST_IRREV
        if any(has3[p] for p in grp):
            cost += COST_THREE
        if any(kinds[p] == 'plog' for p in grp):
            cost += COST_PLOG
        rounds.append((False, grp, cost))
    bins, _ = _lpt([r[2] for r in rounds], nw)
    for w in range(nw):
        mine = sorted(bins[w], key=lambda ix: (not rounds[ix][0], ix))     # pm rounds first
        hdr[w, 2] = sum(1 for ix in mine if rounds[ix][0])
        hdr[w, 3] = len(mine) - hdr[w, 2]
        for ix in mine:
            is_pm, grp, _ = rounds[ix]
            filler = first_pm if is_pm else 0
            recs = []
            for sb in range(nsub):
                if sb < len(grp):
                    recs.append(np.array(rec6[grp[sb]], dtype=np.int64) & 0xFFFFFFFF)
                else:
                    r_ = np.array(rec6[filler], dtype=np.int64) & 0xFFFFFFFF
                    r_[8] |= F_NULL
                    recs.append(r_)
            for q_ in range(4):
                emit(w, [list(r_[4 * q_:4 * q_ + 4]) for r_ in recs])

    # ---------------------------------------------------------------- phase C
    c_main = [expand(red[k]) for k in range(nsp)]
    c_corr = [expand([(p - p_c0, nu) for p, nu in red[k] if p >= p_c0 and corr_rx[p]]) for k in range(nsp)]
    t_list = [list(tcontrib.get(j, [])) for j in range(last)]

    def nrec8(n, coop):
        return -(-(-(-n // coop)) // 8)

    def nrec4(n, coop):
        return -(-(-(-n // coop)) // 4)

    def c_rounds(coop):
        per_round = nsub // coop
        order_c = sorted(range(nsp), key=lambda k: (-(nrec8(len(c_main[k]), coop) * 2 + nrec8(len(c_corr[k]), coop)), k))
        out = []
        for c0 in range(0, nsp, per_round):
            ch = order_c[c0:c0 + per_round]
            nm = max(nrec8(len(c_main[k]), coop) for k in ch)
            nx = max(nrec8(len(c_corr[k]), coop) for k in ch)
            out.append(('C', ch, nm, nx, C6_C_HDR + C6_C_REC * nm + C6_X_REC * nx))
        return out

    def t_rounds(tcoop):
        per_round = nsub // tcoop
        order_t = sorted(range(last), key=lambda j: (-nrec4(len(t_list[j]), tcoop), j))
        out = []
        for c0 in range(0, last, per_round):
            ch = order_t[c0:c0 + per_round]
            n = max(nrec4(len(t_list[j]), tcoop) for j in ch)
            out.append(('T', ch, n, 0, C6_T_HDR + C6_T_REC * n))
        return out

    best = None
    coops = [c for c in (1, 2, 4, 8, 16, 32) if c <= nsub]
    for coop in coops:
        for tcoop in coops:
            rs = c_rounds(coop) + t_rounds(tcoop)
            _, load = _lpt([r[4] for r in rs], nw)
            key = (max(load), sum(load))
            if best is None or key < best[0]:
                best = (key, coop, tcoop, rs)
    _, coop, tcoop, c_all = best
    bins, _ = _lpt([r[4] for r in c_all], nw)
    for w in range(nw):
        mine = sorted(bins[w], key=lambda ix: (c_all[ix][0] != 'C', ix))   # species rounds first
        hdr[w, 4] = sum(1 for ix in mine if c_all[ix][0] == 'C')
        hdr[w, 5] = len(mine) - hdr[w, 4]
        for ix in mine:
            kind, ch, n1, n2, _ = c_all[ix]
            if kind == 'C':
                subs_k = [ch[sb // coop] if sb // coop < len(ch) else None for sb in range(nsub)]
                head = []
                for sb in range(nsub):
                    k = subs_k[sb]
                    if k is None:
                        head.append([NONE32, (n1 << 8) | (n2 << 16), 0, 0])
                    else:
                        head.append([sp_even(k), (1 if sb % coop == 0 else 0) | (n1 << 8) | (n2 << 16)] + _f64_words(sp_w[k]))
                emit(w, head)
                for which, nrec, zero0, row in ((c_main, n1, nr, rxv), (c_corr, n2, ncorr, lambda v: v)):
                    per_sub = []
                    for sb in range(nsub):
                        k = subs_k[sb]
                        lst = [] if k is None else which[k][sb % coop::coop]
                        ent = [row(src) | (sg << 15) for src, sg in lst]
                        ent += [row(zero0 + ((sb + i) & 1)) for i in range(8 * nrec - len(ent))]
                        per_sub.append(ent)
                    if halves:
                        balance_banks(per_sub, lambda v: v & 1)
                    for u in range(nrec):
                        emit(w, [[per_sub[sb][8 * u + 2 * i] | (per_sub[sb][8 * u + 2 * i + 1] << 16) for i in range(4)]
                                 for sb in range(nsub)])
            else:
                subs_j = [ch[sb // tcoop] if sb // tcoop < len(ch) else None for sb in range(nsub)]
                head = []
                for sb in range(nsub):
                    j = subs_j[sb]
                    head.append([0, n1, 0, 0] if j is None else [(j + 1) | ((1 if sb % tcoop == 0 else 0) << 16), n1, 0, 0])
                emit(w, head)
                per_sub = []
                for sb in range(nsub):
                    j = subs_j[sb]
                    lst = [] if j is None else t_list[j][sb % tcoop::tcoop]
                    ent = [src | (rxv(rx_) << 16) for src, rx_ in lst]
                    ent += [(nraw + ((sb + i) & 1)) | (rxv(nr) << 16) for i in range(4 * n1 - len(ent))]
                    per_sub.append(ent)
                if halves:
                    balance_banks(per_sub, lambda v: v & 1)
                for u in range(n1):
                    emit(w, [per_sub[sb][4 * u:4 * u + 4] for sb in range(nsub)])

    # ---------------------------------------------------------------- phase DE
    # Elements by rows.  A sub-group walks a *piece* of one Jacobian row k with W_k a_k, W_k b_k, W_k in
    # registers; the element index follows from the column (e = col * NSP + k + 1), so a record only
    # names columns and signed raw rows.  Record kinds, by the number of entries of an element's list:
    #   K0  16 dense-only elements: sixteen column bytes
    #   K1  5 elements with one entry:   cols in w0 and the low byte of w1, entries in w1.hi, w2, w3
    #   K2  3 elements with two entries: cols in w0, entry pairs in w1, w2, w3
    #   K4 / K6  one element with up to four / six entries: col (and D_VALID) in w0, entries in w1 .. w3
    #   K7  one element of a list longer than six: like K6, the accumulator carried in (D_CIN) / out (D_COUT)
    # col = 0 marks an absent element.  A *segment* = eight pieces (one per sub-group): a header record
    # {D_HDR | owns the T-column element | k << 8, species rows, W_k}, a record of the per-kind record
    # counts (warp-uniform: the maxima over the pieces), then the records kind by kind.
    KINDS = (7, 6, 4, 2, 1, 0)
    PER_REC = {0: 16, 1: 5, 2: 3, 4: 1, 6: 1, 7: 1}
    KCOST = {0: 34.0 + 11.0 * 16, 1: 34.0 + 19.0 * 5, 2: 34.0 + 29.0 * 3, 4: 34.0 + 50.0, 6: 34.0 + 66.0, 7: 34.0 + 72.0}

    def kind_of(n):
        return 0 if n == 0 else 1 if n == 1 else 2 if n == 2 else 4 if n <= 4 else 6 if n <= 6 else 7

    def row_elems(k):
        """Per kind: the elements (col, entries) of row k; a K7 element is a list of its 6-entry parts."""
        by = {kd: [] for kd in KINDS}
        for col in range(1, nsp):
            ent = expand(contrib.get((k, col - 1), []))
            kd = kind_of(len(ent))
            if kd == 7:
                by[7].append((col, [ent[i:i + 6] for i in range(0, len(ent), 6)]))
            else:
                by[kd].append((col, ent))
        return by

    def n_records(by):
        return {kd: (sum(len(parts) for _, parts in by[7]) if kd == 7 else -(-len(by[kd]) // PER_REC[kd])) for kd in KINDS}

    def piece_cost(by):
        nr_ = n_records(by)
        return sum(nr_[kd] * KCOST[kd] for kd in KINDS)

    rows = [row_elems(k) for k in range(last)]
    row_cost = [piece_cost(by) for by in rows]
    total_cost = sum(row_cost)

    def build_de(target):
        """Cut the rows into pieces of about `target` cost, group NSUB pieces of similar shape into a
        segment, deal the segments to the warps; returns (largest warp cost, per-warp segments, loads)."""
        pieces = []
        for k, by in enumerate(rows):
            n_k = max(1, int(round(row_cost[k] / target)))
            parts = [{kd: [] for kd in KINDS} for _ in range(n_k)]
            at = 0
            for kd in KINDS:                       # deal each kind round-robin, continuing where the last one stopped
                for el in by[kd]:
                    parts[at % n_k][kd].append(el)
                    at += 1
            for i, pc in enumerate(parts):
                pieces.append((k, i == 0, pc, n_records(pc)))
        pieces.sort(key=lambda pc: tuple(-pc[3][kd] for kd in KINDS) + (pc[0],))
        segs = []
        for c0 in range(0, len(pieces), nsub):
            grp = pieces[c0:c0 + nsub]
            counts = {kd: max(pc[3][kd] for pc in grp) for kd in KINDS}
            cost = C6_D_HDR + sum(counts[kd] * KCOST[kd] for kd in KINDS)
            segs.append((grp, counts, cost))
        init = [0.0] * nw
        init[0] = C6_DOTS
        bins_, load_ = _lpt([s_[2] for s_ in segs], nw, init)
        return max(load_), [[segs[ix] for ix in sorted(b)] for b in bins_], load_

    best = None
    per_worker = total_cost / (nw * nsub)
    for nseg in range(1, 7):
        for fudge in (0.8, 0.9, 1.0, 1.1, 1.2):
            res = build_de(per_worker / nseg * fudge)
            if best is None or res[0] < best[0]:
                best = res
    _, de_warps, de_load = best

    def ent16(lst, n, sb):
        """n signed 16-bit entries: the list, padded with the zero rows."""
        out = [src | (sg << 15) for src, sg in lst]
        return out + [nraw + ((sb + i) & 1) for i in range(n - len(out))]

    def pack16(e):
        return [e[2 * i] | (e[2 * i + 1] << 16) for i in range(len(e) // 2)]

    for w in range(nw):
        hdr[w, 6] = len(de_warps[w])
        for grp, counts, _ in de_warps[w]:
            head = []
            for sb in range(nsub):
                if sb < len(grp):
                    k, own_t = grp[sb][0], grp[sb][1]
                    head.append([D_HDR | D_VALID | (1 if own_t else 0) | (k << 8), sp_even(k)] + _f64_words(sp_w[k]))
                else:
                    head.append([D_HDR, 0, 0, 0])
            emit(w, head)
            assert all(counts[kd] < 256 for kd in KINDS)
            cw = [counts[7] | (counts[6] << 8) | (counts[4] << 16) | (counts[2] << 24), counts[1] | (counts[0] << 8), 0, 0]
            emit(w, [cw] * nsub)
            for kd in KINDS:
                per = PER_REC[kd]
                for i in range(counts[kd]):
                    rec = []
                    for sb in range(nsub):
                        pc = grp[sb][2] if sb < len(grp) else {k_: [] for k_ in KINDS}
                        if kd == 7:
                            flat = [(col, part, j == 0, j == len(parts) - 1) for col, parts in pc[7] for j, part in enumerate(parts)]
                            if i < len(flat):
                                col, part, first, final = flat[i]
                                x = col | D_VALID | (0 if first else D_CIN) | (0 if final else D_COUT)
                                rec.append([x] + pack16(ent16(part, 6, sb)))
                            else:
                                rec.append([0] + pack16(ent16([], 6, sb)))
                        elif kd in (4, 6):
                            if i < len(pc[kd]):
                                col, ent = pc[kd][i]
                                rec.append([col | D_VALID] + pack16(ent16(ent, 6, sb)))
                            else:
                                rec.append([0] + pack16(ent16([], 6, sb)))
                        else:
                            els = pc[kd][i * per:(i + 1) * per]
                            cols = [c_ for c_, _ in els] + [0] * (per - len(els))
                            if kd == 0:
                                rec.append([cols[4 * j] | (cols[4 * j + 1] << 8) | (cols[4 * j + 2] << 16) | (cols[4 * j + 3] << 24)
                                            for j in range(4)])
                            elif kd == 1:
                                ents = [ent16(e_, 1, sb + j)[0] for j, (_, e_) in enumerate(els)] + \
                                       [ent16([], 1, sb + j)[0] for j in range(per - len(els))]
                                rec.append([cols[0] | (cols[1] << 8) | (cols[2] << 16) | (cols[3] << 24),
                                            cols[4] | (ents[0] << 16), ents[1] | (ents[2] << 16), ents[3] | (ents[4] << 16)])
                            else:
                                pairs = [ent16(e_, 2, sb) for _, e_ in els] + [ent16([], 2, sb) for _ in range(per - len(els))]
                                rec.append([cols[0] | (cols[1] << 8) | (cols[2] << 16)] + [pr_[0] | (pr_[1] << 16) for pr_ in pairs])
                    emit(w, rec)

    # the energy-equation row: records of NSUB columns, dealt to the warps with the least phase-DE work
    e_recs = [list(range(c0, min(c0 + nsub, last))) for c0 in range(0, last, nsub)]
    load = list(de_load)
    e_of = [[] for _ in range(nw)]
    for cols in e_recs:
        w = min(range(nw), key=lambda b: (load[b], b))
        e_of[w].append(cols)
        load[w] += C6_E_REC
    for w in range(nw):
        hdr[w, 7] = len(e_of[w])
        for cols in e_of[w]:
            emit(w, [[cols[sb] + 1 if sb < len(cols) else 0, 0, 0, 0] for sb in range(nsub)])

    # ---------------------------------------------------------------- streams -> chunks
    words_per_chunk = CHB // 4
    allw: List[int] = []
    for w in range(nw):
        n_words = len(streams[w])
        assert n_words % rec_words == 0
        pad = (-n_words) % words_per_chunk
        hdr[w, 0] = len(allw) // words_per_chunk
        hdr[w, 1] = (n_words + pad) // words_per_chunk
        allw += streams[w] + [0] * pad
    P: Dict[str, np.ndarray] = {}
    P['p6_str'] = np.asarray(allw + [0] * words_per_chunk, dtype=np.uint32).view(np.int32)
    P['p6_hdr'] = hdr.astype(np.int32).ravel()

    colfac = [1.0, 0.0]
    for j in range(last):
        colfac += [sp_iw[j], sp_iw[j] * sp_mwf[j]]
    P['p6_colfac'] = np.asarray(colfac, dtype=np.float64)

    L = layout6(nsp, nr, ncorr, nraw, gs, nw)
    waiters = sum(1 for w in range(1, nw) if hdr[w, 7])
    t_sync = 32 * (waiters + 1) if waiters else 0
    P['p6_cfg'] = np.asarray([gs, nt, nw, nsub, L['SP'], L['RX'], L['XC'], L['RAW'], L['ET'], L['SC'], L['PA'], L['CF'],
                              L['ring'], L['mbar'], L['bytes'], t_sync, coop, tcoop, p_c0, ncorr,
                              CHB, NSLOT, chr_, 0], dtype=np.int32)
    return P
