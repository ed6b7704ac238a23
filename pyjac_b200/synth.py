"""Deterministic synthetic Chemkin mechanisms with the *shape* of named mechanisms.

GRI-Mech 3.0, USC-Mech II and the LLNL n-heptane mechanism are not available
offline (SURVEY.md finding 3), so BASELINE.json's configs 2-5 run on mechanisms
generated here: the requested species / reaction counts and reaction-type mix,
element-balanced reactions over a C/H/O/N/Ar species pool, NASA-7 fits that are
continuous at T_mid, rate parameters in the usual Chemkin ranges.  The output is
ordinary Chemkin text, so the reference generator (the oracle) and this package
read identical parameters.  Results on these are labelled "<name>-shaped
synthetic" everywhere.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

ELEMS = ('C', 'H', 'O', 'N', 'AR')

# the 53 species of GRI-Mech 3.0 (names + compositions only; thermo is synthetic)
_GRI_SPECIES = [
    ('H2', (0, 2, 0, 0, 0)), ('H', (0, 1, 0, 0, 0)), ('O', (0, 0, 1, 0, 0)),
    ('O2', (0, 0, 2, 0, 0)), ('OH', (0, 1, 1, 0, 0)), ('H2O', (0, 2, 1, 0, 0)),
    ('HO2', (0, 1, 2, 0, 0)), ('H2O2', (0, 2, 2, 0, 0)), ('C', (1, 0, 0, 0, 0)),
    ('CH', (1, 1, 0, 0, 0)), ('CH2', (1, 2, 0, 0, 0)), ('CH2(S)', (1, 2, 0, 0, 0)),
    ('CH3', (1, 3, 0, 0, 0)), ('CH4', (1, 4, 0, 0, 0)), ('CO', (1, 0, 1, 0, 0)),
    ('CO2', (1, 0, 2, 0, 0)), ('HCO', (1, 1, 1, 0, 0)), ('CH2O', (1, 2, 1, 0, 0)),
    ('CH2OH', (1, 3, 1, 0, 0)), ('CH3O', (1, 3, 1, 0, 0)), ('CH3OH', (1, 4, 1, 0, 0)),
    ('C2H', (2, 1, 0, 0, 0)), ('C2H2', (2, 2, 0, 0, 0)), ('C2H3', (2, 3, 0, 0, 0)),
    ('C2H4', (2, 4, 0, 0, 0)), ('C2H5', (2, 5, 0, 0, 0)), ('C2H6', (2, 6, 0, 0, 0)),
    ('HCCO', (2, 1, 1, 0, 0)), ('CH2CO', (2, 2, 1, 0, 0)), ('HCCOH', (2, 2, 1, 0, 0)),
    ('N', (0, 0, 0, 1, 0)), ('NH', (0, 1, 0, 1, 0)), ('NH2', (0, 2, 0, 1, 0)),
    ('NH3', (0, 3, 0, 1, 0)), ('NNH', (0, 1, 0, 2, 0)), ('NO', (0, 0, 1, 1, 0)),
    ('NO2', (0, 0, 2, 1, 0)), ('N2O', (0, 0, 1, 2, 0)), ('HNO', (0, 1, 1, 1, 0)),
    ('CN', (1, 0, 0, 1, 0)), ('HCN', (1, 1, 0, 1, 0)), ('H2CN', (1, 2, 0, 1, 0)),
    ('HCNN', (1, 1, 0, 2, 0)), ('HCNO', (1, 1, 1, 1, 0)), ('HOCN', (1, 1, 1, 1, 0)),
    ('HNCO', (1, 1, 1, 1, 0)), ('NCO', (1, 0, 1, 1, 0)), ('N2', (0, 0, 0, 2, 0)),
    ('AR', (0, 0, 0, 0, 1)), ('C3H7', (3, 7, 0, 0, 0)), ('C3H8', (3, 8, 0, 0, 0)),
    ('CH2CHO', (2, 3, 1, 0, 0)), ('CH3CHO', (2, 4, 1, 0, 0)),
]

_COLLIDERS = [('H2', 2.0), ('H2O', 6.0), ('CH4', 2.0), ('CO', 1.5), ('CO2', 2.0),
              ('C2H6', 3.0), ('AR', 0.7), ('O2', 0.78), ('N2', 1.0)]


@dataclass
class Shape:
    name: str
    nsp: int
    nr: int
    n_third: int
    n_troe: int        # Troe fall-off (half with T2)
    n_lind: int        # Lindemann fall-off
    n_irrev: int
    n_dup_pairs: int
    tmid_choices: Tuple[float, ...] = (1000.0,)


SHAPES: Dict[str, Shape] = {
    # GRI-Mech 3.0: 53 sp / 325 rxn, 29 fall-off, ~10 +M, 16 irreversible
    'gri30': Shape('gri30', 53, 325, 10, 26, 3, 16, 4),
    # USC-Mech II: 111 sp / 784 rxn
    'usc2': Shape('usc2', 111, 784, 14, 52, 6, 30, 8, (1000.0, 1385.0, 1392.0, 1400.0)),
    # LLNL detailed n-heptane v3-like: ~650 sp / ~2800 rxn
    'nc7': Shape('nc7', 654, 2827, 12, 70, 10, 600, 10, (1000.0, 1382.0, 1391.0, 1400.0)),
    # small shape for quick tests
    'mini': Shape('mini', 20, 60, 4, 5, 2, 6, 2, (1000.0, 1400.0)),
}


def _species_pool(shape: Shape, rng) -> List[Tuple[str, Tuple[int, ...]]]:
    pool = list(_GRI_SPECIES)
    if shape.nsp <= len(pool):
        if shape.nsp == len(pool):
            return pool
        # keep the H2/O2 core + bath gases, then the first others
        core = [s for s in pool if s[0] in ('H2', 'H', 'O', 'O2', 'OH', 'H2O', 'HO2', 'H2O2',
                                           'N2', 'AR', 'CO', 'CO2', 'CH4', 'CH3', 'HCO',
                                           'CH2O', 'C2H6', 'CH2', 'C2H4', 'C2H5')]
        return core[:shape.nsp]
    names = set(n for n, _ in pool)
    # grow larger hydrocarbons / oxygenates CxHyOz, several isomers per formula
    c = 3
    while len(pool) < shape.nsp:
        for h in range(2 * c + 2, max(2 * c - 6, 1), -1):
            for o in range(0, 4):
                for iso in range(3):
                    if len(pool) >= shape.nsp:
                        break
                    nm = 'C%dH%d' % (c, h) + ('O%d' % o if o else '') + ('-%d' % iso if iso else '')
                    if nm in names:
                        continue
                    names.add(nm)
                    pool.append((nm, (c, h, o, 0, 0)))
        c += 1
    return pool


def _thermo(comp, rng, tmid):
    """NASA-7 (lo, hi) continuous in cp, h, s at tmid."""
    natoms = sum(comp)
    nb = max(natoms - 1, 0)
    lin = 1.0 if natoms > 1 else 0.0

    def cp_coeffs(scale):
        a0 = 2.5 + lin * (1.0 + 0.35 * nb) * (1 + 0.1 * rng.uniform(-1, 1))
        a1 = nb * 2.2e-3 * scale * (1 + 0.3 * rng.uniform(-1, 1))
        a2 = -nb * 7.0e-7 * scale * (1 + 0.3 * rng.uniform(-1, 1))
        a3 = nb * 1.1e-10 * scale * (1 + 0.3 * rng.uniform(-1, 1))
        a4 = -nb * 6.5e-15 * scale * (1 + 0.3 * rng.uniform(-1, 1))
        return [a0, a1, a2, a3, a4]

    lo = cp_coeffs(1.6)
    # low branch: stronger curvature, scaled so cp stays positive for T in [300, tmid]
    lo[2] *= 2.0
    lo[3] *= 6.0
    lo[4] *= 30.0
    hi = cp_coeffs(1.0)
    eta = {'C': 85000.0, 'H': 25500.0, 'O': 29200.0, 'N': 56000.0, 'AR': -745.0}
    a5 = sum(n * eta[e] for n, e in zip(comp, ELEMS)) - 41000.0 * nb + 2500.0 * rng.uniform(-1, 1) * lin
    a6 = 4.0 - 0.9 * nb + 1.5 * rng.uniform(-1, 1)
    lo += [a5, a6]

    def cp(a, T):
        return a[0] + T * (a[1] + T * (a[2] + T * (a[3] + T * a[4])))

    def h(a, T):
        return a[5] + T * (a[0] + T * (a[1] / 2 + T * (a[2] / 3 + T * (a[3] / 4 + T * a[4] / 5))))

    def s(a, T):
        return a[0] * math.log(T) + T * (a[1] + T * (a[2] / 2 + T * (a[3] / 3 + T * a[4] / 4))) + a[6]

    hi[0] += cp(lo, tmid) - cp(hi + [0, 0], tmid)
    hi += [0.0, 0.0]
    hi[5] = h(lo, tmid) - h(hi, tmid)
    hi[6] = s(lo, tmid) - s(hi, tmid)
    return lo, hi


def _fmt_thermo(name, comp, lo, hi, tmid) -> str:
    items = [(e, n) for e, n in zip(ELEMS, comp) if n]
    comp_s = ''.join('%-2s%3d' % (e, n) for e, n in items[:4]).ljust(20)
    l1 = '%-18s%-6s%s%s%10.3f%10.3f%10.3f' % (name[:18], 'SYNTH', comp_s, 'G', 300.0, 5000.0, tmid)
    l1 = l1.ljust(79) + '1'
    c = list(hi) + list(lo)
    l2 = ''.join('%15.8E' % v for v in c[0:5]).ljust(79) + '2'
    l3 = ''.join('%15.8E' % v for v in c[5:10]).ljust(79) + '3'
    l4 = ''.join('%15.8E' % v for v in c[10:14]).ljust(79) + '4'
    return '\n'.join([l1, l2, l3, l4]) + '\n'


def _side(names_nus) -> str:
    return '+'.join(('%d' % nu if nu != 1 else '') + nm for nm, nu in names_nus)


def _merge(names) -> List[Tuple[str, int]]:
    out: List[Tuple[str, int]] = []
    for nm in names:
        for k, (n2, nu) in enumerate(out):
            if n2 == nm:
                out[k] = (n2, nu + 1)
                break
        else:
            out.append((nm, 1))
    return out


def generate(shape_name: str = 'gri30', seed: int = 0) -> str:
    """Chemkin text of a synthetic mechanism of the named shape."""
    shape = SHAPES[shape_name]
    rng = np.random.default_rng(seed)
    pool = _species_pool(shape, rng)
    names = [n for n, _ in pool]
    comp = {n: np.array(c) for n, c in pool}
    reactive = [n for n in names if n not in ('AR', 'N2')]

    # ---- thermo
    thermo_txt = 'THERMO ALL\n   300.000  1000.000  5000.000\n'
    h_form: Dict[str, float] = {}
    for n, c in pool:
        tmid = float(rng.choice(shape.tmid_choices))
        lo, hi = _thermo(c, rng, tmid)
        h_form[n] = lo[5]
        thermo_txt += _fmt_thermo(n, c, lo, hi, tmid)
    thermo_txt += 'END\n'

    # ---- element-balanced reaction templates
    by_comp: Dict[Tuple[int, ...], List[str]] = {}
    for n in reactive:
        by_comp.setdefault(tuple(comp[n]), []).append(n)
    pairs: Dict[Tuple[int, ...], List[Tuple[str, str]]] = {}
    for a, b in itertools.combinations_with_replacement(reactive, 2):
        pairs.setdefault(tuple(comp[a] + comp[b]), []).append((a, b))
    exch_keys = [k for k, v in pairs.items() if len(v) >= 2]
    recomb = [(a, b, c) for k, v in pairs.items() if k in by_comp for (a, b) in v for c in by_comp[k]]

    def pick_exchange():
        k = exch_keys[rng.integers(len(exch_keys))]
        v = pairs[k]
        i, j = rng.choice(len(v), size=2, replace=False)
        return list(v[i]), list(v[j])

    def pick_recomb():
        a, b, c = recomb[rng.integers(len(recomb))]
        return [a, b], [c]

    def endo(r, p):
        # endothermicity [cal/mol] of r -> p; an activation energy below it is unphysical
        # and makes kf/Kc astronomically large
        dh = sum(h_form[x] for x in p) - sum(h_form[x] for x in r)
        return max(dh, 0.0) * 1.987

    def arrh(order, kind='elem', e_min=0.0):
        # log10(A) is tied to b so that k(1000 K) stays below collision-limit magnitudes
        if kind == 'third':
            b = round(rng.uniform(-2.0, 0.0), 2)
            A = 10 ** (rng.uniform(13.5, 15.5) - 3.0 * b)
            E = 0.0 if rng.random() < 0.7 else round(rng.uniform(0, 20000), 1)
        elif kind == 'low':
            b = round(rng.uniform(-7.5, -1.0), 2)
            A = 10 ** (rng.uniform(15.0, 19.0) - 3.0 * b)
            E = round(rng.uniform(-1000, 8000), 1)
            if e_min > 0:
                E = round(E + e_min, 1)
        elif kind == 'inf':
            b = round(rng.uniform(-1.0, 1.5), 3)
            A = 10 ** (rng.uniform(11.0, 14.0) - 3.0 * b)
            E = 0.0 if rng.random() < 0.4 else round(rng.uniform(0, 12000), 1)
        else:
            u = rng.random()
            b = 0.0 if u < 0.35 else round(rng.uniform(-1.5, 2.8), 3)
            A = 10 ** (rng.uniform(10.5, 14.0) - 3.0 * b + 3.0 * (order - 2))
            E = 0.0 if rng.random() < 0.3 else round(rng.uniform(-2000, 45000), 1)
        if kind != 'low' and E < e_min:
            E = round(e_min * (1.0 + 0.1 * rng.random()), 1)
        return '%.3E %8.3f %10.2f' % (A, b, E)

    def eff_line():
        k = int(rng.integers(3, 8))
        idx = rng.choice(len(_COLLIDERS), size=k, replace=False)
        items = []
        for i in sorted(idx):
            nm, base = _COLLIDERS[i]
            if nm not in comp:
                continue
            val = round(base * (1 + 0.3 * rng.uniform(-1, 1)), 2)
            if rng.random() < 0.08:
                val = 0.0
            items.append('%s/%.2f/' % (nm, val))
        return ' '.join(items)

    lines: List[str] = []
    seen = set()
    n_plain = shape.nr - shape.n_third - shape.n_troe - shape.n_lind - 2 * shape.n_dup_pairs
    n_irrev_left = shape.n_irrev

    def emit(reac, prod, rev, rate, aux=()):
        arrow = '<=>' if rev else '=>'
        lines.append('%-48s %s' % (_side(_merge(reac)) + arrow + _side(_merge(prod)), rate))
        lines.extend(aux)

    # plain elementary (some termolecular / three-product)
    count = 0
    while count < n_plain:
        r, p = pick_exchange()
        u = rng.random()
        if u < 0.04:
            # add a spectator to both sides -> termolecular, as H+O2+H2O<=>HO2+H2O
            sp = reactive[rng.integers(len(reactive))]
            r, p = r + [sp], p + [sp]
        key = (tuple(sorted(r)), tuple(sorted(p)))
        if key in seen or (key[1], key[0]) in seen or sorted(r) == sorted(p):
            continue
        seen.add(key)
        rev = True
        if n_irrev_left > 0 and rng.random() < 1.5 * shape.n_irrev / shape.nr:
            rev = False
            n_irrev_left -= 1
        emit(r, p, rev, arrh(len(r), 'elem', endo(r, p)))
        count += 1

    # duplicate pairs
    for _ in range(shape.n_dup_pairs):
        r, p = pick_exchange()
        for _k in range(2):
            emit(r, p, True, arrh(2, 'elem', endo(r, p)), (' DUPLICATE',))

    # third-body
    for k in range(shape.n_third):
        r, p = pick_recomb()
        if rng.random() < 0.3:
            r, p = p, r
        aux = (eff_line(),) if k != 1 else ()      # one +M reaction without listed efficiencies
        lines.append('%-48s %s' % (_side(_merge(r)) + '+M<=>' + _side(_merge(p)) + '+M',
                                   arrh(len(r), 'third', endo(r, p))))
        lines.extend(aux)

    # fall-off (Troe with/without T2, Lindemann)
    for k in range(shape.n_troe + shape.n_lind):
        r, p = pick_recomb()
        if rng.random() < 0.15:
            r, p = p, r
        aux = ['     LOW  / %s /' % arrh(len(r), 'low', endo(r, p))]
        if k < shape.n_troe:
            a = round(rng.uniform(0.2, 0.95), 4)
            T3 = round(10 ** rng.uniform(1.7, 3.5), 2)
            T1 = round(10 ** rng.uniform(2.7, 4.0), 2)
            if k % 2 == 0:
                T2 = round(10 ** rng.uniform(3.0, 4.0), 2)
                aux.append('     TROE/ %.4f %.2f %.2f %.2f /' % (a, T3, T1, T2))
            else:
                aux.append('     TROE/ %.4f %.2f %.2f /' % (a, T3, T1))
        if k % 7 != 3:
            aux.append(eff_line())
        lines.append('%-48s %s' % (_side(_merge(r)) + '(+M)<=>' + _side(_merge(p)) + '(+M)',
                                   arrh(len(r), 'inf', endo(r, p))))
        lines.extend(aux)

    # interleave deterministically so reaction types are spread through the list
    blocks: List[List[str]] = []
    for ln in lines:
        if '=' in ln:
            blocks.append([ln])
        else:
            blocks[-1].append(ln)
    order = rng.permutation(len(blocks))
    body = '\n'.join('\n'.join(blocks[i]) for i in order)

    elems_used = [e for k, e in enumerate(ELEMS) if any(c[k] for _, c in pool)]
    txt = 'ELEMENTS\n' + ' '.join(elems_used) + '\nEND\nSPECIES\n'
    for k in range(0, len(names), 6):
        txt += ' '.join('%-12s' % n for n in names[k:k + 6]).rstrip() + '\n'
    txt += 'END\n' + thermo_txt + 'REACTIONS\n' + body + '\nEND\n'
    return txt


def write(shape_name: str, path: str, seed: int = 0) -> str:
    txt = generate(shape_name, seed)
    with open(path, 'w') as fh:
        fh.write(txt)
    return path
