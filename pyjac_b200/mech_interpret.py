"""Chemkin-format mechanism reader (the input side of the hot path).

Replaces ``read_mech`` / ``read_thermo`` of the reference
(pyjac/core/mech_interpret.py:56-883) with an independent implementation that
produces the *same numbers*: Arrhenius ``A`` converted from mol-cm-s to kmol-m-s
by reaction order (``:438-452``), activation energies converted to activation
temperatures (``:42-49``), LOW/HIGH/TROE/SRI/REV auxiliary data (``:468-588``),
third-body efficiencies (``:655-661``), Troe ``T3``/``T1`` of zero bumped to 1e-30
(``:551-557``), explicit-REV reactions split into two irreversible ones
(``:693-713``) and NASA-7 fixed-column thermo blocks (``:735-883``).

PLOG and Chebyshev auxiliary lines are read as the reference reads them (``:589-654``,
``:664-680``); which of them the kernels evaluate is decided in :mod:`pyjac_b200.tables`.
"""
from __future__ import annotations

import copy
import logging
import math
import re
from typing import List, Optional, Tuple

from .chem import ELEM_WT, PA, RU_JOUL, Reaction, Species

_E_FACT = {
    'kelvins': 1.0,
    'evolts': 11595.,
    'cal/mole': 4.184 / RU_JOUL,
    'kcal/mole': 4184. / RU_JOUL,
    'joules/mole': 1. / RU_JOUL,
    'kjoules/mole': 1000.0 / RU_JOUL,
    'joules/kmole': 1. / (RU_JOUL * 1000.),
}
_A_UNITS = ('moles', 'molecules')


class MechanismError(ValueError):
    pass


def _strip(line: str) -> str:
    """Drop a trailing ``!`` comment and surrounding blanks."""
    line = line.strip()
    k = line.find('!')
    if k > 0:
        line = line[:k]
    return line


def _pull_falloff_collider(side: str) -> Tuple[str, Optional[str]]:
    """Find a ``(+M)`` / ``(+SP)`` marker on one side of a reaction string.

    Returns the side with the marker removed and the collider text (``'M'``, a
    species name) or ``None``.  Parentheses that are part of a species name, and
    the ``(+)`` charge idiom, are left alone (mech_interpret.py:238-272).
    """
    pos = 0
    while True:
        a = side.find('(', pos)
        if a < 0:
            return side, None
        b = side.find(')', a)
        if b < 0:
            return side, None
        inner = side[a + 1:b].strip()
        if inner != '+' and inner.startswith('+'):
            return side[:a] + side[b + 1:], inner.replace('+', ' ').strip()
        pos = b + 1


def _split_side(side: str) -> Tuple[List[str], list, bool]:
    """Split ``'2O+M'`` into species names, coefficients and a third-body flag."""
    parts = side.split('+')
    # 'A++B' -> species name ending in '+'
    while '' in parts:
        k = parts.index('')
        if k == 0:
            raise MechanismError('cannot parse reaction side %r' % side)
        parts[k - 1] += '+'
        del parts[k]
    # re-join a species like 'X(+)' that the split tore apart
    k = 0
    while k < len(parts) - 1:
        if parts[k].endswith('(') and parts[k + 1].startswith(')'):
            parts[k] = parts[k] + '+' + parts[k + 1]
            del parts[k + 1]
        else:
            k += 1

    names: List[str] = []
    nus: list = []
    third = False
    for tok in parts:
        tok = tok.strip()
        nu = 1
        if tok[:1].isdigit():
            # the coefficient runs up to the first letter (mech_interpret.py:300-315)
            first_alpha = next((i for i, ch in enumerate(tok) if ch.isalpha()), len(tok))
            num = tok[:first_alpha]
            nu = float(num) if '.' in num else int(num)
            tok = tok[first_alpha:].strip()
        if tok.lower() == 'm':
            third = True
            continue
        if tok in names:
            nus[names.index(tok)] += nu
        else:
            names.append(tok)
            nus.append(nu)
    return names, nus, third


def _aux_numbers(line: str) -> List[str]:
    return line.replace('/', ' ').replace(',', ' ').split()


def read_mech(mech_filename: str, therm_filename: Optional[str] = None):
    """Parse a Chemkin mechanism; returns ``(elems, specs, reacs)``.

    Same return contract as the reference's ``read_mech``
    (mech_interpret.py:56-83): species names are still strings inside the
    reactions; use :class:`pyjac_b200.mechanism.Mechanism` to finalise.
    """
    elems: List[str] = []
    specs: List[Species] = []
    reacs: List[Reaction] = []
    elem_wt = dict(ELEM_WT)

    with open(mech_filename, 'r') as fh:
        raw_lines = fh.readlines()

    section = ''
    units_E = 'cal/mole'
    units_A = 'moles'
    thermo_in_mech = False
    in_cheb = False
    i = 0
    n = len(raw_lines)
    while i < n:
        raw = raw_lines[i]
        i += 1
        if not raw.strip() or raw.lstrip().startswith('!'):
            continue
        line = _strip(raw)
        head = line[:4].lower()

        if head == 'elem' or head == 'spec':
            section = head
            toks = line.split()
            if len(toks) == 1:
                continue
            line = line[line.index(toks[1]):]
        elif head == 'reac':
            section = 'reac'
            units_E, units_A = 'cal/mole', 'moles'
            for unit in line.split()[1:]:
                u = unit.lower()
                if u in _A_UNITS:
                    units_A = u
                elif u in _E_FACT:
                    units_E = u
                else:
                    raise MechanismError('unsupported units %r on REACTIONS line' % unit)
            if units_A == 'molecules':
                raise NotImplementedError('MOLECULES units are not supported')
            continue
        elif head == 'ther':
            # thermo block embedded in the mechanism file: consume it here
            thermo_in_mech = True
            i = _read_thermo_lines(raw_lines, i - 1, specs, elem_wt)
            section = ''
            continue
        elif line[:3].lower() == 'end':
            section = ''
            continue

        if section == 'elem':
            last = ''
            for tok in line.replace('/', ' ').split():
                if tok.isalpha():
                    if tok[:3] == 'end':
                        continue
                    if tok not in elems:
                        elems.append(tok)
                    last = tok
                else:
                    elem_wt[last.lower()] = float(tok)
        elif section == 'spec':
            for tok in line.split():
                if tok[:3] == 'end':
                    continue
                if not any(sp.name == tok for sp in specs):
                    specs.append(Species(tok))
        elif section == 'reac':
            if '=' in line:
                in_cheb = False
                reacs.append(_parse_reaction_line(line, units_A, units_E))
            else:
                _parse_aux_line(line, reacs[-1], units_A, units_E)

    # Chebyshev coefficient count, units of the first coefficient, (n_T, n_P) shape (:664-680)
    for idx, rx in enumerate(reacs):
        if rx.cheb:
            n_t, n_p = rx.cheb_n_temp, rx.cheb_n_pres
            if len(rx.cheb_par) != n_t * n_p:
                raise MechanismError('incorrect number of CHEB coefficients in reaction %d' % idx)
            if not rx.cheb_plim or not rx.cheb_tlim:
                raise MechanismError('Chebyshev reaction %d without PCHEB / TCHEB limits' % idx)
            if units_A == 'moles':
                rx.cheb_par[0] += math.log10(0.001 ** (sum(rx.reac_nu) - 1.))
            rx.cheb_par = [rx.cheb_par[r * n_p:(r + 1) * n_p] for r in range(n_t)]
        if rx.plog and len(rx.plog_par) < 2:
            raise MechanismError('PLOG reaction %d needs at least two pressures' % idx)

    # species named in reactions must exist (mech_interpret.py:682-691)
    known = set(sp.name for sp in specs)
    for idx, rx in enumerate(reacs):
        for nm in rx.reac + rx.prod:
            if nm not in known:
                raise MechanismError('reaction %d contains unknown species %s' % (idx, nm))

    # explicit reverse parameters -> two irreversible reactions (:693-713)
    out: List[Reaction] = []
    for rx in reacs:
        if rx.rev_par:
            back = copy.deepcopy(rx)
            rx.rev = False
            back.A, back.b, back.E = rx.rev_par
            back.rev = False
            back.reac, back.reac_nu = rx.prod[:], rx.prod_nu[:]
            back.prod, back.prod_nu = rx.reac[:], rx.reac_nu[:]
            rx.rev_par = []
            back.rev_par = []
            out.extend([rx, back])
        else:
            out.append(rx)
    reacs = out

    if any(not sp.mw for sp in specs):
        if therm_filename is None:
            raise MechanismError('species without thermo data and no thermo file given')
        with open(therm_filename, 'r') as fh:
            _read_thermo_lines(fh.readlines(), 0, specs, elem_wt)
    missing = [sp.name for sp in specs if not sp.mw]
    if missing:
        raise MechanismError('missing thermo data for ' + ', '.join(missing))
    return elems, specs, reacs


def _parse_reaction_line(line: str, units_A: str, units_E: str) -> Reaction:
    toks = line.split()
    try:
        A, b, E = float(toks[-3]), float(toks[-2]), float(toks[-1])
    except (ValueError, IndexError):
        raise MechanismError('cannot read Arrhenius parameters from %r' % line)
    eqn = line[:line.index(toks[-3])].strip()

    if '<=>' in eqn:
        lhs, rhs = eqn.split('<=>', 1)
        rev = True
    elif '=>' in eqn:
        lhs, rhs = eqn.split('=>', 1)
        rev = False
    else:
        lhs, rhs = eqn.split('=', 1)
        rev = True
    lhs, rhs = lhs.strip(), rhs.strip()

    lhs, col_l = _pull_falloff_collider(lhs)
    rhs, col_r = _pull_falloff_collider(rhs)
    collider = col_r if col_r is not None else col_l
    pdep = collider is not None
    pdep_sp = ''
    thd = False
    if pdep and collider.lower() != 'm':
        pdep_sp = collider

    r_names, r_nu, thd_l = _split_side(lhs)
    p_names, p_nu, thd_r = _split_side(rhs)
    thd = (thd_l or thd_r) and not pdep

    E *= _E_FACT[units_E]
    if units_A == 'moles':
        order = sum(r_nu)
        if thd:
            A /= 1000. ** order
        else:
            A /= 1000. ** (order - 1.)

    rx = Reaction(rev, r_names, r_nu, p_names, p_nu, A, b, E)
    rx.thd_body = thd
    rx.pdep = pdep
    if pdep:
        rx.pdep_sp = pdep_sp
    return rx


def _parse_aux_line(line: str, rx: Reaction, units_A: str, units_E: str) -> None:
    key = line[:3].lower()
    if key == 'dup':
        rx.dup = True
    elif key == 'rev':
        t = _aux_numbers(line)
        A, b, E = float(t[1]), float(t[2]), float(t[3])
        E *= _E_FACT[units_E]
        if units_A == 'moles':
            order = sum(rx.prod_nu)
            if rx.thd_body:
                A /= 1000. ** order
            else:
                A /= 1000. ** (order - 1.)
        if A != 0.0:
            rx.rev_par = [A, b, E]
        else:
            rx.rev = False
    elif key == 'low':
        t = _aux_numbers(line)
        A, b, E = float(t[1]), float(t[2]), float(t[3])
        E *= _E_FACT[units_E]
        if units_A == 'moles':
            A /= 1000. ** sum(rx.reac_nu)
        rx.low = [A, b, E]
    elif key == 'hig':
        t = _aux_numbers(line)
        A, b, E = float(t[1]), float(t[2]), float(t[3])
        E *= _E_FACT[units_E]
        if units_A == 'moles':
            A /= 1000. ** (sum(rx.reac_nu) - 2.)
        rx.high = [A, b, E]
    elif key == 'tro':
        t = _aux_numbers(line)
        a, T3, T1 = float(t[1]), float(t[2]), float(t[3])
        if T3 == 0 or T1 == 0:
            logging.warning('Troe parameters modified to avoid division by zero')
        T3 = 1e-30 if T3 == 0 else T3
        T1 = 1e-30 if T1 == 0 else T1
        rx.troe = True
        rx.troe_par = [a, T3, T1]
        if len(t) > 4:
            rx.troe_par.append(float(t[4]))
    elif key == 'sri':
        t = _aux_numbers(line)
        rx.sri = True
        rx.sri_par = [float(t[1]), float(t[2]), float(t[3])]
        if len(t) > 4:
            rx.sri_par += [float(t[4]), float(t[5])]
    elif key == 'che':
        # CHEB / n_T n_P c.. / on the first line, further coefficients on later CHEB lines
        # (mech_interpret.py:589-606); a Chebyshev reaction is not a fall-off reaction
        t = line.replace('/', ' ').split()
        if not rx.cheb:
            rx.cheb = True
            rx.pdep = False
            rx.cheb_n_temp, rx.cheb_n_pres = int(t[1]), int(t[2])
            rx.cheb_par = [float(v) for v in t[3:]]
        else:
            rx.cheb_par += [float(v) for v in t[1:]]
    elif key == 'pch':
        t = line.replace('/', ' ').split()                                    # :607-620
        rx.cheb_plim = [float(t[1]) * PA, float(t[2]) * PA]
        if len(t) > 3 and t[3].lower() == 'tcheb':
            rx.cheb_tlim = [float(t[4]), float(t[5])]
    elif key == 'tch':
        t = line.replace('/', ' ').split()                                    # :621-631
        rx.cheb_tlim = [float(t[1]), float(t[2])]
        if len(t) > 3 and t[3].lower() == 'pcheb':
            rx.cheb_plim = [float(t[4]) * PA, float(t[5]) * PA]
    elif key == 'plo':
        # PLOG / P(atm) A b E /, one line per pressure (mech_interpret.py:632-654)
        t = line.replace('/', ' ').split()
        if not rx.plog:
            rx.plog = True
            rx.pdep = False
            rx.plog_par = []
        pars = [float(v) for v in t[1:5]]
        pars[0] *= 101325.0
        pars[3] *= _E_FACT[units_E]
        if units_A == 'moles':
            pars[1] /= 1000. ** (sum(rx.reac_nu) - 1.)
        rx.plog_par.append(pars)
    else:
        t = line.replace('/', ' ').split()
        if len(t) % 2:
            raise MechanismError('cannot parse third-body efficiencies %r' % line)
        for k in range(0, len(t), 2):
            rx.thd_body_eff.append([t[k], float(t[k + 1])])


def _chunks(s: str, width: int) -> List[str]:
    return [s[k:k + width] for k in range(0, len(s), width)]


def _read_thermo_lines(lines: List[str], start: int, specs: List[Species], elem_wt) -> int:
    """Read NASA-7 4-line records beginning at the THERMO keyword line ``start``.

    Column layout per mech_interpret.py:794-878.  Returns the index of the first
    line after the block (after ``END`` or once every species has data).
    """
    i = start
    n = len(lines)
    while i < n:
        ln = lines[i]
        i += 1
        if not ln.strip() or ln.lstrip().startswith('!'):
            continue
        if 'thermo' in ln.lower():
            break
    # optional global temperature ranges
    T_ranges = None
    while i < n and (not lines[i].strip() or lines[i].lstrip().startswith('!')):
        i += 1
    if i < n and lines[i].split()[0][:1].isdigit():
        T_ranges = [float(x) for x in lines[i].split()]
        i += 1

    by_name = {sp.name: sp for sp in specs}
    while i < n:
        ln = lines[i]
        i += 1
        if ln[:3].lower() == 'end':
            break
        if not ln.strip() or ln.lstrip().startswith('!'):
            continue
        name = ln[:18].strip()
        if ' ' in name:
            name = name[:name.find(' ')]
        sp = by_name.get(name)
        if sp is None or sp.mw:
            i += 3
            continue
        for item in _chunks(ln[24:44], 5):
            e = item[:2].strip()
            if e == '' or e == '0':
                continue
            cnt = int(float(item[2:].strip()))
            sp.elem.append((e, cnt))
            sp.mw += cnt * elem_wt[e.lower()]
        T_spec = [float(x) for x in ln[45:74].split()]
        T_com = T_spec[2] if len(T_spec) == 3 else T_ranges[1]
        sp.Trange = [T_spec[0], T_com, T_spec[1]]
        c1 = _chunks(lines[i][:75], 15)
        c2 = _chunks(lines[i + 1][:75], 15)
        c3 = _chunks(lines[i + 2][:75], 15)
        i += 3
        sp.hi = [float(c1[0]), float(c1[1]), float(c1[2]), float(c1[3]), float(c1[4]),
                 float(c2[0]), float(c2[1])]
        sp.lo = [float(c2[2]), float(c2[3]), float(c2[4]),
                 float(c3[0]), float(c3[1]), float(c3[2]), float(c3[3])]
        if all(s.mw != 0.0 for s in specs):
            # every species has data: skip the rest of the block
            while i < n:
                head = lines[i].strip()[:4].lower()
                if head[:3] == 'end':
                    i += 1
                    break
                if head == 'reac':
                    break
                i += 1
            break
    return i
