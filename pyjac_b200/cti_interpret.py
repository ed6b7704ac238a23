"""Cantera ``.cti`` mechanism reader that needs no Cantera.

The reference reads ``.cti`` / ``.xml`` files only through ``cantera.Solution``
(pyjac/core/mech_interpret.py:886-1137, ``read_mech_ct``), which is not installable offline.  A ``.cti``
file is a Python script made of calls -- ``units(...)``, ``ideal_gas(...)``, ``species(...)``,
``reaction(...)``, ``three_body_reaction(...)``, ``falloff_reaction(...)`` ... -- so this module runs it
in a namespace that defines exactly those names (no builtins, nothing imported) and collects what they
are given, then converts to the records of :mod:`pyjac_b200.chem` the way ``read_mech_ct`` converts
Cantera's objects:

* kmol-m-s units and activation *temperatures* (Cantera hands pyJac SI values; here the file's
  ``units(length, quantity, act_energy)`` are converted: a pre-exponential factor by
  (length^3 / quantity)^(order - 1), order counting the third body / the low-pressure limit's collider),
* third-body efficiencies listed in the *mechanism's species order* (``handle_effiencies``,
  mech_interpret.py:957-985), a lone collider with default efficiency 0 on a fall-off reaction becomes
  the specific collider ``pdep_sp``,
* fall-off: ``kf`` is the high-pressure limit, ``kf0`` -> ``low``; chemically activated: ``kLow`` is the
  rate, ``kHigh`` -> ``high``; Troe T3 / T1 == 0 -> 1e-30; elementary reactions with A == 0 are dropped.

Deviation: molecular weights come from pyJac's own element table (chem_utilities.py:51-99, what its
Chemkin reader uses), not from Cantera's, so that the same mechanism read from ``.cti`` and from Chemkin
text gives identical tables (tests/test_mechanism.py: data/h2o2.cti against tests/golden/h2o2_n2.inp).
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Tuple

from .chem import ELEM_WT, PA, RU_JOUL, Reaction, Species
from .mech_interpret import MechanismError, _pull_falloff_collider, _split_side

_LENGTH = {'m': 1.0, 'cm': 1.0e-2, 'mm': 1.0e-3, 'km': 1.0e3}
_QUANTITY = {'kmol': 1.0, 'mol': 1.0e-3, 'mole': 1.0e-3, 'gmol': 1.0e-3, 'molec': 1.0 / 6.02214129e26}
# activation energy unit -> activation temperature [K]
_ACT = {'cal/mol': 4.184 / RU_JOUL, 'kcal/mol': 4184.0 / RU_JOUL, 'j/mol': 1.0 / RU_JOUL,
        'kj/mol': 1000.0 / RU_JOUL, 'j/kmol': 1.0 / (RU_JOUL * 1000.0), 'k': 1.0, 'ev': 11604.519}
_PRESSURE = {'pa': 1.0, 'atm': PA, 'bar': 1.0e5, 'torr': PA / 760.0, 'kpa': 1.0e3, 'mpa': 1.0e6}


class _Units:
    def __init__(self):
        self.length, self.quantity, self.act = 'm', 'kmol', 'j/kmol'

    def set(self, length=None, quantity=None, act_energy=None, time='s', mass=None, energy=None, pressure=None):
        if time != 's':
            raise MechanismError('unsupported time unit %r' % time)
        if length:
            self.length = length.lower()
        if quantity:
            self.quantity = quantity.lower()
        if act_energy:
            self.act = act_energy.lower().replace('mole', 'mol')
        for what, tab in ((self.length, _LENGTH), (self.quantity, _QUANTITY), (self.act, _ACT)):
            if what not in tab:
                raise MechanismError('unsupported unit %r' % what)

    def conc_factor(self) -> float:
        """value of one concentration unit of the file in kmol / m^3 (an exact power of ten for the decimal
        units, so that mol-cm-s gives the same 1000^(order - 1) as the Chemkin reader)"""
        p10 = {'m': 0, 'cm': -2, 'mm': -3, 'km': 3}
        q10 = {'kmol': 0, 'mol': -3, 'mole': -3, 'gmol': -3}
        if self.quantity in q10:
            return 10.0 ** (q10[self.quantity] - 3 * p10[self.length])
        return _QUANTITY[self.quantity] / _LENGTH[self.length] ** 3


def _rate(val, u: _Units, order: float) -> List[float]:
    """[A, b, E] or Arrhenius(...) of the file -> [A (kmol-m-s), b, Ta (K)]; values given as (number, 'unit')
    tuples carry their own unit."""
    if isinstance(val, dict):
        val = [val['A'], val['b'], val['E']]
    A, b, E = val
    if isinstance(E, tuple):
        Ta = float(E[0]) * _ACT[E[1].lower().replace('mole', 'mol')]
    else:
        Ta = float(E) * _ACT[u.act]
    if isinstance(A, tuple):
        raise MechanismError('pre-exponential factors with explicit units are not supported')
    return [float(A) / u.conc_factor() ** (order - 1.0), float(b), Ta]


def _pressure(val) -> float:
    if isinstance(val, tuple):
        return float(val[0]) * _PRESSURE[val[1].lower()]
    return float(val)


def _efficiencies(text) -> Dict[str, float]:
    if isinstance(text, dict):
        return {str(k): float(v) for k, v in text.items()}
    out = {}
    for tok in text.split():
        name, _, val = tok.rpartition(':')
        out[name] = float(val)
    return out


def read_mech_cti(filename: str) -> Tuple[List[str], List[Species], List[Reaction]]:
    """(elements, species, reactions) of a ``.cti`` file, records as :func:`mech_interpret.read_mech` returns
    them (species names in the reactions, kmol-m-s units, activation temperatures)."""
    u = _Units()
    phases: List[dict] = []
    species_decl: Dict[str, dict] = {}
    reactions_decl: List[dict] = []

    def units(**kw):
        u.set(**kw)

    def phase(name='', elements='', species='', reactions='all', **kw):
        phases.append({'name': name, 'elements': elements.split(), 'species': species})

    def species(name, atoms='', thermo=None, **kw):
        species_decl[name] = {'atoms': atoms, 'thermo': thermo}

    def NASA(trange, coeffs, p0=None):
        if len(coeffs) != 7:
            raise MechanismError('NASA polynomials need 7 coefficients')
        return ('NASA', [float(v) for v in trange], [float(v) for v in coeffs])

    def unsupported(what):
        def fn(*a, **kw):
            raise MechanismError('unsupported %s in a .cti file (only NASA-7 thermo and gas-phase kinetics are read)' % what)
        return fn

    def snap(kind, equation, order_extra, **kw):
        reactions_decl.append(dict(kind=kind, equation=equation, units=(u.length, u.quantity, u.act), **kw))

    def reaction(equation, kf, id='', order='', options=()):
        snap('elementary', equation, 0, kf=kf, options=options)

    def three_body_reaction(equation, kf, efficiencies='', id='', options=()):
        snap('three_body', equation, 1, kf=kf, efficiencies=efficiencies, options=options)

    def falloff_reaction(equation, kf, kf0, efficiencies='', falloff=None, id='', options=()):
        snap('falloff', equation, 0, kf=kf, kf0=kf0, efficiencies=efficiencies, falloff=falloff, options=options)

    def chemically_activated_reaction(equation, kLow, kHigh, efficiencies='', falloff=None, id='', options=()):
        snap('chem_activated', equation, 0, kLow=kLow, kHigh=kHigh, efficiencies=efficiencies, falloff=falloff, options=options)

    def pdep_arrhenius(equation, *rates, **kw):
        snap('plog', equation, 0, rates=rates, options=kw.get('options', ()))

    def chebyshev_reaction(equation, Tmin, Tmax, Pmin, Pmax, coeffs, **kw):
        snap('cheb', equation, 0, Tmin=Tmin, Tmax=Tmax, Pmin=Pmin, Pmax=Pmax, coeffs=coeffs, options=kw.get('options', ()))

    ns = {
        '__builtins__': {}, 'units': units, 'ideal_gas': phase, 'IdealGas': phase, 'species': species, 'NASA': NASA,
        'NASA9': unsupported('NASA9 thermo'), 'Shomate': unsupported('Shomate thermo'), 'const_cp': unsupported('const_cp thermo'),
        'gas_transport': lambda **kw: None, 'state': lambda **kw: None, 'OneAtm': PA, 'validate': lambda **kw: None,
        'element': lambda **kw: None, 'Arrhenius': lambda A=0.0, b=0.0, E=0.0: {'A': A, 'b': b, 'E': E},
        'Troe': lambda A=0.0, T3=0.0, T1=0.0, T2=None: ('Troe', [A, T3, T1] + ([T2] if T2 is not None else [])),
        'SRI': lambda A=0.0, B=0.0, C=0.0, D=None, E=None: ('SRI', [A, B, C] + ([D, E] if D is not None else [])),
        'Lindemann': lambda: None, 'reaction': reaction, 'three_body_reaction': three_body_reaction,
        'falloff_reaction': falloff_reaction, 'chemically_activated_reaction': chemically_activated_reaction,
        'pdep_arrhenius': pdep_arrhenius, 'chebyshev_reaction': chebyshev_reaction,
        'surface_reaction': unsupported('surface reactions'), 'edge_reaction': unsupported('edge reactions'),
        'ideal_interface': unsupported('interfaces'), 'stoichiometric_solid': unsupported('solids'),
        'True': True, 'False': False, 'None': None,
    }
    with open(filename) as fh:
        src = fh.read()
    try:
        exec(compile(src, filename, 'exec'), ns)             # a .cti file *is* a script of these calls
    except MechanismError:
        raise
    except Exception as exc:
        raise MechanismError('cannot read %s: %s' % (filename, exc))
    return _convert(filename, phases, species_decl, reactions_decl)


def _convert(filename, phases, species_decl, reactions_decl):
    """The collected declarations (of a .cti script or of a YAML document, :mod:`pyjac_b200.yaml_interpret`) ->
    (elements, species, reactions) in pyJac's records and units."""
    if not phases:
        raise MechanismError('no ideal-gas phase in %s' % filename)
    ph = phases[0]
    names = ph['species'].replace(',', ' ').split() if isinstance(ph['species'], str) else list(ph['species'])
    if names == ['all']:
        names = list(species_decl)
    elems = ph['elements']

    specs: List[Species] = []
    for nm in names:
        if nm not in species_decl:
            raise MechanismError('species %s is not declared' % nm)
        d = species_decl[nm]
        sp = Species(nm)
        atoms = d['atoms']
        toks = ['%s:%s' % kv for kv in atoms.items()] if isinstance(atoms, dict) else atoms.replace(',', ' ').split()
        for tok in toks:
            el, _, cnt = tok.partition(':')
            sp.elem.append((el, int(float(cnt))))
            if el.lower() not in ELEM_WT:
                raise MechanismError('unknown element %s' % el)
            sp.mw += ELEM_WT[el.lower()] * float(cnt)
        th = d['thermo']
        th = [th] if th and th[0] == 'NASA' else list(th or [])
        if len(th) != 2:
            raise MechanismError('species %s: two NASA-7 ranges are needed' % nm)
        th.sort(key=lambda t: t[1][0])
        (_, r_lo, lo), (_, r_hi, hi) = th
        sp.lo, sp.hi = lo, hi
        sp.Trange = [r_lo[0], r_lo[1], r_hi[1]]
        specs.append(sp)

    reacs: List[Reaction] = []
    for d in reactions_decl:
        ru = _Units()
        ru.length, ru.quantity, ru.act = d['units']
        eqn = d['equation'].replace(' ', '')
        if '<=>' in eqn:
            lhs, rhs, rev = eqn.split('<=>')[0], eqn.split('<=>', 1)[1], True
        elif '=>' in eqn:
            lhs, rhs, rev = eqn.split('=>')[0], eqn.split('=>', 1)[1], False
        else:
            lhs, rhs, rev = eqn.split('=')[0], eqn.split('=', 1)[1], True
        lhs, col_l = _pull_falloff_collider(lhs)
        rhs, col_r = _pull_falloff_collider(rhs)
        r_names, r_nu, _ = _split_side(lhs)
        p_names, p_nu, _ = _split_side(rhs)
        order = float(sum(r_nu))
        kind = d['kind']
        opts = d.get('options') or ()
        opts = [opts] if isinstance(opts, str) else list(opts)

        def third_bodies(rx: Reaction, eff_text: str, collider: Optional[str]):
            eff = _efficiencies(eff_text)
            if collider is not None and collider.lower() != 'm':
                rx.pdep_sp = collider                         # "A (+ SP) <=> ..."
                return
            for nm in names:                                  # mechanism order, like handle_effiencies
                if nm in eff:
                    rx.thd_body_eff.append([nm, eff[nm]])

        if kind == 'elementary':
            A, b, Ta = _rate(d['kf'], ru, order)
            if A == 0.0:
                continue
            rx = Reaction(rev, r_names, r_nu, p_names, p_nu, A, b, Ta)
        elif kind == 'three_body':
            A, b, Ta = _rate(d['kf'], ru, order + 1.0)
            rx = Reaction(rev, r_names, r_nu, p_names, p_nu, A, b, Ta)
            rx.thd_body = True
            third_bodies(rx, d['efficiencies'], None)
        elif kind in ('falloff', 'chem_activated'):
            collider = col_r if col_r is not None else col_l
            if kind == 'falloff':
                A, b, Ta = _rate(d['kf'], ru, order)
                rx = Reaction(rev, r_names, r_nu, p_names, p_nu, A, b, Ta)
                rx.low = _rate(d['kf0'], ru, order + 1.0)
            else:
                # chemically activated: the low-pressure limit has the reaction's own order, the
                # high-pressure limit one less (Chemkin HIGH, mech_interpret.py:522-530)
                A, b, Ta = _rate(d['kLow'], ru, order)
                rx = Reaction(rev, r_names, r_nu, p_names, p_nu, A, b, Ta)
                rx.high = _rate(d['kHigh'], ru, order - 1.0)
            rx.pdep = True
            rx.pdep_sp = ''
            third_bodies(rx, d['efficiencies'], collider)
            fo = d['falloff']
            if fo and fo[0] == 'Troe':
                par = [float(v) for v in fo[1]]
                if par[1] == 0 or par[2] == 0:
                    logging.warning('Troe parameters modified to avoid division by zero')
                par[1] = 1e-30 if par[1] == 0 else par[1]
                par[2] = 1e-30 if par[2] == 0 else par[2]
                rx.troe, rx.troe_par = True, par
            elif fo and fo[0] == 'SRI':
                rx.sri, rx.sri_par = True, [float(v) for v in fo[1]]
        elif kind == 'plog':
            rx = Reaction(rev, r_names, r_nu, p_names, p_nu, 0.0, 0.0, 0.0)
            rx.plog, rx.plog_par = True, []
            for row in d['rates']:
                A, b, Ta = _rate(list(row[1:4]), ru, order)
                rx.plog_par.append([_pressure(row[0]), A, b, Ta])
        elif kind == 'cheb':
            rx = Reaction(rev, r_names, r_nu, p_names, p_nu, 0.0, 0.0, 0.0)
            rx.cheb = True
            co = [[float(v) for v in row] for row in d['coeffs']]
            rx.cheb_n_temp, rx.cheb_n_pres = len(co), len(co[0])
            rx.cheb_tlim = [float(d['Tmin']), float(d['Tmax'])]
            rx.cheb_plim = [_pressure(d['Pmin']), _pressure(d['Pmax'])]
            # the fitted quantity is log10 k in the file's units: shift the constant term to kmol-m-s
            import math
            co[0][0] -= (order - 1.0) * math.log10(ru.conc_factor())
            rx.cheb_par = co
        else:                                                 # pragma: no cover
            raise MechanismError('unsupported reaction kind %s' % kind)
        rx.dup = any(o.lower().startswith('dup') for o in opts)
        reacs.append(rx)
    return elems, specs, reacs
