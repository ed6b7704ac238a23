"""Cantera YAML mechanism reader that needs no Cantera.

Current Cantera distributes mechanisms as YAML (``gri30.yaml`` ...); the reference reaches them, like ``.cti``
files, only through ``cantera.Solution`` (pyjac/core/mech_interpret.py:886-1137), which is not installable
offline.  This module reads the document with PyYAML and hands the same declarations to the conversion stage of
:mod:`pyjac_b200.cti_interpret` (units, orders, efficiencies in species order, fall-off conventions, pyJac's element
table for the molecular weights), so a mechanism gives identical tables whether it comes as Chemkin text, ``.cti`` or
YAML (tests/test_mechanism.py: tests/golden/mini.yaml against mini.inp).

Read: the top-level ``units`` mapping (length, quantity, activation-energy; time must be s), the first phase of
``phases`` (``elements``, ``species`` as names or ``all``; phases that pull species or reactions from other files are
refused), ``species`` with ``composition`` and two-range ``NASA7`` thermo, and ``reactions`` of type elementary
(default), ``three-body`` (also inferred from ``+ M`` in the equation, as Cantera 3 does), ``falloff`` and
``chemically-activated`` (``Troe`` / ``SRI`` / Lindemann), ``pressure-dependent-Arrhenius`` and ``Chebyshev``, with
``efficiencies``, ``duplicate`` and quantities written with a unit (``1 atm``, ``10 kcal/mol``); ``negative-A`` is
not needed (A < 0 is accepted as it stands).  Refused, loudly: ``default-efficiency`` other than 1, explicit reaction
``orders``, other thermo models and reaction types.
"""
from __future__ import annotations

from typing import List, Tuple

from .chem import Reaction, Species
from .cti_interpret import _Units, _convert
from .mech_interpret import MechanismError


def _quantity(val):
    """number, or 'number unit' -> (number, unit) as the .cti conversion stage takes it."""
    if isinstance(val, str):
        num, _, unit = val.strip().partition(' ')
        return (float(num), unit.strip()) if unit.strip() else float(num)
    return val


def _arrhenius(d) -> list:
    if not isinstance(d, dict) or 'A' not in d:
        raise MechanismError('rate constant %r is not an {A, b, Ea} mapping' % (d,))
    return [_quantity(d['A']), float(d.get('b', 0.0)), _quantity(d.get('Ea', 0.0))]


def read_mech_yaml(filename: str) -> Tuple[List[str], List[Species], List[Reaction]]:
    """(elements, species, reactions) of a Cantera YAML file, records as :func:`mech_interpret.read_mech` returns them."""
    try:
        import yaml
    except ImportError as exc:                                 # pragma: no cover
        raise MechanismError('reading %s needs PyYAML: %s' % (filename, exc))
    with open(filename) as fh:
        try:
            doc = yaml.safe_load(fh)
        except yaml.YAMLError as exc:
            raise MechanismError('cannot read %s: %s' % (filename, exc))
    if not isinstance(doc, dict) or 'phases' not in doc:
        raise MechanismError('%s is not a Cantera YAML mechanism (no phases)' % filename)

    u = _Units()
    un = doc.get('units') or {}
    u.set(length=un.get('length'), quantity=un.get('quantity'), act_energy=un.get('activation-energy'),
          time=un.get('time', 's'))
    units = (u.length, u.quantity, u.act)

    ph = doc['phases'][0]
    if ph.get('thermo', 'ideal-gas') != 'ideal-gas':
        raise MechanismError('phase %s: only ideal-gas phases are read' % ph.get('name'))
    sp_field = ph.get('species', 'all')
    if isinstance(sp_field, list) and any(isinstance(x, dict) for x in sp_field):
        raise MechanismError('phase %s takes species from other sections / files: not supported' % ph.get('name'))
    if ph.get('reactions', 'all') not in ('all', ['all']):
        raise MechanismError('phase %s selects reactions from other sections / files: not supported' % ph.get('name'))
    phases = [{'name': ph.get('name', ''), 'elements': [str(e) for e in ph.get('elements', [])],
               'species': 'all' if sp_field in ('all', ['all']) else [str(x) for x in sp_field]}]

    species_decl = {}
    for sp in doc.get('species') or []:
        th = sp.get('thermo') or {}
        if th.get('model') != 'NASA7':
            raise MechanismError('species %s: only NASA7 thermo is read' % sp.get('name'))
        tr, data = th.get('temperature-ranges', []), th.get('data', [])
        if len(tr) != 3 or len(data) != 2 or any(len(row) != 7 for row in data):
            raise MechanismError('species %s: two NASA-7 ranges are needed' % sp.get('name'))
        species_decl[str(sp['name'])] = {
            'atoms': {str(k): v for k, v in (sp.get('composition') or {}).items()},
            'thermo': (('NASA', [float(tr[0]), float(tr[1])], [float(v) for v in data[0]]),
                       ('NASA', [float(tr[1]), float(tr[2])], [float(v) for v in data[1]]))}

    reactions_decl = []
    for i, rx in enumerate(doc.get('reactions') or []):
        eq = str(rx['equation'])
        kind = rx.get('type', 'elementary')
        if 'orders' in rx:
            raise MechanismError('reaction %d (%s): explicit reaction orders are not supported' % (i, eq))
        if float(rx.get('default-efficiency', 1.0)) != 1.0:
            raise MechanismError('reaction %d (%s): default-efficiency other than 1 is not supported' % (i, eq))
        if kind == 'elementary' and '(+' not in eq.replace(' ', '') and \
                any(tok.strip() == 'M' for side in eq.replace('<=>', '=').replace('=>', '=').split('=') for tok in side.split(' + ')):
            kind = 'three-body'                               # Cantera 3 infers the type from "+ M"
        d = dict(equation=eq, units=units, options=['duplicate'] if rx.get('duplicate') else [])
        eff = rx.get('efficiencies') or {}
        if kind == 'elementary':
            d.update(kind='elementary', kf=_arrhenius(rx['rate-constant']))
        elif kind == 'three-body':
            d.update(kind='three_body', kf=_arrhenius(rx['rate-constant']), efficiencies=eff)
        elif kind in ('falloff', 'chemically-activated'):
            fo = None
            if 'Troe' in rx:
                t = rx['Troe']
                fo = ('Troe', [float(t['A']), float(t['T3']), float(t['T1'])] + ([float(t['T2'])] if 'T2' in t else []))
            elif 'SRI' in rx:
                t = rx['SRI']
                fo = ('SRI', [float(t['A']), float(t['B']), float(t['C'])] + ([float(t['D']), float(t['E'])] if 'D' in t else []))
            hi, lo = _arrhenius(rx['high-P-rate-constant']), _arrhenius(rx['low-P-rate-constant'])
            if kind == 'falloff':
                d.update(kind='falloff', kf=hi, kf0=lo, efficiencies=eff, falloff=fo)
            else:
                d.update(kind='chem_activated', kLow=lo, kHigh=hi, efficiencies=eff, falloff=fo)
        elif kind == 'pressure-dependent-Arrhenius':
            rows = []
            for r_ in rx['rate-constants']:
                a = _arrhenius(r_)
                rows.append([_quantity(r_['P'])] + a)
            d.update(kind='plog', rates=rows)
        elif kind == 'Chebyshev':
            tr, pr = rx['temperature-range'], rx['pressure-range']
            def kelvin(v):
                q = _quantity(v)
                if isinstance(q, tuple):
                    if q[1] != 'K':
                        raise MechanismError('reaction %d (%s): temperature range in %s' % (i, eq, q[1]))
                    return float(q[0])
                return float(q)
            d.update(kind='cheb', Tmin=kelvin(tr[0]), Tmax=kelvin(tr[1]), Pmin=_quantity(pr[0]), Pmax=_quantity(pr[1]),
                     coeffs=rx['data'])
        else:
            raise MechanismError('reaction %d (%s): unsupported type %s' % (i, eq, kind))
        reactions_decl.append(d)
    return _convert(filename, phases, species_decl, reactions_decl)
