"""A finalised mechanism: species permuted so the bath gas is last, reactions indexed.

Reproduces the bookkeeping of the reference's ``create_jacobian`` orchestrator
(pyjac/core/create_jacobian.py:3503-3593): choice of the last species (user value,
else N2 -> Ar -> He, else the final species), the move-to-end permutation
(utils.get_species_mappings, utils.py:55-91) and the names -> indices rewrite
(utils.reassign_species_lists, utils.py:250-277).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

from .chem import ELEM_WT, Reaction, Species
from . import mech_interpret


def species_mappings(num_specs: int, last_species: int):
    """``fwd[i]`` = original index now at position i; ``back[o]`` = new position of
    original index o (utils.py:55-91)."""
    fwd = [i for i in range(num_specs) if i != last_species] + [last_species]
    back = [0] * num_specs
    for new, old in enumerate(fwd):
        back[old] = new
    return fwd, back


def pick_last_species(specs: List[Species], last_spec: Optional[str]) -> int:
    """create_jacobian.py:3503-3542."""
    if last_spec is not None:
        want = last_spec.lower().strip()
        for i, sp in enumerate(specs):
            if sp.name.lower() == want:
                return i
    for name, mw in (('n2', ELEM_WT['n'] * 2.), ('ar', ELEM_WT['ar']), ('he', ELEM_WT['he'])):
        for i, sp in enumerate(specs):
            if sp.name.lower() == name and sp.mw == mw:
                return i
    return len(specs) - 1


@dataclass
class Mechanism:
    elems: List[str]
    specs: List[Species]          # internal (moved-last) order
    reacs: List[Reaction]         # species as indices into ``specs``
    fwd_spec_map: List[int]       # internal position -> original index
    back_spec_map: List[int]      # original index -> internal position
    last_spec_original: int

    @classmethod
    def from_chemkin(cls, mech_name: str, therm_name: Optional[str] = None,
                     last_spec: Optional[str] = None) -> 'Mechanism':
        elems, specs, reacs = mech_interpret.read_mech(mech_name, therm_name)
        if not specs:
            raise mech_interpret.MechanismError('no species found in %s' % mech_name)
        if not reacs:
            raise mech_interpret.MechanismError('no reactions found in %s' % mech_name)
        return cls.finalize(elems, specs, reacs, last_spec)

    @classmethod
    def from_cti(cls, mech_name: str, last_spec: Optional[str] = None) -> 'Mechanism':
        """A Cantera ``.cti`` file, read without Cantera (:mod:`pyjac_b200.cti_interpret`; the reference needs
        ``cantera.Solution`` for this, mech_interpret.py:886-1137)."""
        from . import cti_interpret
        elems, specs, reacs = cti_interpret.read_mech_cti(mech_name)
        if not specs or not reacs:
            raise mech_interpret.MechanismError('no species / reactions found in %s' % mech_name)
        return cls.finalize(elems, specs, reacs, last_spec)

    @classmethod
    def from_yaml(cls, mech_name: str, last_spec: Optional[str] = None) -> 'Mechanism':
        """A Cantera YAML file, read without Cantera (:mod:`pyjac_b200.yaml_interpret`)."""
        from . import yaml_interpret
        elems, specs, reacs = yaml_interpret.read_mech_yaml(mech_name)
        if not specs or not reacs:
            raise mech_interpret.MechanismError('no species / reactions found in %s' % mech_name)
        return cls.finalize(elems, specs, reacs, last_spec)

    @classmethod
    def from_file(cls, mech_name: str, therm_name: Optional[str] = None, last_spec: Optional[str] = None) -> 'Mechanism':
        """By extension, as the reference's create_jacobian does (create_jacobian.py:3476-3489): ``.cti`` / ``.yaml`` ->
        the Cantera-format readers, anything else -> Chemkin."""
        if mech_name.lower().endswith('.cti'):
            return cls.from_cti(mech_name, last_spec)
        if mech_name.lower().endswith(('.yaml', '.yml')):
            return cls.from_yaml(mech_name, last_spec)
        if mech_name.lower().endswith('.xml'):
            raise mech_interpret.MechanismError('Cantera .xml (ctml) files are not read: convert to .cti or .yaml')
        return cls.from_chemkin(mech_name, therm_name, last_spec)

    @classmethod
    def finalize(cls, elems, specs, reacs, last_spec: Optional[str] = None) -> 'Mechanism':
        last = pick_last_species(specs, last_spec)
        fwd, back = species_mappings(len(specs), last)
        specs = [specs[o] for o in fwd]
        index = {sp.name: i for i, sp in enumerate(specs)}
        for rx in reacs:
            rx.reac = [index[nm] for nm in rx.reac]
            rx.prod = [index[nm] for nm in rx.prod]
            rx.thd_body_eff = [(index[nm], alpha) for nm, alpha in rx.thd_body_eff]
            rx.pdep_sp = index[rx.pdep_sp] if rx.pdep_sp != '' else None
        return cls(elems, specs, reacs, fwd, back, last)

    # --- sizes, named as the reference's mechanism.h macros (mech_auxiliary.py:109-176)
    @property
    def NSP(self) -> int:
        return len(self.specs)

    @property
    def NN(self) -> int:
        return len(self.specs) + 1

    @property
    def FWD_RATES(self) -> int:
        return len(self.reacs)

    @property
    def rev_reacs(self) -> List[int]:
        return [i for i, rx in enumerate(self.reacs) if rx.rev]

    @property
    def pdep_reacs(self) -> List[int]:
        return [i for i, rx in enumerate(self.reacs) if rx.thd_body or rx.pdep]

    @property
    def REV_RATES(self) -> int:
        return len(self.rev_reacs)

    @property
    def PRES_MOD_RATES(self) -> int:
        return len(self.pdep_reacs)

    def mechanism_header(self) -> str:
        """Text of a ``mechanism.h`` carrying the macros and ``//last_spec`` comment the
        reference's harnesses regex for (functional_tester/test.py:311-318,358;
        libgen.py:385)."""
        lines = ['#ifndef MECHANISM_h', '#define MECHANISM_h', '',
                 '//last_spec %d' % self.last_spec_original,
                 '/* Species Indexes']
        for i, sp in enumerate(self.specs):
            lines.append('%d  %s' % (i, sp.name))
        lines += ['*/', '',
                  '#define NSP %d' % self.NSP,
                  '#define NN %d' % self.NN,
                  '#define FWD_RATES %d' % self.FWD_RATES,
                  '#define REV_RATES %d' % self.REV_RATES,
                  '#define PRES_MOD_RATES %d' % self.PRES_MOD_RATES,
                  '', '#endif', '']
        return '\n'.join(lines)
