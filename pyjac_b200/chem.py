"""Constants and plain data records for a gas-phase kinetic mechanism.

Mirrors the *schema* of the reference's ``SpecInfo`` / ``ReacInfo``
(pyjac/core/chem_utilities.py:102-254) so that a mechanism parsed here carries the
same numbers the reference generator sees: kmol-m-s units, activation *temperature*
in K, NASA-7 ``lo``/``hi`` coefficient blocks, third-body efficiencies.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

# universal gas constant, J/(kmol K)           -- chem_utilities.py:16
RU = 8314.4621
# J/(mol K)                                    -- chem_utilities.py:17
RU_JOUL = 8.3144621
# one standard atmosphere, Pa                  -- chem_utilities.py:24
PA = 101325.0

# atomic weights, kg/kmol                      -- chem_utilities.py:51-99
ELEM_WT = {
    'h': 1.00794, 'he': 4.00260, 'li': 6.93900, 'be': 9.01220, 'b': 10.81100,
    'c': 12.0110, 'n': 14.00674, 'o': 15.99940, 'f': 18.99840, 'ne': 20.18300,
    'na': 22.98980, 'mg': 24.31200, 'al': 26.98150, 'si': 28.08600,
    'p': 30.97380, 's': 32.06400, 'cl': 35.45300, 'ar': 39.94800,
    'k': 39.10200, 'ca': 40.08000, 'sc': 44.95600, 'ti': 47.90000,
    'v': 50.94200, 'cr': 51.99600, 'mn': 54.93800, 'fe': 55.84700,
    'co': 58.93320, 'ni': 58.71000, 'cu': 63.54000, 'zn': 65.37000,
    'ga': 69.72000, 'ge': 72.59000, 'as': 74.92160, 'se': 78.96000,
    'br': 79.90090, 'kr': 83.80000, 'rb': 85.47000, 'sr': 87.62000,
    'y': 88.90500, 'zr': 91.22000, 'nb': 92.90600, 'mo': 95.94000,
    'tc': 99.00000, 'ru': 101.07000, 'rh': 102.90500, 'pd': 106.40000,
    'ag': 107.87000, 'cd': 112.40000, 'in': 114.82000, 'sn': 118.69000,
    'sb': 121.75000, 'te': 127.60000, 'i': 126.90440, 'xe': 131.30000,
    'cs': 132.90500, 'ba': 137.34000, 'la': 138.91000, 'ce': 140.12000,
    'pr': 140.90700, 'nd': 144.24000, 'pm': 145.00000, 'sm': 150.35000,
    'eu': 151.96000, 'gd': 157.25000, 'tb': 158.92400, 'dy': 162.50000,
    'ho': 164.93000, 'er': 167.26000, 'tm': 168.93400, 'yb': 173.04000,
    'lu': 174.99700, 'hf': 178.49000, 'ta': 180.94800, 'w': 183.85000,
    're': 186.20000, 'os': 190.20000, 'ir': 192.20000, 'pt': 195.09000,
    'au': 196.96700, 'hg': 200.59000, 'tl': 204.37000, 'pb': 207.19000,
    'bi': 208.98000, 'po': 210.00000, 'at': 210.00000, 'rn': 222.00000,
    'fr': 223.00000, 'ra': 226.00000, 'ac': 227.00000, 'th': 232.03800,
    'pa': 231.00000, 'u': 238.03000, 'np': 237.00000, 'pu': 242.00000,
    'am': 243.00000, 'cm': 247.00000, 'bk': 249.00000, 'cf': 251.00000,
    'es': 254.00000, 'fm': 253.00000, 'd': 2.01410, 'e': 5.48578e-4,
}


@dataclass
class Species:
    """One species: name, composition, molecular weight [kg/kmol], NASA-7 fits.

    ``lo``/``hi`` are the 7 coefficients below/above ``T_mid`` (``Trange[1]``);
    same field meaning as SpecInfo (chem_utilities.py:219-254).
    """
    name: str
    elem: List[Tuple[str, int]] = field(default_factory=list)
    mw: float = 0.0
    lo: List[float] = field(default_factory=lambda: [0.0] * 7)
    hi: List[float] = field(default_factory=lambda: [0.0] * 7)
    Trange: List[float] = field(default_factory=lambda: [300.0, 1000.0, 5000.0])


@dataclass
class Reaction:
    """One reaction in kmol-m-s units; ``E`` is an activation temperature [K].

    ``reac``/``prod`` hold species *names* straight out of the parser and species
    *indices* after :func:`pyjac_b200.mechanism.Mechanism.finalize` (the reference does
    the same swap in utils.reassign_species_lists, utils.py:250-277).
    """
    rev: bool
    reac: list
    reac_nu: list
    prod: list
    prod_nu: list
    A: float
    b: float
    E: float
    dup: bool = False
    thd_body: bool = False
    thd_body_eff: list = field(default_factory=list)   # [(species, alpha)]
    pdep: bool = False
    pdep_sp: object = ''                               # '' / name -> None / index
    low: List[float] = field(default_factory=list)
    high: List[float] = field(default_factory=list)
    troe: bool = False
    troe_par: List[float] = field(default_factory=list)
    sri: bool = False
    sri_par: List[float] = field(default_factory=list)
    rev_par: List[float] = field(default_factory=list)
    plog: bool = False
    plog_par: list = field(default_factory=list)       # [[P (Pa), A, b, E (K)], ...] in file order
    cheb: bool = False
    cheb_n_temp: int = 0
    cheb_n_pres: int = 0
    cheb_par: list = field(default_factory=list)       # flat while parsing, (n_temp, n_pres) array after
    cheb_plim: List[float] = field(default_factory=list)   # [Pmin, Pmax] (Pa)
    cheb_tlim: List[float] = field(default_factory=list)   # [Tmin, Tmax] (K)

    def net_nu(self, isp) -> float:
        """Net stoichiometric coefficient of species ``isp`` (utils.get_nu,
        utils.py:94-123)."""
        in_p = isp in self.prod
        in_r = isp in self.reac
        if in_p and in_r:
            return self.prod_nu[self.prod.index(isp)] - self.reac_nu[self.reac.index(isp)]
        if in_p:
            return self.prod_nu[self.prod.index(isp)]
        if in_r:
            return -self.reac_nu[self.reac.index(isp)]
        return 0
