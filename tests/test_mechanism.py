"""Host-side mechanism front-end: parser, last-species permutation, synthetic generator."""
import os

import numpy as np
import pytest

from pyjac_b200 import synth
from pyjac_b200.mechanism import Mechanism, species_mappings


def test_h2o2_sizes(golden_dir):
    m = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
    # NSP 10, FWD 28, REV 28, PRES_MOD 6, last species N2 (original index 9)
    assert (m.NSP, m.FWD_RATES, m.REV_RATES, m.PRES_MOD_RATES) == (10, 28, 28, 6)
    assert m.specs[-1].name == 'N2' and m.last_spec_original == 9
    hdr = m.mechanism_header()
    assert '#define NSP 10' in hdr and '//last_spec 9' in hdr and '#define NN 11' in hdr


def test_units_and_aux(golden_dir):
    m = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
    r0 = m.reacs[0]          # 2O+M<=>O2+M  1.2e17 -1 0 ; third body: A / 1000^order
    assert r0.thd_body and not r0.pdep and r0.A == 1.2e17 / 1000. ** 2
    r2 = m.reacs[2]          # O+H2<=>H+OH 3.87e4 2.7 6260 cal/mol -> K
    assert r2.E == 6260.0 * (4.184 / 8.3144621) and r2.A == 3.87e4 / 1000. ** 1.
    r20 = m.reacs[20]        # 2OH(+M)<=>H2O2(+M) Troe
    assert r20.pdep and r20.troe and r20.pdep_sp is None and len(r20.troe_par) == 4
    assert r20.low[0] == 2.3e18 / 1000. ** 2


def test_last_species_moves(golden_dir):
    m = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'), last_spec='AR')
    assert m.specs[-1].name == 'AR' and m.specs[-2].name == 'N2'
    fwd, back = species_mappings(5, 2)
    assert fwd == [0, 1, 3, 4, 2] and back == [0, 1, 4, 2, 3]


def test_explicit_rev_is_split(golden_dir):
    m = Mechanism.from_chemkin(os.path.join(golden_dir, 'torture.inp'))
    names = [sp.name for sp in m.specs]
    pairs = [(i, rx) for i, rx in enumerate(m.reacs)
             if not rx.rev and sorted(names[k] for k in rx.reac) == ['H2O', 'O']]
    assert len(pairs) == 1                      # the generated backward half of H+HO2<=>O+H2O
    assert pairs[0][1].b == 0.1
    # REV / 0 0 0 / makes the reaction irreversible without adding a partner
    h2o2 = [rx for rx in m.reacs if sorted(names[k] for k in rx.reac) == ['H2', 'O2']]
    assert len(h2o2) == 1 and not h2o2[0].rev


def test_synth_reproduces_committed_gri30(golden_dir):
    txt = synth.generate('gri30', seed=0)
    assert txt == open(os.path.join(golden_dir, 'gri30_syn.inp')).read()


@pytest.mark.parametrize('shape', ['mini', 'gri30', 'usc2'])
def test_synth_shapes(tmp_path, shape):
    p = synth.write(shape, str(tmp_path / (shape + '.inp')))
    m = Mechanism.from_chemkin(p)
    s = synth.SHAPES[shape]
    assert (m.NSP, m.FWD_RATES) == (s.nsp, s.nr)
    assert m.PRES_MOD_RATES == s.n_third + s.n_troe + s.n_lind
    assert m.specs[-1].name == 'N2'


def test_plog_and_chebyshev_lines_are_read(tmp_path, golden_dir):
    """PLOG / CHEB auxiliary data as the reference reads it (mech_interpret.py:589-680)."""
    m = Mechanism.from_chemkin(os.path.join(golden_dir, 'plog.inp'))
    rx = next(r for r in m.reacs if r.plog)
    assert not rx.pdep and len(rx.plog_par) == 3
    assert rx.plog_par[0][0] == pytest.approx(0.1 * 101325.0)
    assert rx.plog_par[1][1] == pytest.approx(3.87e4 / 1000.0)          # bimolecular: cm3/mol -> m3/kmol
    src = open(os.path.join(golden_dir, 'h2o2_n2.inp')).read()
    src = src.replace('O+H2<=>H+OH                              3.870E+04    2.700    6260.00\n',
                      'O+H2(+M)<=>H+OH(+M)                      1.000E+00     .000        .00\n'
                      ' TCHEB / 300.0 2000.0 /  PCHEB / 0.01 100.0 /\n'
                      ' CHEB / 2 3  8.0 0.5 -0.1 /\n CHEB / -1.0 0.2 0.05 /\n')
    p = tmp_path / 'cheb.inp'
    p.write_text(src)
    m = Mechanism.from_chemkin(str(p))
    rx = next(r for r in m.reacs if r.cheb)
    assert not rx.pdep and (rx.cheb_n_temp, rx.cheb_n_pres) == (2, 3)
    assert rx.cheb_tlim == [300.0, 2000.0] and rx.cheb_plim[1] == pytest.approx(100.0 * 101325.0)
    assert rx.cheb_par[0][0] == pytest.approx(8.0 - 3.0) and rx.cheb_par[1] == [-1.0, 0.2, 0.05]


def _same_mechanism(a, b, rel=0.0):
    """Field-by-field equality of two parsed mechanisms (reaction order included).  `rel` is the relative
    tolerance on real-valued fields: 0 where both files print the same decimal numbers."""
    import dataclasses

    def close(x, y, what):
        if isinstance(x, (list, tuple)):
            assert len(x) == len(y), what
            for u, v in zip(x, y):
                close(u, v, what)
        elif isinstance(x, float) or isinstance(y, float):
            assert x == y or abs(x - y) <= rel * max(abs(x), abs(y)), (what, x, y)
        elif isinstance(x, str) and what.endswith('.elem'):
            assert x.lower() == y.lower(), (what, x, y)         # 'Ar' (.cti) / 'AR' (Chemkin)
        else:
            assert x == y, (what, x, y)
    assert [s.name for s in a.specs] == [s.name for s in b.specs]
    assert len(a.reacs) == len(b.reacs)
    for sa, sb in zip(a.specs, b.specs):
        for f in dataclasses.fields(sa):
            va, vb = getattr(sa, f.name), getattr(sb, f.name)
            if f.name == 'elem':    # composition: the order of the file's element list
                va, vb = sorted([e[0].lower(), e[1]] for e in va), sorted([e[0].lower(), e[1]] for e in vb)
            close(va, vb, 'species %s.%s' % (sa.name, f.name))
    for i, (ra, rb) in enumerate(zip(a.reacs, b.reacs)):
        for f in dataclasses.fields(ra):
            va, vb = getattr(ra, f.name), getattr(rb, f.name)
            if f.name == 'thd_body_eff':
                va, vb = sorted(va), sorted(vb)
            if f.name == 'A' and (ra.plog or ra.cheb):
                continue            # not used: the .cti entry has no separate Arrhenius triple
            close(va, vb, 'reaction %d.%s' % (i, f.name))


def test_cti_reader_matches_chemkin_twin(golden_dir):
    """tests/golden/mini.cti and mini.inp state one mechanism (every reaction class the path knows) in the two
    formats: the Cantera-free reader must hand the tables the same numbers."""
    a = Mechanism.from_file(os.path.join(golden_dir, 'mini.cti'))
    b = Mechanism.from_file(os.path.join(golden_dir, 'mini.inp'))
    _same_mechanism(a, b)
    from pyjac_b200 import tables
    Ta, Tb = tables.build(a, 8, 256, False), tables.build(b, 8, 256, False)
    assert sorted(Ta) == sorted(Tb)
    for k in Ta:
        assert np.array_equal(np.asarray(Ta[k]), np.asarray(Tb[k])), k


def test_cti_reader_on_the_reference_h2o2_file(golden_dir):
    """data/h2o2.cti of the reference is the ck2cti conversion of the mechanism tests/golden/h2o2_n2.inp was taken
    from."""
    cti = '/root/reference/data/h2o2.cti'
    if not os.path.exists(cti):
        pytest.skip('reference checkout not present')
    a = Mechanism.from_file(cti)
    b = Mechanism.from_chemkin(os.path.join(golden_dir, 'h2o2_n2.inp'))
    _same_mechanism(a, b, rel=1e-15)


def test_negative_pre_exponential(golden_dir):
    from pyjac_b200 import tables
    m = Mechanism.from_chemkin(os.path.join(golden_dir, 'nega.inp'))
    assert sum(1 for rx in m.reacs if rx.A < 0) == 8          # 7 written + the split REV half
    T = tables.build(m, 8, 256, False)
    assert int(((np.asarray(T['rx_flags']) & tables.F_NEGA) != 0).sum()) == 8


def test_yaml_reader_matches_chemkin_twin(golden_dir):
    """tests/golden/mini.yaml states the mini mechanism in Cantera's YAML layout (mol-cm-s, activation temperatures):
    the Cantera-free reader must hand the tables the same numbers as the Chemkin text and the .cti file."""
    a = Mechanism.from_file(os.path.join(golden_dir, 'mini.yaml'))
    b = Mechanism.from_file(os.path.join(golden_dir, 'mini.inp'))
    c = Mechanism.from_file(os.path.join(golden_dir, 'mini.cti'))
    _same_mechanism(a, b)
    _same_mechanism(a, c)
    from pyjac_b200 import tables
    Ta, Tb = tables.build(a, 8, 256, False), tables.build(b, 8, 256, False)
    assert sorted(Ta) == sorted(Tb)
    for k in Ta:
        assert np.array_equal(np.asarray(Ta[k]), np.asarray(Tb[k])), k


def test_yaml_reader_units_and_refusals(tmp_path, golden_dir):
    from pyjac_b200.mech_interpret import MechanismError
    txt = open(os.path.join(golden_dir, 'mini.yaml')).read()
    # the same file in kcal/mol with one energy written with its own unit, and a three-body type left to inference
    alt = txt.replace('activation-energy: K', 'activation-energy: kcal/mol')
    alt = alt.replace('Ea: 3150.14}', 'Ea: 3150.14 K}').replace('  type: three-body\n', '')
    p = tmp_path / 'alt.yaml'
    p.write_text(alt)
    a = Mechanism.from_file(str(p))
    b = Mechanism.from_file(os.path.join(golden_dir, 'mini.inp'))
    assert a.reacs[0].E == b.reacs[0].E and a.reacs[1].thd_body and a.reacs[1].A == b.reacs[1].A
    assert abs(a.reacs[10].E - b.reacs[10].E * 4184.0 / 8.3144621) <= 1e-9 * abs(a.reacs[10].E)
    for bad, what in ((txt.replace('model: NASA7', 'model: Shomate', 1), 'NASA7'),
                      (txt.replace('  duplicate: true\n', '  duplicate: true\n  orders: {HO2: 1.5}\n', 1), 'orders'),
                      (txt.replace('thermo: ideal-gas', 'thermo: Redlich-Kwong'), 'ideal-gas'),
                      (txt.replace('time: s', 'time: ms'), 'time')):
        q = tmp_path / 'bad.yaml'
        q.write_text(bad)
        with pytest.raises(MechanismError, match=what):
            Mechanism.from_file(str(q))
