"""Form factors for every element (reference: xraylib.FF_Rayl for any Z,
master_kernel.py:14-36).  The embedded Waasmaier-Kirfel table is checked for
internal consistency; only carbon has a golden vector in the reference
(c60_scat.txt, tests/test_host_logic.py)."""
import numpy as np
import pytest

from pyiid_b200 import ElasticScatter, formfactors, structures, ase_shim


def test_table_covers_z_1_to_98_and_sums_to_z():
    assert sorted(formfactors.WK95) == list(range(1, 99))
    for z, (a, b, c) in formfactors.WK95.items():
        assert len(a) == 5 and len(b) == 5
        assert abs(sum(a) + c - z) < 0.045, z   # f(0) = number of electrons


def test_every_row_is_positive_monotone_and_smooth_in_z():
    q = np.linspace(0., 25., 251)            # the default experiment's range
    f = {z: formfactors.form_factor(z, q) for z in formfactors.WK95}
    for z, fz in f.items():
        assert fz.min() > 0 and np.all(np.diff(fz) <= 1e-9), z
    # a wrong width b_i shows up as a kink of f_Z(Q)/Z against Z.  Beyond the
    # second period the second difference in Z stays small; at low Q the largest
    # ones are the physical 4s1 anomalies of Cr (24) and Cu (29)
    for k, bound in ((25, 0.03), (50, 0.02), (100, 0.01), (200, 0.01)):
        v = np.array([f[z][k] / z for z in range(1, 99)])
        d2 = np.abs(v[2:] - 2 * v[1:-1] + v[:-2])
        assert d2[10:].max() < bound, (k, int(d2[10:].argmax()) + 12)
    v = np.array([f[z][25] / z for z in range(1, 99)])
    d2 = np.abs(v[2:] - 2 * v[1:-1] + v[:-2])
    assert set(np.argsort(d2[10:])[-2:] + 12) == {24, 29}


def test_any_element_can_be_wrapped():
    """A Pd/Ag/Ni/Fe particle gets its scatter-factor arrays (the reference
    accepts any Z)."""
    Atoms = ase_shim.Atoms
    pos = structures.fcc_sphere_positions(40, 3.9)
    numbers = np.array([46, 47, 28, 26] * 10)
    atoms = Atoms(numbers=numbers, positions=pos)
    scat = ElasticScatter.__new__(ElasticScatter)   # no device needed for the wrap
    scat.exp = {'qmin': 0., 'qmax': 25., 'qbin': .1, 'rmin': 0., 'rmax': 40., 'rstep': .01,
                'sampling': 'full'}
    scat.pdf_qbin = np.pi / (40. + 6 * 2 * np.pi / 25.)
    scat._wrap_atoms(atoms)
    sf = atoms.get_array('F(Q) scatter')
    assert sf.shape == (40, 250) and sf.dtype == np.float32
    for z in (46, 47, 28, 26):
        rows = sf[numbers == z]
        assert np.all(rows == rows[0]) and abs(rows[0][0] - z) < 0.05
    assert formfactors.source(46) in ('table', 'xraylib')


def test_user_tables_and_functions_take_precedence(tmp_path):
    path = tmp_path / 'f0.txt'
    path.write_text('# test\n99 Es 50 20 15 10 3 1  1 2 3 4 5\n')
    assert formfactors.source(99) is None
    formfactors.load_table(str(path))
    try:
        assert abs(formfactors.form_factor(99, np.zeros(1))[0] - 99) < 1e-12
        formfactors.register_form_factor(99, lambda q: np.full(q.shape, 7.0))
        assert formfactors.form_factor(99, np.arange(3.))[1] == 7.0
        assert formfactors.source(99) == 'registered'
    finally:
        formfactors._custom.pop(99, None)
        formfactors.WK95.pop(99, None)
    with pytest.raises(ValueError):
        bad = tmp_path / 'bad.txt'
        bad.write_text('1 H 1 2 3\n')
        formfactors.read_table(str(bad))
