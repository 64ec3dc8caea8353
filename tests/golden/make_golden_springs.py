"""Generate tests/golden/springs.npz from the REFERENCE's spring_calc.py.

Run in the build container only (needs /root/reference):

    PYTHONPATH=. python tests/golden/make_golden_springs.py

Every stored output is what the reference's own module-level functions
(pyiid/calc/spring_calc.py: spring_nrg, spring_force, atomwise_spring_nrg and
their com_ / att_ twins) return, loaded through oracle/ref_shim.spring().  The
voxel functions are not run: they raise under the installed numpy (float array
shape, spring_calc.py:152); see oracle/spring.py.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from pyiid_b200 import structures  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# reference tests/__init__.py:216-218 plus a cut-off inside the Au55 shells
KWARGS = [dict(k=100, rt=5., sp_type='rep'), dict(k=100, rt=1., sp_type='com'),
          dict(k=100, rt=1., sp_type='att'), dict(k=10, rt=3.1, sp_type='rep'),
          dict(k=7.5, rt=2.9, sp_type='att'), dict(k=3, rt=4.25, sp_type='com')]


def main():
    if not ref_shim.available():
        raise SystemExit('reference mount not found')
    ref = ref_shim.spring()
    funcs = {'rep': (ref.spring_nrg, ref.spring_force, ref.atomwise_spring_nrg),
             'com': (ref.com_spring_nrg, ref.com_spring_force,
                     ref.atomwise_com_spring_nrg),
             'att': (ref.att_spring_nrg, ref.att_spring_force,
                     ref.atomwise_att_spring_nrg)}
    rs = np.random.RandomState(20161018)
    a1, a2 = structures.atomic_square()
    ico = structures.icosahedron('Au', 2)
    ico.set_positions(ico.get_positions() + rs.normal(0, 0.05, (55, 3)))
    alloy = structures.alloy_sphere(37, seed=3)   # Au/Pt: mass-weighted centre
    systems = {'au4_square': a1, 'au4_small': a2,
               'au10_random': structures.random_atoms(10, 1),
               'au55_ico': ico, 'aupt37_alloy': alloy}
    out = {'kwargs': np.array([(kw['sp_type'], kw['k'], kw['rt']) for kw in KWARGS],
                              dtype=object).astype(str)}
    for name, atoms in systems.items():
        out[name + '/positions'] = atoms.get_positions()
        out[name + '/com'] = atoms.get_center_of_mass()
        for i, kw in enumerate(KWARGS):
            nrg, force, atomwise = funcs[kw['sp_type']]
            with contextlib.redirect_stdout(io.StringIO()):  # the reference prints
                e = nrg(atoms, kw['k'], kw['rt'])
                f = force(atoms, kw['k'], kw['rt'])
                a = atomwise(atoms, kw['k'], kw['rt'])
            out['%s/%d/energy' % (name, i)] = np.float64(e)
            out['%s/%d/forces' % (name, i)] = np.asarray(f, np.float64)
            out['%s/%d/atomwise' % (name, i)] = np.asarray(a, np.float64)
    np.savez_compressed(os.path.join(HERE, 'springs.npz'), **out)
    print('wrote springs.npz with', len(out), 'arrays')


if __name__ == '__main__':
    main()
