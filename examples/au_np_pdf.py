"""The reference's example workflow (examples/Au_NP_PDF.py) on the B200 path,
written against the reference's import paths (resolved by the `pyiid` alias
package of this repository).

    python examples/au_np_pdf.py [n_nuts_iterations]
"""
import os
import sys
from copy import deepcopy as dc

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyiid.experiments.elasticscatter import ElasticScatter  # noqa: E402
from pyiid.calc.calc_1d import Calc1D  # noqa: E402
from pyiid.sim.nuts_hmc import NUTSCanonicalEnsemble  # noqa: E402
from pyiid_b200 import structures  # noqa: E402


def main(iterations=5):
    # the reference builds ase.cluster.Octahedron('Au', 2); ASE is optional
    # here, so take the 55-atom Mackay icosahedron instead
    atoms = structures.icosahedron('Au', 2)
    scat = ElasticScatter()
    pdf = scat.get_pdf(atoms)
    # dilate the atoms so that they do not match the PDF
    atoms2 = dc(atoms)
    atoms2.positions *= 1.05
    calc = Calc1D(target_data=pdf, exp_function=scat.get_pdf,
                  exp_grad_function=scat.get_grad_pdf, conv=100, potential='rw')
    atoms2.set_calculator(calc)
    e0 = atoms2.get_potential_energy()
    print('start: Rw*100 =', e0, ' max |force| =', np.abs(atoms2.get_forces()).max())
    np.random.seed(0)
    ensemble = NUTSCanonicalEnsemble(atoms2, temperature=1000, verbose=False,
                                     escape_level=8, seed=0)
    traj, metadata = ensemble.run(iterations)
    pe = [a.get_potential_energy() for a in traj]
    print('NUTS:', metadata, ' lowest Rw*100 =', min(pe))
    return e0, pe, metadata


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
