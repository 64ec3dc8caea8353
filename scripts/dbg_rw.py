import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from pyiid_b200 import ElasticScatter, structures
from pyiid_b200.calc import _potential, _contract
a1, a2 = structures.atomic_square()
scat = ElasticScatter(precision='fp64')
target = scat.get_pdf(a1)
gc = scat.get_pdf(a2)
gp = scat.get_grad_pdf(a2)
ogp = oracle.experiment_grad_pdf(a2.get_positions(), a2.get_array('PDF scatter'), oracle.DEFAULT_EXP, 'fp64')
print('grad_pdf err', np.abs(gp-ogp).max()/np.abs(ogp).max())
for pot in (0, 1):
    v, s, c = _potential(gc, target, pot, True)
    print('pot', pot, 'value', v, 'scale', s, 'c[:3]', c[:3], 'c norm', np.linalg.norm(c))
    f = _contract(gp, c)
    of = (oracle.wrap_grad_rw if pot == 0 else oracle.wrap_grad_chi_sq)(ogp, gc, target)
    print(' f', f[0], 'of', of[0], 'ratio', f[0]/of[0])
    print(' oracle rw', oracle.wrap_rw(gc, target), oracle.get_scale(target, gc))
print('--- more')
from pyiid_b200.backend import pdf_matrix
g = scat.grad(a2, scat.pdf_qbin, 'PDF')
og = oracle.wrap_fq_grad(a2.get_positions(), a2.get_array('PDF scatter'), scat.pdf_qbin, 'fp64')
print('grad on pdf grid err', np.abs(g-og).max()/np.abs(og).max(), g.shape, g.dtype)
T = pdf_matrix(330, .01, scat.pdf_qbin, scat.get_r(), 0.0)
ref = g.reshape(12, 330).dot(T.T)
be = scat.backend
mine = be.grad_pdf(g)
print('grad_pdf vs numpy', np.abs(mine.reshape(12,-1)-ref).max()/np.abs(ref).max())
print('numpy vs oracle', np.abs(ref.reshape(4,3,-1)-ogp).max()/np.abs(ogp).max())
rnd = np.random.RandomState(0).normal(size=(100, 3, 330))
mine = be.grad_pdf(rnd); ref = rnd.reshape(300, 330).dot(T.T)
print('random grad_pdf vs numpy', np.abs(mine.reshape(300,-1)-ref).max()/np.abs(ref).max())
