"""Full-size (bench workload) consistency: FP32 mode against FP64 mode on the
same float32-rounded positions, Pt 50 000 atoms, F(Q) and the full gradient."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
atoms = structures.fcc_sphere('Pt', n)
pos32 = atoms.get_positions().astype(np.float32).astype(np.float64)
out = {}
for prec in ('fp32', 'fp64'):
    scat = ElasticScatter(precision=prec)
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    t = time.perf_counter()
    g, f = be.grad_fq(pos32, with_fq=True)
    print(prec, 'grad+F wall s %.3f' % (time.perf_counter() - t), flush=True)
    out[prec] = (g, f)
g32, f32 = out['fp32']; g64, f64 = out['fp64']
nf = np.abs(f32 - f64).max() / np.abs(f64).max()
ng = np.abs(g32 - g64).max() / np.abs(g64).max()
rowsum = np.abs(g32.sum(axis=0, dtype=np.float64)).max() / np.abs(g32).max()
print('N = %d: F(Q) fp32 vs fp64 normalised max err %.3e; grad %.3e; |sum_i grad| / max|grad| %.3e' % (n, nf, ng, rowsum))
