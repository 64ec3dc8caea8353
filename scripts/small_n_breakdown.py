"""Per-stage device time of the fused energy+forces sequence at small N
(back-to-back repetitions of each stage, CUDA events on the handle's stream)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyiid_b200 import ElasticScatter, structures
shells = int(sys.argv[1]) if len(sys.argv) > 1 else 5
scat = ElasticScatter()
ideal = structures.icosahedron('Au', shells)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', shells)
atoms.positions *= 1.05
scat._ensure_wrapped(atoms)
be = scat._load(atoms, scat.pdf_qbin, 'PDF')
be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
be.energy_forces(atoms.get_positions(), target, 'rw', 100.)
lib, h = be.lib, be.h
dev = 'cuda:%d' % be.device
K = 200
with be._on_stream():
    pos = torch.from_numpy(atoms.get_positions()).to(dev)
    S = torch.zeros(be.nq, dtype=torch.float64, device=dev)
    F = torch.zeros_like(S); wq = torch.zeros_like(S)
    G = torch.zeros(be.nr, dtype=torch.float64, device=dev)
    tg = torch.from_numpy(target).to(dev)
    o4 = torch.zeros(4, dtype=torch.float64, device=dev)
    fo = torch.zeros((be.n, 3), dtype=torch.float64, device=dev)
    stages = {
        'fq_partial (prep+memset+F kernel)': lambda: lib.iid_fq_partial(h, pos.data_ptr(), S.data_ptr(), None),
        'fq_finish': lambda: lib.iid_fq_finish(h, S.data_ptr(), F.data_ptr(), None),
        'fq_to_gr': lambda: lib.iid_fq_to_gr(h, F.data_ptr(), G.data_ptr(), None),
        'potential+wq': lambda: lib.iid_potential(h, G.data_ptr(), tg.data_ptr(), 0, 100.0, o4.data_ptr(), wq.data_ptr(), None),
        'force_partial (prep+memset+force kernel)': lambda: lib.iid_force_partial(h, pos.data_ptr(), wq.data_ptr(), fo.data_ptr(), None),
    }
    for name, fn in stages.items():
        for _ in range(5): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(K): assert fn() == 0
        e1.record(); torch.cuda.synchronize()
        print('%-45s %.1f us' % (name, e0.elapsed_time(e1) / K * 1e3))
print(be.sizes())
