"""One plain `python` process, every GPU of the box: what the reference's
set_processor('Multi-GPU') does with one Python thread per device
(gpu_wrappers/gpu_wrap.py:119-156, 287-314), here the C layer's multi-device
handle (iid_create_multi).  Prints the time of get_grad_fq / get_fq / Rw energy
+ forces of the Pt 50 000-atom bench workload and the per-GPU utilisation
nvidia-smi saw meanwhile.

    python scripts/multi_gpu_one_process.py [atoms]
"""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
atoms = structures.fcc_sphere('Pt', n)
scat = ElasticScatter()
print('processor', scat.processor, flush=True)

util = []
stop = threading.Event()


def watch():
    while not stop.is_set():
        try:
            out = subprocess.run(['nvidia-smi', '--query-gpu=utilization.gpu',
                                  '--format=csv,noheader,nounits'], capture_output=True,
                                 text=True, timeout=5).stdout.split()
            util.append([int(x) for x in out])
        except Exception:
            pass
        time.sleep(0.05)


g = scat.get_grad_fq(atoms)  # structure upload, pinned output pool
for _ in range(3):           # two live results: the pool's steady state
    g2 = scat.get_grad_fq(atoms)
res = {'atoms': n, 'devices': scat.backend.devices()}
reps = 5
t = time.perf_counter()
for _ in range(reps):
    g2 = scat.get_grad_fq(atoms)
res['get_grad_fq_ms'] = 1e3 * (time.perf_counter() - t) / reps
res['bit_reproducible'] = bool(np.array_equal(g, g2))
f = scat.get_fq(atoms)
t = time.perf_counter()
for _ in range(reps):
    f = scat.get_fq(atoms)
res['get_fq_ms'] = 1e3 * (time.perf_counter() - t) / reps
# utilisation: a separate loop (polling nvidia-smi perturbs the timing above)
th = threading.Thread(target=watch, daemon=True)
th.start()
for _ in range(reps):
    g2 = scat.get_grad_fq(atoms)
stop.set()
th.join()
if util:
    res['gpu_util_max_pct'] = np.max(np.array(util), axis=0).tolist()
one = ElasticScatter(device=0)
g1 = one.get_grad_fq(atoms)
for _ in range(3):
    g3 = one.get_grad_fq(atoms)
t = time.perf_counter()
for _ in range(reps):
    g3 = one.get_grad_fq(atoms)
res['one_gpu_get_grad_fq_ms'] = 1e3 * (time.perf_counter() - t) / reps
res['nerr_vs_one_gpu'] = float(np.abs(g2 - g1).max() / np.abs(g1).max())
res['pairq_per_s'] = n * (n - 1) / 2 * g.shape[2] / (res['get_grad_fq_ms'] * 1e-3)
print(json.dumps(res))
