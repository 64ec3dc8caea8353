"""Host-side pieces of bench.py that need no GPU: the reference arm's fixed
CPU sample (oracle only -- it must never map the product library) and the
shared config of the two arms."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_reference_arm_pieces_load_only_the_oracle():
    code = r'''
import sys, json
sys.path.insert(0, %r)
import bench, oracle
oracle.build()
pos, sf, qbin = bench.host_workload(1500)
cs = bench.CpuSample(pos, sf, qbin, pairs=1 << 13)
tf, tg = cs.run()
tf2, tg2 = cs.run()            # same slice, same preallocated buffers
d = cs.describe(tf, tg)
maps = open('/proc/self/maps').read()
print(json.dumps({'product': 'libiid_b200' in maps, 'oracle': 'libiid_oracle' in maps,
                  'value': d['value'], 'kind': d['kind'], 'm': cs.m, 'k0': cs.k0,
                  'cfg': bench.workload_config(1) == bench.workload_config(world=1)}))
''' % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res['oracle'] and not res['product']
    assert res['kind'] == 'port' and res['value'] > 0 and res['m'] == 1 << 13 and res['cfg']
    assert res['k0'] % 2048 == 0


def test_cpu_sample_equals_the_full_pair_sum_on_a_small_case():
    """The slice the sample times is the reference's pair range arithmetic:
    two half slices add up to the whole F(Q) pair sum."""
    sys.path.insert(0, ROOT)
    import oracle
    rs = np.random.RandomState(0)
    pos = rs.rand(60, 3) * 12
    sf = np.full((60, 40), 7.5, np.float32)
    k = 60 * 59 // 2
    whole = oracle.fq_pairsum(pos, sf, 0.1, 'fp32', (0, k), 64, 2)
    a = oracle.fq_pairsum(pos, sf, 0.1, 'fp32', (0, 1024), 64, 2)
    b = oracle.fq_pairsum(pos, sf, 0.1, 'fp32', (1024, k), 64, 2)
    assert np.allclose(a + b, whole, rtol=1e-12, atol=1e-9)
    ws = np.empty((2, 60, 3, 40), np.float32)
    g1 = oracle.grad_pairsum(pos, sf, 0.1, 'fp32', (0, k), 64, 2, workspace=ws)
    g2 = oracle.grad_pairsum(pos, sf, 0.1, 'fp32', (0, k), 64, 1)
    assert np.abs(g1 - g2).max() <= 1e-5 * np.abs(g2).max()
