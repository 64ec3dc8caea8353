"""bench.py -- headline benchmark of the elastic-scattering hot path.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (reference CPU algorithm)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], the one the metric is quoted on): a
50 000-atom Pt nanoparticle (random-perturbed fcc sphere, sigma 0.05 A, seed 0),
F(Q) + grad F(Q) on the default experiment grid (qmax 25, qbin 0.1 -> 250 Q
bins), FP32 mode.  One step = one evaluation of F(Q) [250] and the full
per-atom gradient [N, 3, 250].  Metric: pair*Q evaluations per second with
pair*Q = N(N-1)/2 * 250 unique pairs.  With N GPUs the
pair-tile work list is sharded over the ranks (strong scaling) and the partial
F(Q) / gradient arrays are all-reduced over NCCL.

The JSON line carries: `value` (device-resident inputs, CUDA-event timed
region over K steps, max over ranks), `e2e` (same metric through the public
host-buffer API, H2D of positions and D2H of F(Q) + gradient inside the timed
region), `roofline` (dominant kernel against the FP32-FMA/SFU bound of
SURVEY.md section 8d), `cpu_baseline` (the oracle's C port of the reference's
numba CPU path on this box's host cores, bounded sample), `clocks`,
`gpu_launches`.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'pair*Q evaluations/s, F(Q)+grad F(Q)'
UNIT = 'pair*Q/s'
N_ATOMS = 50000
SYMBOL = 'Pt'
# Algorithmic work per unique pair*Q for F + grad F:
# 2 transcendentals + 20 flop ~ 2 SFU + 12 FP32 instructions, so the bound is
# min(SFU/2, FP32/12) = 8 pair*Q per clock per SM.
PAIRQ_PER_CLK_PER_SM = 8.0
# Sharded run against the single-rank evaluation, normalised max-norm.  The two
# cut the rows into different pieces, i.e. they add the same float32 terms in
# different partial sums; each is within 2e-6 of the float64 mode at this size
# (profiles/r2_fullsize_consistency.txt), so their difference stays below twice
# that.  F(Q) is summed in float64: 1e-12.
NERR_SHARDED = 5e-6
FLOP_PER_PAIRQ = 22.0


# --- torch.distributed helpers (also exercised by tests/test_dist_cpu.py) ------
def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _reduce_scalar(x, op_name):
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return float(x)
    import torch
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(x):
    return _reduce_scalar(x, 'MAX')


def sum_over_ranks(x):
    return _reduce_scalar(x, 'SUM')


def barrier():
    dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        dist.barrier()


# --- clocks ---------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,'
             'clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.proc = None
        self.lines = []
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# --- workload -------------------------------------------------------------------
def build_workload(n_atoms):
    from pyiid_b200 import ElasticScatter, structures
    atoms = structures.fcc_sphere(SYMBOL, n_atoms, sigma=0.05, seed=0)
    scat = ElasticScatter(precision='fp32', device=int(os.environ.get('LOCAL_RANK', '0')))
    scat._ensure_wrapped(atoms)
    return atoms, scat


def host_workload(n_atoms):
    """Positions [N,3] float64 and the per-atom scatter-factor array [N,250]
    float32 of the bench workload, built WITHOUT the product library (the
    reference arm must not map libiid_b200.so)."""
    from pyiid_b200 import structures, formfactors
    pos = structures.fcc_sphere_positions(n_atoms, structures.A_PT, sigma=0.05, seed=0)
    qbin, nq = 0.1, 250
    row = np.zeros((1, nq), np.float32)
    formfactors.get_scatter_array(row, [78], qbin)
    return pos, np.repeat(row, n_atoms, axis=0), qbin


class CpuSample(object):
    """The oracle's C port of the reference's numba CPU path
    (cpu_wrappers/flat_multi_cpu_wrap.py: Pool over pair chunks of atomic_fq +
    atomic_grad_fq, materialised K x Q / K x 3 x Q chunk arrays) on ONE fixed
    slice of the workload's pair list -- 2^24 pairs from the middle of the
    k range -- with all host cores.  The result array and the per-thread
    accumulators (what each Pool task returns in the reference) are allocated
    once, so repeated timed runs measure the arithmetic, not page faults."""
    PAIRS = 1 << 24
    CHUNK = 1 << 11

    def __init__(self, pos, sf, qbin, pairs=None, ws=None):
        import oracle
        self.oracle = oracle
        self.pos, self.sf, self.qbin = pos, sf, qbin
        self.cores = os.cpu_count() or 1
        # explicit thread count: torchrun exports OMP_NUM_THREADS=1, which must
        # not throttle the CPU baseline
        self.threads = max(1, min(self.cores, 32))
        n, nq = sf.shape
        self.n, self.nq = n, nq
        self.k_total = n * (n - 1) // 2
        m = min(pairs or self.PAIRS, self.k_total)
        self.k0 = ((self.k_total - m) // 2) // self.CHUNK * self.CHUNK
        self.m = m
        self.out = np.zeros((n, 3, nq), np.float32)
        # touched once here: no page faults inside the timed runs
        self.ws = ws if ws is not None else np.ones((self.threads, n, 3, nq), np.float32)

    def run(self):
        """(seconds F(Q) pass, seconds grad F(Q) pass) of the fixed slice."""
        o, rng = self.oracle, (self.k0, self.k0 + self.m)
        t0 = time.perf_counter()
        o.fq_pairsum(self.pos, self.sf, self.qbin, 'fp32', rng, self.CHUNK, self.threads)
        t1 = time.perf_counter()
        o.grad_pairsum(self.pos, self.sf, self.qbin, 'fp32', rng, self.CHUNK, self.threads,
                       out=self.out, workspace=self.ws)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    def describe(self, tf, tg):
        return {'value': self.m * self.nq / (tf + tg), 'unit': UNIT, 'cores': self.threads,
                'kind': 'port', 'host_cores': self.cores,
                'sample': ('%d of %d pairs (k in [%d, %d)) of the %d-atom workload x %d Q bins, '
                           'F(Q) pass %.2f s + grad F(Q) pass %.2f s, chunks of %d pairs, '
                           'materialised K x Q / K x 3 x Q intermediates as the reference does, '
                           'result and per-thread accumulators preallocated; scaled linearly in '
                           'pairs' % (self.m, self.k_total, self.k0, self.k0 + self.m, self.n,
                                      self.nq, tf, tg, self.CHUNK)),
                'seconds': tf + tg}


def cpu_baseline_sample(pos, sf, qbin):
    """One timed run of the fixed CPU sample (after a short untimed one)."""
    cs = CpuSample(pos, sf, qbin)
    CpuSample(pos, sf, qbin, pairs=cs.threads * cs.CHUNK, ws=cs.ws).run()  # warms the threads
    tf, tg = cs.run()
    return cs.describe(tf, tg)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle C port; the
    reference is numba Python and cannot travel to the GPU box) on the host
    cores.  Every step is the SAME fixed sample of the workload (CpuSample);
    only the oracle library is loaded."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    pos, sf, qbin = host_workload(N_ATOMS)
    steps = max(1, args.steps)
    # ONE sample size for the whole run: 2^24 pairs unless this box's cores would
    # need more than ~4.5 minutes for all the steps (probe: 2^20 pairs), never
    # below 2^22 pairs
    probe = CpuSample(pos, sf, qbin, pairs=1 << 20)
    probe.run()
    rate = probe.m / sum(probe.run())
    budget_s = 270.0 / (steps + min(args.warmup, 1))
    pairs = int(min(CpuSample.PAIRS, max(1 << 22, rate * budget_s)))
    pairs = pairs // CpuSample.CHUNK * CpuSample.CHUNK
    ws = probe.ws
    del probe
    cs = CpuSample(pos, sf, qbin, pairs=pairs, ws=ws)
    for _ in range(min(args.warmup, 1)):  # compiled C, nothing to warm but the caches
        cs.run()
    secs = []
    last = (0.0, 0.0)
    for _ in range(steps):
        last = cs.run()
        secs.append(sum(last))
    mean_s = float(np.mean(secs))
    value = cs.m * cs.nq / mean_s
    base = cs.describe(*last)
    base.update(value=value, seconds=mean_s,
                spread='min %.2f s, max %.2f s over %d steps' % (min(secs), max(secs), steps))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * mean_s, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(world=args.gpus),
        'cpu_baseline': base,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def workload_config(world, note=None):
    cfg = {'workload': '%s %d-atom fcc nanoparticle (sigma 0.05 A, seed 0), F(Q)+grad F(Q), '
                       'qmax 25, qbin 0.1 (250 Q bins), FP32 mode' % (SYMBOL, N_ATOMS),
           'atoms': N_ATOMS, 'q_bins': 250, 'pairs': N_ATOMS * (N_ATOMS - 1) // 2,
           'parallelism': 'gradient rows (i-tiles) dealt to %d GPU(s), no gradient collective; '
                          'NCCL all-reduce of the 250 F(Q) pair sums' % world
                          if world > 1 else 'single GPU',
           'l2': 'the 150 MB output (> 126 MB L2) is rewritten every step; inputs are 0.6 MB and '
                 'the kernel is FP32-pipe bound, so no explicit flush'}
    if note:
        cfg['note'] = note
    return cfg


def hmc_extra():
    """configs[1]: Au561 icosahedron, Calc1D Rw energy+forces under NUTS."""
    from pyiid_b200 import ElasticScatter, Calc1D, structures
    from pyiid_b200 import sim
    ideal = structures.icosahedron('Au', 5)
    scat = ElasticScatter(precision='fp32', device=_dev())
    target = scat.get_pdf(ideal)
    atoms = structures.icosahedron('Au', 5)
    atoms.positions *= 1.05
    calc = Calc1D(target_data=target, exp_function=scat.get_pdf,
                  exp_grad_function=scat.get_grad_pdf, conv=100, potential='rw')
    atoms.set_calculator(calc)
    atoms.get_forces()
    be = scat.pdf_backend
    pos = atoms.get_positions()
    target = calc.target_data  # the calculator's read-only copy: recognised by identity
    for _ in range(3):
        be.energy_forces(pos, target, 'rw', 100.)
    t = time.perf_counter()
    reps = 200
    for i in range(reps):
        be.energy_forces(pos + 1e-6 * i, target, 'rw', 100.)
    evals_per_s = reps / (time.perf_counter() - t)
    out = {'workload': 'Au561 icosahedron x1.05 vs ideal PDF, Calc1D(conv=100, rw), '
                       'NUTS(T=1000, escape_level=8, seed=0)',
           'energy_force_evals_per_s': evals_per_s,
           'pairq_per_eval': 561 * 560 // 2 * 330}
    # native leapfrog on device-resident states: one call per step, and a chain of
    # 16 steps inside one cooperative launch (what a NUTS subtree asks for)
    a0 = atoms.copy()
    a0.set_calculator(calc)
    a0.set_momenta(np.random.RandomState(0).normal(0, 1, (len(a0), 3)))
    a0.get_forces()
    dsys = sim._DeviceSystem(a0)
    st0 = dsys.state_of(a0)
    slots = [dsys.pool.take() for _ in range(16)]
    for n_chain, key in ((1, 'leapfrog_native_us_single_step'),
                         (16, 'leapfrog_native_us_per_step_chain_of_16')):
        for _ in range(6):
            be.leapfrog_chain(st0.slot, slots[:n_chain], 1e-3, True, target, 'rw', 100.)
        n0, t = be.launch_count(), time.perf_counter()
        reps = 100
        for _ in range(reps):
            be.leapfrog_chain(st0.slot, slots[:n_chain], 1e-3, True, target, 'rw', 100.)
        out[key] = (time.perf_counter() - t) / reps / n_chain * 1e6
        out['launches_per_' + ('step' if n_chain == 1 else 'chain_of_16')] = \
            (be.launch_count() - n0) / reps
    for s_ in slots:
        dsys.pool.give(s_)
    del st0, dsys
    # default: tree states resident on the device (one native call per leapfrog);
    # then the array-level host path and the Atoms-level (reference-style) path
    for fast, dev, tag in ((True, True, ''), (True, False, '_array_level_path'),
                           (False, False, '_atoms_level_path')):
        np.random.seed(0)
        a = atoms.copy()
        a.set_calculator(calc)
        ens = sim.NUTSCanonicalEnsemble(a, temperature=1000, escape_level=8, seed=0, fast=fast,
                                        device_states=dev)
        ens.run(1)  # graph capture, state slots, step-size adaptation start
        lf0, t = ens.leapfrogs, time.perf_counter()
        iters = 6 if fast else 3
        ens.run(iters)
        dt = time.perf_counter() - t
        out['hmc_leapfrog_steps_per_s' + tag] = (ens.leapfrogs - lf0) / dt
        out['nuts_iterations_per_s' + tag] = iters / dt
    return out


def _dev():
    """Extras run on this rank's GPU (never the one-process multi-GPU handle)."""
    return int(os.environ.get('LOCAL_RANK', '0'))


def config3_extra():
    """configs[2]: Au 10 000-atom particle, every pair sum, FP32 and FP64 mode
    (kernel-only CUDA-event times) with a roofline per mode, plus the fused Rw
    energy+forces wall time."""
    from pyiid_b200 import ElasticScatter, structures
    atoms = structures.fcc_sphere('Au', 10000, sigma=0.05, seed=0)
    ideal = structures.fcc_sphere('Au', 10000, sigma=0.0)
    pos = atoms.get_positions()
    n, nq = 10000, 250
    pairq = n * (n - 1) // 2 * nq
    out = {'workload': 'Au 10000-atom fcc particle; F(Q)/grad on 250 bins, energy+forces on the '
                       '330-bin PDF grid; kernel ms by CUDA events',
           'pairq_250': pairq}
    sm = ctypes.c_int(0)
    peaks = None
    for prec in ('fp32', 'fp64'):
        scat = ElasticScatter(precision=prec, device=_dev())
        scat._ensure_wrapped(atoms)
        be = scat._load(atoms, scat.exp['qbin'], 'fq')
        if peaks is None:
            peaks = be.measure_peaks()
            be.lib.iid_device_info(be.device, ctypes.byref(sm), None, None)
            out['measured_pipe_peaks_lane_fma_per_s'] = peaks
        be.set_timing(True)
        for name, fn in (('fq', be.fq), ('fq_grad', be.grad_fq)):
            ts = []
            for _ in range(3):
                fn(pos)
                ts.append(be.last_kernel_ms()[0])
            out['%s_kernel_ms_%s' % (name, prec)] = min(ts)
        if prec == 'fp32':
            # the default F(Q) pass at this size is the radial pair histogram
            # (O(N^2 + K Q)); the direct O(N^2 Q) kernel for the roofline of SURVEY 8d
            be.set_option('fq_hist', 0)
            ts = []
            for _ in range(3):
                be.fq(pos)
                ts.append(be.last_kernel_ms()[0])
            out['fq_direct_kernel_ms_fp32'] = min(ts)
            be.set_option('fq_hist', 1)
        be.set_timing(False)
        bp = scat._load(atoms, scat.pdf_qbin, 'PDF')
        bp.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
        target = bp.pdf(ideal.get_positions())
        ts = []
        for _ in range(4):
            t = time.perf_counter()
            bp.energy_forces(pos, target, 'rw', 100.)
            ts.append(time.perf_counter() - t)
        out['energy_forces_wall_ms_%s' % prec] = 1e3 * min(ts)
    # rooflines.  Algorithmic bounds: F only 16 pair*Q/clk/SM (SFU),
    # F + grad 8; FP64 mode: the same counts on the 64 lane/clk DFMA pipe, i.e.
    # the kernel's 8 (F + grad, 6.2 above the diagonal) / ~2.1 (F only) DFMA-class
    # instructions per ordered / unordered pair*bin against the MEASURED DFMA rate.
    clk = 1965e6
    try:
        clk = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['sm_max_mhz']) * 1e6
    except (OSError, ValueError, KeyError):
        pass
    sms = sm.value or 148
    roof = {}
    t = out['fq_direct_kernel_ms_fp32'] * 1e-3
    roof['fq_fp32'] = {'achieved': pairq / t, 'peak': sms * clk * 16, 'unit': UNIT,
                       'frac': pairq / t / (sms * clk * 16),
                       'note': 'the direct O(N^2 Q) kernel (fq_hist = 0).  Above 1: the SFU count '
                               'of the bound is not executed, sin comes from FP32 recurrences',
                       'executed_lane_fma_frac_of_measured_ffma2': pairq * 2.6 / t / peaks['ffma2']}
    t = out['fq_kernel_ms_fp32'] * 1e-3
    roof['fq_hist_fp32'] = {'achieved': pairq / t, 'unit': UNIT, 'peak': None, 'frac': None,
                            'pairs_per_s': 0.5 * n * (n - 1) / t,
                            'note': 'the shipped F(Q) pass at this size: radial pair histogram, '
                                    'O(N^2 + K Q) -- no pair*Q bound applies; it is bound by 16 '
                                    '32-bit shared-memory atomics per pair (8-point stencil on the fine grid; '
                                    'the atomic unit issues one per ~6 cycles per SM), '
                                    '%.1f SM cycles per pair' % (t * sms * clk / (0.5 * n * (n - 1)))}
    t = out['fq_grad_kernel_ms_fp32'] * 1e-3
    roof['fq_grad_fp32'] = {'achieved': pairq / t, 'peak': sms * clk * 8, 'unit': UNIT,
                            'frac': pairq / t / (sms * clk * 8),
                            'executed_lane_fma_frac_of_measured_ffma2':
                                0.5 * n * n * nq * (8.25 + 6.7) / t / peaks['ffma2']}
    t = out['fq_kernel_ms_fp64'] * 1e-3
    roof['fq_fp64'] = {'achieved': pairq / t, 'unit': UNIT,
                       'peak': peaks['dfma'] / 2.1, 'frac': pairq * 2.1 / t / peaks['dfma'],
                       'peak_how': 'measured DFMA rate / 2.1 DFMA-class instructions per pair*bin'}
    t = out['fq_grad_kernel_ms_fp64'] * 1e-3
    roof['fq_grad_fp64'] = {'achieved': pairq / t, 'unit': UNIT,
                            'peak': peaks['dfma'] / (8 + 6.2),
                            'frac': 0.5 * n * n * nq * (8 + 6.2) / t / peaks['dfma'],
                            'peak_how': 'measured DFMA rate / (8 + 6.2) DFMA-class instructions '
                                        'per unique pair*bin (square walk: F + grad below the '
                                        'diagonal, gradient only above it)',
                            'nominal_dfma_peak': sms * clk * 64}
    out['roofline'] = roof
    return out


def config1_extra():
    """configs[0]: Au55 icosahedron, get_pdf + get_grad_pdf (examples/Au_NP_PDF.py
    :14-49 workflow) -- wall time on the B200 beside the oracle's CPU time for
    the same two calls (the reference's own CPU-runnable case)."""
    import oracle
    from pyiid_b200 import ElasticScatter, structures
    atoms = structures.icosahedron('Au', 2)
    scat = ElasticScatter(precision='fp32', device=_dev())
    for _ in range(2):
        pdf = scat.get_pdf(atoms)
        gpdf = scat.get_grad_pdf(atoms)
    t = time.perf_counter()
    reps = 5
    for _ in range(reps):
        pdf = scat.get_pdf(atoms)
        gpdf = scat.get_grad_pdf(atoms)
    gpu_s = (time.perf_counter() - t) / reps
    pos = atoms.get_positions()
    sp = atoms.get_array('PDF scatter')
    exp = oracle.DEFAULT_EXP
    t = time.perf_counter()
    o_pdf = oracle.experiment_pdf(pos, sp, exp)
    o_gpdf = oracle.experiment_grad_pdf(pos, sp, exp)
    cpu_s = time.perf_counter() - t

    def nerr(a, b):
        return float(np.abs(np.asarray(a, float) - b).max() / np.abs(b).max())

    return {'workload': 'Au55 Mackay icosahedron, get_pdf + get_grad_pdf, 330-bin PDF grid, '
                        '4000 r points, host arrays in and out',
            'b200_ms': 1e3 * gpu_s, 'cpu_oracle_ms': 1e3 * cpu_s,
            'cpu_kind': 'port (oracle: C pair sums, numpy FFT per gradient row as '
                        'master_kernel.grad_pdf does), one core',
            'nerr_pdf_vs_oracle': nerr(pdf, o_pdf), 'nerr_grad_pdf_vs_oracle': nerr(gpdf, o_gpdf)}


def config5_extra(world):
    """configs[4]: 100 000-atom Au/Pt alloy, F(Q) -> G(r) + Rw on the 330-bin
    PDF grid, all ranks (collective).  Returns the record on every rank."""
    import torch
    import torch.distributed as dist
    from pyiid_b200 import ElasticScatter, structures
    n = 100000
    atoms = structures.alloy_sphere(n)
    ideal = structures.alloy_sphere(n, sigma=0.0)
    res = {}
    for prec in ('fp32', 'fp64'):
        scat = ElasticScatter(precision=prec)
        target = scat.get_pdf(ideal)
        scat._ensure_wrapped(atoms)
        be = scat.pdf_backend
        pos = atoms.get_positions()
        if prec == 'fp64':  # same float32-rounded coordinates as the FP32 mode sees
            pos = pos.astype(np.float32).astype(np.float64)
        for _ in range(2):
            e, scale, _, _ = be.energy_forces(pos, target, 'rw', 1.0, want_forces=False)
        dist.barrier()
        torch.cuda.synchronize()
        k = 5 if prec == 'fp32' else 2
        t = time.perf_counter()
        for _ in range(k):
            e, scale, _, _ = be.energy_forces(pos, target, 'rw', 1.0, want_forces=False)
        dt = max_over_ranks((time.perf_counter() - t) / k)
        res[prec] = (dt, float(e), float(scale), be.nq)
    dt, e, scale, nq = res['fp32']
    pairq = n * (n - 1) // 2 * nq
    sms, clk = 148, 1965e6
    return {'workload': 'Au/Pt %d-atom random alloy (seed 1), F(Q)->G(r)+Rw, %d-bin PDF grid, host '
                        'positions in, Rw out; pair-triangle items dealt to the ranks, NCCL '
                        'all-reduce of the F(Q) pair sums' % (n, nq),
            'n_gpus': world, 'ms_per_evaluation': 1e3 * dt, 'pairq_per_s': pairq / dt,
            'frac_of_f_only_bound': pairq / dt / (world * sms * clk * 16),
            'note': 'FP32 mode runs the F(Q) pass as a radial pair histogram (O(N^2 + K Q), DESIGN '
                    '4.1b): the pair*Q bound of SURVEY 8d does not apply to it (fraction above 1)',
            'rw_fp32': e, 'scale_fp32': scale,
            'ms_per_evaluation_fp64': 1e3 * res['fp64'][0], 'rw_fp64': res['fp64'][1],
            'rw_fp32_vs_fp64_rel': abs(e - res['fp64'][1]) / abs(res['fp64'][1])}


_JSON_FD = None


def _reserve_stdout():
    """Keep the real stdout for the ONE JSON line: everything else that writes
    to file descriptor 1 while the bench runs (NCCL's version banner, nvcc, C
    libraries) goes to stderr instead."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--atoms', type=int, default=N_ATOMS, help=argparse.SUPPRESS)
    ap.add_argument('--no-extras', action='store_true')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import __graft_entry__ as entry
    import torch
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        if rank == 0:
            entry.build()  # the other ranks load the library only after this
        dist.barrier()
    else:
        entry.build()
    warmup = max(3, args.warmup)
    n_atoms = args.atoms
    atoms, scat = build_workload(n_atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    be.sync_shard()
    pos = atoms.get_positions()
    n, nq = be.n, be.nq
    pairq = n * (n - 1) // 2 * nq
    lib, h = be.lib, be.h

    # ---- device-resident timed region --------------------------------------
    # every rank computes the gradient rows of its i-tiles into its own device
    # array (rows of the other ranks stay zero: no gradient collective); the
    # 250 F(Q) pair sums are all-reduced
    dev = torch.device('cuda', be.device)
    with torch.cuda.device(dev), be._on_stream():
        pos_d = torch.from_numpy(pos).to(dev)
        g_d = torch.zeros((n, 3, nq), dtype=torch.float32, device=dev)
        s_d = torch.zeros(nq, dtype=torch.float64, device=dev)
        f_d = torch.zeros(nq, dtype=torch.float64, device=dev)

        def step():
            rc = lib.iid_grad_fq_partial(h, pos_d.data_ptr(), g_d.data_ptr(), s_d.data_ptr(), None)
            assert rc == 0, lib.iid_last_error()
            if world > 1:
                dist.all_reduce(s_d)
            rc = lib.iid_fq_finish(h, s_d.data_ptr(), f_d.data_ptr(), None)
            assert rc == 0, lib.iid_last_error()

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        barrier()
        be.set_timing(True)
        sampler = ClockSampler(be.device)
        if rank == 0:
            sampler.start()
        launches0 = be.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
        barrier()
        elapsed_ms = ev0.elapsed_time(ev1)
        launches = be.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        # per-launch duration of the dominant kernel (CUDA events on the launching
        # stream, recorded by the library around the launch)
        ks = []
        for _ in range(3):
            step()
            ks.append(be.last_kernel_ms()[0])
        be.set_timing(False)
        f_host = f_d.cpu().numpy()
        g_max = float(g_d.abs().max().item())

        # ---- parity of the sharded run (world > 1): the rows of all ranks,
        # put together, against ONE rank evaluating the whole pair list
        check = {'fq_max': float(np.abs(f_host).max()), 'grad_max': g_max}
        if world > 1:
            g_all = g_d.clone()
            dist.all_reduce(g_all)  # disjoint rows: the sum is the assembled gradient
            if rank == 0:
                assert lib.iid_set_shard(h, 0, 1) == 0
                g_one = torch.empty_like(g_d)
                s_one = torch.zeros_like(s_d)
                f_one = torch.zeros_like(f_d)
                assert lib.iid_grad_fq_partial(h, pos_d.data_ptr(), g_one.data_ptr(),
                                               s_one.data_ptr(), None) == 0
                assert lib.iid_fq_finish(h, s_one.data_ptr(), f_one.data_ptr(), None) == 0
                check['nerr_grad_vs_single'] = float(
                    ((g_all - g_one).abs().max() / g_one.abs().max()).item())
                check['nerr_fq_vs_single'] = float(
                    ((f_d - f_one).abs().max() / f_one.abs().max()).item())
                g_one_host = g_one.cpu().numpy()
                del g_one
                assert lib.iid_set_shard(h, rank, world) == 0
            del g_all

    elapsed_ms = max_over_ranks(elapsed_ms)
    value = pairq * args.steps / (elapsed_ms * 1e-3)
    kernel_ms_avg = max_over_ranks(float(np.mean(ks)))

    # ---- end to end through the public host-buffer API ------------------------
    # host positions in, one host array [N,3,250] + F(Q) out.  One GPU: the
    # kernel stores its rows straight into a pinned array from the pool.
    # Several ranks: all ranks store their rows into ONE shared host array
    # (root_only; the single array the reference's multi-GPU path assembles).
    # Warm-up with the reference pattern of the timed loop: the previous result
    # is still alive while the next one is produced, so the output pool reaches
    # its steady state (two buffers) before the clock starts.
    for _ in range(3):
        g_host, f_host2 = be.grad_fq(pos, with_fq=True, root_only=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g_host, f_host2 = be.grad_fq(pos, with_fq=True, root_only=True)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = pairq * args.steps / e2e_s
    h2d = pos.nbytes
    d2h = g_host.nbytes + f_host2.nbytes
    if world > 1 and rank == 0:
        den = np.abs(g_one_host).max()
        check['nerr_e2e_grad_vs_single'] = float(np.abs(g_host - g_one_host).max() / den)
        del g_one_host

    extras_multi = None
    if world > 1 and not args.no_extras:
        try:
            extras_multi = config5_extra(world)  # collective: every rank takes part
        except Exception as exc:
            extras_multi = {'error': repr(exc)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    sm_count = ctypes.c_int(0)
    khz = ctypes.c_int(0)
    lib.iid_device_info(be.device, ctypes.byref(sm_count), ctypes.byref(khz), None)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    sm_max_mhz = float(peaks.get('sm_max_mhz') or (clocks or {}).get('sm_max_mhz') or khz.value / 1e3)
    peak_pairq = sm_count.value * sm_max_mhz * 1e6 * PAIRQ_PER_CLK_PER_SM
    achieved = (pairq / world) / (kernel_ms_avg * 1e-3)
    pipes = be.measure_peaks()
    # executed packed-FP32 lane operations of the square walk: lower triangle +
    # diagonal with F(Q) 8.25, upper triangle 6.7 FFMA2-class instructions per
    # (ordered pair, two bins) -- DESIGN.md section 4.1
    lane_fma = 0.5 * n * n * nq * (8.25 + 6.7) / world
    roofline = {
        'bound': 'fp32_fma+sfu (no tensor cores, HBM negligible: SURVEY.md 8d)',
        'kernel': 'iid::debye2_kernel<32, MODE_GRAD, 256, 1, 16, CHEB, 2> (row jobs)',
        'achieved': achieved, 'peak': peak_pairq, 'unit': UNIT, 'frac': achieved / peak_pairq,
        'peak_how': '%d SMs x %.0f MHz (max SM clock, MEASURED_PEAKS.json / nvidia-smi) x 8 pair*Q/clk/SM '
                    '= min(SFU 16/clk / 2, FP32 128/clk / 12); algorithmic count, our kernel replaces '
                    'the SFU sin/cos by FP32 recurrences' % (sm_count.value, sm_max_mhz),
        'kernel_ms': kernel_ms_avg,
        'achieved_tflops': achieved * FLOP_PER_PAIRQ / 1e12,
        'peak_tflops_fp32': sm_count.value * 128 * 2 * sm_max_mhz * 1e6 / 1e12,
        # executed work against the MEASURED packed-FFMA2 issue rate of this GPU
        # (iid_measure_peaks): what the FP32 pipe itself would allow
        'pipe': {'executed_lane_fma_per_s': lane_fma / (kernel_ms_avg * 1e-3),
                 'measured_ffma2_lane_fma_per_s': pipes['ffma2'],
                 'measured_ffma_lane_fma_per_s': pipes['ffma'],
                 'frac_of_measured_ffma2': lane_fma / (kernel_ms_avg * 1e-3) / pipes['ffma2']},
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from
        # profiles/ (ncu --set full); algorithmic bytes are ~150 MB of output
        'traffic': TRAFFIC_BYTES,
        'algorithmic_bytes': n * 3 * nq * 4 + n * 24,
    }
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': warmup, 'ms_per_step': elapsed_ms / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(world) if n_atoms == N_ATOMS else
        dict(workload_config(world), atoms=n_atoms, note='non-default --atoms'),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * e2e_s / args.steps,
                'api': 'Backend.grad_fq (ElasticScatter.grad = wrap_fq_grad) + F(Q): host numpy '
                       'positions in, host arrays out; the kernel stores the gradient rows '
                       'straight into the pinned host array'
                       + (' shared by the %d ranks (root_only)' % world if world > 1 else '')},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'clocks': clocks,
        'check': check,
    }
    failed = None
    if world > 1:
        bad = {k: v for k, v in check.items() if k.startswith('nerr_') and not v < NERR_SHARDED}
        if bad:
            failed = 'sharded run differs from the single-rank evaluation: %r' % bad
            line['check']['failed'] = failed
        if extras_multi is not None:
            line['extras'] = {'cfg5_aupt100k': extras_multi}
    if world == 1:
        sf = atoms.get_array('F(Q) scatter')
        line['cpu_baseline'] = cpu_baseline_sample(pos, sf, scat.exp['qbin'])
        if not args.no_extras:
            line['extras'] = {}
            for key, fn in (('hmc_au561', hmc_extra), ('au10k_modes', config3_extra),
                            ('cfg1_au55', config1_extra)):
                try:
                    line['extras'][key] = fn()
                except Exception as exc:  # keep the headline line even if an extra fails
                    line['extras'][key + '_error'] = repr(exc)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        sys.stderr.write('bench: ' + failed + '\n')
        sys.exit(3)


# dram bytes (read + write) of one MODE_GRAD launch at the bench workload, from
# the ncu --set full capture summarised in profiles/; None until measured there.
TRAFFIC_BYTES = 157.7e6  # profiles/r2_ncu_full_grad50k_metrics.csv: 5.2 MB read + 152.6 MB written

if __name__ == '__main__':
    main()
