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
    scat = ElasticScatter(precision='fp32')
    scat._ensure_wrapped(atoms)
    return atoms, scat


def cpu_baseline_sample(atoms, scat, seconds_target=16.0):
    """Time the oracle's C port of the reference's numba CPU path
    (cpu_wrappers/flat_multi_cpu_wrap.py: Pool over pair chunks of atomic_fq +
    atomic_grad_fq) on a bounded slice of the SAME workload's pair list, all
    host cores."""
    import oracle
    cores = os.cpu_count() or 1
    # explicit thread count: torchrun exports OMP_NUM_THREADS=1, which must not
    # throttle the CPU baseline (the C port passes it to `num_threads`)
    threads = max(1, min(cores, 32))
    pos = atoms.get_positions()
    sf = atoms.get_array('F(Q) scatter')
    n, nq = sf.shape
    k_total = n * (n - 1) // 2
    chunk = 1 << 11
    # calibrate on a small slice, then size the sample for ~seconds_target
    k0 = (k_total // 3) // chunk * chunk
    m = min(threads * chunk * 2, k_total - k0)

    def run(m):
        t = time.perf_counter()
        oracle.fq_pairsum(pos, sf, scat.exp['qbin'], 'fp32', (k0, k0 + m), chunk, threads)
        t1 = time.perf_counter()
        oracle.grad_pairsum(pos, sf, scat.exp['qbin'], 'fp32', (k0, k0 + m), chunk, threads)
        t2 = time.perf_counter()
        return t1 - t, t2 - t1

    tf, tg = run(m)
    m2 = m
    # grow the slice until it takes about seconds_target (the first, tiny
    # slice is dominated by allocation, so re-estimate the rate as we go)
    for _ in range(6):
        if tf + tg >= 0.6 * seconds_target:
            break
        rate = m2 / max(tf + tg, 1e-9)
        nxt = int(min(k_total - k0, max(2 * m2, rate * seconds_target)))
        nxt = max(chunk, nxt // chunk * chunk)
        if nxt <= m2:
            break
        m2 = nxt
        tf, tg = run(m2)
    value = m2 * nq / (tf + tg)
    return {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'host_cores': cores,
            'sample': ('%d of %d pairs (k in [%d, %d)) of the %d-atom workload x %d Q bins, '
                       'F(Q) pass %.2f s + grad F(Q) pass %.2f s, chunks of %d pairs, '
                       'materialised K x Q / K x 3 x Q intermediates as the reference does; '
                       'scaled linearly in pairs' % (m2, k_total, k0, k0 + m2, n, nq, tf, tg, chunk)),
            'seconds': tf + tg}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle C port; the
    reference is numba Python and cannot travel to the GPU box) on the host
    cores, each step a bounded sample of the workload."""
    if rank != 0:
        return
    atoms, scat = build_workload(N_ATOMS)
    steps = max(1, args.steps)
    per_step = max(2.0, min(12.0, 60.0 / (steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline_sample(atoms, scat, 1.0)
    vals, secs, last = [], [], None
    for _ in range(steps):
        last = cpu_baseline_sample(atoms, scat, per_step)
        vals.append(last['value'])
        secs.append(last['seconds'])
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * float(np.mean(secs)), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(world=1, note='CPU sample per step'),
        'cpu_baseline': dict(last, value=value),
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def workload_config(world, note=None):
    cfg = {'workload': '%s %d-atom fcc nanoparticle (sigma 0.05 A, seed 0), F(Q)+grad F(Q), '
                       'qmax 25, qbin 0.1 (250 Q bins), FP32 mode' % (SYMBOL, N_ATOMS),
           'atoms': N_ATOMS, 'q_bins': 250, 'pairs': N_ATOMS * (N_ATOMS - 1) // 2,
           'parallelism': 'pair-tile sharding over %d GPU(s), NCCL all-reduce of F(Q) and grad'
                          % world if world > 1 else 'single GPU',
           'l2': 'output 150 MB (> 126 MB L2) is re-zeroed and rewritten every step; inputs are '
                 '0.6 MB and the kernel is FP32-pipe bound, so no explicit flush'}
    if note:
        cfg['note'] = note
    return cfg


def hmc_extra():
    """configs[1]: Au561 icosahedron, Calc1D Rw energy+forces under NUTS."""
    from pyiid_b200 import ElasticScatter, Calc1D, structures
    from pyiid_b200 import sim
    ideal = structures.icosahedron('Au', 5)
    scat = ElasticScatter(precision='fp32')
    target = scat.get_pdf(ideal)
    atoms = structures.icosahedron('Au', 5)
    atoms.positions *= 1.05
    calc = Calc1D(target_data=target, exp_function=scat.get_pdf,
                  exp_grad_function=scat.get_grad_pdf, conv=100, potential='rw')
    atoms.set_calculator(calc)
    atoms.get_forces()
    be = scat.pdf_backend
    pos = atoms.get_positions()
    for _ in range(3):
        be.energy_forces(pos, target, 'rw', 100.)
    t = time.perf_counter()
    reps = 50
    for i in range(reps):
        be.energy_forces(pos + 1e-6 * i, target, 'rw', 100.)
    evals_per_s = reps / (time.perf_counter() - t)
    out = {'workload': 'Au561 icosahedron x1.05 vs ideal PDF, Calc1D(conv=100, rw), '
                       'NUTS(T=1000, escape_level=8, seed=0)',
           'energy_force_evals_per_s': evals_per_s,
           'pairq_per_eval': 561 * 560 // 2 * 330}
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


def config3_extra():
    """configs[2]: Au 10 000-atom particle, every pair sum, FP32 and FP64 mode
    (kernel-only CUDA-event times) plus the fused Rw energy+forces wall time."""
    from pyiid_b200 import ElasticScatter, structures
    atoms = structures.fcc_sphere('Au', 10000, sigma=0.05, seed=0)
    ideal = structures.fcc_sphere('Au', 10000, sigma=0.0)
    pos = atoms.get_positions()
    out = {'workload': 'Au 10000-atom fcc particle; F(Q)/grad on 250 bins, energy+forces on the '
                       '330-bin PDF grid; kernel ms by CUDA events',
           'pairq_250': 10000 * 9999 // 2 * 250}
    for prec in ('fp32', 'fp64'):
        scat = ElasticScatter(precision=prec)
        scat._ensure_wrapped(atoms)
        be = scat._load(atoms, scat.exp['qbin'], 'fq')
        be.set_timing(True)
        for name, fn in (('fq', be.fq), ('fq_grad', be.grad_fq)):
            ts = []
            for _ in range(3):
                fn(pos)
                ts.append(be.last_kernel_ms()[0])
            out['%s_kernel_ms_%s' % (name, prec)] = min(ts)
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
    return out


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

    import __graft_entry__ as entry
    if args.impl == 'reference':
        if rank == 0:
            entry.build()
        run_reference(args, rank, world)
        return

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
                dist.all_reduce(g_d)
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
        g_check = float(g_d.abs().max().item())

    elapsed_ms = max_over_ranks(elapsed_ms)
    value = pairq * args.steps / (elapsed_ms * 1e-3)
    kernel_ms_avg = max_over_ranks(float(np.mean(ks)))

    # ---- end to end through the public host-buffer API ------------------------
    # with several ranks the reduced gradient is delivered to rank 0's host
    # array (root_only), as the reference's one-process multi-GPU path does
    # warm-up with the reference pattern of the timed loop: the previous result
    # is still alive while the next one is produced, so the pinned-output pool
    # reaches its steady state (two buffers) before the clock starts
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
    d2h = (g_host.nbytes if g_host is not None else 0) + f_host2.nbytes

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
    roofline = {
        'bound': 'fp32_fma+sfu (no tensor cores, HBM negligible: SURVEY.md 8d)',
        'kernel': 'iid::debye2_kernel<32, MODE_GRAD, 256, 1, 16, CHEB>',
        'achieved': achieved, 'peak': peak_pairq, 'unit': UNIT, 'frac': achieved / peak_pairq,
        'peak_how': '%d SMs x %.0f MHz (max SM clock, MEASURED_PEAKS.json / nvidia-smi) x 8 pair*Q/clk/SM '
                    '= min(SFU 16/clk / 2, FP32 128/clk / 12); algorithmic count, our kernel replaces '
                    'the SFU sin/cos by FP32 recurrences' % (sm_count.value, sm_max_mhz),
        'kernel_ms': kernel_ms_avg,
        'achieved_tflops': achieved * FLOP_PER_PAIRQ / 1e12,
        'peak_tflops_fp32': sm_count.value * 128 * 2 * sm_max_mhz * 1e6 / 1e12,
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
                'api': 'ElasticScatter.grad (wrap_fq_grad) + F(Q), host numpy in/out'
                       + (' (reduced to rank 0: Backend.grad_fq(root_only=True))' if world > 1 else '')},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'clocks': clocks,
        'check': {'fq_max': float(np.abs(f_host).max()), 'grad_max': g_check},
    }
    if world == 1:
        line['cpu_baseline'] = cpu_baseline_sample(atoms, scat)
        if not args.no_extras:
            line['extras'] = {}
            for key, fn in (('hmc_au561', hmc_extra), ('au10k_modes', config3_extra)):
                try:
                    line['extras'][key] = fn()
                except Exception as exc:  # keep the headline line even if an extra fails
                    line['extras'][key + '_error'] = repr(exc)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dram bytes (read + write) of one MODE_GRAD launch at the bench workload, from
# the ncu --set full capture summarised in profiles/; None until measured there.
TRAFFIC_BYTES = 730.3e6  # profiles/r1_ncu_full_grad50k_metrics.csv: 292.4 MB read + 437.9 MB write

if __name__ == '__main__':
    main()
