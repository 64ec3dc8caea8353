"""CPU oracle for the pyIID elastic-scattering hot path -- TEST INFRASTRUCTURE.

This package restates the reference's algorithm on the CPU so that the CUDA
path can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import it;
the product package ``pyiid_b200`` never does.

Two halves:

* ``iid_oracle.c`` (built into ``libiid_oracle.so`` by ``oracle/Makefile``):
  the O(N^2 Q) flattened-pair Debye sums and their gradient, operation by
  operation as in the reference's numba CPU kernels
  (``pyiid/experiments/elasticscatter/kernels/cpu_flat.py:14-194``,
  ``kernels/cpu_experimental.py:8-15``, ``kernels/__init__.py:15-19``), driven
  like ``cpu_wrappers/flat_serial_cpu_wrap.py:13-133`` (one thread) or
  ``cpu_wrappers/flat_multi_cpu_wrap.py:22-134`` (all cores).
* numpy restatements (this file) of the float64 host stage
  ``kernels/master_kernel.py``: F(Q)->G(r) (``get_pdf_at_qmin :39-104``,
  ``fft_fq_to_gr :108-128``, ``fft_gr_to_fq :131-203``), Rw / chi^2 / scale
  (``:206-266, 388-396``) and their gradients (``:293-375``).

Parity pin: ``oracle/ref_shim.py`` loads the reference's own kernels from
``/root/reference`` (build container only); ``tests/golden/make_golden.py``
stores their outputs and ``tests/test_oracle.py`` checks this oracle against
those fixtures and against the reference's own known-answer tests.

Normaliser: the oracle uses the float64 mean the reference keeps in a comment
(``flat_multi_cpu_wrap.py:55``); ``na_mode=1`` reproduces the shipped float32
``np.mean`` (``:54``) so the as-is deviation can be reported.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libiid_oracle.so')
_lib = None

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int


def build(force=False):
    """Compile ``libiid_oracle.so`` with the Makefile beside this file."""
    src = os.path.join(_HERE, 'iid_oracle.c')
    if (force or not os.path.exists(_LIB_PATH) or
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libiid_oracle.so'],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for suf, cr in (('f32', ctypes.c_float), ('f64', ctypes.c_double)):
            f = getattr(_lib, 'oracle_fq_pairsum_' + suf)
            f.restype = _c_int
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _c_i64, _c_i64, cr,
                          _c_i64, _c_i64, _c_i64, _c_int, ctypes.c_void_p]
            g = getattr(_lib, 'oracle_grad_pairsum_' + suf)
            g.restype = _c_int
            g.argtypes = f.argtypes
            gw = getattr(_lib, 'oracle_grad_pairsum_ws_' + suf)
            gw.restype = _c_int
            gw.argtypes = f.argtypes + [ctypes.c_void_p]
            h = getattr(_lib, 'oracle_pair_internals_' + suf)
            h.restype = _c_int
            h.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _c_i64, _c_i64, cr,
                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                          ctypes.c_void_p]
            nrm = getattr(_lib, 'oracle_normaliser_' + suf)
            nrm.restype = _c_int
            nrm.argtypes = [ctypes.c_void_p, _c_i64, _c_i64, _c_int,
                            ctypes.c_void_p]
        _lib.oracle_max_threads.restype = _c_int
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def _prep(positions, scatter, precision):
    dt = np.float32 if precision == 'fp32' else np.float64
    # flat_multi_cpu_wrap.py:11-19 setup_cpu_calc: positions and scatter
    # factors are cast to the working precision
    q = np.ascontiguousarray(np.asarray(positions, dtype=np.float64).astype(dt))
    scat = np.ascontiguousarray(np.asarray(scatter).astype(dt))
    suf = 'f32' if precision == 'fp32' else 'f64'
    return q, scat, dt, suf


def normaliser(scatter, precision='fp32', na_mode=0):
    """na[Q] = mean over pairs of f_i f_j, times N
    (flat_multi_cpu_wrap.py:52-55)."""
    dt = np.float32 if precision == 'fp32' else np.float64
    scat = np.ascontiguousarray(np.asarray(scatter).astype(dt))
    n, nq = scat.shape
    na = np.zeros(nq, np.float64)
    suf = 'f32' if precision == 'fp32' else 'f64'
    rc = getattr(lib(), 'oracle_normaliser_' + suf)(
        scat.ctypes.data, n, nq, int(na_mode), na.ctypes.data)
    if rc:
        raise MemoryError('oracle normaliser')
    return na


def fq_pairsum(positions, scatter, qbin, precision='fp32', k_range=None,
               chunk=0, nthreads=1):
    """float64 S[Q] = sum over pairs of f_i f_j sin(Q r)/r for pairs
    [k_begin, k_end) of the flattened i>j list (cpu_atomics.py:60-78)."""
    q, scat, dt, suf = _prep(positions, scatter, precision)
    n, nq = scat.shape
    k0, k1 = (0, n * (n - 1) // 2) if k_range is None else k_range
    out = np.zeros(nq, np.float64)
    rc = getattr(lib(), 'oracle_fq_pairsum_' + suf)(
        q.ctypes.data, scat.ctypes.data, n, nq, dt(qbin), int(k0), int(k1),
        int(chunk), int(nthreads), out.ctypes.data)
    if rc:
        raise MemoryError('oracle fq')
    return out


def grad_pairsum(positions, scatter, qbin, precision='fp32', k_range=None,
                 chunk=0, nthreads=1, out=None, workspace=None):
    """rtn[N,3,Q] scatter-summed pair gradients before normalisation
    (cpu_atomics.py:81-102).  ``out`` / ``workspace`` ([nthreads, N, 3, Q]):
    preallocated result and per-thread accumulators for repeated timed calls."""
    q, scat, dt, suf = _prep(positions, scatter, precision)
    n, nq = scat.shape
    k0, k1 = (0, n * (n - 1) // 2) if k_range is None else k_range
    if out is None:
        out = np.zeros((n, 3, nq), dt)
    assert out.dtype == dt and out.shape == (n, 3, nq) and out.flags.c_contiguous
    wptr = None
    if workspace is not None:
        assert workspace.dtype == dt and workspace.size >= max(1, nthreads) * n * 3 * nq
        wptr = workspace.ctypes.data
    rc = getattr(lib(), 'oracle_grad_pairsum_ws_' + suf)(
        q.ctypes.data, scat.ctypes.data, n, nq, dt(qbin), int(k0), int(k1),
        int(chunk), int(nthreads), out.ctypes.data, wptr)
    if rc:
        raise MemoryError('oracle grad')
    return out


def pair_internals(positions, scatter, qbin, precision='fp32'):
    """d[K,3], r[K], norm[K,Q], omega[K,Q] (cpu_flat.py:14-96)."""
    q, scat, dt, suf = _prep(positions, scatter, precision)
    n, nq = scat.shape
    k = n * (n - 1) // 2
    d = np.zeros((k, 3), dt)
    r = np.zeros(k, dt)
    norm = np.zeros((k, nq), dt)
    omega = np.zeros((k, nq), dt)
    getattr(lib(), 'oracle_pair_internals_' + suf)(
        q.ctypes.data, scat.ctypes.data, n, nq, dt(qbin), d.ctypes.data,
        r.ctypes.data, norm.ctypes.data, omega.ctypes.data)
    return d, r, norm, omega


def wrap_fq(positions, scatter, qbin, precision='fp32', na_mode=0,
            nthreads=1, chunk=0):
    """F(Q) as flat_serial_cpu_wrap.wrap_fq :13-69 returns it."""
    dt = np.float32 if precision == 'fp32' else np.float64
    s = fq_pairsum(positions, scatter, qbin, precision, None, chunk, nthreads)
    final = s.astype(dt)                       # :61-62
    na = normaliser(scatter, precision, na_mode)
    if na_mode == 1:
        na = na.astype(np.float32)
    with np.errstate(all='ignore'):
        final = np.nan_to_num(final / na)      # :65
    return 2 * final                           # :69


def wrap_fq_grad(positions, scatter, qbin, precision='fp32', na_mode=0,
                 nthreads=1, chunk=0):
    """grad F(Q) [N,3,Q] as flat_serial_cpu_wrap.wrap_fq_grad :72-133."""
    rtn = grad_pairsum(positions, scatter, qbin, precision, None, chunk,
                       nthreads)
    n = rtn.shape[0]
    if n < 2:
        return rtn
    na = normaliser(scatter, precision, na_mode)
    if na_mode == 1:
        na = na.astype(np.float32)
    with np.errstate(all='ignore'):
        rtn = np.nan_to_num(rtn / na)          # :129
    return rtn


# --- atomic displacement parameters (the reference's unused kernels) ----------

def pair_indices(n):
    """(i, j), i > j, of the flattened pair list k = i (i - 1) / 2 + j
    (kernels/__init__.py:15-19 k_to_ij)."""
    i, j = np.tril_indices(n, -1)
    return i, j


def adp_tau(adps, nq, qbin, precision='fp32'):
    """tau[K, Q] = exp(-(u_i^2 + u_j^2) Q^2 / 2) for isotropic mean-square
    displacements adps[N] -- the Debye-Waller array the reference's
    get_adp_fq / get_adp_grad_fq take as an INPUT (the reference itself builds
    none: its wrappers pass adps = None, flat_multi_cpu_wrap.py:18-19)."""
    dt = np.float32 if precision == 'fp32' else np.float64
    u = np.asarray(adps, dtype=np.float64)
    i, j = pair_indices(len(u))
    sv = (dt(qbin) * np.arange(nq).astype(dt)).astype(np.float64)  # cpu_flat.py:93
    return np.exp(-0.5 * (u[i] + u[j])[:, None] * sv[None, :] ** 2).astype(dt)


def wrap_adp_fq(positions, scatter, adps, qbin, precision='fp32'):
    """F(Q) with fq = norm * omega * tau (kernels/cpu_nxn.py:114-121 get_adp_fq),
    summed and normalised as flat_serial_cpu_wrap.wrap_fq :52-69 (na from norm
    alone).  numpy over all K pairs: small structures only."""
    dt = np.float32 if precision == 'fp32' else np.float64
    d, r, norm, omega = pair_internals(positions, scatter, qbin, precision)
    tau = adp_tau(adps, norm.shape[1], qbin, precision)
    fq = (norm * omega).astype(dt) * tau
    final = fq.sum(axis=0, dtype=np.float64).astype(dt)
    na = normaliser(scatter, precision, 0)
    with np.errstate(all='ignore'):
        final = np.nan_to_num(final / na)
    return 2 * final


def wrap_adp_grad_fq(positions, scatter, adps, qbin, precision='fp32'):
    """grad F(Q) [N, 3, Q] with grad = norm * (tau * grad_omega + omega *
    grad_tau) (kernels/cpu_flat.py:156-174 get_adp_grad_fq), grad_tau = 0 for
    position-independent displacements; grad_omega as cpu_flat.py:117-131,
    scatter-sum as cpu_experimental.py:8-15, / na as
    flat_serial_cpu_wrap.py:129."""
    dt = np.float32 if precision == 'fp32' else np.float64
    d, r, norm, omega = pair_internals(positions, scatter, qbin, precision)
    n, nq = np.asarray(scatter).shape
    tau = adp_tau(adps, nq, qbin, precision)
    sv = dt(qbin) * np.arange(nq).astype(dt)
    with np.errstate(all='ignore'):
        rr = r[:, None]
        a = ((sv[None, :] * np.cos(sv[None, :] * rr).astype(dt)).astype(dt) - omega) / \
            (rr * rr).astype(dt)
    go = (a.astype(dt)[:, None, :] * d[:, :, None]).astype(dt)       # get_grad_omega
    g = (norm[:, None, :] * (tau[:, None, :] * go).astype(dt)).astype(dt)  # grad_tau = 0
    i, j = pair_indices(n)
    rtn = np.zeros((n, 3, nq), dt)
    np.subtract.at(rtn, i, g)
    np.add.at(rtn, j, g)
    na = normaliser(scatter, precision, 0)
    with np.errstate(all='ignore'):
        rtn = np.nan_to_num(rtn / na)
    return rtn


# --- float64 host stage: kernels/master_kernel.py ---------------------------

def fft_gr_to_fq(g, rstep, rmin):
    """master_kernel.py:131-203: odd extension into a 4*npad2 real array at
    even slots, inverse complex FFT, imaginary part of the even outputs."""
    padrmin = int(round(rmin / rstep))
    npad1 = padrmin + len(g)
    npad2 = (1 << int(math.ceil(math.log(npad1, 2)))) * 2
    npad4 = 4 * npad2
    gpadc = np.zeros(npad4)
    gpadc[:2 * len(g):2] = g[:]
    gpadc[-2:-2 * len(g) + 1:-2] = -1 * g[1:]
    gpadcfft = np.fft.ifft(gpadc)
    f = np.zeros(npad2, dtype=complex)
    f[:] = gpadcfft[:npad2 * 2:2] * npad2 * rstep
    return f.imag


def fft_fq_to_gr(f, qbin, qmin):
    """master_kernel.py:108-128."""
    g = fft_gr_to_fq(f, qbin, qmin)
    g *= 2.0 / math.pi
    return g


def get_pdf_at_qmin(fpad, rstep, qstep, rgrid, qmin):
    """master_kernel.py:39-104.  Like the reference this zeroes the low-Q bins
    of the array it is handed."""
    fpad[:int(math.ceil(qmin / qstep))] = 0.0
    nfromdr = int(math.ceil(math.pi / rstep / qstep))
    if nfromdr > int(len(fpad)):
        fpad2 = np.zeros(nfromdr)
        fpad2[:len(fpad)] = fpad
        fpad = fpad2
    gpad = fft_fq_to_gr(fpad, qstep, qmin)
    drpad = math.pi / (len(gpad) * qstep)
    axdrp = rgrid / drpad / 2
    aiplo = axdrp.astype(int)
    aiphi = aiplo + 1
    awphi = axdrp - aiplo
    awplo = 1.0 - awphi
    pdf0 = awplo * gpad[aiplo] + awphi * gpad[aiphi]
    return (pdf0 * 2).real


def grad_pdf(grad_fq, rstep, qstep, rgrid, qmin):
    """master_kernel.py:276-290 without the Pool: one transform per (atom,
    direction) row."""
    n = len(grad_fq)
    out = np.zeros((n, 3, len(rgrid)))
    for tx in range(n):
        for tz in range(3):
            out[tx, tz] = get_pdf_at_qmin(
                np.array(grad_fq[tx, tz], dtype=np.float64), rstep, qstep,
                rgrid, qmin)
    return out


def get_scale(target, calculated):
    """master_kernel.py:388-389."""
    return np.dot(calculated.T, target) / np.dot(calculated.T, calculated)


def get_grad_scale(target, calculated, grad_calculated, tx, tz):
    """master_kernel.py:392-396."""
    a = get_scale(target, calculated)
    return (-2 * a * np.dot(calculated, grad_calculated[tx, tz, :]) +
            np.dot(target, grad_calculated[tx, tz, :])) / np.dot(calculated,
                                                                 calculated)


def get_rw(gobs, gcalc, weight=None):
    """master_kernel.py:206-236."""
    if weight is None:
        weight = np.ones(gcalc.shape)
    with np.errstate(all='ignore'):
        scale = get_scale(gobs, gcalc)
    if scale <= 0:
        return 1, 1
    top = np.sum(weight * (gobs - scale * gcalc) ** 2)
    bottom = np.sum(weight * gobs ** 2)
    return np.sqrt(top / bottom).real, scale


def get_chi_sq(gobs, gcalc):
    """master_kernel.py:239-266."""
    with np.errstate(all='ignore'):
        scale = get_scale(gobs, gcalc)
    if scale <= 0:
        scale = 1
    return np.sum((gobs - scale * gcalc) ** 2), scale


def get_grad_rw(grad_pdf_arr, gcalc, gobs, rw, scale):
    """master_kernel.py:293-347 (returns the [N,3] array it fills)."""
    n = len(grad_pdf_arr)
    grad_rw = np.zeros((n, 3))
    for tx in range(n):
        for tz in range(3):
            if scale <= 0:
                grad_a = 0
                scale = 1
            else:
                grad_a = get_grad_scale(gobs, gcalc, grad_pdf_arr, tx, tz)
            diff = gobs - scale * gcalc
            grad_rw[tx, tz] = -1 * rw / np.dot(diff, diff) * np.sum(
                (scale * grad_pdf_arr[tx, tz, :] + gcalc * grad_a) * diff)
    return grad_rw


def get_grad_chi_sq(grad_pdf_arr, gcalc, gobs, scale):
    """master_kernel.py:350-375."""
    n = len(grad_pdf_arr)
    grad = np.zeros((n, 3))
    for tx in range(n):
        for tz in range(3):
            grad_a = get_grad_scale(gobs, gcalc, grad_pdf_arr, tx, tz)
            if scale <= 0:
                grad_a = 0
            grad[tx, tz] = -2 * np.sum(
                (scale * grad_pdf_arr[tx, tz, :] + gcalc * grad_a) *
                (gobs - scale * gcalc))
    return grad


# --- calc/__init__.py:10-105 -------------------------------------------------

def wrap_rw(gcalc, gobs):
    return get_rw(gobs, gcalc, weight=None)


def wrap_chi_sq(gcalc, gobs):
    return get_chi_sq(gobs, gcalc)


def wrap_grad_rw(grad_gcalc, gcalc, gobs):
    rw, scale = wrap_rw(gcalc, gobs)
    return get_grad_rw(grad_gcalc, gcalc, gobs, rw, scale)


def wrap_grad_chi_sq(grad_gcalc, gcalc, gobs):
    chi_sq, scale = wrap_chi_sq(gcalc, gobs)
    return get_grad_chi_sq(grad_gcalc, gcalc, gobs, scale)


# --- experiment-level re-drive (elasticscatter/__init__.py) -----------------

DEFAULT_EXP = dict(qmin=0.0, qmax=25, qbin=.1, rmin=0.0, rmax=40.0, rstep=.01,
                   sampling='full')


def pdf_qbin(exp):
    """elasticscatter/__init__.py:203-204."""
    return np.pi / (exp['rmax'] + 6 * 2 * np.pi / exp['qmax'])


def r_grid(exp):
    """elasticscatter/__init__.py:549-558."""
    return np.arange(exp['rmin'], exp['rmax'], exp['rstep'])


def experiment_fq(positions, scatter_fq, exp, precision='fp32', **kw):
    """ElasticScatter.get_fq :304-341 without noise."""
    fq = wrap_fq(positions, scatter_fq, exp['qbin'], precision, **kw)
    return fq[int(np.floor(exp['qmin'] / exp['qbin'])):]


def experiment_grad_fq(positions, scatter_fq, exp, precision='fp32', **kw):
    """ElasticScatter.get_grad_fq :477-496."""
    g = wrap_fq_grad(positions, scatter_fq, exp['qbin'], precision, **kw)
    return g[:, :, int(np.floor(exp['qmin'] / exp['qbin'])):]


def experiment_pdf(positions, scatter_pdf, exp, precision='fp32', **kw):
    """ElasticScatter.get_pdf :343-391 without noise."""
    qb = pdf_qbin(exp)
    fq = wrap_fq(positions, scatter_pdf, qb, precision, **kw)
    return get_pdf_at_qmin(np.array(fq, dtype=np.float64), exp['rstep'], qb,
                           r_grid(exp), exp['qmin'])


def experiment_grad_pdf(positions, scatter_pdf, exp, precision='fp32', **kw):
    """ElasticScatter.get_grad_pdf :498-524."""
    qb = pdf_qbin(exp)
    g = wrap_fq_grad(positions, scatter_pdf, qb, precision, **kw)
    g = np.array(g)
    g[:, :, :int(exp['qmin'] / qb)] = 0.
    return grad_pdf(g, exp['rstep'], qb, r_grid(exp), exp['qmin'])


def calc1d_energy_forces(positions, scatter_pdf, exp, target, potential='rw',
                         conv=1., precision='fp32', **kw):
    """Calc1D.calculate_energy / calculate_forces (calc/calc_1d.py:78-95)
    driven with exp_function=get_pdf, exp_grad_function=get_grad_pdf."""
    gcalc = experiment_pdf(positions, scatter_pdf, exp, precision, **kw)
    ggrad = experiment_grad_pdf(positions, scatter_pdf, exp, precision, **kw)
    if potential == 'rw':
        e, scale = wrap_rw(gcalc, target)
        f = wrap_grad_rw(ggrad, gcalc, target)
    elif potential == 'chi_sq':
        e, scale = wrap_chi_sq(gcalc, target)
        f = wrap_grad_chi_sq(ggrad, gcalc, target)
    else:
        raise NotImplementedError('Potential not implemented')
    return e * conv, f * conv, scale
