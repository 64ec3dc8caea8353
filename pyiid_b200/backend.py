"""Host side of the B200 processor: owns the native handle and the caches.

Replaces the reference's processor wrappers
(``pyiid/experiments/elasticscatter/gpu_wrappers/gpu_wrap.py:119-314`` and the
chunk workers ``atomics/gpu_atomics.py:89-280``): instead of chunking the pair
list by free memory and farming chunks to one Python thread per GPU, one
process drives one GPU through the C ABI (``include/iid_b200.h``), and with
``torch.distributed`` initialised each rank computes its slice of the
pair-tile work list and the partial F(Q) / force / gradient arrays are
all-reduced over NCCL.

PyTorch is used only when world_size > 1 (device buffers for the collective);
the single-GPU path talks to the library with numpy host buffers.
"""
import atexit
import collections
import ctypes
import math
import os

import numpy as np

from . import _lib
from . import hostmem
from ._lib import IID_FP32, IID_FP64, IID_POT_RW, IID_POT_CHI_SQ, check

POTENTIALS = {'rw': IID_POT_RW, 'chi_sq': IID_POT_CHI_SQ}
SPRING_TYPES = {'rep': _lib.IID_SPRING_REP, 'com': _lib.IID_SPRING_COM,
                'att': _lib.IID_SPRING_ATT}
SPRING_TYPES.update({v: v for v in list(SPRING_TYPES.values())})


def _dist_state():
    """(rank, world) of an initialised torch.distributed group, else (0, 1)."""
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def pdf_matrix(nq, rstep, qstep, rgrid, qmin):
    """Dense T[R, Q] with ``G = T @ F`` equal to the reference's
    ``get_pdf_at_qmin(F, rstep, qstep, rgrid, qmin)``
    (``kernels/master_kernel.py:39-104`` with ``fft_fq_to_gr :108-128`` and
    ``fft_gr_to_fq :131-203``).

    The reference zeroes F below ceil(qmin/qstep), zero-pads it to
    L = max(len F, ceil(pi/rstep/qstep)), odd-extends it into a 4*npad2 real
    array on the even slots, takes ``np.fft.ifft`` and keeps
    ``imag[::2]*npad2*qstep*2/pi``, i.e.
    ``g[i] = (qstep/pi) * sum_k F[k] sin(2 pi k i / npad2)``; it then
    interpolates linearly at ``rgrid/(2 drpad)``, ``drpad = pi/(npad2 qstep)``,
    and doubles.  The phases k*i mod npad2 are reduced in integers, so the
    matrix is exact to float64 rounding.
    """
    rgrid = np.asarray(rgrid, dtype=np.float64)
    kmin = int(math.ceil(qmin / qstep))
    nfromdr = int(math.ceil(math.pi / rstep / qstep))
    length = max(int(nq), nfromdr)
    padrmin = int(round(qmin / qstep))
    npad1 = padrmin + length
    npad2 = (1 << int(math.ceil(math.log(npad1, 2)))) * 2
    drpad = math.pi / (npad2 * qstep)
    axdrp = rgrid / drpad / 2
    ilo = axdrp.astype(int)
    whi = axdrp - ilo
    wlo = 1.0 - whi
    if len(ilo) and ilo.max() + 1 >= npad2:
        raise IndexError('rgrid reaches beyond the padded transform length')
    k = np.arange(int(nq), dtype=np.int64)
    ph_lo = (ilo[:, None].astype(np.int64) * k[None, :]) % npad2
    ph_hi = ((ilo[:, None].astype(np.int64) + 1) * k[None, :]) % npad2
    w = 2.0 * np.pi / npad2
    t = wlo[:, None] * np.sin(w * ph_lo) + whi[:, None] * np.sin(w * ph_hi)
    t *= 2.0 * qstep / math.pi
    t[:, :kmin] = 0.0
    return np.ascontiguousarray(t)


def element_table(scatter_array, numbers=None):
    """Split a per-atom scatter-factor array [N, Q] into a per-element table
    [E, Q] and an index [N].  Rows are grouped by atomic number when that is
    consistent (the usual case: ``_wrap_atoms`` broadcasts one row per
    element), otherwise by unique rows."""
    scat = np.asarray(scatter_array)
    n = scat.shape[0]
    if numbers is not None and len(numbers) == n:
        numbers = np.asarray(numbers)
        zs, first, inv = np.unique(numbers, return_index=True, return_inverse=True)
        table = scat[first]
        if np.array_equal(table[inv], scat):
            return (np.ascontiguousarray(table, dtype=np.float64),
                    np.ascontiguousarray(inv, dtype=np.int32))
    table, inv = np.unique(scat, axis=0, return_inverse=True)
    return (np.ascontiguousarray(table, dtype=np.float64),
            np.ascontiguousarray(inv.reshape(-1), dtype=np.int32))


ADP_CLASSES_MAX = 32


def adp_tables(table, idx, adps, qbin):
    """Per-(element, displacement) classes: pair table f(Q) exp(-u^2 Q^2 / 2),
    normaliser table f(Q), class index per atom."""
    nq = table.shape[1]
    pairs = np.stack([idx.astype(np.float64), adps], axis=1)
    classes, inv = np.unique(pairs, axis=0, return_inverse=True)
    if len(classes) > ADP_CLASSES_MAX:
        raise ValueError('%d distinct (element, displacement) classes; at most %d '
                         '(atoms are tiled by class: quantise the displacements)'
                         % (len(classes), ADP_CLASSES_MAX))
    q = np.arange(nq, dtype=np.float64) * qbin
    norm = np.ascontiguousarray(table[classes[:, 0].astype(int)], dtype=np.float64)
    pair = np.ascontiguousarray(norm * np.exp(-0.5 * classes[:, 1:2] * q[None, :] ** 2))
    return pair, norm, np.ascontiguousarray(inv.reshape(-1), dtype=np.int32)


def visible_devices():
    """Number of CUDA devices the library sees (0 without a GPU)."""
    cnt = ctypes.c_int(0)
    try:
        if _lib.load().iid_device_count(ctypes.byref(cnt)) != 0:
            return 0
    except _lib.IIDError:
        return 0
    return cnt.value


class Backend(object):
    """One native handle for one precision: on one GPU, or -- ``device='multi'``
    -- over every GPU of the box from this one process (``iid_create_multi``,
    the reference's ``gpu_wrap.py`` thread-per-GPU farm)."""
    _instances = {}

    MAX_PER_KEY = 4  # live handles per (precision, device, Q grid)

    @classmethod
    def close_all(cls):
        """Destroy every native handle (registered with atexit)."""
        for pool in list(cls._instances.values()):
            for inst in pool.values():
                inst.close()
        cls._instances.clear()

    def close(self):
        if self.h is not None and self.h.value:
            self.lib.iid_destroy(self.h)
            self.h = ctypes.c_void_p()

    @classmethod
    def get(cls, precision='fp32', device=None, slot='fq', owner=None):
        """The handle of ``owner`` (an ElasticScatter object's id) for this
        precision, device and Q grid.  Every owner gets its own handle, so two
        ElasticScatter objects with different structures or experiments do not
        re-upload on every alternate call; beyond MAX_PER_KEY the least
        recently used handle is handed on."""
        if device == 'multi' and _dist_state()[1] > 1:
            device = None  # one process per GPU already: this rank's device
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', '0'))
            try:
                lib = _lib.load()
                cnt = ctypes.c_int(0)
                if lib.iid_device_count(ctypes.byref(cnt)) == 0 and cnt.value:
                    device %= cnt.value
            except _lib.IIDError:
                raise
        # one handle per Q grid in use ('fq' / 'pdf'), so that alternating
        # get_fq / get_pdf calls do not re-upload the structure
        device = device if device == 'multi' else int(device)
        key = (precision, device, slot)
        pool = cls._instances.setdefault(key, collections.OrderedDict())
        inst = pool.pop(owner, None)
        if inst is None:
            if len(pool) < cls.MAX_PER_KEY:
                inst = cls(precision, device)
            else:
                _, inst = pool.popitem(last=False)
        pool[owner] = inst  # most recently used last
        return inst

    def __init__(self, precision='fp32', device=0):
        if precision not in ('fp32', 'fp64'):
            raise ValueError("precision must be 'fp32' or 'fp64'")
        self.lib = _lib.load()
        self.precision = precision
        self.multi = device == 'multi'
        self.device = 0 if self.multi else device
        self.h = ctypes.c_void_p()
        prec = IID_FP32 if precision == 'fp32' else IID_FP64
        if self.multi:
            check(self.lib.iid_create_multi(0, prec, ctypes.byref(self.h)))
        else:
            check(self.lib.iid_create(device, prec, ctypes.byref(self.h)))
        self.gdtype = np.float32 if precision == 'fp32' else np.float64
        self._skey = None
        self._tkey = None
        self._last_scat = None
        self._last_qbin = None
        self._last_numbers = None
        self._last_adps = None
        self._last_target = None
        self._target_key = None
        self._restraints = []
        self._chain = None
        self.restraint_energy = 0.0
        self.n = self.nq = self.nr = 0
        self.rank, self.world = 0, 1
        self._tensors = {}
        self._shared = {}
        self._ext = None
        self.sync_shard()

    # -- sharding -----------------------------------------------------------
    def devices(self):
        """(devices behind the handle, devices the current structure uses)."""
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        check(self.lib.iid_handle_devices(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def sync_shard(self):
        if self.multi:
            return  # shards over its own devices
        rank, world = _dist_state()
        if (rank, world) != (self.rank, self.world):
            check(self.lib.iid_set_shard(self.h, rank, world))
            self.rank, self.world = rank, world

    # -- cached state -------------------------------------------------------
    def set_structure(self, scatter_array, numbers, qbin, adps=None):
        """``adps``: isotropic mean-square displacements <u^2> per atom (A^2) or
        None.  Their Debye-Waller factor tau_ij(Q) = exp(-(u_i^2 + u_j^2) Q^2 / 2)
        = t_i t_j multiplies the pair term, not the normaliser
        (``get_adp_fq``, kernels/cpu_nxn.py:114-121): the pair sums get one table
        row f t per (element, displacement) class, na the plain f."""
        scat = np.asarray(scatter_array)
        if adps is not None:
            adps = np.ascontiguousarray(adps, dtype=np.float64).reshape(-1)
            if adps.shape[0] != scat.shape[0]:
                raise ValueError('adps: one mean-square displacement per atom')
            if not np.all(np.isfinite(adps)) or np.any(adps < 0):
                raise ValueError('adps: mean-square displacements must be finite and >= 0')
            if not np.any(adps):
                adps = None
        # fast path: the very same (shared, read-only) table object as last time
        if scat is self._last_scat and float(qbin) == self._last_qbin and \
                np.array_equal(numbers, self._last_numbers) and \
                (adps is None) == (self._last_adps is None) and \
                (adps is None or np.array_equal(adps, self._last_adps)):
            return
        table, idx = element_table(scat, numbers)
        norm = None
        if adps is not None:
            table, norm, idx = adp_tables(table, idx, adps, float(qbin))
        key = (scat.shape, float(qbin), idx.tobytes(), table.tobytes(),
               None if norm is None else norm.tobytes())
        self._last_scat, self._last_qbin = scat, float(qbin)
        self._last_numbers = np.array(numbers)
        self._last_adps = None if adps is None else adps.copy()
        if key == self._skey:
            return
        n, nq = scat.shape
        check(self.lib.iid_set_structure_norm(
            self.h, n, idx.ctypes.data, table.shape[0], table.ctypes.data,
            None if norm is None else norm.ctypes.data, nq, float(qbin)))
        if nq != self.nq:
            self._tkey = None
            self.nr = 0
        self._skey = key
        self._target_key = None
        self._last_target = None
        self._sampler_key = None  # the native state slots are sized per structure
        self.n, self.nq = n, nq
        self._tensors = {}

    def set_transform(self, rstep, qstep, rgrid, qmin):
        rgrid = np.asarray(rgrid, dtype=np.float64)
        key = (self.nq, float(rstep), float(qstep), float(qmin), rgrid.tobytes())
        if key == self._tkey:
            return
        t = pdf_matrix(self.nq, rstep, qstep, rgrid, qmin)
        check(self.lib.iid_set_transform(self.h, t.shape[0], t.shape[1],
                                         t.ctypes.data))
        self._tkey = key
        self._target_key = None
        self._last_target = None
        self.nr = t.shape[0]
        self._tensors = {}

    @staticmethod
    def _pos(positions):
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        if pos.ndim != 2 or pos.shape[1] != 3:
            raise ValueError('positions must be [N, 3]')
        return pos

    # -- torch plumbing for world > 1 ----------------------------------------
    def _t(self, name, shape, dtype):
        import torch
        t = self._tensors.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.zeros(shape, dtype=dtype, device='cuda:%d' % self.device)
            self._tensors[name] = t
        return t

    def _stream(self):
        """NULL = the handle's own stream (see _on_stream)."""
        return None

    def _on_stream(self):
        """Context that makes the handle's stream torch's current stream, so
        torch copies / NCCL collectives and our kernels are ordered on ONE
        stream (torch's default stream 0 does not order with it)."""
        import torch
        if self._ext is None:
            sp = ctypes.c_void_p()
            check(self.lib.iid_get_stream(self.h, ctypes.byref(sp)))
            self._ext = torch.cuda.ExternalStream(sp.value, device=self.device)
        return torch.cuda.stream(self._ext)

    def _upload(self, pos):
        import torch
        t = self._t('pos', (self.n, 3), torch.float64)
        t.copy_(torch.from_numpy(pos))
        return t

    # -- the three bound callables + fused paths ------------------------------
    def fq(self, positions):
        """F(Q) [nq] float64 (flat_multi_cpu_wrap.wrap_fq :22-60)."""
        pos = self._pos(positions)
        self.sync_shard()
        out = np.empty(self.nq, np.float64)
        if self.world == 1:
            check(self.lib.iid_fq_host(self.h, pos.ctypes.data, out.ctypes.data))
            return out
        import torch
        import torch.distributed as dist
        with torch.cuda.device(self.device), self._on_stream():
            p = self._upload(pos)
            s = self._t('S', (self.nq,), torch.float64)
            f = self._t('F', (self.nq,), torch.float64)
            st = self._stream()
            check(self.lib.iid_fq_partial(self.h, p.data_ptr(), s.data_ptr(), st))
            dist.all_reduce(s)
            check(self.lib.iid_fq_finish(self.h, s.data_ptr(), f.data_ptr(), st))
            return f.cpu().numpy()

    def grad_fq(self, positions, with_fq=False, root_only=False):
        """grad F(Q) [N,3,nq] in the kernel precision
        (flat_multi_cpu_wrap.wrap_fq_grad :63-102).

        One process: the array is pinned, mapped host memory from a pool and
        the kernel stores its rows straight into it (``hostmem.pinned_empty``;
        pageable memory + staged download when the pool is exhausted).

        One process per GPU (torch.distributed): every rank computes the rows
        of its i-tiles.  Default: each rank returns its own full copy
        (all-reduce of the disjoint rows + one download per rank).
        ``root_only``: all ranks write their rows into ONE shared host array
        (``hostmem.SharedOutput``) -- the single array the reference's
        one-process multi-GPU path assembles (gpu_wrap.py:159-194, 287-314) --
        without any gradient collective; every rank gets a view of it, valid
        until the call after the next."""
        pos = self._pos(positions)
        self.sync_shard()
        f = np.empty(self.nq, np.float64)
        if self.world == 1:
            g = hostmem.pinned_empty((self.n, 3, self.nq), self.gdtype)
            if g is None:
                g = np.empty((self.n, 3, self.nq), self.gdtype)
            check(self.lib.iid_grad_fq_host(self.h, pos.ctypes.data, g.ctypes.data,
                                            f.ctypes.data))
            return (g, f) if with_fq else g
        import torch
        import torch.distributed as dist
        tdt = torch.float32 if self.precision == 'fp32' else torch.float64
        shared = self._shared_out() if root_only else None
        with torch.cuda.device(self.device), self._on_stream():
            p = self._upload(pos)
            s = self._t('S', (self.nq,), torch.float64)
            ft = self._t('F', (self.nq,), torch.float64)
            st = self._stream()
            if shared is not None:
                g, gptr = shared.next()
            else:
                gt = self._t('G', (self.n, 3, self.nq), tdt)
                gt.zero_()  # rows of the other ranks' atoms are not written here
                gptr = gt.data_ptr()
            check(self.lib.iid_grad_fq_partial(self.h, p.data_ptr(), gptr, s.data_ptr(), st))
            dist.all_reduce(s)
            check(self.lib.iid_fq_finish(self.h, s.data_ptr(), ft.data_ptr(), st))
            if shared is None:
                dist.all_reduce(gt)
                g = np.empty((self.n, 3, self.nq), self.gdtype)
                check(self.lib.iid_download_host(self.h, gt.data_ptr(), g.ctypes.data,
                                                 g.nbytes))
            f = ft.cpu().numpy()  # synchronises this rank's stream: its rows are written
            if shared is not None:
                dist.barrier()    # ... and so are everybody else's
        return (g, f) if with_fq else g

    def _shared_out(self):
        """The shared output arrays of this structure (collective on first
        use); None when they cannot be set up on every rank."""
        import torch.distributed as dist
        key = (self.n, self.nq)
        so = self._shared.get(key)
        if so is None:
            so = hostmem.SharedOutput((self.n, 3, self.nq), self.gdtype, dist, self.device)
            self._shared[key] = so
        return so if so.ok else None

    def pdf(self, positions, with_fq=False):
        """G(r) [nr] float64 = get_pdf_at_qmin(F(Q)) (master_kernel.py:39-104)."""
        pos = self._pos(positions)
        self.sync_shard()
        if self.nr == 0:
            raise _lib.IIDError('set_transform has not been called')
        g = np.empty(self.nr, np.float64)
        f = np.empty(self.nq, np.float64)
        if self.world == 1:
            check(self.lib.iid_pdf_host(self.h, pos.ctypes.data, g.ctypes.data,
                                        f.ctypes.data))
            return (g, f) if with_fq else g
        import torch
        import torch.distributed as dist
        with torch.cuda.device(self.device), self._on_stream():
            p = self._upload(pos)
            s = self._t('S', (self.nq,), torch.float64)
            ft = self._t('F', (self.nq,), torch.float64)
            gr = self._t('Gr', (self.nr,), torch.float64)
            st = self._stream()
            check(self.lib.iid_fq_partial(self.h, p.data_ptr(), s.data_ptr(), st))
            dist.all_reduce(s)
            check(self.lib.iid_fq_finish(self.h, s.data_ptr(), ft.data_ptr(), st))
            check(self.lib.iid_fq_to_gr(self.h, ft.data_ptr(), gr.data_ptr(), st))
            g = gr.cpu().numpy()
            f = ft.cpu().numpy()
        return (g, f) if with_fq else g

    def energy_forces(self, positions, target, potential='rw', conv=1.,
                      want_forces=True, want_pdf=False):
        """Fused Calc1D evaluation (calc/calc_1d.py:78-95 with
        exp_function=get_pdf, exp_grad_function=get_grad_pdf): returns
        (energy, scale, forces[N,3] or None, pdf or None) from ONE F(Q) pass
        and ONE force pass, never forming the N x 3 x Q / N x 3 x R arrays."""
        if potential not in POTENTIALS:
            raise NotImplementedError('Potential not implemented')
        pos = self._pos(positions)
        self.sync_shard()
        if self.nr == 0:
            raise _lib.IIDError('set_transform has not been called')
        target = np.ascontiguousarray(target, dtype=np.float64)
        if target.shape != (self.nr,):
            raise ValueError('target must have the r-grid length %d' % self.nr)
        out = np.zeros(4, np.float64)
        forces = np.empty((self.n, 3), np.float64) if want_forces else None
        pdf = np.empty(self.nr, np.float64) if want_pdf else None
        pot = POTENTIALS[potential]
        if self.world == 1:
            tptr, tkey = self._target_ptr(target)
            check(self.lib.iid_energy_forces_host(
                self.h, pos.ctypes.data, tptr, pot, float(conv), out.ctypes.data,
                forces.ctypes.data if want_forces else None,
                pdf.ctypes.data if want_pdf else None))
            self._target_done(target, tkey)
            if self._restraints:
                self._read_restraint_energy()
            return out[0], out[1], forces, pdf
        import torch
        import torch.distributed as dist
        with torch.cuda.device(self.device), self._on_stream():
            p = self._upload(pos)
            s = self._t('S', (self.nq,), torch.float64)
            ft = self._t('F', (self.nq,), torch.float64)
            gr = self._t('Gr', (self.nr,), torch.float64)
            tg = self._t('target', (self.nr,), torch.float64)
            o4 = self._t('out4', (4,), torch.float64)
            wq = self._t('wq', (self.nq,), torch.float64)
            fo = self._t('force', (self.n, 3), torch.float64)
            tg.copy_(torch.from_numpy(target))
            st = self._stream()
            check(self.lib.iid_fq_partial(self.h, p.data_ptr(), s.data_ptr(), st))
            dist.all_reduce(s)
            check(self.lib.iid_fq_finish(self.h, s.data_ptr(), ft.data_ptr(), st))
            check(self.lib.iid_fq_to_gr(self.h, ft.data_ptr(), gr.data_ptr(), st))
            check(self.lib.iid_potential(self.h, gr.data_ptr(), tg.data_ptr(), pot,
                                         float(conv), o4.data_ptr(),
                                         wq.data_ptr() if want_forces else None, st))
            if want_forces:
                check(self.lib.iid_force_partial(self.h, p.data_ptr(), wq.data_ptr(),
                                                 fo.data_ptr(), st))
            if self._restraints:
                es = self._t('e_sp', (len(self._restraints),), torch.float64)
                fs = self._t('force_sp', (self.n, 3), torch.float64)
                for i, (ty, kk, rt) in enumerate(self._restraints):
                    check(self.lib.iid_spring_partial(
                        self.h, p.data_ptr(), self.n, ty, kk, rt, None,
                        es.data_ptr() + 8 * i,
                        fs.data_ptr() if want_forces else None, None, st))
                    if want_forces:
                        fo += fs
                dist.all_reduce(es)
                self.restraint_energy = float(es.sum().item())
            if want_forces:
                dist.all_reduce(fo)
                forces = fo.cpu().numpy()
            out = o4.cpu().numpy()
            if want_pdf:
                pdf = gr.cpu().numpy()
        return out[0], out[1], forces, pdf

    def _target_ptr(self, target):
        """(pointer or None, key): None when this target is already resident on
        the device.  A READ-ONLY array object seen before is recognised by
        identity (Calc1D keeps its target that way: the sampler's hot path);
        a writable one may have been modified in place since, so its content
        is compared with the private copy of what was uploaded."""
        if target is self._last_target and not target.flags.writeable:
            return None, self._target_key
        if self._target_key is not None and target.shape == self._target_key.shape and \
                np.array_equal(target, self._target_key):
            return None, self._target_key
        return target.ctypes.data, target

    def _target_done(self, target, tkey):
        """The native call succeeded: remember what is resident."""
        if tkey is not self._target_key:
            self._target_key = np.array(tkey, dtype=np.float64)  # private copy
        self._last_target = target

    # -- device-resident sampler states (pyiid/sim/__init__.py:10-38) ------------
    def sampler_setup(self, n_slots, masses, cell_centre):
        """Allocate ``n_slots`` phase-space slots (q, p, f) on the device."""
        m = np.ascontiguousarray(masses, dtype=np.float64).reshape(-1)
        c = np.ascontiguousarray(cell_centre, dtype=np.float64).reshape(3)
        if m.shape != (self.n,):
            raise ValueError('masses must have one entry per atom')
        check(self.lib.iid_sampler_setup(self.h, int(n_slots), m.ctypes.data, c.ctypes.data))
        self.n_slots = int(n_slots)

    def state_upload(self, slot, q, p, f):
        q, p, f = (np.ascontiguousarray(a, dtype=np.float64).reshape(self.n, 3)
                   for a in (q, p, f))
        check(self.lib.iid_state_upload(self.h, int(slot), q.ctypes.data, p.ctypes.data,
                                        f.ctypes.data))

    def state_download(self, slot, want=('q', 'p', 'f')):
        out = {k: np.empty((self.n, 3), np.float64) for k in want}
        check(self.lib.iid_state_download(
            self.h, int(slot), *[out[k].ctypes.data if k in out else None for k in 'qpf']))
        return out

    def leapfrog(self, src, dst, step, center, target, potential='rw', conv=1.):
        """One kick-drift-kick step from slot ``src`` into slot ``dst`` on the
        device; returns (energy, scale, restraint energy, kinetic energy, q, p)
        of the new state (q, p are host mirrors for the U-turn test)."""
        if potential not in POTENTIALS:
            raise NotImplementedError('Potential not implemented')
        if target is not self._last_target:  # a new target object: validate it once
            target = np.ascontiguousarray(target, dtype=np.float64)
            if target.shape != (self.nr,):
                raise ValueError('target must have the r-grid length %d' % self.nr)
            self.sync_shard()
        if self.world != 1:
            raise _lib.IIDError('the device-resident leapfrog needs world == 1')
        tptr, tkey = self._target_ptr(target)
        out = np.empty(9, np.float64)
        q = np.empty((self.n, 3), np.float64)
        p = np.empty((self.n, 3), np.float64)
        check(self.lib.iid_leapfrog_host(
            self.h, int(src), int(dst), float(step), int(bool(center)), tptr,
            POTENTIALS[potential], float(conv), out.ctypes.data, q.ctypes.data, p.ctypes.data))
        self._target_done(target, tkey)
        return out[0], out[1], out[4], out[5], q, p

    def leapfrog_chain(self, src, dsts, step, center, target, potential='rw', conv=1.):
        """``len(dsts)`` consecutive leapfrog steps src -> dsts[0] -> dsts[1] ...
        queued behind each other with one synchronisation
        (``iid_leapfrog_chain_host``); returns a list of
        (energy, scale, restraint energy, kinetic energy, q, p), one per step."""
        if potential not in POTENTIALS:
            raise NotImplementedError('Potential not implemented')
        if target is not self._last_target:
            target = np.ascontiguousarray(target, dtype=np.float64)
            if target.shape != (self.nr,):
                raise ValueError('target must have the r-grid length %d' % self.nr)
            self.sync_shard()
        if self.world != 1:
            raise _lib.IIDError('the device-resident leapfrog needs world == 1')
        m = len(dsts)
        tptr, tkey = self._target_ptr(target)
        d = np.asarray(dsts, dtype=np.int32)
        out = np.empty((m, 9), np.float64)
        q = np.empty((m, self.n, 3), np.float64)
        p = np.empty((m, self.n, 3), np.float64)
        check(self.lib.iid_leapfrog_chain_host(
            self.h, int(src), d.ctypes.data, m, float(step), int(bool(center)), tptr,
            POTENTIALS[potential], float(conv), out.ctypes.data, q.ctypes.data, p.ctypes.data))
        self._target_done(target, tkey)
        return [(out[i, 0], out[i, 1], out[i, 4], out[i, 5], q[i], p[i]) for i in range(m)]

    def leapfrog_chain_begin(self, src, dsts, step, center, target, potential='rw', conv=1.):
        """Enqueue ``len(dsts)`` consecutive leapfrog steps and return at once
        (``iid_leapfrog_chain_begin``); :meth:`leapfrog_chain_next` hands out
        the steps in order as the device completes them.  Any other call on
        this backend first waits for the chain and drops what was not
        collected.  Returns the chain's id."""
        if potential not in POTENTIALS:
            raise NotImplementedError('Potential not implemented')
        if target is not self._last_target:
            target = np.ascontiguousarray(target, dtype=np.float64)
            if target.shape != (self.nr,):
                raise ValueError('target must have the r-grid length %d' % self.nr)
            self.sync_shard()
        if self.world != 1:
            raise _lib.IIDError('the device-resident leapfrog needs world == 1')
        tptr, tkey = self._target_ptr(target)
        d = np.asarray(dsts, dtype=np.int32)
        cid = ctypes.c_int64(0)
        check(self.lib.iid_leapfrog_chain_begin(
            self.h, int(src), d.ctypes.data, len(d), float(step), int(bool(center)), tptr,
            POTENTIALS[potential], float(conv), ctypes.byref(cid)))
        self._target_done(target, tkey)
        # result rows of the whole chain in three allocations (this runs once per
        # leapfrog of a sampler: per-step allocations and pointer look-ups count)
        m = len(d)
        out = np.empty((m, 9), np.float64)
        q = np.empty((m, self.n, 3), np.float64)
        p = np.empty((m, self.n, 3), np.float64)
        self._chain = [cid.value, 0, out, q, p, out.ctypes.data, q.ctypes.data, p.ctypes.data,
                       self.n * 24]
        return cid.value

    def leapfrog_chain_next(self, chain_id):
        """The next completed step of chain ``chain_id``: (energy, scale,
        restraint energy, kinetic energy, q, p); raises
        :class:`_lib.ChainDropped` when another call dropped the chain."""
        c = self._chain
        if c is None or c[0] != chain_id:
            raise _lib.ChainDropped('this leapfrog chain is not in flight (any more)')
        i = c[1]
        rc = self.lib.iid_leapfrog_chain_next(self.h, chain_id, c[5] + 72 * i, c[6] + c[8] * i,
                                              c[7] + c[8] * i)
        if rc:
            self._chain = None
            check(rc)
        c[1] = i + 1
        o = c[2][i]
        return o[0], o[1], o[4], o[5], c[3][i], c[4][i]

    # -- spring restraints (calc/spring_calc.py) -------------------------------
    def set_restraints(self, springs):
        """rep / att springs [(sp_type, k, rt), ...] evaluated inside
        energy_forces (same CUDA graph, forces summed on the device)."""
        springs = [(SPRING_TYPES[t], float(k), float(rt)) for t, k, rt in springs]
        if springs != self._restraints:
            ty = np.array([s[0] for s in springs], np.int32)
            kk = np.array([s[1] for s in springs], np.float64)
            rt = np.array([s[2] for s in springs], np.float64)
            check(self.lib.iid_set_restraints(self.h, len(springs), ty.ctypes.data,
                                              kk.ctypes.data, rt.ctypes.data))
            self._restraints = springs
        self.restraint_energy = 0.0

    def _read_restraint_energy(self):
        e = ctypes.c_double(0.0)
        check(self.lib.iid_get_restraint_energy(self.h, ctypes.byref(e)))
        self.restraint_energy = e.value

    def spring(self, positions, sp_type, k, rt, com=None, want_energy=True,
               want_forces=False, want_atomwise=False):
        """(energy, forces[N,3], atomwise[N]) of one spring restraint; entries
        not asked for are None."""
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        n = len(pos)
        ty = SPRING_TYPES[sp_type]
        cm = None if com is None else np.ascontiguousarray(com, dtype=np.float64)
        if ty == _lib.IID_SPRING_COM and (cm is None or cm.shape != (3,)):
            raise ValueError('the com spring needs the centre of mass')
        cptr = None if cm is None else cm.ctypes.data
        self.sync_shard()
        if self.world == 1:
            e = ctypes.c_double(0.0)
            f = np.zeros((n, 3), np.float64) if want_forces else None
            a = np.zeros(n, np.float64) if want_atomwise else None
            check(self.lib.iid_spring_host(
                self.h, pos.ctypes.data, n, ty, float(k), float(rt), cptr,
                ctypes.byref(e) if want_energy else None,
                f.ctypes.data if want_forces else None,
                a.ctypes.data if want_atomwise else None))
            return (e.value if want_energy else None), f, a
        import torch
        import torch.distributed as dist
        with torch.cuda.device(self.device), self._on_stream():
            dev = 'cuda:%d' % self.device
            p = torch.from_numpy(pos).to(dev)
            buf = torch.zeros(4 * n + 1, dtype=torch.float64, device=dev)
            base = buf.data_ptr()
            check(self.lib.iid_spring_partial(
                self.h, p.data_ptr(), n, ty, float(k), float(rt), cptr,
                base + 32 * n if want_energy else None,
                base if want_forces else None,
                base + 24 * n if want_atomwise else None, self._stream()))
            dist.all_reduce(buf)
            out = buf.cpu().numpy()
        return (float(out[4 * n]) if want_energy else None,
                out[:3 * n].reshape(n, 3).copy() if want_forces else None,
                out[3 * n:4 * n].copy() if want_atomwise else None)

    def spring_voxels(self, positions, sp_type, k, rt, resolution, shape, com=None):
        """Energy added by a probe atom at every voxel centre."""
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        nx, ny, nz = (int(v) for v in shape)
        cm = None if com is None else np.ascontiguousarray(com, dtype=np.float64)
        vox = np.zeros((nx, ny, nz), np.float64)
        check(self.lib.iid_spring_voxel_host(
            self.h, pos.ctypes.data, len(pos), SPRING_TYPES[sp_type], float(k),
            float(rt), None if cm is None else cm.ctypes.data, float(resolution),
            nx, ny, nz, vox.ctypes.data))
        return vox

    def grad_pdf(self, grad_fq):
        """[rows.., R] = grad_fq[rows.., Q] . T^T on the device
        (master_kernel.grad_pdf :276-290)."""
        import torch
        if self.nr == 0:
            raise _lib.IIDError('set_transform has not been called')
        g = np.ascontiguousarray(grad_fq, dtype=self.gdtype)
        if g.shape[-1] != self.nq:
            raise ValueError('grad_fq last axis must be %d' % self.nq)
        rows = int(np.prod(g.shape[:-1]))
        with torch.cuda.device(self.device), self._on_stream():
            gin = torch.from_numpy(g.reshape(rows, self.nq)).to('cuda:%d' % self.device)
            out = torch.empty((rows, self.nr), dtype=torch.float64,
                              device='cuda:%d' % self.device)
            check(self.lib.iid_grad_pdf(self.h, gin.data_ptr(), rows, out.data_ptr(),
                                        self._stream()))
            res = out.cpu().numpy()
        return res.reshape(g.shape[:-1] + (self.nr,))

    def grad_pdf_of(self, positions, qmin_bin=0):
        """grad G(r) [N,3,R] float64 straight from the positions: the full
        gradient of F(Q) stays on the device between the pair sum and the
        contraction with T (ElasticScatter.get_grad_pdf :498-524 does
        grad -> host -> grad_pdf -> device again when the two callables are
        bound separately).  One GPU, one shard; returns None otherwise."""
        if self.multi or self.world != 1 or self.nr == 0:
            return None
        import torch
        pos = self._pos(positions)
        tdt = torch.float32 if self.precision == 'fp32' else torch.float64
        dev = 'cuda:%d' % self.device
        with torch.cuda.device(self.device), self._on_stream():
            p = torch.from_numpy(pos).to(dev)
            g = torch.empty((self.n, 3, self.nq), dtype=tdt, device=dev)
            check(self.lib.iid_grad_fq_partial(self.h, p.data_ptr(), g.data_ptr(), None,
                                               self._stream()))
            if qmin_bin > 0:
                g[:, :, :qmin_bin] = 0
            out = torch.empty((self.n * 3, self.nr), dtype=torch.float64, device=dev)
            check(self.lib.iid_grad_pdf(self.h, g.data_ptr(), self.n * 3, out.data_ptr(),
                                        self._stream()))
            res = out.cpu().numpy()
        return res.reshape(self.n, 3, self.nr)

    def gr_from_fq(self, fq):
        """G(r) = T F for a host F(Q) (get_pdf's noise branch)."""
        if self.nr == 0:
            raise _lib.IIDError('set_transform has not been called')
        f = np.ascontiguousarray(fq, dtype=np.float64)
        if f.shape != (self.nq,):
            raise ValueError('F(Q) must have %d bins' % self.nq)
        g = np.empty(self.nr, np.float64)
        check(self.lib.iid_fq_to_gr_host(self.h, f.ctypes.data, g.ctypes.data))
        return g

    def set_option(self, key, value):
        """Native tunables, see iid_set_option in include/iid_b200.h."""
        check(self.lib.iid_set_option(self.h, key.encode(), int(value)))

    # -- instrumentation ------------------------------------------------------
    def launch_count(self):
        c = ctypes.c_int64(0)
        check(self.lib.iid_launch_count(self.h, ctypes.byref(c)))
        return c.value

    def set_timing(self, on):
        check(self.lib.iid_set_timing(self.h, int(bool(on))))

    def last_kernel_ms(self):
        ms = ctypes.c_float(0)
        pq = ctypes.c_double(0)
        check(self.lib.iid_last_kernel_ms(self.h, ctypes.byref(ms), ctypes.byref(pq)))
        return ms.value, pq.value

    def measure_peaks(self):
        """Measured lane-FMA/s of the scalar FFMA, packed FFMA2 and DFMA pipes."""
        out = np.zeros(3, np.float64)
        check(self.lib.iid_measure_peaks(self.h, out.ctypes.data))
        return {'ffma': out[0], 'ffma2': out[1], 'dfma': out[2]}

    def sizes(self):
        v = [ctypes.c_int64(0) for _ in range(5)]
        check(self.lib.iid_get_sizes(self.h, *[ctypes.byref(x) for x in v]))
        return dict(zip(('n', 'nq', 'nr', 'items_fq', 'items_grad'),
                        [x.value for x in v]))


atexit.register(Backend.close_all)
