"""Host output arrays the GPU writes directly.

The full-gradient kernel owns its rows and stores each of them once, in
coalesced 128-byte pieces; when the destination is pinned, mapped host memory
those stores ARE the device-to-host transfer (150 MB spread over the 187 ms of
the Pt 50k pass = 0.8 GB/s of posted PCIe writes).  This module hands out
numpy arrays backed by such memory:

* ``pinned_empty`` -- one process (one GPU or the one-process multi-GPU
  handle): buffers from ``iid_host_alloc`` kept in a pool.  Every call of
  ``wrap_fq_grad`` must return a FRESH array (the reference's tests assert
  ``ans1 is not ans2``, ``pyiid/tests/test_scatter.py:43``): a buffer goes back
  to the pool only when the array handed out over it, and every view of it,
  has been garbage collected.
* ``SharedOutput`` -- one process per GPU (torchrun): a POSIX shared-memory
  segment that every rank maps and registers with its own CUDA context, so the
  ranks' kernels write their (disjoint) rows into ONE host array, as the
  reference's one-process multi-GPU path assembles one array
  (``gpu_wrappers/gpu_wrap.py:159-194``).  No 150 MB collective, no copy.
"""
import ctypes
import os
import weakref

import numpy as np

from . import _lib
from ._lib import check

# pinned memory is a limited resource: beyond this the callers fall back to
# pageable arrays + staged download
POOL_LIMIT_BYTES = int(os.environ.get('IID_PINNED_POOL_MB', '4096')) << 20


class _Pool(object):
    def __init__(self):
        self.free = {}       # nbytes -> [ptr, ...]
        self.total = 0

    def take(self, nbytes):
        lst = self.free.get(nbytes)
        if lst:
            return lst.pop()
        if self.total + nbytes > POOL_LIMIT_BYTES:
            self.trim()
            if self.total + nbytes > POOL_LIMIT_BYTES:
                return None
        p = ctypes.c_void_p()
        rc = _lib.load().iid_host_alloc(nbytes, ctypes.byref(p))
        if rc != 0 or not p.value:
            return None
        self.total += nbytes
        return p.value

    def give(self, nbytes, ptr):
        self.free.setdefault(nbytes, []).append(ptr)

    def trim(self):
        """Free every idle buffer."""
        lib = _lib.load()
        for nbytes, lst in self.free.items():
            for ptr in lst:
                lib.iid_host_free(ctypes.c_void_p(ptr))
                self.total -= nbytes
        self.free = {}


_pool = _Pool()


def pinned_empty(shape, dtype):
    """Uninitialised array in pinned, mapped host memory, or None when the
    pool is exhausted (caller falls back to pageable memory)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes == 0:
        return None
    ptr = _pool.take(nbytes)
    if ptr is None:
        return None
    # the array and all its views keep `buf` alive through their base chain;
    # when the last of them dies the buffer returns to the pool
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    weakref.finalize(buf, _pool.give, nbytes, ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def is_pinned(arr):
    """True if the GPU can write `arr` directly."""
    p = ctypes.c_void_p()
    return _lib.load().iid_host_device_pointer(
        ctypes.c_void_p(arr.ctypes.data), ctypes.byref(p)) == 0


class SharedOutput(object):
    """A shared host array [shape] that every rank of a torch.distributed
    group writes its rows into.  Two segments alternate, so the array returned
    by one call stays valid until the call after the next.

    The segments are memfd files (not limited by the size of /dev/shm) that
    rank 0 creates; the other ranks open them through /proc/<pid>/fd (all
    ranks of one node share a PID namespace under torchrun) and fall back to
    ``multiprocessing.shared_memory`` names if that is not permitted.  Every
    rank maps them and registers the mapping with its CUDA context.
    ``ok`` is False on every rank if any rank failed to set this up."""

    def __init__(self, shape, dtype, dist, device=None):
        import mmap
        import torch
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        self.dist = dist
        self.turn = 0
        self.segs = []
        self._keep = []
        rank = dist.get_rank()
        lib = _lib.load()
        good = 1
        for k in range(2):
            spec = [None]
            if rank == 0:
                try:
                    fd = os.memfd_create('iid_b200_grad_%d' % k)
                    os.ftruncate(fd, self.nbytes)
                    spec[0] = ('memfd', os.getpid(), fd)
                except OSError:
                    from multiprocessing import shared_memory
                    shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
                    self._keep.append(shm)
                    spec[0] = ('shm', shm.name, 0)
            dist.broadcast_object_list(spec, src=0)
            try:
                kind, a, b = spec[0]
                if kind == 'memfd':
                    if rank != 0:
                        fd = os.open('/proc/%d/fd/%d' % (a, b), os.O_RDWR)
                    mm = mmap.mmap(fd, self.nbytes)
                    self._keep.append((fd, mm))
                    arr = np.frombuffer(mm, dtype=self.dtype).reshape(self.shape)
                else:
                    from multiprocessing import shared_memory
                    if rank != 0:
                        shm = shared_memory.SharedMemory(name=a)
                        self._keep.append(shm)
                    arr = np.ndarray(self.shape, self.dtype, buffer=self._keep[-1].buf)
                check(lib.iid_host_register(ctypes.c_void_p(arr.ctypes.data), self.nbytes))
                dev = ctypes.c_void_p()
                check(lib.iid_host_device_pointer(ctypes.c_void_p(arr.ctypes.data),
                                                  ctypes.byref(dev)))
                self.segs.append((arr, dev.value))
            except (OSError, ValueError, _lib.IIDError):
                good = 0
        flag = torch.tensor([good], dtype=torch.int32,
                            device='cuda' if dist.get_backend() == 'nccl' else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = bool(flag.item())

    def next(self):
        """(fresh view of the array, device alias) of the segment this call
        writes."""
        arr, dev = self.segs[self.turn]
        self.turn ^= 1
        return arr.view(), dev
