"""ctypes binding of the C ABI declared in ``include/iid_b200.h``.

The shared library ``libiid_b200.so`` is built in-tree by
``__graft_entry__.build()`` (nvcc, sm_100a).  There is no fallback: if the
library is missing, or no sm_100 device is present, every compute entry point
raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IID_LIB_PATH: a differently built copy of the library (developer A/B runs)
LIB_PATH = os.environ.get('IID_LIB_PATH') or os.path.join(_HERE, 'libiid_b200.so')

IID_FP32, IID_FP64 = 0, 1
IID_POT_RW, IID_POT_CHI_SQ = 0, 1
IID_SPRING_REP, IID_SPRING_COM, IID_SPRING_ATT = 0, 1, 2
IID_MAX_RESTRAINTS = 4
IID_LF_CHAIN = 64

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double
_pi64 = ctypes.POINTER(ctypes.c_int64)
_pint = ctypes.POINTER(ctypes.c_int)

# name -> argtypes; every function returns int (see iid_b200.h)
SIGNATURES = {
    'iid_version': [],
    'iid_device_count': [_pint],
    'iid_device_info': [_int, _pint, _pint, _pint],
    'iid_create': [_int, _int, ctypes.POINTER(_vp)],
    'iid_create_multi': [_int, _int, ctypes.POINTER(_vp)],
    'iid_handle_devices': [_vp, _pint, _pint],
    'iid_destroy': [_vp],
    'iid_get_stream': [_vp, ctypes.POINTER(_vp)],
    'iid_synchronize': [_vp],
    'iid_set_shard': [_vp, _int, _int],
    'iid_set_structure': [_vp, _i64, _vp, _i64, _vp, _i64, _dbl],
    'iid_set_structure_norm': [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _dbl],
    'iid_set_transform': [_vp, _i64, _i64, _vp],
    'iid_stencil_weights': [_dbl, _vp, _vp, _vp, _vp],
    'iid_hist_stencil_weights': [_int, _dbl, _vp, _vp, _vp, _vp],
    'iid_plan_shard': [_i64, _vp, _i64, _int, _int, _int, _int, _pi64, _pi64,
                       _pi64, _pi64],
    'iid_plan_rows': [_i64, _vp, _i64, _int, _int, _int, _int, _pi64, _pi64, _pi64,
                      _vp, _i64, _vp, _i64, _vp, _i64],
    'iid_get_sizes': [_vp, _pi64, _pi64, _pi64, _pi64, _pi64],
    'iid_fq_partial': [_vp, _vp, _vp, _vp],
    'iid_fq_finish': [_vp, _vp, _vp, _vp],
    'iid_grad_fq_partial': [_vp, _vp, _vp, _vp, _vp],
    'iid_force_partial': [_vp, _vp, _vp, _vp, _vp],
    'iid_fq_to_gr': [_vp, _vp, _vp, _vp],
    'iid_potential': [_vp, _vp, _vp, _int, _dbl, _vp, _vp, _vp],
    'iid_grad_pdf': [_vp, _vp, _i64, _vp, _vp],
    'iid_fq_host': [_vp, _vp, _vp],
    'iid_grad_fq_host': [_vp, _vp, _vp, _vp],
    'iid_pdf_host': [_vp, _vp, _vp, _vp],
    'iid_energy_forces_host': [_vp, _vp, _vp, _int, _dbl, _vp, _vp, _vp],
    'iid_rw_host': [_vp, _vp, _vp, _i64, _int, _dbl, _vp, _vp],
    'iid_contract_host': [_vp, _vp, _int, _i64, _i64, _vp, _vp],
    'iid_fq_to_gr_host': [_vp, _vp, _vp],
    'iid_download_host': [_vp, _vp, _vp, _i64],
    'iid_host_alloc': [_i64, ctypes.POINTER(_vp)],
    'iid_host_free': [_vp],
    'iid_host_register': [_vp, _i64],
    'iid_host_unregister': [_vp],
    'iid_host_device_pointer': [_vp, ctypes.POINTER(_vp)],
    'iid_spring_partial': [_vp, _vp, _i64, _int, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp],
    'iid_spring_host': [_vp, _vp, _i64, _int, _dbl, _dbl, _vp, _vp, _vp, _vp],
    'iid_spring_voxel_host': [_vp, _vp, _i64, _int, _dbl, _dbl, _vp, _dbl, _i64, _i64, _i64, _vp],
    'iid_set_restraints': [_vp, _int, _vp, _vp, _vp],
    'iid_get_restraint_energy': [_vp, _vp],
    'iid_sampler_setup': [_vp, _i64, _vp, _vp],
    'iid_state_upload': [_vp, _int, _vp, _vp, _vp],
    'iid_state_download': [_vp, _int, _vp, _vp, _vp],
    'iid_leapfrog_host': [_vp, _int, _int, _dbl, _int, _vp, _int, _dbl, _vp, _vp, _vp],
    'iid_leapfrog_chain_host': [_vp, _int, _vp, _int, _dbl, _int, _vp, _int, _dbl, _vp, _vp, _vp],
    'iid_leapfrog_chain_begin': [_vp, _int, _vp, _int, _dbl, _int, _vp, _int, _dbl, _pi64],
    'iid_leapfrog_chain_next': [_vp, _i64, _vp, _vp, _vp],
    'iid_set_option': [_vp, ctypes.c_char_p, _i64],
    'iid_launch_count': [_vp, _pi64],
    'iid_last_kernel_ms': [_vp, ctypes.POINTER(ctypes.c_float),
                           ctypes.POINTER(_dbl)],
    'iid_set_timing': [_vp, _int],
    'iid_measure_peaks': [_vp, _vp],
}

_lib = None


class IIDError(RuntimeError):
    pass


class ChainDropped(IIDError):
    """iid_leapfrog_chain_next (IID_E_NOCHAIN): the chain was dropped by another
    call on the handle; its steps have to be asked for again."""


def load():
    """Load ``libiid_b200.so``; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IIDError(
            'pyiid_b200: %s is missing. Build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` from the '
            'repository root. There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = _int
        fn.argtypes = args
    lib.iid_last_error.restype = ctypes.c_char_p
    lib.iid_last_error.argtypes = []
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().iid_last_error()
        if rc == -6:
            raise ChainDropped(msg.decode() if msg else '?')
        raise IIDError('iid_b200 error %d: %s' % (
            rc, msg.decode() if msg else '?'))
