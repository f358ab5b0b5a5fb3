"""ctypes binding of libdpc_b200.so (the C ABI declared in include/dpc_b200.h).

There is NO fallback: if the shared library is missing or an entry point is absent, importing the
product path fails loudly.  Build it with `python -m dpc_b200.build` (or __graft_entry__.build()).
"""
import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_uint64, c_float, c_void_p, c_char_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libdpc_b200.so')


class ConvGeom(ctypes.Structure):
    """mirror of `dpc_conv_geom`"""
    _fields_ = [(n, c_int32) for n in
                ('NB', 'Ti', 'Hi', 'Wi', 'Ci', 'To', 'Ho', 'Wo', 'Co', 'kT', 'kH', 'kW',
                 'sT', 'sH', 'sW', 'pT', 'pH', 'pW')]


P = c_void_p
_SIGS = {
    'dpc_abi_version': (c_int, []),
    'dpc_last_error': (c_char_p, []),
    'dpc_launch_count': (c_int64, []),
    'dpc_split_bf16': (c_int, [P, P, P, c_int64, P]),
    'dpc_pack_conv_weight_bf16': (c_int, [P, P, P, P, P, c_int, c_int, c_int, P]),
    'dpc_score_matmul_tc': (c_int, [c_int, c_int, c_int, P, P, P, P, c_int, P, P]),
    'dpc_split_f16': (c_int, [P, P, P, c_int64, P]),
    'dpc_gemm_nt_split_tc': (c_int, [c_int, c_int, c_int, P, P, P, P, c_int, P, c_int, P]),
    'dpc_conv3d_fwd_tc': (c_int, [POINTER(ConvGeom), P, P, P, P, P, P, P]),
    'dpc_conv3d_dgrad_tc': (c_int, [POINTER(ConvGeom), P, P, P, P, P, c_int, P]),
    'dpc_conv3d_dgrad_bnred_tc': (c_int, [POINTER(ConvGeom), P, P, P, P, P, c_int, P, P, P, P, P, P]),
    'dpc_conv3d_wgrad_tc': (c_int, [POINTER(ConvGeom), P, P, P, P, P, P, P]),
    'dpc_stem_conv_fwd_tc': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_stem_conv_wgrad_tc': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_stem_s2d_pack': (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_stem_conv_wgrad_s2d': (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_stem_pool_supported': (c_int, [c_int, c_int]),
    'dpc_stem_s2d_wpack': (c_int, [P, P, P]),
    'dpc_stem_pool_fwd': (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_stem_pool_finalize': (c_int, [P, P, P, P, P, P, P, P, P, c_int64, P]),
    'dpc_stem_pool_bwd_reduce': (c_int, [P, P, P, P, P, P, P, P, c_int64, P]),
    'dpc_stem_pool_bwd': (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_stem_pool_bwd_wgrad': (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_bn_stats': (c_int, [P, c_int64, c_int, P, P, P, c_float, P]),
    'dpc_bn_finalize': (c_int, [P, c_int64, c_int, c_float, P, P, P]),
    'dpc_bn_apply_fwd': (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, c_int, P, P, P, c_int64, c_int, P]),
    'dpc_bn_bwd': (c_int, [P, P, P, c_int, P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, P]),
    'dpc_bn_bwd_apply': (c_int, [P, P, P, c_int, P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, P]),
    'dpc_bn_relu_maxpool_fwd': (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_bn_relu_maxpool_bwd': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_stem_tail_bwd': (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    'dpc_pool_split_fwd':(c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_pool_split_bwd': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    'dpc_gemm_f32': (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, P, c_int, P, c_int, c_float, P, c_int, P]),
    'dpc_gather_rows': (c_int, [P, P, c_int64, c_int, c_int64, c_int64, c_int64, P]),
    'dpc_scatter_rows': (c_int, [P, P, c_int64, c_int, c_int64, c_int64, c_int64, c_int, P]),
    'dpc_colsum': (c_int, [P, c_int64, c_int, P, c_int, P]),
    'dpc_gru_gates_zr': (c_int, [P, P, c_int, P, P, P, P, P, P, P, c_int64, c_int, P]),
    'dpc_gru_out': (c_int, [P, c_int, P, P, P, P, P, P, P, c_float, c_uint64, c_uint64, c_int64, c_int, P]),
    'dpc_gru_bwd_out': (c_int, [P, P, P, P, P, P, P, P, c_int64, c_int, P]),
    'dpc_gru_bwd_zr': (c_int, [P, P, P, P, P, P, P, c_int64, c_int, P]),
    'dpc_head_chain_pack': (c_int, [P] * 10),
    'dpc_head_chain_fwd': (c_int, [P] * 10 + [c_int, c_int, c_int, c_int, c_float, c_uint64] + [P] * 11),
    'dpc_head_chain_bwd': (c_int, [P] * 6 + [c_int, c_int, c_int, c_int] + [P] * 13),
    'dpc_bias_relu': (c_int, [P, P, P, c_int, c_int64, c_int, P]),
    'dpc_relu_bwd': (c_int, [P, P, P, c_int, c_int64, P]),
    'dpc_nce_mask_fill': (c_int, [P, c_int, c_int, c_int, P]),
    'dpc_nce_ce_fwd': (c_int, [P, c_int, c_int, P, P, P]),
    'dpc_nce_ce_bwd': (c_int, [P, P, P, P, c_int, c_int, P]),
    'dpc_bn_running_update': (c_int, [P, P, c_int64, c_float, c_float, P, P, c_int, P]),
    'dpc_bn_rstd_from_var': (c_int, [P, c_float, P, c_int, P]),
    'dpc_relu_pool_fwd': (c_int, [P, P, c_int, c_int, c_int64, P]),
    'dpc_relu_pool_bwd': (c_int, [P, P, P, c_int, c_int, c_int64, P]),
    'dpc_dropout_fwd': (c_int, [P, P, P, c_float, c_uint64, c_uint64, c_int64, P]),
    'dpc_mul': (c_int, [P, P, P, c_int64, P]),
    'dpc_comm_unique_id': (c_int, [P]),
    'dpc_comm_init': (c_int, [P, c_int, c_int, P]),
    'dpc_flat_allreduce': (c_int, [P, P, c_int64, P]),
    'dpc_comm_destroy': (c_int, [P]),
    'dpc_augment_clips': (c_int, [P, P, P, P, P, P] + [c_int] * 9 + [P]),
    'dpc_adam_step': (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, c_float, P]),
}

EXPORTS = tuple(_SIGS.keys())


class DpcLibError(RuntimeError):
    pass


class _Lib:
    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise DpcLibError(
                'libdpc_b200.so not found at %s: the CUDA extension is required (no CPU fallback). '
                'Build it with `python -m dpc_b200.build`.' % path)
        self._dll = ctypes.CDLL(path)
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError:
                raise DpcLibError('libdpc_b200.so does not export %s (stale build?)' % name)
            fn.restype = res
            fn.argtypes = args
            if res is c_int and name not in ('dpc_abi_version', 'dpc_stem_pool_supported'):
                setattr(self, name[4:], self._checked(fn, name))
            else:
                setattr(self, name[4:], fn)
        if self.abi_version() != 1:
            raise DpcLibError('libdpc_b200.so ABI version %d != 1' % self.abi_version())

    def _checked(self, fn, name):
        dll = self._dll

        def call(*a):
            rc = fn(*a)
            if rc != 0:
                raise DpcLibError('%s failed (%d): %s' % (name, rc, dll.dpc_last_error().decode()))
        call.__name__ = name
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()
