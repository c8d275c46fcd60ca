"""ctypes binding of libb2groth.so (include/b2groth.h).  There is NO CPU fallback: if the shared library is missing,
or no CUDA device is present, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('B2G_LIB') or os.path.join(_HERE, 'libb2groth.so')   # B2G_LIB: tuning builds only

B2G_OK, B2G_E_DOMAIN, B2G_E_SHAPE, B2G_E_DEVICE, B2G_E_INPUT = 0, -1, -2, -3, -4
PARTIAL_BYTES = 768
IPC_HANDLE_BYTES = 80          # B2G_IPC_HANDLE_BYTES: cudaIpcMemHandle_t + arena capacity
REDUCTION_CIRCOM, REDUCTION_LIBSNARK = 0, 1


class B2gError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b2groth error {code}: {msg}")
        self.code = code


class PolynomialDegreeTooLarge(B2gError):
    """SynthesisError::PolynomialDegreeTooLarge (/root/reference/src/circom/qap.rs:31,66)."""


class PkDesc(C.Structure):
    _fields_ = [('n_vars', C.c_uint32), ('n_public', C.c_uint32), ('domain_size', C.c_uint32), ('reserved', C.c_uint32)] + \
               [(k, C.c_void_p) for k in ('alpha_g1', 'beta_g1', 'delta_g1', 'beta_g2', 'delta_g2', 'a_query', 'b_g1_query',
                                          'b_g2_query', 'l_query', 'h_query')]


class MatDesc(C.Structure):
    _fields_ = [('num_constraints', C.c_uint32), ('num_inputs', C.c_uint32), ('n_vars', C.c_uint32), ('reduction', C.c_uint32)] + \
               [(k, C.c_void_p) for k in ('a_rowptr', 'a_col', 'a_val', 'b_rowptr', 'b_col', 'b_val', 'c_rowptr', 'c_col', 'c_val')]


EXPORTS = ['b2g_last_error', 'b2g_version', 'b2g_device_count', 'b2g_ctx_create', 'b2g_ctx_destroy', 'b2g_ctx_prepare', 'b2g_pk_load', 'b2g_pk_free',
           'b2g_matrices_load', 'b2g_matrices_free', 'b2g_witness_map', 'b2g_prove', 'b2g_prove_submit', 'b2g_prove_wait', 'b2g_host_register', 'b2g_host_unregister', 'b2g_prove_partial', 'b2g_prove_finish',
           'b2g_p2p_export', 'b2g_p2p_import', 'b2g_p2p_connect_local', 'b2g_prove_sharded_p2p', 'b2g_msm_g1', 'b2g_msm_g2', 'b2g_ntt', 'b2g_fixed_base_g1', 'b2g_fixed_base_g2', 'b2g_test_op', 'b2g_last_timings',
           'b2g_bench_device', 'b2g_bench_msm', 'b2g_launch_count']

_lib = None


def lib():
    """Load libb2groth.so; raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C circom_compat_b200/csrc` "
                              "(there is no CPU fallback for the proving path)")
        L = C.CDLL(LIB_PATH)
        L.b2g_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(L, name)  # AttributeError if a declared symbol is not exported
        vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
        L.b2g_ctx_create.argtypes = [i, i, i, C.POINTER(vp)]
        L.b2g_ctx_destroy.argtypes = [vp]
        L.b2g_ctx_prepare.argtypes = [vp, vp, vp]
        L.b2g_pk_load.argtypes = [vp, C.POINTER(PkDesc), C.POINTER(vp)]
        L.b2g_pk_free.argtypes = [vp]
        L.b2g_matrices_load.argtypes = [vp, C.POINTER(MatDesc), C.POINTER(vp)]
        L.b2g_matrices_free.argtypes = [vp]
        L.b2g_witness_map.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_uint32)]
        L.b2g_prove.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b2g_prove_submit.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b2g_prove_wait.argtypes = [vp]
        L.b2g_host_register.argtypes = [vp, sz]
        L.b2g_host_unregister.argtypes = [vp]
        L.b2g_prove_partial.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b2g_prove_finish.argtypes = [vp, vp, vp, i, vp, vp, vp]
        L.b2g_p2p_export.argtypes = [vp, vp]
        L.b2g_p2p_import.argtypes = [vp, vp, i]
        L.b2g_p2p_connect_local.argtypes = [vp, i]
        L.b2g_prove_sharded_p2p.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b2g_msm_g1.argtypes = [vp, vp, vp, sz, i, vp]
        L.b2g_msm_g2.argtypes = [vp, vp, vp, sz, i, vp]
        L.b2g_ntt.argtypes = [vp, vp, i, i]
        L.b2g_fixed_base_g1.argtypes = [vp, vp, sz, vp]
        L.b2g_fixed_base_g2.argtypes = [vp, vp, sz, vp]
        L.b2g_test_op.argtypes = [vp, i, vp, vp, sz, vp]
        L.b2g_last_timings.argtypes = [vp, vp]
        L.b2g_bench_device.argtypes = [vp, vp, vp, i, C.POINTER(C.c_float)]
        L.b2g_bench_msm.argtypes = [vp, vp, vp, i, i, C.POINTER(C.c_float)]
        L.b2g_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.b2g_device_count.argtypes = [C.POINTER(C.c_int)]
        _lib = L
    return _lib


def check(rc: int):
    if rc != B2G_OK:
        msg = lib().b2g_last_error().decode(errors='replace')
        if rc == B2G_E_DOMAIN:
            raise PolynomialDegreeTooLarge(rc, msg)
        raise B2gError(rc, msg)
