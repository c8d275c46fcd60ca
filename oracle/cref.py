"""ctypes front-end of oracle/libcref.so (the C restatement of the ark-groth16 0.5 CPU prover path).

TEST INFRASTRUCTURE ONLY - see oracle/cref.c.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_MASK = (1 << 64) - 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, 'libcref.so')
    srcs = [os.path.join(_HERE, f) for f in ('cref.c', 'cref_curve.inc', 'Makefile')]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['make', '-C', _HERE, 'libcref.so'], stdout=subprocess.DEVNULL)
    return so


class CrefKey(C.Structure):
    _fields_ = [('n_vars', C.c_uint32), ('n_public', C.c_uint32), ('domain_size', C.c_uint32),
                ('num_constraints', C.c_uint32)] + \
               [(k, C.c_void_p) for k in ('alpha_g1', 'beta_g1', 'delta_g1', 'beta_g2', 'delta_g2', 'a_query',
                                          'b_g1_query', 'l_query', 'h_query', 'b_g2_query',
                                          'a_rowptr', 'a_col', 'a_val', 'b_rowptr', 'b_col', 'b_val')]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.cref_last_phase_seconds.argtypes = [C.c_void_p]
    return _LIB


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


# ----------------------------------------------------------------------------- int <-> limb helpers
def ints_to_limbs(vals) -> np.ndarray:
    """list of python ints (< 2^256) -> (n, 4) uint64 little-endian limbs."""
    buf = b''.join(int(v).to_bytes(32, 'little') for v in vals)
    return np.frombuffer(buf, dtype='<u8').reshape(-1, 4).copy()


def limbs_to_ints(arr: np.ndarray):
    b = np.ascontiguousarray(arr, dtype='<u8').tobytes()
    return [int.from_bytes(b[i:i + 32], 'little') for i in range(0, len(b), 32)]


def fr_to_mont(canon: np.ndarray) -> np.ndarray:
    canon = np.ascontiguousarray(canon, dtype=np.uint64); out = np.empty_like(canon)
    lib().cref_fr_from_canon(_p(out), _p(canon), C.c_size_t(canon.size // 4)); return out


def fr_from_mont(mont: np.ndarray) -> np.ndarray:
    mont = np.ascontiguousarray(mont, dtype=np.uint64); out = np.empty_like(mont)
    lib().cref_fr_to_canon(_p(out), _p(mont), C.c_size_t(mont.size // 4)); return out


def fq_to_mont(canon: np.ndarray) -> np.ndarray:
    canon = np.ascontiguousarray(canon, dtype=np.uint64); out = np.empty_like(canon)
    lib().cref_fq_from_canon(_p(out), _p(canon), C.c_size_t(canon.size // 4)); return out


def fq_from_mont(mont: np.ndarray) -> np.ndarray:
    mont = np.ascontiguousarray(mont, dtype=np.uint64); out = np.empty_like(mont)
    lib().cref_fq_to_canon(_p(out), _p(mont), C.c_size_t(mont.size // 4)); return out


def fr_mul(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a); lib().cref_fr_mul(_p(out), _p(a), _p(b), C.c_size_t(a.size // 4)); return out


def fq_mul(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a); lib().cref_fq_mul(_p(out), _p(a), _p(b), C.c_size_t(a.size // 4)); return out


# ----------------------------------------------------------------------------- kernels
def ntt(data_mont: np.ndarray, inverse: bool = False, nthreads: int = 0) -> np.ndarray:
    d = np.ascontiguousarray(data_mont, dtype=np.uint64).copy()
    n = d.size // 4
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().cref_ntt(_p(d), C.c_int(log_n), C.c_int(int(inverse)), C.c_int(nthreads))
    assert rc == 0
    return d


def witness_map(m, num_inputs, n_vars, a_csr, b_csr, w_mont, nthreads: int = 0) -> np.ndarray:
    """a_csr/b_csr = (rowptr u32[m+1], col u32[nnz], val u64[nnz,4] Montgomery)."""
    n = 1
    while n < m + num_inputs:
        n <<= 1
    h = np.zeros((n, 4), dtype=np.uint64)
    w = np.ascontiguousarray(w_mont, dtype=np.uint64)
    rc = lib().cref_witness_map(C.c_uint32(m), C.c_uint32(num_inputs), C.c_uint32(n_vars),
                                _p(a_csr[0]), _p(a_csr[1]), _p(a_csr[2]), _p(b_csr[0]), _p(b_csr[1]), _p(b_csr[2]),
                                _p(w), _p(h), C.c_int(nthreads))
    if rc < 0:
        raise ValueError("PolynomialDegreeTooLarge")
    return h


def witness_map_libsnark(m, num_inputs, a_csr, b_csr, c_csr, w_mont, nthreads: int = 0) -> np.ndarray:
    n = 1
    while n < m + num_inputs:
        n <<= 1
    h = np.zeros((n, 4), dtype=np.uint64)
    w = np.ascontiguousarray(w_mont, dtype=np.uint64)
    rc = lib().cref_witness_map_libsnark(C.c_uint32(m), C.c_uint32(num_inputs), _p(a_csr[0]), _p(a_csr[1]), _p(a_csr[2]),
                                         _p(b_csr[0]), _p(b_csr[1]), _p(b_csr[2]), _p(c_csr[0]), _p(c_csr[1]), _p(c_csr[2]),
                                         _p(w), _p(h), C.c_int(nthreads))
    if rc < 0:
        raise ValueError("PolynomialDegreeTooLarge")
    return h


def msm_g1(bases: np.ndarray, scalars_canon: np.ndarray, nthreads: int = 0):
    bases = np.ascontiguousarray(bases, dtype=np.uint64); sc = np.ascontiguousarray(scalars_canon, dtype=np.uint64)
    n = min(bases.size // 8, sc.size // 4)
    out = np.zeros(8, dtype=np.uint64)
    lib().cref_msm_g1(_p(bases), _p(sc), C.c_size_t(n), _p(out), C.c_int(nthreads))
    return out


def msm_g2(bases: np.ndarray, scalars_canon: np.ndarray, nthreads: int = 0):
    bases = np.ascontiguousarray(bases, dtype=np.uint64); sc = np.ascontiguousarray(scalars_canon, dtype=np.uint64)
    n = min(bases.size // 16, sc.size // 4)
    out = np.zeros(16, dtype=np.uint64)
    lib().cref_msm_g2(_p(bases), _p(sc), C.c_size_t(n), _p(out), C.c_int(nthreads))
    return out


def fixed_base_g1(scalars_canon: np.ndarray, nthreads: int = 0) -> np.ndarray:
    sc = np.ascontiguousarray(scalars_canon, dtype=np.uint64); n = sc.size // 4
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().cref_fixed_base_g1(_p(sc), C.c_size_t(n), _p(out), C.c_int(nthreads)); return out


def fixed_base_g2(scalars_canon: np.ndarray, nthreads: int = 0) -> np.ndarray:
    sc = np.ascontiguousarray(scalars_canon, dtype=np.uint64); n = sc.size // 4
    out = np.zeros((n, 16), dtype=np.uint64)
    lib().cref_fixed_base_g2(_p(sc), C.c_size_t(n), _p(out), C.c_int(nthreads)); return out


def mul_g1(p, k_canon):
    p = np.ascontiguousarray(p, dtype=np.uint64); k = np.ascontiguousarray(k_canon, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64); lib().cref_mul_g1(_p(p), _p(k), _p(out)); return out


def mul_g2(p, k_canon):
    p = np.ascontiguousarray(p, dtype=np.uint64); k = np.ascontiguousarray(k_canon, dtype=np.uint64)
    out = np.zeros(16, dtype=np.uint64); lib().cref_mul_g2(_p(p), _p(k), _p(out)); return out


def add_g1(p, q):
    p = np.ascontiguousarray(p, dtype=np.uint64); q = np.ascontiguousarray(q, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64); lib().cref_add_g1(_p(p), _p(q), _p(out)); return out


def add_g2(p, q):
    p = np.ascontiguousarray(p, dtype=np.uint64); q = np.ascontiguousarray(q, dtype=np.uint64)
    out = np.zeros(16, dtype=np.uint64); lib().cref_add_g2(_p(p), _p(q), _p(out)); return out


# ----------------------------------------------------------------------------- zkey -> arrays (src/zkey.rs layouts)
def zkey_arrays(data: bytes) -> dict:
    """Raw array view of a snarkjs .zkey, the layout both libcref and the product C ABI consume:
    points stay Montgomery LE exactly as stored (src/zkey.rs:327-368); coefficients are converted from
    v*R^2 to the Montgomery residue v*R (src/zkey.rs:320-325) and packed as CSR over the first
    num_constraints rows (src/zkey.rs:151-196)."""
    assert data[:4] == b'zkey'
    nsec = struct.unpack_from('<I', data, 8)[0]
    pos = 12
    sec = {}
    for _ in range(nsec):
        sid, slen = struct.unpack_from('<IQ', data, pos); pos += 12
        sec.setdefault(sid, (pos, slen)); pos += slen
    p = sec[2][0]
    n8q = struct.unpack_from('<I', data, p)[0]; p += 4 + n8q
    n8r = struct.unpack_from('<I', data, p)[0]; p += 4 + n8r
    n_vars, n_public, domain = struct.unpack_from('<III', data, p); p += 12

    def arr(off, count, words):
        return np.frombuffer(data, dtype='<u8', count=count * words, offset=off).reshape(count, words).copy()

    out = dict(n_vars=n_vars, n_public=n_public, domain_size=domain)
    out['alpha_g1'] = arr(p, 1, 8); p += 64
    out['beta_g1'] = arr(p, 1, 8); p += 64
    out['beta_g2'] = arr(p, 1, 16); p += 128
    out['gamma_g2'] = arr(p, 1, 16); p += 128
    out['delta_g1'] = arr(p, 1, 8); p += 64
    out['delta_g2'] = arr(p, 1, 16); p += 128
    out['ic'] = arr(sec[3][0], n_public + 1, 8)
    out['a_query'] = arr(sec[5][0], n_vars, 8)
    out['b_g1_query'] = arr(sec[6][0], n_vars, 8)
    out['b_g2_query'] = arr(sec[7][0], n_vars, 16)
    out['l_query'] = arr(sec[8][0], n_vars - n_public - 1, 8)
    out['h_query'] = arr(sec[9][0], domain, 8)
    p = sec[4][0]
    ncoef = struct.unpack_from('<I', data, p)[0]; p += 4
    rec = np.frombuffer(data, dtype=np.dtype([('m', '<u4'), ('c', '<u4'), ('s', '<u4'), ('v', '<u8', (4,))]),
                        count=ncoef, offset=p)
    max_c = int(rec['c'].max()) if ncoef else 0
    m = max_c - n_public
    out['num_constraints'] = m
    one = np.zeros((1, 4), dtype=np.uint64); one[0, 0] = 1
    for mi, name in ((0, 'a'), (1, 'b')):
        sel = rec[(rec['m'] == mi) & (rec['c'] < m)]
        order = np.argsort(sel['c'], kind='stable')
        sel = sel[order]
        counts = np.bincount(sel['c'], minlength=m)[:m]
        rowptr = np.zeros(m + 1, dtype=np.uint32); rowptr[1:] = np.cumsum(counts)
        vals = np.ascontiguousarray(sel['v'], dtype=np.uint64).reshape(-1, 4)
        # raw = v*R^2 ; Montgomery form of v is v*R = raw * R^-1  == montmul(raw, 1)
        vals = fr_mul(vals, np.repeat(one, len(vals), axis=0)) if len(vals) else vals
        out[name + '_csr'] = (rowptr, np.ascontiguousarray(sel['s'], dtype=np.uint32), vals)
    return out


def make_key(za: dict):
    """CrefKey over the arrays of zkey_arrays() (keeps references alive on the returned object)."""
    k = CrefKey()
    k.n_vars, k.n_public, k.domain_size, k.num_constraints = za['n_vars'], za['n_public'], za['domain_size'], za['num_constraints']
    for name in ('alpha_g1', 'beta_g1', 'delta_g1', 'beta_g2', 'delta_g2', 'a_query', 'b_g1_query', 'l_query', 'h_query', 'b_g2_query'):
        setattr(k, name, za[name].ctypes.data)
    k.a_rowptr, k.a_col, k.a_val = (x.ctypes.data for x in za['a_csr'])
    k.b_rowptr, k.b_col, k.b_val = (x.ctypes.data for x in za['b_csr'])
    k._keep = za
    return k


def prove(za: dict, r: int, s: int, w_mont: np.ndarray, nthreads: int = 0, want_h: bool = False):
    """create_proof_with_reduction_and_matrices on the CPU.  Returns 256 proof bytes (and h if asked)."""
    key = make_key(za)
    rr = ints_to_limbs([r % R_MOD]); ss = ints_to_limbs([s % R_MOD])
    w = np.ascontiguousarray(w_mont, dtype=np.uint64)
    out = np.zeros(256, dtype=np.uint8)
    h = np.zeros((za['domain_size'], 4), dtype=np.uint64) if want_h else None
    rc = lib().cref_prove(C.byref(key), _p(rr), _p(ss), _p(w), _p(out), _p(h) if want_h else None, C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("cref_prove failed")
    return (out.tobytes(), h) if want_h else out.tobytes()


def last_phase_seconds():
    buf = (C.c_double * 8)()
    lib().cref_last_phase_seconds(buf)
    return dict(zip(('witness_map', 'msm_h', 'msm_l', 'msm_a', 'msm_b1', 'msm_b2', 'glue'), list(buf)[:7]))
