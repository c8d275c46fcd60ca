"""Host-side snarkjs .zkey reader with the reference's shape: read_zkey(f) -> (ProvingKey, ConstraintMatrices)
(/root/reference/src/zkey.rs:53-60).  Parsing stays on the host (north star); the point sections of a zkey are already
arrays of Montgomery little-endian coordinates (zkey.rs:327-368) - exactly the device layout - so they are handed to
b2g_pk_load as zero-copy numpy views."""
from __future__ import annotations

import io
import struct
from dataclasses import dataclass, field

import numpy as np

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_MONT_R = 1 << 256


def fr_to_mont(values) -> np.ndarray:
    """Field elements (python ints, reduced mod r) -> (n, 4) uint64 Montgomery limbs, what `Fr::from` yields in
    arkworks memory (src/witness/witness_calculator.rs:163-179)."""
    buf = b''.join(((int(v) % R_MOD) * _MONT_R % R_MOD).to_bytes(32, 'little') for v in values)
    return np.frombuffer(buf, dtype='<u8').reshape(-1, 4).copy()


_R_INV = pow(_MONT_R, -1, R_MOD)


def fr_from_mont(arr: np.ndarray):
    b = np.ascontiguousarray(arr, dtype='<u8').tobytes()
    return [int.from_bytes(b[i:i + 32], 'little') * _R_INV % R_MOD for i in range(0, len(b), 32)]


@dataclass
class ProvingKey:
    """ProvingKey<Bn254> as assembled at src/zkey.rs:103-133; arrays keep the zkey byte layout."""
    n_vars: int
    n_public: int
    domain_size: int
    alpha_g1: np.ndarray
    beta_g1: np.ndarray
    beta_g2: np.ndarray
    gamma_g2: np.ndarray
    delta_g1: np.ndarray
    delta_g2: np.ndarray
    gamma_abc_g1: np.ndarray        # IC
    a_query: np.ndarray
    b_g1_query: np.ndarray
    b_g2_query: np.ndarray
    l_query: np.ndarray
    h_query: np.ndarray
    _device: dict = field(default_factory=dict, repr=False)


@dataclass
class ConstraintMatrices:
    """ConstraintMatrices<Fr> (src/zkey.rs:181-193) with a and b in CSR form; c is empty on the zkey route."""
    num_instance_variables: int
    num_witness_variables: int
    num_constraints: int
    a_num_non_zero: int
    b_num_non_zero: int
    c_num_non_zero: int
    a: tuple                        # (rowptr u32[m+1], col u32[nnz], val u64[nnz,4] Montgomery)
    b: tuple
    c: tuple = None                 # only on the R1CS route (LibsnarkReduction); the zkey route has none (zkey.rs:188-192)
    _device: dict = field(default_factory=dict, repr=False)

    @property
    def n_vars(self) -> int:
        return self.num_instance_variables + self.num_witness_variables - 0


def _sections(data):
    if len(data) < 12 or bytes(data[:4]) != b'zkey':
        raise ValueError("not a zkey file")
    nsec = struct.unpack_from('<I', data, 8)[0]
    pos, sec = 12, {}
    for _ in range(nsec):                                           # src/zkey.rs:73-101
        if pos + 12 > len(data):
            raise ValueError("malformed zkey: truncated section table")
        sid, slen = struct.unpack_from('<IQ', data, pos)
        pos += 12
        sec.setdefault(sid, (pos, slen))
        pos += slen
    return sec


def csr_from_coo(rows, cols, vals_mont, m):
    order = np.argsort(rows, kind='stable')
    rows, cols, vals_mont = rows[order], cols[order], vals_mont[order]
    rowptr = np.zeros(m + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum(np.bincount(rows, minlength=m)[:m])
    return rowptr, np.ascontiguousarray(cols, dtype=np.uint32), np.ascontiguousarray(vals_mont, dtype=np.uint64).reshape(-1, 4)


def read_zkey(src):
    """src: path, bytes or a binary file object.  Returns (ProvingKey, ConstraintMatrices)."""
    if isinstance(src, (bytes, bytearray, memoryview)):
        data = bytes(src)
    elif isinstance(src, (str, bytes)) or hasattr(src, '__fspath__'):
        data = np.fromfile(src, dtype=np.uint8).tobytes()
    elif isinstance(src, io.IOBase) or hasattr(src, 'read'):
        data = src.read()
    else:
        raise TypeError("read_zkey expects a path, bytes or a binary reader")
    sec = _sections(data)
    for sid in (2, 3, 4, 5, 6, 7, 8, 9):
        if sid not in sec:
            raise ValueError(f"malformed zkey: section {sid} missing")
        if sec[sid][0] + sec[sid][1] > len(data):
            raise ValueError(f"malformed zkey: section {sid} is truncated")
    p = sec[2][0]                                                   # header, src/zkey.rs:282-318
    n8q = struct.unpack_from('<I', data, p)[0]; p += 4
    q = int.from_bytes(data[p:p + n8q], 'little'); p += n8q
    n8r = struct.unpack_from('<I', data, p)[0]; p += 4
    r = int.from_bytes(data[p:p + n8r], 'little'); p += n8r
    if n8q != 32 or n8r != 32 or q != Q_MOD or r != R_MOD:
        raise ValueError("only BN254 zkeys are supported")
    n_vars, n_public, domain = struct.unpack_from('<III', data, p); p += 12

    def arr(off, count, words):
        return np.frombuffer(data, dtype='<u8', count=count * words, offset=off).reshape(count, words)

    alpha_g1 = arr(p, 1, 8); p += 64
    beta_g1 = arr(p, 1, 8); p += 64
    beta_g2 = arr(p, 1, 16); p += 128
    gamma_g2 = arr(p, 1, 16); p += 128
    delta_g1 = arr(p, 1, 8); p += 64
    delta_g2 = arr(p, 1, 16); p += 128
    pk = ProvingKey(n_vars, n_public, domain, alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2,
                    arr(sec[3][0], n_public + 1, 8), arr(sec[5][0], n_vars, 8), arr(sec[6][0], n_vars, 8),
                    arr(sec[7][0], n_vars, 16), arr(sec[8][0], n_vars - n_public - 1, 8), arr(sec[9][0], domain, 8))
    # coefficients, src/zkey.rs:151-196
    p = sec[4][0]
    ncoef = struct.unpack_from('<I', data, p)[0]; p += 4
    rec = np.frombuffer(data, dtype=np.dtype([('m', '<u4'), ('c', '<u4'), ('s', '<u4'), ('v', '<u8', (4,))]), count=ncoef, offset=p)
    max_c = int(rec['c'].max()) if ncoef else 0
    m = max_c - n_public                                            # src/zkey.rs:171
    if m < 0:
        raise ValueError("malformed zkey: no constraints")
    mats = []
    for mi in (0, 1):
        sel = rec[(rec['m'] == mi) & (rec['c'] < m)]               # public-input rows are dropped (zkey.rs:172-175)
        # stored value = v * R^2 (zkey.rs:320-325); the Montgomery residue of v is v*R = stored * R^-1
        raw = np.ascontiguousarray(sel['v']).tobytes()
        vals = b''.join((int.from_bytes(raw[i:i + 32], 'little') * _R_INV % R_MOD).to_bytes(32, 'little')
                        for i in range(0, len(raw), 32))
        vals = np.frombuffer(vals, dtype='<u8').reshape(-1, 4) if vals else np.zeros((0, 4), dtype=np.uint64)
        mats.append(csr_from_coo(sel['c'].astype(np.int64), sel['s'], vals, m))
    cm = ConstraintMatrices(n_public + 1, n_vars - n_public - 1, m, len(mats[0][1]), len(mats[1][1]), 0, mats[0], mats[1])
    return pk, cm
