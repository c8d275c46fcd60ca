"""Big-integer oracle for the Groth16/BN254 prover hot path of ark-circom 0.5.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it may be imported by
tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference leg, and by
nothing else.  The product path (circom_compat_b200/) never imports it and fails loudly without
its CUDA library.

What is restated here (file:line are relative to /root/reference):
  * read_zkey                              src/zkey.rs:53-60, 73-101, 151-196, 282-368
  * CircomReduction::witness_map_from_matrices   src/circom/qap.rs:23-88
  * CircomReduction::h_query_scalars       src/circom/qap.rs:90-105
  * Groth16::create_proof_with_reduction_and_matrices -> create_proof_with_assignment
        call sites src/zkey.rs:903-912, benches/groth16.rs:52-61.  The body lives in the
        un-vendored crate ark-groth16 0.5.0 (prover.rs; Cargo.toml:24-32 pins ^0.5.0, Cargo.lock is
        git-ignored) and is restated from its published algorithm: SURVEY.md section 3.4.
  * VariableBaseMSM::msm_bigint            ark-ec 0.5.0 (un-vendored) - any correct MSM gives the
        same affine point, so this oracle uses plain windowed sums.
  * Radix2EvaluationDomain fft/ifft        ark-poly 0.5.0 (un-vendored); natural order in/out,
        omega_n = 5^((r-1)/n), ifft scales by 1/n.
  * Groth16 verify (pairing check)         ark-groth16 0.5.0 verifier.rs, restated; used to make sure
        the restated prover still satisfies e(A,B) = e(alpha,beta) e(IC.x,gamma) e(C,delta).

Parity pinning: the reference's own tests pin the zkey encodings (src/zkey.rs:398-463, 519-779) and
assert only `verified == true` for proofs (src/zkey.rs:872,918).  Proof BYTES are not pinned by the
reference ("parity unpinned" for bytes); they are pinned here by determinism of the group law plus
the verifier equation, and by self-derived golden vectors under tests/golden/.
"""
from __future__ import annotations

import hashlib
import struct

# ----------------------------------------------------------------------------- constants
# src/witness/witness_calculator.rs:328-332 (r), zkey header field q (src/zkey.rs:292)
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq
MONT_R = 1 << 256
R_INV_Q = pow(MONT_R, -1, Q_MOD)
R_INV_R = pow(MONT_R, -1, R_MOD)
TWO_ADICITY = 28
FR_GENERATOR = 5
ROOT_2_28 = pow(FR_GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)

G1_GEN = (1, 2)
# src/zkey.rs:443-463
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


# ----------------------------------------------------------------------------- Fq / Fq2
class _Fq:
    """Prime field Fq as plain ints."""
    zero = 0
    one = 1

    @staticmethod
    def add(a, b): return (a + b) % Q_MOD
    @staticmethod
    def sub(a, b): return (a - b) % Q_MOD
    @staticmethod
    def mul(a, b): return (a * b) % Q_MOD
    @staticmethod
    def sqr(a): return (a * a) % Q_MOD
    @staticmethod
    def neg(a): return (-a) % Q_MOD
    @staticmethod
    def inv(a): return pow(a, -1, Q_MOD)
    @staticmethod
    def is_zero(a): return a == 0
    @staticmethod
    def small(k): return k % Q_MOD


class _Fq2:
    """Fq2 = Fq[u]/(u^2+1); elements are (c0, c1)."""
    zero = (0, 0)
    one = (1, 0)

    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)
    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)
    @staticmethod
    def sqr(a):
        return ((a[0] + a[1]) * (a[0] - a[1]) % Q_MOD, 2 * a[0] * a[1] % Q_MOD)
    @staticmethod
    def neg(a): return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)
    @staticmethod
    def inv(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, Q_MOD)
        return (a[0] * d % Q_MOD, (-a[1]) * d % Q_MOD)
    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0
    @staticmethod
    def small(k): return (k % Q_MOD, 0)


FQ, FQ2 = _Fq, _Fq2
G1_B = 3
G2_B = _Fq2.mul((3, 0), _Fq2.inv((9, 1)))  # 3/(9+u)


# ----------------------------------------------------------------------------- curve (generic)
# Affine points are (x, y) or None for infinity.  Jacobian (X, Y, Z) internally.
def _jac_double(F, P):
    X, Y, Z = P
    if F.is_zero(Z):
        return P
    A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
    t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
    D = F.add(t, t)
    E = F.add(F.add(A, A), A)
    Fv = F.sqr(E)
    X3 = F.sub(Fv, F.add(D, D))
    C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
    YZ = F.mul(Y, Z)
    return (X3, Y3, F.add(YZ, YZ))


def _jac_add(F, P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if F.is_zero(Z1):
        return Q
    if F.is_zero(Z2):
        return P
    Z1Z1 = F.sqr(Z1); Z2Z2 = F.sqr(Z2)
    U1 = F.mul(X1, Z2Z2); U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(F.mul(Y1, Z2), Z2Z2); S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        if S1 == S2:
            return _jac_double(F, P)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, U1)
    Rr = F.sub(S2, S1)
    HH = F.sqr(H); HHH = F.mul(H, HH)
    V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.add(V, V))
    Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def _to_jac(F, P):
    return (F.one, F.one, F.zero) if P is None else (P[0], P[1], F.one)


def _to_affine(F, P):
    X, Y, Z = P
    if F.is_zero(Z):
        return None
    zi = F.inv(Z); zi2 = F.sqr(zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


class Curve:
    def __init__(self, F, b):
        self.F, self.b = F, b

    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        return F.sqr(P[1]) == F.add(F.mul(F.sqr(P[0]), P[0]), self.b)

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        return _to_affine(F, _jac_add(F, _to_jac(F, P), _to_jac(F, Q)))

    def sum(self, pts):
        F = self.F
        acc = _to_jac(F, None)
        for P in pts:
            acc = _jac_add(F, acc, _to_jac(F, P))
        return _to_affine(F, acc)

    def mul(self, P, k):
        F = self.F
        k %= R_MOD
        acc = _to_jac(F, None)
        if P is None or k == 0:
            return None
        base = _to_jac(F, P)
        for bit in bin(k)[2:]:
            acc = _jac_double(F, acc)
            if bit == '1':
                acc = _jac_add(F, acc, base)
        return _to_affine(F, acc)

    def msm(self, bases, scalars, c=None):
        """sum_i scalars[i]*bases[i] over min(len) terms (msm_bigint truncation rule, ark-ec 0.5.0)."""
        F = self.F
        n = min(len(bases), len(scalars))
        if n == 0:
            return None
        if c is None:
            c = 3 if n < 32 else max(3, int(n.bit_length() * 0.6))
        total = _to_jac(F, None)
        nwin = (254 + c - 1) // c
        jb = [_to_jac(F, P) for P in bases[:n]]
        for w in range(nwin - 1, -1, -1):
            for _ in range(c):
                total = _jac_double(F, total)
            buckets = {}
            for i in range(n):
                d = (scalars[i] >> (w * c)) & ((1 << c) - 1)
                if d and bases[i] is not None:
                    buckets[d] = _jac_add(F, buckets[d], jb[i]) if d in buckets else jb[i]
            run = _to_jac(F, None); acc = _to_jac(F, None)
            for d in range((1 << c) - 1, 0, -1):
                if d in buckets:
                    run = _jac_add(F, run, buckets[d])
                acc = _jac_add(F, acc, run)
            total = _jac_add(F, total, acc)
        return _to_affine(F, total)


G1 = Curve(_Fq, G1_B)
G2 = Curve(_Fq2, G2_B)


# ----------------------------------------------------------------------------- Fr domain (ark-poly Radix2)
def domain_size_for(k: int) -> int:
    """Radix2EvaluationDomain::new(k).size(): next power of two, None above 2^28 (qap.rs:30-31)."""
    n = 1
    while n < k:
        n <<= 1
    if n > (1 << TWO_ADICITY):
        raise ValueError("PolynomialDegreeTooLarge")
    return n


def root_of_unity(n: int) -> int:
    assert n & (n - 1) == 0 and n <= (1 << TWO_ADICITY)
    return pow(ROOT_2_28, (1 << TWO_ADICITY) // n, R_MOD)


def _bitrev_permute(a):
    n = len(a)
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j |= bit
        if i < j:
            a[i], a[j] = a[j], a[i]


def fft(a, inverse=False):
    """In-place natural-order radix-2 (i)NTT over Fr; the inverse includes the 1/n factor."""
    n = len(a)
    if n == 1:
        return a
    w_n = root_of_unity(n)
    if inverse:
        w_n = pow(w_n, -1, R_MOD)
    _bitrev_permute(a)
    length = 2
    while length <= n:
        w_len = pow(w_n, n // length, R_MOD)
        half = length >> 1
        tw = [1] * half
        for k in range(1, half):
            tw[k] = tw[k - 1] * w_len % R_MOD
        for s in range(0, n, length):
            for k in range(half):
                u = a[s + k]; v = a[s + k + half] * tw[k] % R_MOD
                a[s + k] = (u + v) % R_MOD
                a[s + k + half] = (u - v) % R_MOD
        length <<= 1
    if inverse:
        ninv = pow(n, -1, R_MOD)
        for i in range(n):
            a[i] = a[i] * ninv % R_MOD
    return a


def distribute_powers(a, g):
    p = 1
    for i in range(len(a)):
        a[i] = a[i] * p % R_MOD
        p = p * g % R_MOD


def witness_map_from_matrices(mat_a, mat_b, num_inputs, num_constraints, w):
    """src/circom/qap.rs:23-88, same order of operations.

    mat_a / mat_b: list (len >= num_constraints) of rows [(coeff, index), ...] - the arkworks
    ConstraintMatrices tuple order (src/zkey.rs:168).  Returns h, length = domain size.
    """
    n = domain_size_for(num_constraints + num_inputs)              # qap.rs:30-32
    a = [0] * n
    b = [0] * n
    for i in range(num_constraints):                               # qap.rs:37-44
        a[i] = sum(c * w[j] for c, j in mat_a[i]) % R_MOD
        b[i] = sum(c * w[j] for c, j in mat_b[i]) % R_MOD
    for j in range(num_inputs):                                    # qap.rs:46-50
        a[num_constraints + j] = w[j] % R_MOD
    c = [0] * n
    for i in range(num_constraints):                               # qap.rs:52-58
        c[i] = a[i] * b[i] % R_MOD
    fft(a, inverse=True); fft(b, inverse=True)                     # qap.rs:60-61
    g = root_of_unity(2 * n)                                       # qap.rs:63-68
    distribute_powers(a, g); distribute_powers(b, g)               # qap.rs:69-70
    fft(a); fft(b)                                                 # qap.rs:72-73
    ab = [x * y % R_MOD for x, y in zip(a, b)]                     # qap.rs:75
    fft(c, inverse=True); distribute_powers(c, g); fft(c)          # qap.rs:79-81
    return [(x - y) % R_MOD for x, y in zip(ab, c)]                # qap.rs:83-85


def h_query_scalars(max_power, t, delta_inverse):
    """src/circom/qap.rs:90-105."""
    scalars = [delta_inverse * pow(t, i, R_MOD) % R_MOD for i in range(2 * max_power + 1)]
    n = domain_size_for(len(scalars))
    scalars += [0] * (n - len(scalars))
    fft(scalars, inverse=True)
    return scalars[1::2]


# ----------------------------------------------------------------------------- zkey reader
class ZKey:
    """Parsed snarkjs .zkey (src/zkey.rs).  Points are affine canonical ints (None = infinity)."""


def _fq_from_mont(b: bytes) -> int:
    return int.from_bytes(b, 'little') * R_INV_Q % Q_MOD           # zkey.rs:327-332


def _g1_from(b: bytes):
    x = _fq_from_mont(b[0:32]); y = _fq_from_mont(b[32:64])        # zkey.rs:340-349
    if x == 0 and y == 0:
        return None
    P = (x, y)
    if not G1.on_curve(P):
        raise ValueError("G1 point not on curve")                  # G1Affine::new panics
    return P


def _g2_from(b: bytes):
    x = (_fq_from_mont(b[0:32]), _fq_from_mont(b[32:64]))          # zkey.rs:334-338, 351-360
    y = (_fq_from_mont(b[64:96]), _fq_from_mont(b[96:128]))
    if x == (0, 0) and y == (0, 0):
        return None
    P = (x, y)
    if not G2.on_curve(P):
        raise ValueError("G2 point not on curve")
    return P


def read_sections(data: bytes):
    magic = data[0:4]                                              # zkey.rs:73-101
    version, nsec = struct.unpack_from('<II', data, 4)
    pos = 12
    sections = {}
    for _ in range(nsec):
        sid, slen = struct.unpack_from('<IQ', data, pos)
        pos += 12
        sections.setdefault(sid, []).append((pos, slen))
        pos += slen
    return magic, version, sections


def read_zkey(data: bytes, decode_points: bool = True) -> ZKey:
    magic, version, sections = read_sections(data)
    if magic != b'zkey':
        raise ValueError("not a zkey")
    z = ZKey()
    z.raw = data
    z.sections = sections
    p, _ = sections[2][0]                                          # header, zkey.rs:282-318
    n8q = struct.unpack_from('<I', data, p)[0]; p += 4
    z.q = int.from_bytes(data[p:p + n8q], 'little'); p += n8q
    n8r = struct.unpack_from('<I', data, p)[0]; p += 4
    z.r = int.from_bytes(data[p:p + n8r], 'little'); p += n8r
    z.n_vars, z.n_public, z.domain_size = struct.unpack_from('<III', data, p); p += 12
    z.alpha_g1 = _g1_from(data[p:p + 64]); p += 64
    z.beta_g1 = _g1_from(data[p:p + 64]); p += 64
    z.beta_g2 = _g2_from(data[p:p + 128]); p += 128
    z.gamma_g2 = _g2_from(data[p:p + 128]); p += 128
    z.delta_g1 = _g1_from(data[p:p + 64]); p += 64
    z.delta_g2 = _g2_from(data[p:p + 128]); p += 128

    def g1s(sec, num):
        p0, _ = sections[sec][0]
        return [_g1_from(data[p0 + 64 * i: p0 + 64 * i + 64]) for i in range(num)]

    def g2s(sec, num):
        p0, _ = sections[sec][0]
        return [_g2_from(data[p0 + 128 * i: p0 + 128 * i + 128]) for i in range(num)]

    z.ic = g1s(3, z.n_public + 1)
    if decode_points:
        z.a_query = g1s(5, z.n_vars)
        z.b_g1_query = g1s(6, z.n_vars)
        z.b_g2_query = g2s(7, z.n_vars)
        z.l_query = g1s(8, z.n_vars - z.n_public - 1)
        z.h_query = g1s(9, z.domain_size)

    # coefficients, zkey.rs:151-196.  value is v*R^2; the reader strips one R and keeps the result
    # as a Montgomery residue, i.e. the field element is raw * R^-2 (zkey.rs:320-325).
    p, _ = sections[4][0]
    ncoef = struct.unpack_from('<I', data, p)[0]; p += 4
    mats = [[[] for _ in range(z.domain_size)] for _ in range(2)]
    max_c = 0
    for _ in range(ncoef):
        m, c, s = struct.unpack_from('<III', data, p); p += 12
        v = int.from_bytes(data[p:p + 32], 'little') * R_INV_R * R_INV_R % R_MOD; p += 32
        max_c = max(max_c, c)
        mats[m][c].append((v, s))
    z.num_constraints = max_c - z.n_public                         # zkey.rs:171
    z.mat_a = mats[0][:z.num_constraints]
    z.mat_b = mats[1][:z.num_constraints]
    z.num_inputs = z.n_public + 1                                  # zkey.rs:182
    return z


# ----------------------------------------------------------------------------- prover (ark-groth16 0.5.0 restated)
def prove_from_parts(z, r, s, h, w):
    """create_proof_with_assignment.  Returns affine (A, B, C)."""
    li = z.num_inputs
    h_acc = G1.msm(z.h_query, h)
    l_acc = G1.msm(z.l_query, w[li:])

    def coeff(curve, initial, query, vk_param, delta, k):
        acc = curve.msm(query[1:], w[1:])
        return curve.sum([curve.mul(delta, k), query[0], acc, vk_param])

    A = coeff(G1, None, z.a_query, z.alpha_g1, z.delta_g1, r)
    B1 = coeff(G1, None, z.b_g1_query, z.beta_g1, z.delta_g1, s) if r != 0 else None
    B2 = coeff(G2, None, z.b_g2_query, z.beta_g2, z.delta_g2, s)
    rs_delta = G1.mul(z.delta_g1, r * s % R_MOD)
    C = G1.sum([G1.mul(A, s), G1.mul(B1, r), G1.neg(rs_delta), l_acc, h_acc])
    return A, B2, C


def prove(z, r, s, w):
    """Groth16::<Bn254, CircomReduction>::create_proof_with_reduction_and_matrices (zkey.rs:903-912)."""
    h = witness_map_from_matrices(z.mat_a, z.mat_b, z.num_inputs, z.num_constraints, w)
    return prove_from_parts(z, r % R_MOD, s % R_MOD, h, [x % R_MOD for x in w])


def proof_to_bytes(A, B, C) -> bytes:
    """256-byte uncompressed view used at the C ABI: A.x A.y B.x.c0 B.x.c1 B.y.c0 B.y.c1 C.x C.y,
    canonical little-endian, infinity = zeros."""
    def f(v): return int(v).to_bytes(32, 'little')
    out = b''
    out += (f(A[0]) + f(A[1])) if A else bytes(64)
    out += (f(B[0][0]) + f(B[0][1]) + f(B[1][0]) + f(B[1][1])) if B else bytes(128)
    out += (f(C[0]) + f(C[1])) if C else bytes(64)
    return out


# ----------------------------------------------------------------------------- pairing / verifier
# Fq12 = Fq[w]/(w^12 - 18 w^6 + 82); untwist psi(x, y) = (x' w^2, y' w^3) with u = w^6 - 9.
_F12_DEG = 12


def _f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for i in range(22, 11, -1):
        v = t[i]
        if v:
            t[i - 6] += 18 * v
            t[i - 12] -= 82 * v
    return [x % Q_MOD for x in t[:12]]


def _f12_one():
    return [1] + [0] * 11


def _f12_pow(a, e):
    res = _f12_one()
    base = a
    while e:
        if e & 1:
            res = _f12_mul(res, base)
        base = _f12_mul(base, base)
        e >>= 1
    return res


def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _f12_inv(a):
    # extended Euclid over Fq[w] against the modulus polynomial
    lm, hm = [1] + [0] * 12, [0] * 13
    low = list(a) + [0]
    high = [82, 0, 0, 0, 0, 0, (-18) % Q_MOD, 0, 0, 0, 0, 0, 1]
    while _poly_deg(low):
        # r = high / low (polynomial rounded division)
        dl, dh = _poly_deg(low), _poly_deg(high)
        temp = list(high)
        quo = [0] * 13
        linv = pow(low[dl], -1, Q_MOD)
        for i in range(dh - dl, -1, -1):
            quo[i] = temp[dl + i] * linv % Q_MOD
            for c in range(dl + 1):
                temp[c + i] = (temp[c + i] - low[c] * quo[i]) % Q_MOD
        nm = list(hm); new = list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * quo[j]) % Q_MOD
                new[i + j] = (new[i + j] - low[i] * quo[j]) % Q_MOD
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, Q_MOD)
    return [x * inv0 % Q_MOD for x in lm[:12]]


class _Fq12:
    zero = [0] * 12
    one = _f12_one()
    @staticmethod
    def add(a, b): return [(x + y) % Q_MOD for x, y in zip(a, b)]
    @staticmethod
    def sub(a, b): return [(x - y) % Q_MOD for x, y in zip(a, b)]
    mul = staticmethod(_f12_mul)
    @staticmethod
    def sqr(a): return _f12_mul(a, a)
    @staticmethod
    def neg(a): return [(-x) % Q_MOD for x in a]
    inv = staticmethod(_f12_inv)
    @staticmethod
    def is_zero(a): return not any(a)


def _untwist(Q):
    (x0, x1), (y0, y1) = Q
    nx = [0] * 12; ny = [0] * 12
    # x' = (x0 - 9 x1) + x1 w^6, times w^2
    nx[2] = (x0 - 9 * x1) % Q_MOD; nx[8] = x1
    ny[3] = (y0 - 9 * y1) % Q_MOD; ny[9] = y1
    return (nx, ny)


def _embed(P):
    return ([P[0]] + [0] * 11, [P[1]] + [0] * 11)


def _aff12_double(P):
    F = _Fq12
    x, y = P
    m = F.mul(F.mul([3] + [0] * 11, F.sqr(x)), F.inv(F.add(y, y)))
    nx = F.sub(F.sqr(m), F.add(x, x))
    ny = F.sub(F.mul(m, F.sub(x, nx)), y)
    return (nx, ny), m


def _aff12_add(P, Q):
    F = _Fq12
    x1, y1 = P; x2, y2 = Q
    m = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    nx = F.sub(F.sub(F.sqr(m), x1), x2)
    ny = F.sub(F.mul(m, F.sub(x1, nx)), y1)
    return (nx, ny), m


ATE_LOOP = 29793968203157093288


def miller_loop(Q2, P1):
    """Optimal-ate Miller loop f_{6x+2,Q}(P) with the two Frobenius line corrections."""
    F = _Fq12
    if Q2 is None or P1 is None:
        return F.one
    Q = _untwist(Q2); P = _embed(P1)
    xt, yt = P
    Rp = Q
    f = F.one

    def line(m, A):
        return F.sub(F.mul(m, F.sub(xt, A[0])), F.sub(yt, A[1]))

    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        newR, m = _aff12_double(Rp)
        f = F.mul(F.sqr(f), line(m, Rp))
        Rp = newR
        if (ATE_LOOP >> i) & 1:
            newR, m = _aff12_add(Rp, Q)
            f = F.mul(f, line(m, Rp))
            Rp = newR
    Q1 = (_f12_pow(Q[0], Q_MOD), _f12_pow(Q[1], Q_MOD))
    nQ2 = (_f12_pow(Q1[0], Q_MOD), F.neg(_f12_pow(Q1[1], Q_MOD)))
    newR, m = _aff12_add(Rp, Q1)
    f = F.mul(f, line(m, Rp)); Rp = newR
    _, m = _aff12_add(Rp, nQ2)
    f = F.mul(f, line(m, Rp))
    return f


def final_exponentiation(f):
    return _f12_pow(f, (Q_MOD ** 12 - 1) // R_MOD)


def verify(z, public_inputs, proof) -> bool:
    """Groth16 verifier equation (ark-groth16 0.5.0 verify_with_processed_vk, restated):
    e(A,B) * e(-alpha,beta) * e(-IC.x, gamma) * e(-C, delta) == 1."""
    A, B, C = proof
    if not (G1.on_curve(A) and G2.on_curve(B) and G1.on_curve(C)):
        return False
    acc = z.ic[0]
    for x, P in zip(public_inputs, z.ic[1:]):
        acc = G1.add(acc, G1.mul(P, x))
    f = miller_loop(B, A)
    f = _f12_mul(f, miller_loop(z.beta_g2, G1.neg(z.alpha_g1)))
    f = _f12_mul(f, miller_loop(z.gamma_g2, G1.neg(acc)))
    f = _f12_mul(f, miller_loop(z.delta_g2, G1.neg(C)))
    return final_exponentiation(f) == _f12_one()


# ----------------------------------------------------------------------------- squaring-chain circuit + trapdoor setup
def chain_witness(n_vars: int, a: int = 3):
    """Witness of the reference's bench circuit family (test-vectors/complex-circuit/
    complex-circuit.circom.template): wires [1, c, a, b0, b1, ...], b0=a^2, b_i=b_{i-1}^2, c=b_last."""
    w = [0] * n_vars
    w[0] = 1; w[2] = a % R_MOD
    for k in range(3, n_vars):
        w[k] = w[k - 1] * w[k - 1] % R_MOD
    w[1] = w[n_vars - 1] * w[n_vars - 1] % R_MOD if n_vars > 3 else w[2] * w[2] % R_MOD
    return w


def chain_matrices(n_vars: int):
    """R1CS of the chain as the circom compiler emits it (decoded from the 10000-constraint
    fixture): constraint k: (-w[k+2]) * (w[k+2]) = (-w[k+3]); the last one targets wire 1.
    m = n_vars - 2 constraints.  Returns (A rows, B rows, C rows) in (coeff, index) order."""
    m = n_vars - 2
    A = [[(R_MOD - 1, k + 2)] for k in range(m)]
    B = [[(1, k + 2)] for k in range(m)]
    C = [[(R_MOD - 1, k + 3 if k + 3 < n_vars else 1)] for k in range(m)]
    return A, B, C


def sha_stream_fr(seed: int, count: int, tag: bytes = b'b2g'):
    out = []
    ctr = 0
    while len(out) < count:
        d = hashlib.sha256(tag + struct.pack('<QQ', seed, ctr)).digest()
        ctr += 1
        v = int.from_bytes(d, 'little') % R_MOD
        if v:
            out.append(v)
    return out


def lagrange_at(n: int, tau: int):
    """L_j(tau) for the size-n radix-2 domain, j = 0..n-1."""
    w = root_of_unity(n)
    zt = (pow(tau, n, R_MOD) - 1) % R_MOD
    ninv = pow(n, -1, R_MOD)
    out = []
    wj = 1
    for _ in range(n):
        out.append(zt * ninv % R_MOD * wj % R_MOD * pow((tau - wj) % R_MOD, -1, R_MOD) % R_MOD)
        wj = wj * w % R_MOD
    return out


def trapdoor_setup_scalars(A, B, C, n_vars, num_inputs, tau, alpha, beta, delta):
    """Discrete logs of every proving-key base for a circuit given as matrices (snarkjs flavour:
    public-input rows appended to A, H from qap.rs:90-105).  Returns dict of scalar lists."""
    m = len(A)
    n = domain_size_for(m + num_inputs)
    L = lagrange_at(n, tau)
    a_t = [0] * n_vars; b_t = [0] * n_vars; c_t = [0] * n_vars
    for i in range(m):
        for coef, j in A[i]:
            a_t[j] = (a_t[j] + coef * L[i]) % R_MOD
        for coef, j in B[i]:
            b_t[j] = (b_t[j] + coef * L[i]) % R_MOD
        for coef, j in C[i]:
            c_t[j] = (c_t[j] + coef * L[i]) % R_MOD
    for j in range(num_inputs):
        a_t[j] = (a_t[j] + L[m + j]) % R_MOD
    dinv = pow(delta, -1, R_MOD)
    l_t = [(beta * a_t[i] + alpha * b_t[i] + c_t[i]) * dinv % R_MOD for i in range(num_inputs, n_vars)]
    ic_t = [(beta * a_t[i] + alpha * b_t[i] + c_t[i]) % R_MOD for i in range(num_inputs)]  # gamma = 1
    h_t = h_query_scalars(n - 1, tau, dinv)
    return dict(n=n, a=a_t, b=b_t, c=c_t, l=l_t, ic=ic_t, h=h_t)


def trapdoor_expected_dlogs(sc, w, h, r, s, alpha, beta, delta, num_inputs):
    """dlog(A), dlog(B), dlog(C) of the proof for a trapdoor-known key."""
    da = (alpha + sum(x * y for x, y in zip(w, sc['a'])) + r * delta) % R_MOD
    db = (beta + sum(x * y for x, y in zip(w, sc['b'])) + s * delta) % R_MOD
    dc = (sum(x * y for x, y in zip(w[num_inputs:], sc['l'])) + sum(x * y for x, y in zip(h, sc['h']))
          + s * da + r * db - r * s % R_MOD * delta) % R_MOD
    return da, db, dc


# ----------------------------------------------------------------------------- ark-serialize decompression (independent check)
def _fq_sqrt(a):
    # q = 3 mod 4
    r = pow(a, (Q_MOD + 1) // 4, Q_MOD)
    return r if r * r % Q_MOD == a % Q_MOD else None


def _fq2_sqrt(a):
    # complex method for u^2 = -1: sqrt(a0 + a1 u)
    a0, a1 = a
    if a1 == 0:
        r = _fq_sqrt(a0)
        if r is not None:
            return (r, 0)
        r = _fq_sqrt((-a0) % Q_MOD)
        return (0, r)
    alpha = _fq_sqrt((a0 * a0 + a1 * a1) % Q_MOD)
    if alpha is None:
        return None
    inv2 = pow(2, -1, Q_MOD)
    delta = (a0 + alpha) * inv2 % Q_MOD
    x0 = _fq_sqrt(delta)
    if x0 is None:
        delta = (a0 - alpha) * inv2 % Q_MOD
        x0 = _fq_sqrt(delta)
    x1 = a1 * pow(2 * x0, -1, Q_MOD) % Q_MOD
    return (x0, x1)


def decompress_proof(data: bytes):
    """Inverse of ark-serialize's compressed Proof<Bn254> encoding (flags: bit 7 = y is the larger root, bit 6 = infinity)."""
    def g1(b):
        flags = b[31] & 0xC0
        x = int.from_bytes(bytes(b[:31]) + bytes([b[31] & 0x3F]), 'little')
        if flags & 0x40:
            return None
        y = _fq_sqrt((x * x * x + 3) % Q_MOD)
        neg = (Q_MOD - y) % Q_MOD
        big, small = (y, neg) if y > neg else (neg, y)
        return (x, big if flags & 0x80 else small)

    def g2(b):
        flags = b[63] & 0xC0
        x = (int.from_bytes(b[:32], 'little'), int.from_bytes(bytes(b[32:63]) + bytes([b[63] & 0x3F]), 'little'))
        if flags & 0x40:
            return None
        y = _fq2_sqrt(_Fq2.add(_Fq2.mul(_Fq2.sqr(x), x), G2_B))
        neg = _Fq2.neg(y)
        big, small = (y, neg) if (y[1], y[0]) > (neg[1], neg[0]) else (neg, y)
        return (x, big if flags & 0x80 else small)
    return g1(data[0:32]), g2(data[32:96]), g1(data[96:128])


# ----------------------------------------------------------------------------- LibsnarkReduction + R1CS route
def libsnark_witness_map_from_matrices(mat_a, mat_b, mat_c, num_inputs, num_constraints, w):
    """LibsnarkReduction::witness_map_from_matrices, ark-groth16 0.5.0 r1cs_to_qap.rs (un-vendored; the default QAP of
    `Groth16<Bn254>` used by /root/reference/tests/groth16.rs:9,25-35).  Restated: a, b, c evaluations (c from the real C
    matrix) -> ifft -> coset fft with offset g = Fr::GENERATOR = 5 -> (a*b - c) / Z(g) -> coset ifft.  Returns the n
    coefficients of h (the top one is 0)."""
    n = domain_size_for(num_constraints + num_inputs)
    a = [0] * n; b = [0] * n; c = [0] * n
    for i in range(num_constraints):
        a[i] = sum(k * w[j] for k, j in mat_a[i]) % R_MOD
        b[i] = sum(k * w[j] for k, j in mat_b[i]) % R_MOD
        c[i] = sum(k * w[j] for k, j in mat_c[i]) % R_MOD
    for j in range(num_inputs):
        a[num_constraints + j] = w[j] % R_MOD
    g = FR_GENERATOR
    for v in (a, b, c):
        fft(v, inverse=True); distribute_powers(v, g); fft(v)
    zinv = pow((pow(g, n, R_MOD) - 1) % R_MOD, -1, R_MOD)
    ab = [(x * y - z) * zinv % R_MOD for x, y, z in zip(a, b, c)]
    fft(ab, inverse=True)
    distribute_powers(ab, pow(g, -1, R_MOD))
    return ab


def libsnark_h_query_scalars(n, t, delta_inverse):
    """LibsnarkReduction::h_query_scalars: tau^i * Z(tau) / delta, i < n - 1"""
    zt = (pow(t, n, R_MOD) - 1) * delta_inverse % R_MOD
    return [pow(t, i, R_MOD) * zt % R_MOD for i in range(n - 1)]


def read_r1cs(data: bytes):
    """src/circom/r1cs_reader.rs:54-249 -> (num_inputs, n_wires, constraints[(A, B, C)] with rows of (coeff, index))"""
    assert data[:4] == b'r1cs' and struct.unpack_from('<I', data, 4)[0] == 1
    nsec = struct.unpack_from('<I', data, 8)[0]
    pos, off = 12, {}
    for _ in range(nsec):
        t, sz = struct.unpack_from('<IQ', data, pos); pos += 12
        off[t] = pos; pos += sz
    p = off[1]
    assert struct.unpack_from('<I', data, p)[0] == 32 and int.from_bytes(data[p + 4:p + 36], 'little') == R_MOD
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_cons = struct.unpack_from('<IIIIQI', data, p + 36)
    p = off[2]
    cons = []
    for _ in range(n_cons):
        trip = []
        for _k in range(3):
            n = struct.unpack_from('<I', data, p)[0]; p += 4
            row = []
            for _j in range(n):
                wire = struct.unpack_from('<I', data, p)[0]
                row.append((int.from_bytes(data[p + 4:p + 36], 'little'), wire)); p += 36
            trip.append(row)
        cons.append(tuple(trip))
    return 1 + n_pub_in + n_pub_out, n_wires, cons
