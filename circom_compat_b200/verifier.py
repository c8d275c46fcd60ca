"""Groth16 verification over BN254 on the host: the product-side counterpart of the entry points the reference calls
right after proving (/root/reference/src/zkey.rs:868-870, 914-916; tests/groth16.rs:33-35):

    pvk = Groth16.process_vk(vk)                                   <- GrothBn::process_vk(&params.vk)
    ok  = Groth16.verify_with_processed_vk(pvk, inputs, proof)     <- GrothBn::verify_with_processed_vk(&pvk, &inputs, &proof)
    ok  = Groth16.verify(vk, inputs, proof)                        <- SNARK::verify (tests/groth16.rs:33)

ark-groth16 0.5.0 semantics: prepared_inputs = gamma_abc_g1[0] + sum_i x_i * gamma_abc_g1[i + 1]; accept iff
    e(A, B) * e(prepared_inputs, -gamma) * e(C, -delta) == e(alpha, beta)           (one multi-Miller loop + final exp)
and `MalformedVerifyingKey` when len(inputs) + 1 != len(gamma_abc_g1).  The verifier is milliseconds of host work and is
not part of the accelerated path; it exists so that flows written against the reference's API run
unchanged.  It shares nothing with oracle/ (the tests' checker keeps its own, differently built, pairing): this one is
the optimal ate pairing on the Fq2 -> Fq6 -> Fq12 tower (u^2 = -1, v^3 = 9 + u, w^2 = v) with sparse line
multiplication and affine line functions on the sextic D-twist.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

from .zkey import Q_MOD, R_MOD

P = Q_MOD
_X = 4965661367192848881                      # BN parameter x: p = 36x^4 + 36x^3 + 24x^2 + 6x + 1
ATE_LOOP_COUNT = 6 * _X + 2                   # 29793968203157093288
_MONT_R_INV = pow(1 << 256, -1, P)


class MalformedVerifyingKey(ValueError):
    """SynthesisError::MalformedVerifyingKey (ark-groth16 prepare_inputs)"""


# ---------------------------------------------------------------------------------------------- Fq2 = Fq[u] / (u^2 + 1)
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2_conj(a): return (a[0], (-a[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return ((a[0] + a[1]) * (a[0] - a[1]) % P, 2 * a[0] * a[1] % P)
def f2_scale(a, k): return (a[0] * k % P, a[1] * k % P)
def f2_mul_xi(a): return ((9 * a[0] - a[1]) % P, (9 * a[1] + a[0]) % P)          # * (9 + u)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


F2_ZERO, F2_ONE = (0, 0), (1, 0)
XI = (9, 1)

# ---------------------------------------------------------------------------------------------- Fq6 = Fq2[v] / (v^3 - xi)
F6_ZERO, F6_ONE = (F2_ZERO, F2_ZERO, F2_ZERO), (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(a, b): return (f2_add(a[0], b[0]), f2_add(a[1], b[1]), f2_add(a[2], b[2]))
def f6_sub(a, b): return (f2_sub(a[0], b[0]), f2_sub(a[1], b[1]), f2_sub(a[2], b[2]))
def f6_neg(a): return (f2_neg(a[0]), f2_neg(a[1]), f2_neg(a[2]))
def f6_mul_v(a): return (f2_mul_xi(a[2]), a[0], a[1])                              # * v


def f6_mul(a, b):
    t0, t1, t2 = f2_mul(a[0], b[0]), f2_mul(a[1], b[1]), f2_mul(a[2], b[2])
    c0 = f2_add(t0, f2_mul_xi(f2_sub(f2_sub(f2_mul(f2_add(a[1], a[2]), f2_add(b[1], b[2])), t1), t2)))
    c1 = f2_add(f2_sub(f2_sub(f2_mul(f2_add(a[0], a[1]), f2_add(b[0], b[1])), t0), t1), f2_mul_xi(t2))
    c2 = f2_add(f2_sub(f2_sub(f2_mul(f2_add(a[0], a[2]), f2_add(b[0], b[2])), t0), t2), t1)
    return (c0, c1, c2)


def f6_inv(a):
    c0 = f2_sub(f2_sqr(a[0]), f2_mul_xi(f2_mul(a[1], a[2])))
    c1 = f2_sub(f2_mul_xi(f2_sqr(a[2])), f2_mul(a[0], a[1]))
    c2 = f2_sub(f2_sqr(a[1]), f2_mul(a[0], a[2]))
    t = f2_inv(f2_add(f2_mul(a[0], c0), f2_mul_xi(f2_add(f2_mul(a[2], c1), f2_mul(a[1], c2)))))
    return (f2_mul(c0, t), f2_mul(c1, t), f2_mul(c2, t))


# ---------------------------------------------------------------------------------------------- Fq12 = Fq6[w] / (w^2 - v)
F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):
    t0, t1 = f6_mul(a[0], b[0]), f6_mul(a[1], b[1])
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a[0], a[1]), f6_add(b[0], b[1])), t0), t1)
    return (f6_add(t0, f6_mul_v(t1)), c1)


def f12_sqr(a):
    return f12_mul(a, a)


def f12_conj(a): return (a[0], f6_neg(a[1]))                                       # the p^6-power Frobenius


def f12_inv(a):
    t = f6_inv(f6_sub(f6_mul(a[0], a[0]), f6_mul_v(f6_mul(a[1], a[1]))))
    return (f6_mul(a[0], t), f6_neg(f6_mul(a[1], t)))


def f12_pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == '1':
            r = f12_mul(r, a)
    return r


def f12_mul_line(f, l0, l1, l3):
    """f * (l0 + l1 w + l3 w^3) with l0 in Fq, l1, l3 in Fq2: as a tower element ((l0, 0, 0), (l1, l3, 0))."""
    return f12_mul(f, (((l0 % P, 0), F2_ZERO, F2_ZERO), (l1, l3, F2_ZERO)))


# ---------------------------------------------------------------------------------------------- curve points (affine, None = infinity)
TWIST_B = f2_mul((3, 0), f2_inv(XI))                                               # E': y^2 = x^3 + 3 / (9 + u)
# Frobenius on the twist: pi(x, y) = (conj(x) * xi^((p-1)/3), conj(y) * xi^((p-1)/2)); pi^2(x, y) = (x * xi^((p^2-1)/3), y * xi^((p^2-1)/2))
_G12 = f2_pow(XI, (P - 1) // 3)
_G13 = f2_pow(XI, (P - 1) // 2)
_G22 = f2_pow(XI, (P * P - 1) // 3)
_G23 = f2_pow(XI, (P * P - 1) // 2)


def g1_on_curve(pt) -> bool:
    return pt is None or (pt[1] * pt[1] - pt[0] * pt[0] * pt[0] - 3) % P == 0


def g2_on_curve(pt) -> bool:
    return pt is None or f2_sub(f2_sqr(pt[1]), f2_add(f2_mul(f2_sqr(pt[0]), pt[0]), TWIST_B)) == F2_ZERO


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, P) % P
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, P) % P
    x = (lam * lam - a[0] - b[0]) % P
    return (x, (lam * (a[0] - x) - a[1]) % P)


def g1_mul(pt, k):
    acc = None
    k %= R_MOD
    while k:
        if k & 1:
            acc = g1_add(acc, pt)
        pt = g1_add(pt, pt)
        k >>= 1
    return acc


def g1_neg(pt):
    return None if pt is None else (pt[0], (-pt[1]) % P)


def g2_neg(pt):
    return None if pt is None else (pt[0], f2_neg(pt[1]))


# ---------------------------------------------------------------------------------------------- pairing
def _line_and_step(t, q, xp, yp):
    """Line through t and q (tangent when they coincide) on the twist, evaluated at the G1 point (xp, yp), and t + q.
    Untwisting (x', y') -> (x' w^2, y' w^3) turns slope lambda into lambda * w, hence
    l(P) = yp - (lambda xp) w + (lambda x_t - y_t) w^3."""
    if t[0] == q[0] and t[1] == q[1]:
        lam = f2_mul(f2_scale(f2_sqr(t[0]), 3), f2_inv(f2_scale(t[1], 2)))
    else:
        lam = f2_mul(f2_sub(q[1], t[1]), f2_inv(f2_sub(q[0], t[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), t[0]), q[0])
    y3 = f2_sub(f2_mul(lam, f2_sub(t[0], x3)), t[1])
    return (yp, f2_scale(lam, (-xp) % P), f2_sub(f2_mul(lam, t[0]), t[1])), (x3, y3)


def miller_loop(pairs) -> tuple:
    """prod_i f_{6x+2, Q_i}(P_i) * (the two Frobenius lines), pairs = [(P in G1, Q in G2)]; infinity on either side contributes 1."""
    pairs = [(p, q) for p, q in pairs if p is not None and q is not None]
    f = F12_ONE
    ts = [q for _, q in pairs]
    bits = bin(ATE_LOOP_COUNT)[3:]
    for bit in bits:
        f = f12_sqr(f)
        for i, (p, q) in enumerate(pairs):
            l, ts[i] = _line_and_step(ts[i], ts[i], p[0], p[1])
            f = f12_mul_line(f, *l)
        if bit == '1':
            for i, (p, q) in enumerate(pairs):
                l, ts[i] = _line_and_step(ts[i], q, p[0], p[1])
                f = f12_mul_line(f, *l)
    for i, (p, q) in enumerate(pairs):
        q1 = (f2_mul(f2_conj(q[0]), _G12), f2_mul(f2_conj(q[1]), _G13))
        q2 = (f2_mul(q[0], _G22), f2_neg(f2_mul(q[1], _G23)))                     # -pi^2(Q)
        l, t = _line_and_step(ts[i], q1, p[0], p[1])
        f = f12_mul_line(f, *l)
        l, _ = _line_and_step(t, q2, p[0], p[1])
        f = f12_mul_line(f, *l)
    return f


_HARD_EXP = (P ** 6 + 1) // R_MOD


def final_exponentiation(f):
    """f^((p^12 - 1) / r) = (f^(p^6 - 1))^((p^6 + 1) / r); the first factor is conj(f) / f."""
    return f12_pow(f12_mul(f12_conj(f), f12_inv(f)), _HARD_EXP)


def pairing(p, q):
    return final_exponentiation(miller_loop([(p, q)]))


# ---------------------------------------------------------------------------------------------- Groth16 verifier
def _mont_words_to_ints(arr) -> List[int]:
    import numpy as np
    raw = np.ascontiguousarray(arr, dtype='<u8').tobytes()
    return [int.from_bytes(raw[i:i + 32], 'little') * _MONT_R_INV % P for i in range(0, len(raw), 32)]


def _g1_from_words(arr):
    x, y = _mont_words_to_ints(arr)
    return None if x == 0 and y == 0 else (x, y)


def _g2_from_words(arr):
    x0, x1, y0, y1 = _mont_words_to_ints(arr)
    return None if (x0, x1, y0, y1) == (0, 0, 0, 0) else ((x0, x1), (y0, y1))


@dataclass
class VerifyingKey:
    """VerifyingKey<Bn254> (params.vk, src/zkey.rs:103-119): canonical affine coordinates, None = infinity."""
    alpha_g1: tuple
    beta_g2: tuple
    gamma_g2: tuple
    delta_g2: tuple
    gamma_abc_g1: list

    @staticmethod
    def from_proving_key(pk) -> 'VerifyingKey':
        return VerifyingKey(_g1_from_words(pk.alpha_g1), _g2_from_words(pk.beta_g2), _g2_from_words(pk.gamma_g2), _g2_from_words(pk.delta_g2),
                            [_g1_from_words(p) for p in pk.gamma_abc_g1])


@dataclass
class PreparedVerifyingKey:
    """PreparedVerifyingKey<Bn254>: vk, e(alpha, beta), -gamma, -delta (ark-groth16 prepare_verifying_key)."""
    vk: VerifyingKey
    alpha_g1_beta_g2: tuple
    gamma_g2_neg: tuple
    delta_g2_neg: tuple


def prepare_verifying_key(vk) -> PreparedVerifyingKey:
    if not isinstance(vk, VerifyingKey):
        vk = VerifyingKey.from_proving_key(vk)
    for pt in [vk.alpha_g1] + list(vk.gamma_abc_g1):
        if not g1_on_curve(pt):
            raise ValueError("verifying key: G1 point not on the curve")
    for pt in (vk.beta_g2, vk.gamma_g2, vk.delta_g2):
        if not g2_on_curve(pt):
            raise ValueError("verifying key: G2 point not on the curve")
    return PreparedVerifyingKey(vk, pairing(vk.alpha_g1, vk.beta_g2), g2_neg(vk.gamma_g2), g2_neg(vk.delta_g2))


def prepare_inputs(pvk: PreparedVerifyingKey, public_inputs: Sequence[int]):
    ic = pvk.vk.gamma_abc_g1
    if len(public_inputs) + 1 != len(ic):
        raise MalformedVerifyingKey(f"{len(public_inputs)} public inputs for a key with {len(ic) - 1}")
    acc = ic[0]
    for x, b in zip(public_inputs, ic[1:]):
        acc = g1_add(acc, g1_mul(b, int(x)))
    return acc


def _proof_points(proof):
    a, b, c = (proof.a, proof.b, proof.c) if hasattr(proof, 'a') else proof
    a = None if a is None or tuple(a) == (0, 0) else (int(a[0]), int(a[1]))
    c = None if c is None or tuple(c) == (0, 0) else (int(c[0]), int(c[1]))
    if b is not None:
        b = ((int(b[0][0]), int(b[0][1])), (int(b[1][0]), int(b[1][1])))
        if b == ((0, 0), (0, 0)):
            b = None
    return a, b, c


def verify_with_processed_vk(pvk: PreparedVerifyingKey, public_inputs: Sequence[int], proof) -> bool:
    """proof: a groth16.Proof (or an (A, B, C) tuple of canonical affine coordinates).  Points that are not on the curve
    cannot be constructed in arkworks (deserialisation fails); here they make the proof invalid."""
    a, b, c = _proof_points(proof)
    if not (g1_on_curve(a) and g1_on_curve(c) and g2_on_curve(b)):
        return False
    prepared = prepare_inputs(pvk, public_inputs)
    f = miller_loop([(a, b), (prepared, pvk.gamma_g2_neg), (c, pvk.delta_g2_neg)])
    return final_exponentiation(f) == pvk.alpha_g1_beta_g2


def verify(vk, public_inputs: Sequence[int], proof) -> bool:
    return verify_with_processed_vk(prepare_verifying_key(vk), public_inputs, proof)
