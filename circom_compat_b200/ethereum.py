"""Output formats of a proof / verifying key.

  * Ethereum tuples  <- /root/reference/src/ethereum.rs: G1 (:20-54), G2 (:56-95, `as_tuple` emits c1 BEFORE c0 :82-86),
    Proof (:98-128), VerifyingKey (:130-174), Inputs (:10-18); U256 = big-endian canonical integer (:185-189).  Both
    directions, like the reference (`From<&G1Affine> for G1` and `From<G1> for G1Affine`, ... :26-95, :110-128, :151-174,
    `u256_to_point` / `point_to_u256` :176-189); tests/test_host.py mirrors its convert_* tests (:195-279).
  * ark-serialize 0.5 (un-vendored; restated): compressed = x little-endian with flags in the top bits of the last byte
    (bit 7: y is the lexicographically larger of {y, -y}; bit 6: infinity), uncompressed = x then y, infinity flag on the
    last byte of y.  Fq2 compares c1 first, then c0; its flags sit in the last byte of c1.
Host-side only (ms-scale formatting of 3 points)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

from .zkey import Q_MOD, R_MOD

FLAG_NEG, FLAG_INF = 0x80, 0x40


def _fq_from_mont_words(arr) -> List[int]:
    import numpy as np
    b = np.ascontiguousarray(arr, dtype='<u8').tobytes()
    rinv = pow(1 << 256, -1, Q_MOD)
    return [int.from_bytes(b[i:i + 32], 'little') * rinv % Q_MOD for i in range(0, len(b), 32)]


def point_to_u256(v: int, modulus: int = Q_MOD) -> bytes:
    """point_to_u256 (src/ethereum.rs:185-189): the canonical integer of a field element as a big-endian 32-byte U256"""
    return (int(v) % modulus).to_bytes(32, 'big')


def u256_to_point(word, modulus: int = Q_MOD) -> int:
    """u256_to_point (src/ethereum.rs:176-181): `F::from_bigint(..).expect(..)` - a value >= the modulus panics in the
    reference; here it raises"""
    v = int.from_bytes(word, 'big') if isinstance(word, (bytes, bytearray)) else int(word)
    if not 0 <= v < modulus:
        raise ValueError("U256 is not a canonical field element")
    return v


@dataclass(frozen=True)
class G1:
    x: int
    y: int

    def as_tuple(self) -> Tuple[int, int]:
        return (self.x, self.y)

    @staticmethod
    def from_tuple(t) -> 'G1':
        return G1(u256_to_point(t[0]), u256_to_point(t[1]))

    def to_affine(self):
        """From<G1> for G1Affine (src/ethereum.rs:26-39): (0, 0) is the point at infinity; returns (x, y) or None"""
        return None if self.x == 0 and self.y == 0 else (self.x, self.y)

    @staticmethod
    def from_affine(pt) -> 'G1':
        """From<&G1Affine> for G1 (src/ethereum.rs:46-54)"""
        return G1(0, 0) if pt is None else G1(int(pt[0]), int(pt[1]))


@dataclass(frozen=True)
class G2:
    x: Tuple[int, int]          # (c0, c1)
    y: Tuple[int, int]

    def as_tuple(self):
        # NB: c1 first (src/ethereum.rs:82-86)
        return ([self.x[1], self.x[0]], [self.y[1], self.y[0]])

    @staticmethod
    def from_tuple(t) -> 'G2':
        """inverse of as_tuple: the tuple carries c1 BEFORE c0"""
        return G2((u256_to_point(t[0][1]), u256_to_point(t[0][0])), (u256_to_point(t[1][1]), u256_to_point(t[1][0])))

    def to_affine(self):
        """From<G2> for G2Affine (src/ethereum.rs:61-80); ((c0, c1), (c0, c1)) or None for infinity"""
        return None if self.x == (0, 0) and self.y == (0, 0) else (tuple(self.x), tuple(self.y))

    @staticmethod
    def from_affine(pt) -> 'G2':
        """From<&G2Affine> for G2 (src/ethereum.rs:88-95)"""
        return G2((0, 0), (0, 0)) if pt is None else G2((int(pt[0][0]), int(pt[0][1])), (int(pt[1][0]), int(pt[1][1])))


@dataclass(frozen=True)
class Proof:
    a: G1
    b: G2
    c: G1

    def as_tuple(self):
        return (self.a.as_tuple(), self.b.as_tuple(), self.c.as_tuple())

    @staticmethod
    def from_proof(proof) -> 'Proof':
        """proof: circom_compat_b200.Proof (canonical affine coordinates; zeros = infinity)"""
        return Proof(G1(*proof.a), G2(proof.b[0], proof.b[1]), G1(*proof.c))

    @staticmethod
    def from_tuple(t) -> 'Proof':
        return Proof(G1.from_tuple(t[0]), G2.from_tuple(t[1]), G1.from_tuple(t[2]))

    def to_proof(self):
        """From<Proof> for ark_groth16::Proof<Bn254> (src/ethereum.rs:120-128): back to the prover's Proof (256-byte
        canonical little-endian view: A.x, A.y, B.x.c0, B.x.c1, B.y.c0, B.y.c1, C.x, C.y; infinity = zeros)"""
        from .groth16 import Proof as ArkProof
        words = [self.a.x, self.a.y, self.b.x[0], self.b.x[1], self.b.y[0], self.b.y[1], self.c.x, self.c.y]
        return ArkProof(b''.join(int(w).to_bytes(32, 'little') for w in words))

    def calldata(self) -> bytes:
        """abi.encode(uint[2] a, uint[2][2] b, uint[2] c) as the snarkjs / tests/verifier.sol verifier expects"""
        a, b, c = self.as_tuple()
        words = [a[0], a[1], b[0][0], b[0][1], b[1][0], b[1][1], c[0], c[1]]
        return b''.join(int(w).to_bytes(32, 'big') for w in words)


@dataclass(frozen=True)
class VerifyingKey:
    alpha1: G1
    beta2: G2
    gamma2: G2
    delta2: G2
    ic: List[G1]

    def as_tuple(self):
        return (self.alpha1.as_tuple(), self.beta2.as_tuple(), self.gamma2.as_tuple(), self.delta2.as_tuple(), [i.as_tuple() for i in self.ic])

    @staticmethod
    def from_proving_key(pk) -> 'VerifyingKey':
        def g1(arr):
            v = _fq_from_mont_words(arr)
            return G1(v[0], v[1])

        def g2(arr):
            v = _fq_from_mont_words(arr)
            return G2((v[0], v[1]), (v[2], v[3]))
        return VerifyingKey(g1(pk.alpha_g1), g2(pk.beta_g2), g2(pk.gamma_g2), g2(pk.delta_g2), [g1(p) for p in pk.gamma_abc_g1])


    @staticmethod
    def from_tuple(t) -> 'VerifyingKey':
        return VerifyingKey(G1.from_tuple(t[0]), G2.from_tuple(t[1]), G2.from_tuple(t[2]), G2.from_tuple(t[3]), [G1.from_tuple(i) for i in t[4]])

    @staticmethod
    def from_verifying_key(vk) -> 'VerifyingKey':
        """From<ark_groth16::VerifyingKey<Bn254>> for VerifyingKey (src/ethereum.rs:151-161); vk = verifier.VerifyingKey"""
        return VerifyingKey(G1.from_affine(vk.alpha_g1), G2.from_affine(vk.beta_g2), G2.from_affine(vk.gamma_g2), G2.from_affine(vk.delta_g2),
                            [G1.from_affine(p) for p in vk.gamma_abc_g1])

    def to_verifying_key(self):
        """From<VerifyingKey> for ark_groth16::VerifyingKey<Bn254> (src/ethereum.rs:163-174): the object the host verifier
        (verifier.process_vk / verify) takes"""
        from .verifier import VerifyingKey as ArkVk
        return ArkVk(self.alpha1.to_affine(), self.beta2.to_affine(), self.gamma2.to_affine(), self.delta2.to_affine(), [p.to_affine() for p in self.ic])


def inputs(public_inputs) -> List[int]:
    """Inputs(&[Fr]) -> Vec<U256> (src/ethereum.rs:10-18): canonical integers of w[1..num_inputs]"""
    return [int(x) % R_MOD for x in public_inputs]


# ---------------------------------------------------------------------------------------------- ark-serialize
def _fq_is_neg(y: int) -> bool:
    return y > (Q_MOD - y) % Q_MOD


def _fq2_is_neg(y) -> bool:
    n = ((Q_MOD - y[0]) % Q_MOD, (Q_MOD - y[1]) % Q_MOD)
    return (y[1], y[0]) > (n[1], n[0])


def _g1_bytes(p: G1, compressed: bool) -> bytes:
    inf = p.x == 0 and p.y == 0
    if compressed:
        b = bytearray(p.x.to_bytes(32, 'little'))
        b[31] |= FLAG_INF if inf else (FLAG_NEG if _fq_is_neg(p.y) else 0)
        return bytes(b)
    # Compress::No: x plain, then y.serialize_with_flags(item.to_flags()) - ark-ec 0.5 writes the YIsNegative bit on y
    # here too (short_weierstrass SWCurveConfig::serialize_with_mode), not only the infinity flag
    b = bytearray(p.x.to_bytes(32, 'little') + p.y.to_bytes(32, 'little'))
    b[63] |= FLAG_INF if inf else (FLAG_NEG if _fq_is_neg(p.y) else 0)
    return bytes(b)


def _g2_bytes(p: G2, compressed: bool) -> bytes:
    inf = p.x == (0, 0) and p.y == (0, 0)
    xb = p.x[0].to_bytes(32, 'little') + p.x[1].to_bytes(32, 'little')
    if compressed:
        b = bytearray(xb)
        b[63] |= FLAG_INF if inf else (FLAG_NEG if _fq2_is_neg(p.y) else 0)
        return bytes(b)
    b = bytearray(xb + p.y[0].to_bytes(32, 'little') + p.y[1].to_bytes(32, 'little'))
    b[127] |= FLAG_INF if inf else (FLAG_NEG if _fq2_is_neg(p.y) else 0)      # flags ride on the last byte of y.c1
    return bytes(b)


def serialize_compressed(proof: Proof) -> bytes:
    """Proof<Bn254>::serialize_compressed: 32 + 64 + 32 = 128 bytes"""
    return _g1_bytes(proof.a, True) + _g2_bytes(proof.b, True) + _g1_bytes(proof.c, True)


def serialize_uncompressed(proof: Proof) -> bytes:
    """Proof<Bn254>::serialize_uncompressed: 64 + 128 + 64 = 256 bytes.  Each point is x || y with the SWFlags of the
    point (bit 7 = y is the larger of {y, -y}, bit 6 = infinity) OR-ed into the last byte of y, exactly as in the
    compressed form they ride on x; clear the top two bits of bytes 63 / 191 / 255 to recover the raw coordinates."""
    return _g1_bytes(proof.a, False) + _g2_bytes(proof.b, False) + _g1_bytes(proof.c, False)
