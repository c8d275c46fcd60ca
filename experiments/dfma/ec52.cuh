// ec52.cuh - ROUND-2 CANDIDATE: XYZZ mixed addition for BN254 G1 on the FP64 pipe (see fp52.cuh, fp52_model.py).
//
// Domains: table points arrive exactly as libb2groth stores them, residues x~ = x * 2^256 mod p (canonical, 8 x u32).  The
// accumulator keeps X, Y in the 2^260 Montgomery domain and ZZ, ZZZ with an extra factor 16 (ZZ = zz * 16 * 2^260), so that
// U2 = mont(x~, ZZ) = x zz 2^260 lands in the X domain without converting the table point.  No modular reduction anywhere
// (almost-Montgomery products, signed limbs); three limb normalisations per addition (P, R, X3) keep every product split
// inside its precondition: magnitudes stay below |X| < 2.1 p, |Y| < 1.1 p, |P| < 2.6 p, |R| < 1.6 p (fp52_model.py).
//
// Exceptional case P = 0 (mod p) - equal x-coordinates, i.e. a doubling or a cancellation - is NOT handled here: a cheap
// necessary condition (the balanced low limb of P is one of 0, +-p_0, +-2p_0) makes madd52 return false, and the caller hands
// the whole run of sorted entries to the integer kernel, which has the complete addition law.  False positives cost a
// re-run with probability 2^-50 per addition; false negatives are impossible.
#pragma once
#include "fp52.cuh"

namespace b2g52 {

struct Pt52 { fe52 X, Y, ZZ, ZZZ; };

// first point of a run: acc = (x, y), zz = zzz = 1
__device__ __forceinline__ void from_affine52(Pt52& acc, const fe52& x2, const fe52& y2) {
    const fe52 k = k264();
    acc.X = mont_mul(x2, k); acc.Y = mont_mul(y2, k);
    acc.ZZ = k; acc.ZZZ = k;
}

// acc += (x2, y2) (limb-normalised residues in the 2^256 domain; y2 may have been negated by the caller); returns false
// when the addition must be redone by the integer kernel
__device__ __forceinline__ bool madd52(Pt52& acc, const fe52& x2, const fe52& y2) {
    const fe52 U2 = mont_mul(x2, acc.ZZ), S2 = mont_mul(y2, acc.ZZZ);
    const fe52 P = normalize(sub(U2, acc.X)), R = normalize(sub(S2, acc.Y));
    const double l0 = fabs(P.l[0]);
    if (l0 == 0.0 || l0 == 154029749239111.0 || l0 == 308059498478222.0) return false;
    const fe52 PP = mont_sqr(P), PPP = mont_mul(P, PP), Q = mont_mul(acc.X, PP), RR = mont_sqr(R);
    const fe52 X3 = normalize(sub(sub(RR, PPP), add(Q, Q)));
    acc.Y = mont_mul_sub(R, sub(Q, X3), acc.Y, PPP);
    acc.X = X3;
    acc.ZZ = mont_mul(acc.ZZ, PP);
    acc.ZZZ = mont_mul(acc.ZZZ, PPP);
    return true;
}

// -> the product's representation: X, Y, ZZ, ZZZ as canonical residues in the 2^256 Montgomery domain (8 x u32 each).
// `reduce3(w)` must subtract p from the 8-word value while it is >= p, up to three times.
template <class Reduce>
__device__ __forceinline__ void store52(const Pt52& acc, uint32_t* out /* 32 words */, Reduce reduce3) {
    const fe52 c256 = k256(), c252 = k252();
    to_u32_plus_2p(mont_mul(acc.X, c256), out);        reduce3(out);
    to_u32_plus_2p(mont_mul(acc.Y, c256), out + 8);    reduce3(out + 8);
    to_u32_plus_2p(mont_mul(acc.ZZ, c252), out + 16);  reduce3(out + 16);
    to_u32_plus_2p(mont_mul(acc.ZZZ, c252), out + 24); reduce3(out + 24);
}

}  // namespace b2g52
