// fq2_52.cuh - ROUND-2 CANDIDATE: Fq2 = Fq[u]/(u^2 + 1) and the G2 XYZZ mixed addition on the FP64 pipe.
// Model and bounds: fp52_model.py (mul2 / sqr2 / mul_sub2 / madd52_g2).  Every Fq2 product is a pair of signed sums of limb
// products with one reduction each: c0 = a0 b0 - a1 b1, c1 = a0 b1 + a1 b0 (2 x 75 splits); the square is c0 = a0^2 - a1^2
// (two triangles, 55 splits), c1 = (2 a0) a1 (50 splits); a b - c d is 2 x 125 splits.  Same domains and the same
// equal-x replay rule as ec52.cuh.
#pragma once
#include "fp52.cuh"

namespace b2g52 {

struct fe52x2 { fe52 c0, c1; };

// The six wide sums below are ~350 - 840 instructions each; fully inlined a G2 mixed addition is 9 000 instructions (144 KB,
// past the instruction cache).  -DB2G52_FQ2_CALL=__noinline__ turns them into real calls (operands by reference in local
// memory, 54 KB of code in total); which one wins is a measurement for round 2.
#ifndef B2G52_FQ2_CALL
#define B2G52_FQ2_CALL __forceinline__
#endif
__device__ B2G52_FQ2_CALL fe52 fq2_mul_c0(const fe52x2& a, const fe52x2& b) { return mont_sum<2, 2u, false>(a.c0, b.c0, a.c1, b.c1, a.c0, b.c0, a.c0, b.c0); }
__device__ B2G52_FQ2_CALL fe52 fq2_mul_c1(const fe52x2& a, const fe52x2& b) { return mont_sum<2, 0u, false>(a.c0, b.c1, a.c1, b.c0, a.c0, b.c0, a.c0, b.c0); }
__device__ B2G52_FQ2_CALL fe52 fq2_sqr_c0(const fe52x2& a) { return mont_sum<2, 2u, true>(a.c0, a.c0, a.c1, a.c1, a.c0, a.c0, a.c0, a.c0); }
__device__ B2G52_FQ2_CALL fe52 fq2_sqr_c1(const fe52x2& a) { return mont_mul(add(a.c0, a.c0), a.c1); }
__device__ B2G52_FQ2_CALL fe52 fq2_mul_sub_c0(const fe52x2& a, const fe52x2& b, const fe52x2& c, const fe52x2& d) {
    return mont_sum<4, 6u, false>(a.c0, b.c0, a.c1, b.c1, c.c0, d.c0, c.c1, d.c1);      // + - - +
}
__device__ B2G52_FQ2_CALL fe52 fq2_mul_sub_c1(const fe52x2& a, const fe52x2& b, const fe52x2& c, const fe52x2& d) {
    return mont_sum<4, 12u, false>(a.c0, b.c1, a.c1, b.c0, c.c0, d.c1, c.c1, d.c0);     // + + - -
}
__device__ __forceinline__ fe52x2 mul2(const fe52x2& a, const fe52x2& b) { fe52x2 r; r.c0 = fq2_mul_c0(a, b); r.c1 = fq2_mul_c1(a, b); return r; }
// both components of a must be limb-normalised (|l| <= 2^51)
__device__ __forceinline__ fe52x2 sqr2(const fe52x2& a) { fe52x2 r; r.c0 = fq2_sqr_c0(a); r.c1 = fq2_sqr_c1(a); return r; }
__device__ __forceinline__ fe52x2 mul_sub2(const fe52x2& a, const fe52x2& b, const fe52x2& c, const fe52x2& d) {
    fe52x2 r; r.c0 = fq2_mul_sub_c0(a, b, c, d); r.c1 = fq2_mul_sub_c1(a, b, c, d); return r;
}
__device__ __forceinline__ fe52x2 add2(const fe52x2& a, const fe52x2& b) { fe52x2 r; r.c0 = add(a.c0, b.c0); r.c1 = add(a.c1, b.c1); return r; }
__device__ __forceinline__ fe52x2 sub2(const fe52x2& a, const fe52x2& b) { fe52x2 r; r.c0 = sub(a.c0, b.c0); r.c1 = sub(a.c1, b.c1); return r; }
__device__ __forceinline__ fe52x2 neg2(const fe52x2& a) { fe52x2 r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
__device__ __forceinline__ fe52x2 norm2(const fe52x2& a) { fe52x2 r; r.c0 = normalize(a.c0); r.c1 = normalize(a.c1); return r; }

struct Pt52x2 { fe52x2 X, Y, ZZ, ZZZ; };

__device__ __forceinline__ void from_affine52_g2(Pt52x2& acc, const fe52x2& x2, const fe52x2& y2) {
    fe52x2 k; k.c0 = k264();
    #pragma unroll
    for (int i = 0; i < 5; i++) k.c1.l[i] = 0.0;
    acc.X.c0 = mont_mul(x2.c0, k.c0); acc.X.c1 = mont_mul(x2.c1, k.c0);
    acc.Y.c0 = mont_mul(y2.c0, k.c0); acc.Y.c1 = mont_mul(y2.c1, k.c0);
    acc.ZZ = k; acc.ZZZ = k;
}

__device__ __forceinline__ bool cand52(double l) {
    const double a = fabs(l);
    return a == 0.0 || a == 154029749239111.0 || a == 308059498478222.0;
}

// acc += (x2, y2); false = equal x-coordinates (or a 2^-100 false alarm): replay the run with the integer kernel
__device__ __forceinline__ bool madd52_g2(Pt52x2& acc, const fe52x2& x2, const fe52x2& y2) {
    const fe52x2 U2 = mul2(x2, acc.ZZ), S2 = mul2(y2, acc.ZZZ);
    const fe52x2 P = norm2(sub2(U2, acc.X)), R = norm2(sub2(S2, acc.Y));
    if (cand52(P.c0.l[0]) && cand52(P.c1.l[0])) return false;
    const fe52x2 PP = sqr2(P), PPP = mul2(P, PP), Q = mul2(acc.X, PP), RR = sqr2(R);
    const fe52x2 X3 = norm2(sub2(sub2(RR, PPP), add2(Q, Q)));
    acc.Y = mul_sub2(R, sub2(Q, X3), acc.Y, PPP);
    acc.X = X3;
    acc.ZZ = mul2(acc.ZZ, PP);
    acc.ZZZ = mul2(acc.ZZZ, PPP);
    return true;
}

// -> the product's representation (8 Fq residues: X.c0, X.c1, Y.c0, Y.c1, ZZ.c0, ZZ.c1, ZZZ.c0, ZZZ.c1; 64 words)
template <class Reduce>
__device__ __forceinline__ void store52_g2(const Pt52x2& acc, uint32_t* out, Reduce reduce3) {
    const fe52 c256 = k256(), c252 = k252();
    to_u32_plus_2p(mont_mul(acc.X.c0, c256), out);          reduce3(out);
    to_u32_plus_2p(mont_mul(acc.X.c1, c256), out + 8);      reduce3(out + 8);
    to_u32_plus_2p(mont_mul(acc.Y.c0, c256), out + 16);     reduce3(out + 16);
    to_u32_plus_2p(mont_mul(acc.Y.c1, c256), out + 24);     reduce3(out + 24);
    to_u32_plus_2p(mont_mul(acc.ZZ.c0, c252), out + 32);    reduce3(out + 32);
    to_u32_plus_2p(mont_mul(acc.ZZ.c1, c252), out + 40);    reduce3(out + 40);
    to_u32_plus_2p(mont_mul(acc.ZZZ.c0, c252), out + 48);   reduce3(out + 48);
    to_u32_plus_2p(mont_mul(acc.ZZZ.c1, c252), out + 56);   reduce3(out + 56);
}

}  // namespace b2g52
