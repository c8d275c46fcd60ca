// fp52.cuh - ROUND-2 CANDIDATE (not part of libb2groth.so, not used by tests or bench): BN254 Fq on the FP64 pipe.
//
// Why: the MSM accumulation kernels are bound by the IMAD.WIDE (fmaheavy) pipe.  A B200 SM also has
// an FP64 pipe (one DFMA per 2 cycles per scheduler, 37 TFLOP/s) that the prover leaves idle.  Warps multiplying with the
// routine below use the FP64 and ALU pipes only, so they can share an SM with warps of the integer kernel.
//
// Exact model, operand-range rules and the cost estimate: fp52_model.py (same directory).
// Value = sum l[i] * 2^(52 i), five signed integer-valued doubles; Montgomery radix R = 2^260; "almost-Montgomery":
// |a| < A p, |b| < B p  =>  |output| < (0.0118 A B + 0.5) p with no final subtraction (1.26 p for A = B = 8); a whole XYZZ
// mixed addition keeps every operand below 3 p (fp52_model.py: madd_bounds).  A 52 x 52 product is split exactly by two DFMAs and a DADD:
//     hf = fma_rd(a, b, M)                 M = 1.5 * 2^104 (ulp 2^52):  hf = M + H,  H = a*b rounded down to a multiple of 2^52
//     lo = fma_rn(a, b, (M + 2^52) - hf)   = a*b - H + 2^52, in [2^52, 2^53), exact
// The raw bit patterns of hf and lo are offset-binary integers (H / 2^52 and a*b - H), so column sums are plain 64-bit
// integer additions of the register pairs; the offsets are folded into the columns' initial values.
// Precondition of every split: -2^103 <= a*b < 2^103  (limbs |.| <= 2^51 against limbs |.| <= 2^52).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2g52 {

struct fe52 { double l[5]; };

__device__ __forceinline__ double mk(unsigned long long bits) { return __longlong_as_double((long long)bits); }
#define B2G52_M    mk(0x4678000000000000ull)    /* 1.5 * 2^104                    */
#define B2G52_MK   mk(0x4678000000000001ull)    /* 1.5 * 2^104 + 2^52             */
#define B2G52_C52B mk(0x4338000000000000ull)    /* 1.5 * 2^52: int <-> double     */
#define B2G52_RAW_H 0x4678000000000000ll        /* raw(hf) = RAW_H + H / 2^52     */
#define B2G52_RAW_L 0x4330000000000000ll        /* raw(lo) = RAW_L + (a*b - H)    */
#define B2G52_RAW_B 0x4338000000000000ll        /* raw(1.5 * 2^52 + v) = RAW_B + v */

// p as balanced limbs and -p^-1 mod 2^52 (balanced): printed by fp52_model.py (PL, PINV_B)
__device__ __forceinline__ double p_limb(int j) {
    return j == 0 ? 154029749239111.0 : j == 1 ? -1945555279752254.0 : j == 2 ? 423691504025963.0
         : j == 3 ? -1685982885422232.0 : 53207371014450.0;
}
#define B2G52_PINV 571208714576777.0

// one exact product split; adds (SIGN = +1) or subtracts (-1) the two halves into the integer columns lo_col (weight of
// a*b) and hi_col (weight * 2^52).  All column arithmetic is modulo 2^64.
template <int SIGN>
__device__ __forceinline__ void split_acc(double a, double b, unsigned long long& lo_col, unsigned long long& hi_col) {
    const double hf = __fma_rd(a, b, B2G52_M);
    const double t = __dadd_rn(B2G52_MK, -hf);              // 2^52 - H, exact
    const double lo = __fma_rn(a, b, t);
    if (SIGN > 0) { hi_col += (unsigned long long)__double_as_longlong(hf); lo_col += (unsigned long long)__double_as_longlong(lo); }
    else          { hi_col -= (unsigned long long)__double_as_longlong(hf); lo_col -= (unsigned long long)__double_as_longlong(lo); }
}
// low half only (q = l * p' mod 2^52) as a balanced double in [-2^51, 2^51]: round-to-nearest split, no offset
__device__ __forceinline__ double split_low(double a, double b) {
    const double hf = __fma_rn(a, b, B2G52_M);
    return __fma_rn(a, b, __dadd_rn(B2G52_M, -hf));
}
// signed 64-bit column value -> balanced low limb (as a double, |l| <= 2^51); the carry is added to `next`
__device__ __forceinline__ double column_low(unsigned long long t, unsigned long long& next) {
    const long long c = ((long long)t + (1ll << 51)) >> 52;
    const long long l = (long long)(t - ((unsigned long long)c << 52));
    next += (unsigned long long)c;
    return __dadd_rn(__longlong_as_double(l + B2G52_RAW_B), -B2G52_C52B);
}

// (sum_k s_k * A_k * B_k + q p) / 2^260 with ONE reduction; N <= 4 terms, s_k = -1 where bit k of NEG is set.
// SQUARE: every term is a square A_k^2 (B_k ignored): upper triangle with doubled cross terms, needs |A_k.l[i]| <= 2^51.
// Every limb product fed to a split must lie in [-2^103, 2^103].  Columns start at minus the sum of the raw-pattern offsets
// they are going to receive.  Mirrored line by line by fp52_model.py: mont_sum_gpu.
template <int N, unsigned NEG, bool SQUARE>
__device__ __forceinline__ fe52 mont_sum(const fe52& a0, const fe52& b0, const fe52& a1, const fe52& b1,
                                         const fe52& a2, const fe52& b2, const fe52& a3, const fe52& b3) {
    unsigned long long T[10];
    #pragma unroll
    for (int k = 0; k < 10; k++) {
        int nlo = 0, nhi = 0;
        #pragma unroll
        for (int i = 0; i < 5; i++)
            #pragma unroll
            for (int j = 0; j < 5; j++) {
                const bool in_prod = !SQUARE || j >= i;
                #pragma unroll
                for (int t = 0; t < N; t++) {
                    const int s = ((NEG >> t) & 1u) ? -1 : 1;
                    if (in_prod && i + j == k) nlo += s;
                    if (in_prod && i + j + 1 == k) nhi += s;
                }
                if (i + j == k) nlo++;                                        // reduction round i, limb j
                if (i + j + 1 == k) nhi++;
            }
        T[k] = 0ull - ((unsigned long long)(long long)nlo * (unsigned long long)B2G52_RAW_L + (unsigned long long)(long long)nhi * (unsigned long long)B2G52_RAW_H);
    }
    #pragma unroll
    for (int t = 0; t < N; t++) {
        const fe52& a = t == 0 ? a0 : t == 1 ? a1 : t == 2 ? a2 : a3;
        const fe52& b = t == 0 ? b0 : t == 1 ? b1 : t == 2 ? b2 : b3;
        const bool neg = (NEG >> t) & 1u;
        #pragma unroll
        for (int i = 0; i < 5; i++) {
            if (SQUARE) {
                if (neg) split_acc<-1>(a.l[i], a.l[i], T[2 * i], T[2 * i + 1]); else split_acc<1>(a.l[i], a.l[i], T[2 * i], T[2 * i + 1]);
                const double d2 = __dadd_rn(a.l[i], a.l[i]);
                #pragma unroll
                for (int j = i + 1; j < 5; j++) { if (neg) split_acc<-1>(d2, a.l[j], T[i + j], T[i + j + 1]); else split_acc<1>(d2, a.l[j], T[i + j], T[i + j + 1]); }
            } else {
                #pragma unroll
                for (int j = 0; j < 5; j++) { if (neg) split_acc<-1>(a.l[i], b.l[j], T[i + j], T[i + j + 1]); else split_acc<1>(a.l[i], b.l[j], T[i + j], T[i + j + 1]); }
            }
        }
    }
    #pragma unroll
    for (int i = 0; i < 5; i++) {
        // the one offset pre-subtracted from column i that has not arrived yet is round i's own lo(q p_0)
        unsigned long long dummy = 0;
        const double l = column_low(T[i] + (unsigned long long)B2G52_RAW_L, dummy);
        const double q = split_low(l, B2G52_PINV);
        #pragma unroll
        for (int j = 0; j < 5; j++) split_acc<1>(q, p_limb(j), T[i + j], T[i + j + 1]);
        T[i + 1] += (unsigned long long)((long long)T[i] >> 52);              // column i is now an exact multiple of 2^52
    }
    fe52 r;
    #pragma unroll
    for (int k = 5; k < 9; k++) r.l[k - 5] = column_low(T[k], T[k + 1]);
    r.l[4] = __dadd_rn(__longlong_as_double((long long)T[9] + B2G52_RAW_B), -B2G52_C52B);
    return r;
}
__device__ __forceinline__ fe52 mont_mul(const fe52& a, const fe52& b) { return mont_sum<1, 0u, false>(a, b, a, b, a, b, a, b); }
__device__ __forceinline__ fe52 mont_sqr(const fe52& a) { return mont_sum<1, 0u, true>(a, a, a, a, a, a, a, a); }            // |a.l[i]| <= 2^51
__device__ __forceinline__ fe52 mont_mul_sub(const fe52& a, const fe52& b, const fe52& c, const fe52& d) { return mont_sum<2, 2u, false>(a, b, c, d, a, b, a, b); }

__device__ __forceinline__ fe52 add(const fe52& a, const fe52& b) {
    fe52 r;
    #pragma unroll
    for (int i = 0; i < 5; i++) r.l[i] = __dadd_rn(a.l[i], b.l[i]);
    return r;
}
__device__ __forceinline__ fe52 sub(const fe52& a, const fe52& b) {
    fe52 r;
    #pragma unroll
    for (int i = 0; i < 5; i++) r.l[i] = __dadd_rn(a.l[i], -b.l[i]);
    return r;
}
// carry-propagate to balanced limbs, entirely on the FP64 pipe: c = (v + M) - M is v rounded to a multiple of 2^52
__device__ __forceinline__ fe52 normalize(const fe52& a) {
    fe52 r;
    double carry = 0.0;
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const double v = __dadd_rn(a.l[i], carry);
        const double c = __dadd_rn(__dadd_rn(v, B2G52_M), -B2G52_M);
        r.l[i] = __dadd_rn(v, -c);
        carry = __dmul_rn(c, mk(0x3cb0000000000000ull));      // 2^-52
    }
    r.l[4] = __dadd_rn(a.l[4], carry);
    return r;
}

__device__ __forceinline__ fe52 neg(const fe52& a) {
    fe52 r;
    #pragma unroll
    for (int i = 0; i < 5; i++) r.l[i] = -a.l[i];
    return r;
}
__device__ __forceinline__ fe52 k264() { fe52 r; r.l[0] = -1390697610713478.0; r.l[1] = -323933227188290.0; r.l[2] = -1721143775100325.0; r.l[3] = -504184215139471.0; r.l[4] = 14813684363143.0; return r; }   // 2^264 mod p
__device__ __forceinline__ fe52 k256() { fe52 r; r.l[0] = -770148746195555.0; r.l[1] = 720577144020278.0; r.l[2] = -2118457520129813.0; r.l[3] = -577284827629832.0; r.l[4] = 15438121638408.0; return r; }    // 2^256 mod p
__device__ __forceinline__ fe52 k252() { fe52 r; r.l[0] = 0.0; r.l[1] = 0.0; r.l[2] = 0.0; r.l[3] = 0.0; r.l[4] = 17592186044416.0; return r; }                                                             // 2^252

// canonical 8 x u32 (little-endian, < 2^256) -> five balanced limbs (integer shifts, then the 1.5 * 2^52 bit trick)
__device__ __forceinline__ fe52 from_u32(const uint32_t* x) {
    unsigned long long w[4];
    #pragma unroll
    for (int i = 0; i < 4; i++) w[i] = (unsigned long long)x[2 * i] | ((unsigned long long)x[2 * i + 1] << 32);
    const unsigned long long mask = (1ull << 52) - 1;
    long long chunk[5];
    chunk[0] = (long long)(w[0] & mask);
    chunk[1] = (long long)(((w[0] >> 52) | (w[1] << 12)) & mask);
    chunk[2] = (long long)(((w[1] >> 40) | (w[2] << 24)) & mask);
    chunk[3] = (long long)(((w[2] >> 28) | (w[3] << 36)) & mask);
    chunk[4] = (long long)(w[3] >> 16);
    fe52 r; long long carry = 0;
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        long long t = chunk[i] + carry;
        carry = (t + (1ll << 51)) >> 52;                     // 0 or 1
        t -= (long long)((unsigned long long)carry << 52);
        r.l[i] = __dadd_rn(__longlong_as_double(t + B2G52_RAW_B), -B2G52_C52B);
    }
    r.l[4] = __dadd_rn(__longlong_as_double(chunk[4] + carry + B2G52_RAW_B), -B2G52_C52B);
    return r;
}
// five signed limbs, |value| < 2 p  ->  8 x u32 holding value + 2p in (0, 4p); the caller subtracts p up to three times
__device__ __forceinline__ void to_u32_plus_2p(const fe52& a, uint32_t* out) {
    const long long two_p[5] = {308059498478222ll, -3891110559504508ll, 847383008051926ll, -3371965770844464ll, 106414742028900ll};
    long long carry = 0; unsigned long long chunk[5];
    #pragma unroll
    for (int i = 0; i < 5; i++) {
        const long long v = __double_as_longlong(__dadd_rn(a.l[i], B2G52_C52B)) - B2G52_RAW_B;      // exact for normalised limbs, |l| <= 2^51
        const long long t = v + two_p[i] + carry;
        if (i < 4) { chunk[i] = (unsigned long long)t & ((1ull << 52) - 1); carry = t >> 52; } else chunk[i] = (unsigned long long)t;
    }
    unsigned long long w[4];
    w[0] = chunk[0] | (chunk[1] << 52);
    w[1] = (chunk[1] >> 12) | (chunk[2] << 40);
    w[2] = (chunk[2] >> 24) | (chunk[3] << 28);
    w[3] = (chunk[3] >> 36) | (chunk[4] << 16);
    #pragma unroll
    for (int i = 0; i < 4; i++) { out[2 * i] = (uint32_t)w[i]; out[2 * i + 1] = (uint32_t)(w[i] >> 32); }
}

}  // namespace b2g52
