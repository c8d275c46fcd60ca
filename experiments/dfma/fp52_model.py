"""Exact model of a double-precision (DFMA) Montgomery multiplier for BN254 Fq -- ROUND-2 CANDIDATE, not product code.

Idea: the accumulation kernels are bound by the IMAD.WIDE (fmaheavy) pipe.  A B200 SM also has an FP64 pipe that issues
a DFMA every 2 cycles per scheduler (37 TFLOP/s), idle today.  A 52 x 52 -> 104-bit product can be split exactly into a
high and a low half with two DFMAs:  hf = fma(a, b, M) with M = 1.5 * 2^104 rounds a*b to a multiple H of 2^52,
lo = fma(a, b, -(hf - M)) = a*b - H exactly.  Warps running this variant use the FP64 + ALU pipes, warps running the
integer variant use fmaheavy, so the two kinds can share an SM.

Representation: 5 signed ("balanced") limbs of 52 bits, value = sum l_i 2^(52 i), each limb an integer-valued double;
R = 2^260.  "Almost-Montgomery": |a| < A p, |b| < B p gives |output| < (0.0118 A B + 0.5) p with NO final subtraction
(p / 2^260 = 0.0118; the 0.5 is the balanced quotient), e.g. 1.26 p for A = B = 8.  That is enough for a whole XYZZ mixed
addition to run without a single modular reduction: iterating the formulas gives |X| < 2.2 p, |Y| < 1.1 p, |P| < 2.8 p,
|R| < 1.7 p, every product operand below 3 p (madd_bounds()).

Every floating-point step below is exact integer arithmetic in disguise; this file checks that claim with Python
integers (float(int) is correctly rounded, so fma on integer-valued doubles is float(a*b + c)).
"""
import random

W = 52
L = 5
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 1 << (W * L)
M = 1.5 * 2.0 ** 104            # product-splitting magic: ulp(M) = 2^52
HALF = 1 << (W - 1)


def fma(a, b, c):
    """IEEE fma (round to nearest even) for integer-valued doubles."""
    for v in (a, b, c):
        assert float(v) == v and v == int(v)
    return float(int(a) * int(b) + int(c))


def fadd(a, b):
    return float(int(a) + int(b))


def balanced(x, n=L):
    """integer -> n balanced limbs (|l_i| <= 2^51 for i < n-1), as doubles"""
    out = []
    for i in range(n - 1):
        l = ((x + HALF) % (1 << W)) - HALF
        out.append(float(l)); x = (x - l) >> W
    out.append(float(x))
    assert abs(x) < 2 ** 53
    return out


def value(l):
    return sum(int(v) << (W * i) for i, v in enumerate(l))


PL = balanced(P)
PINV = (-pow(P, -1, 1 << W)) % (1 << W)
PINV_B = float(((PINV + HALF) % (1 << W)) - HALF)          # balanced -p^-1 mod 2^52


def split(a, b):
    """a*b = H + lo, H a multiple of 2^52, |lo| <= 2^51; requires |a*b| <= 2^103.  2 DFMA + 1 DADD."""
    assert abs(int(a) * int(b)) <= 1 << 103, "split precondition"
    hf = fma(a, b, M)
    assert 2.0 ** 104 <= hf <= 2.0 ** 105
    H = fadd(hf, -M)                      # exact: both multiples of 2^52
    lo = fma(a, b, -H)
    assert int(H) + int(lo) == int(a) * int(b) and abs(lo) <= HALF and int(H) % (1 << W) == 0
    return H, lo


def norm_column(t):
    """int64 column -> (balanced low limb as double, carry).  Integer ops on the GPU (ALU pipe)."""
    assert abs(t) < 1 << 62
    l = ((t + HALF) % (1 << W)) - HALF
    return float(l), (t - l) >> W


def mont_mul(a, b, count=None):
    """a, b: 5 limbs each with |a_i * b_j| <= 2^103.  Returns 5 balanced limbs of (a*b + q*p) / 2^260.
    GPU cost: 25 + 5 * (1 + 5) = 55 splits = 110 DFMA + 55 DADD, ~110 column terms (integer adds), 10 column normalisations."""
    T = [0] * (2 * L)                     # int64 columns (sums of raw double bit patterns on the GPU)
    for i in range(L):
        for j in range(L):
            H, lo = split(a[i], b[j])
            T[i + j] += int(lo); T[i + j + 1] += int(H) >> W
    for i in range(L):
        l, _ = norm_column(T[i])          # balanced low 52 bits of the running column
        _, q = split(l, PINV_B)           # q = l * (-p^-1) mod 2^52, balanced: only the low half is used
        for j in range(L):
            H, lo = split(q, PL[j])
            T[i + j] += int(lo); T[i + j + 1] += int(H) >> W
        assert T[i] % (1 << W) == 0       # column i has been cancelled: what is left is a carry
        T[i + 1] += T[i] >> W; T[i] = 0
    out = []
    carry = 0
    for k in range(L, 2 * L - 1):
        l, carry = norm_column(T[k] + carry); out.append(l)
    top = T[2 * L - 1] + carry
    assert abs(top) < 1 << 52
    out.append(float(top))
    if count is not None:
        count['max_col'] = max(count.get('max_col', 0), max(abs(t) for t in T))
    return out


def mont_sqr(a):
    """15 splits instead of 25 for the product phase: cross terms use the doubled limb 2 a_i (exact; needs |a_i| <= 2^51)."""
    T = [0] * (2 * L)
    for i in range(L):
        H, lo = split(a[i], a[i]); T[2 * i] += int(lo); T[2 * i + 1] += int(H) >> W
        for j in range(i + 1, L):
            H, lo = split(fadd(a[i], a[i]), a[j]); T[i + j] += int(lo); T[i + j + 1] += int(H) >> W
    for i in range(L):
        l, _ = norm_column(T[i])
        _, q = split(l, PINV_B)
        for j in range(L):
            H, lo = split(q, PL[j]); T[i + j] += int(lo); T[i + j + 1] += int(H) >> W
        assert T[i] % (1 << W) == 0
        T[i + 1] += T[i] >> W; T[i] = 0
    out, carry = [], 0
    for k in range(L, 2 * L - 1):
        l, carry = norm_column(T[k] + carry); out.append(l)
    out.append(float(T[2 * L - 1] + carry))
    return out


def ladd(a, b): return [fadd(x, y) for x, y in zip(a, b)]          # limb-wise, no carries (DADD x 5)
def lsub(a, b): return [fadd(x, -y) for x, y in zip(a, b)]


def normalize(a):
    """carry-propagate to balanced limbs (|l_i| <= 2^51, i < 4); FP64 version: c = (l + M) - M, l -= c, next += c * 2^-52"""
    out, carry = [], 0.0
    for i in range(L - 1):
        v = fadd(a[i], carry)
        c = fadd(fadd(v, M), -M)          # nearest multiple of 2^52 (ties to even multiple)
        out.append(fadd(v, -c)); carry = float(int(c) >> W)
        assert abs(out[-1]) <= HALF
    out.append(fadd(a[L - 1], carry))
    return out


def to_mont(x):    return balanced(x * R % P)
def from_mont(l):  return value(l) * pow(R, -1, P) % P


# ---------------------------------------------------------------------------------------------------------------------
# Line-by-line mirror of fp52.cuh (round-down split, raw bit patterns summed as wrapping 64-bit integers, offsets
# pre-subtracted from the columns).  Pins the offset bookkeeping of mont_mul_impl.
import struct
MASK64 = (1 << 64) - 1
RAW_H, RAW_L, RAW_B = 0x4678000000000000, 0x4330000000000000, 0x4338000000000000


def raw(d):        return struct.unpack('<Q', struct.pack('<d', d))[0]
def from_raw(b):   return struct.unpack('<d', struct.pack('<Q', b & MASK64))[0]
def s64(x):        x &= MASK64; return x - (1 << 64) if x >> 63 else x
MK = from_raw(0x4678000000000001)


def fma_rd(a, b, c):
    """fma rounded toward -inf, for the one use here: c = M, result in [2^104, 2^105]"""
    x = int(a) * int(b) + int(c)
    assert (1 << 104) <= x <= (1 << 105)
    return float((x >> 52) << 52)


def split_acc(a, b, T, lo_k, hi_k):
    assert -(1 << 103) <= int(a) * int(b) <= (1 << 103)
    hf = fma_rd(a, b, M)
    t = fadd(MK, -hf)
    lo = fma(a, b, t)
    assert 2.0 ** 52 <= lo < 2.0 ** 53
    T[hi_k] = (T[hi_k] + raw(hf)) & MASK64
    T[lo_k] = (T[lo_k] + raw(lo)) & MASK64


def split_low(a, b):
    """balanced low half: round-to-nearest split, lo = a*b - H in [-2^51, 2^51]"""
    hf = fma(a, b, M)
    return fma(a, b, fadd(M, -hf))


def column_low(t):
    """t: wrapped 64-bit column -> (balanced low limb as double, carry)"""
    t = s64(t)
    c = (t + (1 << 51)) >> 52
    l = t - (c << 52)
    return fadd(from_raw(l + RAW_B), -from_raw(RAW_B)), c


def split_sub(a, b, T, lo_k, hi_k):
    assert -(1 << 103) <= int(a) * int(b) <= (1 << 103)
    hf = fma_rd(a, b, M)
    lo = fma(a, b, fadd(MK, -hf))
    T[hi_k] = (T[hi_k] - raw(hf)) & MASK64
    T[lo_k] = (T[lo_k] - raw(lo)) & MASK64


def mont_core_gpu(mode, a, b, c=None, d=None):
    """mirror of mont_core<MODE>: 0 a*b, 1 a*a, 2 a*b - c*d"""
    T = []
    for k in range(10):
        nlo = nhi = 0
        for i in range(5):
            for j in range(5):
                in_prod = mode == 0 or (mode == 1 and j >= i)
                nlo += (in_prod and i + j == k) + (i + j == k)
                nhi += (in_prod and i + j + 1 == k) + (i + j + 1 == k)
        T.append((-(nlo * RAW_L + nhi * RAW_H)) & MASK64)
    for i in range(5):
        if mode == 1:
            split_acc(a[i], a[i], T, 2 * i, 2 * i + 1)
            a2 = fadd(a[i], a[i])
            for j in range(i + 1, 5): split_acc(a2, a[j], T, i + j, i + j + 1)
        else:
            for j in range(5):
                split_acc(a[i], b[j], T, i + j, i + j + 1)
                if mode == 2: split_sub(c[i], d[j], T, i + j, i + j + 1)
    for i in range(5):
        l, _ = column_low(T[i] + RAW_L)
        q = split_low(l, PINV_B)
        for j in range(5): split_acc(q, PL[j], T, i + j, i + j + 1)
        assert s64(T[i]) % (1 << 52) == 0
        T[i + 1] = (T[i + 1] + (s64(T[i]) >> 52)) & MASK64
    r = []
    for k in range(5, 9):
        l, c_ = column_low(T[k]); T[k + 1] = (T[k + 1] + c_) & MASK64; r.append(l)
    t9 = s64(T[9]); assert abs(t9) < 1 << 51
    r.append(fadd(from_raw(t9 + RAW_B), -from_raw(RAW_B)))
    return r


def mont_mul_gpu(a, b, square=False):
    return mont_core_gpu(1 if square else 0, a, b)


# ---------------------------------------------------------------------------------------------------------------------
# XYZZ mixed addition on top of the almost-Montgomery products (mirrors ec52.cuh).
# Domains: table points arrive as the product stores them, residues x~ = x * 2^256 mod p (canonical, 8 x u32); the
# accumulator keeps X, Y in the R = 2^260 domain and ZZ, ZZZ with an extra factor 16 (ZZ = zz * 16 * 2^260), which makes
# U2 = mont(x~, ZZ) = x zz 2^260 land in the X domain without converting the table point.  No modular reduction anywhere;
# three limb normalisations per addition (P, R, X3) keep every split inside its precondition.
B = 3                                                       # the curve y^2 = x^3 + 3
K264 = balanced(pow(2, 264, P)); K256 = balanced(pow(2, 256, P)); K252 = balanced(pow(2, 252, P))
CAND = {0.0, float(int(PL[0])), float(2 * int(PL[0]))}     # |low limb| of 0, +-p, +-2p in balanced form


def mul_sub_fused(a, b, c, d):
    """(a*b - c*d + q*p) / 2^260 with one reduction: the second product's raw patterns are subtracted, offsets cancel"""
    T = [0] * (2 * L)
    for i in range(L):
        for j in range(L):
            H, lo = split(a[i], b[j]); T[i + j] += int(lo); T[i + j + 1] += int(H) >> W
            H, lo = split(c[i], d[j]); T[i + j] -= int(lo); T[i + j + 1] -= int(H) >> W
    for i in range(L):
        l, _ = norm_column(T[i])
        _, q = split(l, PINV_B)
        for j in range(L):
            H, lo = split(q, PL[j]); T[i + j] += int(lo); T[i + j + 1] += int(H) >> W
        assert T[i] % (1 << W) == 0
        T[i + 1] += T[i] >> W; T[i] = 0
    out, carry = [], 0
    for k in range(L, 2 * L - 1):
        l, carry = norm_column(T[k] + carry); out.append(l)
    out.append(float(T[2 * L - 1] + carry))
    return out


class Acc52:
    def __init__(self): self.inf = True; self.redo = False; self.X = self.Y = self.ZZ = self.ZZZ = None


class _Impl:
    """the two restatements of the field layer: 'rn' (balanced halves, big-int columns) and 'gpu' (mirror of fp52.cuh)"""
    def __init__(self, gpu):
        self.mul = (lambda a, b: mont_core_gpu(0, a, b)) if gpu else mont_mul
        self.sqr = (lambda a: mont_core_gpu(1, a, a)) if gpu else mont_sqr
        self.mul_sub = (lambda a, b, c, d: mont_core_gpu(2, a, b, c, d)) if gpu else mul_sub_fused


IMPL = _Impl(False)


def madd52(acc, xt, yt, negate=False):
    """acc += (x, y) given as residues xt = x * 2^256 mod p, yt likewise (0, 0 = infinity); negate: add (x, -y)"""
    if xt == 0 and yt == 0: return
    x2, y2 = balanced(xt), balanced(yt)
    if negate: y2 = [-v for v in y2]
    if acc.inf:
        acc.X, acc.Y = IMPL.mul(x2, K264), IMPL.mul(y2, K264)
        acc.ZZ, acc.ZZZ = list(K264), list(K264)
        acc.inf = False
        return
    U2, S2 = IMPL.mul(x2, acc.ZZ), IMPL.mul(y2, acc.ZZZ)
    Pn, Rn = normalize(lsub(U2, acc.X)), normalize(lsub(S2, acc.Y))
    if abs(Pn[0]) in CAND:                      # necessary for P = 0 mod p: hand the whole run to the integer kernel
        acc.redo = True
        return
    PP = IMPL.sqr(Pn); PPP = IMPL.mul(Pn, PP); Q = IMPL.mul(acc.X, PP)
    RR = IMPL.sqr(Rn)
    X3 = normalize(lsub(lsub(RR, PPP), ladd(Q, Q)))
    Y3 = IMPL.mul_sub(Rn, lsub(Q, X3), acc.Y, PPP)
    acc.ZZ, acc.ZZZ = IMPL.mul(acc.ZZ, PP), IMPL.mul(acc.ZZZ, PPP)
    acc.X, acc.Y = X3, Y3
    for v, bound in ((acc.X, 2.2), (acc.Y, 1.2), (acc.ZZ, 0.6), (acc.ZZZ, 0.6)):
        assert abs(value(v)) < bound * P


def store52(acc):
    """-> (X, Y, ZZ, ZZZ) as canonical residues in the product's 2^256 Montgomery form"""
    c = lambda v: value(v) % P
    return (c(IMPL.mul(acc.X, K256)), c(IMPL.mul(acc.Y, K256)), c(IMPL.mul(acc.ZZ, K252)), c(IMPL.mul(acc.ZZZ, K252)))


def _ec_check():
    rng = random.Random(7)
    R256 = pow(2, 256, P); R256i = pow(R256, -1, P)
    def rand_point():
        while True:
            x = rng.randrange(P); y2 = (x ** 3 + B) % P
            y = pow(y2, (P + 1) // 4, P)
            if y * y % P == y2: return x, (y if rng.random() < 0.5 else P - y)
    def aff_add(p, q):
        if p is None: return q
        (x1, y1), (x2, y2) = p, q
        if x1 == x2:
            if (y1 + y2) % P == 0: return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
        x3 = (lam * lam - x1 - x2) % P
        return x3, (lam * (x1 - x3) - y1) % P
    for trial in range(40):
        acc, ref = Acc52(), None
        for k in range(24):
            x, y = rand_point(); neg = rng.random() < 0.5
            madd52(acc, x * R256 % P, y * R256 % P, neg)
            ref = aff_add(ref, (x, (P - y) % P if neg else y))
        assert not acc.redo
        X, Y, ZZ, ZZZ = (v * R256i % P for v in store52(acc))
        assert ZZ ** 3 % P == ZZZ ** 2 % P
        assert (X * pow(ZZ, -1, P) % P, Y * pow(ZZZ, -1, P) % P) == ref
    # the exceptional case is caught: adding the same point twice
    acc = Acc52(); x, y = rand_point()
    madd52(acc, x * R256 % P, y * R256 % P); madd52(acc, x * R256 % P, y * R256 % P)
    assert acc.redo
    acc = Acc52(); madd52(acc, x * R256 % P, y * R256 % P); madd52(acc, x * R256 % P, y * R256 % P, True)
    assert acc.redo
    print("ec52 model ok (%s field layer)" % ("fp52.cuh mirror" if IMPL.mul is not mont_mul else "round-to-nearest"))


# ---------------------------------------------------------------------------------------------------------------------
# mirrors of from_u32 / to_u32_plus_2p (fp52.cuh): 8 x u32 <-> five balanced limbs through 64-bit shifts
def from_u32_gpu(words):
    w = [words[2 * i] | (words[2 * i + 1] << 32) for i in range(4)]
    mask = (1 << 52) - 1
    chunk = [w[0] & mask, ((w[0] >> 52) | (w[1] << 12)) & mask, ((w[1] >> 40) | (w[2] << 24)) & mask,
             ((w[2] >> 28) | (w[3] << 36)) & mask, w[3] >> 16]
    out, carry = [], 0
    for i in range(4):
        t = chunk[i] + carry
        carry = (t + (1 << 51)) >> 52
        t -= carry << 52
        out.append(fadd(from_raw(t + RAW_B), -from_raw(RAW_B)))
    out.append(fadd(from_raw(chunk[4] + carry + RAW_B), -from_raw(RAW_B)))
    return out


def to_u32_plus_2p_gpu(a):
    two_p = [2 * int(x) for x in PL]
    carry, chunk = 0, []
    for i in range(5):
        v = s64(raw(fadd(a[i], from_raw(RAW_B))) - RAW_B)
        assert v == int(a[i])
        t = v + two_p[i] + carry
        if i < 4: chunk.append(t & ((1 << 52) - 1)); carry = t >> 52
        else:
            assert 0 <= t < 1 << 48
            chunk.append(t)
    w = [(chunk[0] | (chunk[1] << 52)) & MASK64, ((chunk[1] >> 12) | (chunk[2] << 40)) & MASK64,
         ((chunk[2] >> 24) | (chunk[3] << 28)) & MASK64, ((chunk[3] >> 36) | (chunk[4] << 16)) & MASK64]
    words = []
    for x in w: words += [x & 0xffffffff, x >> 32]
    return words


def _conv_check():
    rng = random.Random(3)
    for _ in range(3000):
        x = rng.choice([0, 1, P - 1, (1 << 256) - 1, rng.randrange(1 << 256), rng.randrange(P)])
        words = [(x >> (32 * i)) & 0xffffffff for i in range(8)]
        l = from_u32_gpu(words)
        assert value(l) == x and all(abs(v) <= HALF for v in l[:4])
        v = rng.randrange(-2 * P + 1, 2 * P)
        back = to_u32_plus_2p_gpu(normalize(balanced(v)))
        assert sum(wd << (32 * i) for i, wd in enumerate(back)) == v + 2 * P
    print("conversion mirrors ok")


# ---------------------------------------------------------------------------------------------------------------------
# Fq2 = Fq[u] / (u^2 + 1) and the G2 mixed addition (mirrors fq2_52.cuh).  Every Fq2 product is a pair of signed sums of
# limb products with ONE reduction each: c0 = a0 b0 - a1 b1, c1 = a0 b1 + a1 b0 (2 x 75 splits); the square is
# c0 = a0^2 - a1^2 (two triangles, 55 splits), c1 = (2 a0) a1 (50 splits); a*b - c*d is 2 x 125 splits.
def mont_sum_gpu(terms):
    """mirror of mont_sum<...>: terms = [(sign, a, b)] with b is None meaning the square of a (upper triangle, doubled cross terms).
    Returns (sum sign * a * b + q p) / 2^260 with one reduction."""
    T = []
    for k in range(10):
        nlo = nhi = 0
        for i in range(5):
            for j in range(5):
                for sign, a, b in terms:
                    in_prod = (b is not None) or j >= i
                    nlo += sign * (in_prod and i + j == k); nhi += sign * (in_prod and i + j + 1 == k)
                nlo += (i + j == k); nhi += (i + j + 1 == k)
        T.append((-(nlo * RAW_L + nhi * RAW_H)) & MASK64)
    for sign, a, b in terms:
        f = split_acc if sign > 0 else split_sub
        for i in range(5):
            if b is None:
                f(a[i], a[i], T, 2 * i, 2 * i + 1)
                a2 = fadd(a[i], a[i])
                for j in range(i + 1, 5): f(a2, a[j], T, i + j, i + j + 1)
            else:
                for j in range(5): f(a[i], b[j], T, i + j, i + j + 1)
    for i in range(5):
        l, _ = column_low(T[i] + RAW_L)
        q = split_low(l, PINV_B)
        for j in range(5): split_acc(q, PL[j], T, i + j, i + j + 1)
        assert s64(T[i]) % (1 << 52) == 0
        T[i + 1] = (T[i + 1] + (s64(T[i]) >> 52)) & MASK64
    r = []
    for k in range(5, 9):
        l, c_ = column_low(T[k]); T[k + 1] = (T[k + 1] + c_) & MASK64; r.append(l)
    t9 = s64(T[9]); assert abs(t9) < 1 << 51
    r.append(fadd(from_raw(t9 + RAW_B), -from_raw(RAW_B)))
    return r


def mul2(a, b):       return (mont_sum_gpu([(1, a[0], b[0]), (-1, a[1], b[1])]), mont_sum_gpu([(1, a[0], b[1]), (1, a[1], b[0])]))
def sqr2(a):          return (mont_sum_gpu([(1, a[0], None), (-1, a[1], None)]), mont_sum_gpu([(1, ladd(a[0], a[0]), a[1])]))
def mul_sub2(a, b, c, d):
    return (mont_sum_gpu([(1, a[0], b[0]), (-1, a[1], b[1]), (-1, c[0], d[0]), (1, c[1], d[1])]),
            mont_sum_gpu([(1, a[0], b[1]), (1, a[1], b[0]), (-1, c[0], d[1]), (-1, c[1], d[0])]))
def add2(a, b):       return (ladd(a[0], b[0]), ladd(a[1], b[1]))
def sub2(a, b):       return (lsub(a[0], b[0]), lsub(a[1], b[1]))
def norm2(a):         return (normalize(a[0]), normalize(a[1]))
def val2(a):          return (value(a[0]), value(a[1]))


def madd52_g2(acc, xt, yt, negate=False):
    """G2 counterpart of madd52; xt, yt = (c0, c1) residue pairs in the 2^256 domain"""
    if xt == (0, 0) and yt == (0, 0): return
    x2 = (balanced(xt[0]), balanced(xt[1])); y2 = (balanced(yt[0]), balanced(yt[1]))
    if negate: y2 = ([-v for v in y2[0]], [-v for v in y2[1]])
    if acc.inf:
        k = (list(K264), [0.0] * 5)
        acc.X, acc.Y = mul2(x2, k), mul2(y2, k)
        acc.ZZ, acc.ZZZ = k, k
        acc.inf = False
        return
    U2, S2 = mul2(x2, acc.ZZ), mul2(y2, acc.ZZZ)
    Pn, Rn = norm2(sub2(U2, acc.X)), norm2(sub2(S2, acc.Y))
    if abs(Pn[0][0]) in CAND and abs(Pn[1][0]) in CAND:
        acc.redo = True
        return
    PP = sqr2(Pn); PPP = mul2(Pn, PP); Q = mul2(acc.X, PP); RR = sqr2(Rn)
    X3 = norm2(sub2(sub2(RR, PPP), add2(Q, Q)))
    Y3 = mul_sub2(Rn, sub2(Q, X3), acc.Y, PPP)
    acc.ZZ, acc.ZZZ = mul2(acc.ZZ, PP), mul2(acc.ZZZ, PPP)
    acc.X, acc.Y = X3, Y3
    for v in (acc.X, acc.Y, acc.ZZ, acc.ZZZ):
        assert max(abs(value(v[0])), abs(value(v[1]))) < 3 * P


def _g2_check():
    rng = random.Random(11)
    R256 = pow(2, 256, P); R256i = pow(R256, -1, P); rinv = pow(R, -1, P)
    # Fq2 helpers on integers
    f2mul = lambda a, b: ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    f2sub = lambda a, b: ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
    def f2inv(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, P); return (a[0] * d % P, -a[1] * d % P)
    # field-level checks against integers
    for _ in range(300):
        a = (balanced(rng.randrange(-2 * P, 2 * P)), balanced(rng.randrange(-2 * P, 2 * P)))
        b = (balanced(rng.randrange(-2 * P, 2 * P)), balanced(rng.randrange(-2 * P, 2 * P)))
        c = (balanced(rng.randrange(-P, P)), balanced(rng.randrange(-P, P)))
        d = (balanced(rng.randrange(-P, P)), balanced(rng.randrange(-P, P)))
        m = val2(mul2(a, b)); e = f2mul(val2(a), val2(b))
        assert (m[0] - e[0] * rinv) % P == 0 and (m[1] - e[1] * rinv) % P == 0
        q = val2(sqr2(a)); e = f2mul(val2(a), val2(a))
        assert (q[0] - e[0] * rinv) % P == 0 and (q[1] - e[1] * rinv) % P == 0
        ms = val2(mul_sub2(a, b, c, d)); e = f2sub(f2mul(val2(a), val2(b)), f2mul(val2(c), val2(d)))
        assert (ms[0] - e[0] * rinv) % P == 0 and (ms[1] - e[1] * rinv) % P == 0
    # curve-level: the twist y^2 = x^3 + 3 / (9 + u); points by cofactor-free trial (any point of E'(Fq2) works for the formulas)
    bt = f2mul((3, 0), f2inv((9, 1)))
    def f2sqrt(a):
        # Fq2 square root by the complex method (p = 3 mod 4)
        if a == (0, 0): return (0, 0)
        n = (a[0] * a[0] + a[1] * a[1]) % P
        s = pow(n, (P + 1) // 4, P)
        if s * s % P != n: return None
        for sgn in (1, -1):
            t = (a[0] + sgn * s) * pow(2, -1, P) % P
            x0 = pow(t, (P + 1) // 4, P)
            if x0 * x0 % P == t and x0:
                x1 = a[1] * pow(2 * x0, -1, P) % P
                if f2mul((x0, x1), (x0, x1)) == a: return (x0, x1)
        return None
    def rand_point():
        while True:
            x = (rng.randrange(P), rng.randrange(P))
            rhs = tuple((u + v) % P for u, v in zip(f2mul(f2mul(x, x), x), bt))
            y = f2sqrt(rhs)
            if y is not None: return x, y
    def aff_add(p, q):
        if p is None: return q
        (x1, y1), (x2, y2) = p, q
        assert x1 != x2
        lam = f2mul(f2sub(y2, y1), f2inv(f2sub(x2, x1)))
        x3 = f2sub(f2sub(f2mul(lam, lam), x1), x2)
        return x3, f2sub(f2mul(lam, f2sub(x1, x3)), y1)
    for trial in range(6):
        acc, ref = Acc52(), None
        for k in range(10):
            x, y = rand_point(); neg = rng.random() < 0.5
            madd52_g2(acc, tuple(v * R256 % P for v in x), tuple(v * R256 % P for v in y), neg)
            ref = aff_add(ref, (x, tuple((P - v) % P for v in y) if neg else y))
        assert not acc.redo
        k256 = (list(K256), [0.0] * 5); k252 = (list(K252), [0.0] * 5)
        X = tuple(v * R256i % P for v in val2(mul2(acc.X, k256))); Y = tuple(v * R256i % P for v in val2(mul2(acc.Y, k256)))
        ZZ = tuple(v * R256i % P for v in val2(mul2(acc.ZZ, k252))); ZZZ = tuple(v * R256i % P for v in val2(mul2(acc.ZZZ, k252)))
        assert (f2mul(X, f2inv(ZZ)), f2mul(Y, f2inv(ZZZ))) == ref
    print("fq2 / g2 model ok")


def madd_bounds(rounds=30):
    """fixed point of the XYZZ mixed-addition magnitudes (in units of p) under almost-Montgomery products"""
    f = lambda a, b: 0.0118 * a * b + 0.5
    bx = by = zz = zzz = 1.0
    for _ in range(rounds):
        u2, s2 = f(1, zz), f(1, zzz)
        p_, r_ = u2 + bx, s2 + by
        pp = f(p_, p_); ppp = f(p_, pp); q = f(bx, pp)
        x3 = f(r_, r_) + ppp + 2 * q
        y3 = f(r_, q + x3) + f(by, ppp)
        bx, by, zz, zzz = x3, y3, f(zz, pp), f(zzz, ppp)
    return dict(X=bx, Y=by, P=p_, R=r_, ZZ=zz, ZZZ=zzz)


def _check():
    rng = random.Random(52)
    rinv = pow(R, -1, P)
    ext = [HALF, -HALF, HALF - 1, -HALF + 1, 0, 1, -1]
    cnt = {}
    def rnd_elem(bound):
        if rng.random() < 0.3:
            l = [float(rng.choice(ext)) for _ in range(L - 1)] + [float(rng.randrange(-(1 << 48), 1 << 48))]
            if abs(value(l)) < bound: return l
        return balanced(rng.randrange(-bound + 1, bound))
    for it in range(4000):
        a, b = rnd_elem(8 * P), rnd_elem(8 * P)
        r = mont_mul(a, b, cnt)
        assert (value(r) - value(a) * value(b) * rinv) % P == 0
        assert abs(value(r)) < 3 * P // 2 + 1, "almost-Montgomery bound"
        assert all(abs(x) <= HALF for x in r[:-1])
        r2 = mont_sqr(a)
        assert (value(r2) - value(a) ** 2 * rinv) % P == 0 and abs(value(r2)) < 3 * P // 2 + 1
        g = mont_mul_gpu(a, b)
        assert (value(g) - value(a) * value(b) * rinv) % P == 0 and abs(value(g)) < 13 * P // 10
        assert all(abs(x) <= HALF for x in g[:-1])
        g2 = mont_mul_gpu(a, a, square=True)
        assert (value(g2) - value(a) ** 2 * rinv) % P == 0
        # one un-normalised operand (difference of two normalised values) against a normalised one
        c = rnd_elem(4 * P); d = lsub(a if abs(value(a)) < 4 * P else rnd_elem(4 * P), c)
        if abs(value(d)) < 8 * P:
            r3 = mont_mul(d, b)
            assert (value(r3) - value(d) * value(b) * rinv) % P == 0
            g3 = mont_mul_gpu(d, b)
            assert (value(g3) - value(d) * value(b) * rinv) % P == 0
            assert value(normalize(d)) == value(d)
    b = madd_bounds()
    assert max(b.values()) < 3.0
    print("fp52 model ok; largest |column| = 2^%d; madd magnitudes / p:" % cnt['max_col'].bit_length(), {k: round(v, 2) for k, v in b.items()})


if __name__ == '__main__':
    _check()
    _conv_check()
    _ec_check()
    IMPL = _Impl(True)
    _ec_check()
    _g2_check()
