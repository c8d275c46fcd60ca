#!/usr/bin/env python
"""CPU model of the index logic of ntt_pass8_kernel (circom_compat_b200/csrc/ntt.cu): which thread holds which tile element in
which register position, the position rotations between steps, the field changes through the swizzled shared-memory tile,
the twiddle exponents, and the pass schedule of ntt_domain_create.  Python integers instead of Montgomery limbs, so it
checks everything about the kernel except the field arithmetic itself (which the radix-2 kernel and the MSM share).
Checked against the DFT definition (forward), its inverse, and the fused H -> g*H chain the witness map runs
(/root/reference/src/circom/qap.rs:60-73).  Run by tests/test_host.py; `python tools/ntt8_model.py` prints the schedules."""
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
def root(logm):  # primitive 2^logm-th root
    g = pow(5, (R - 1) >> 28, R)
    for _ in range(28 - logm): g = g * g % R
    return g
def brev(x, bits):
    r = 0
    for i in range(bits): r |= ((x >> i) & 1) << (bits - 1 - i)
    return r
def tile_global_index(loc, tile_id, cols_log, sb, k):
    t = loc >> cols_log; col = loc & ((1 << cols_log) - 1)
    o = (tile_id << cols_log) | col
    lo = o & ((1 << sb) - 1); high = o >> sb
    return (high << (sb + k)) | (t << sb) | lo
def sw(loc): return loc ^ (loc >> 3)
def rotl3(p): return ((p << 1) | (p >> 2)) & 7
def rotr3(p): return ((p >> 1) | (p << 2)) & 7
def elem_of(p, k):
    for _ in range(k): p = rotl3(p)
    return p
def field_loc(tid, e, f): return ((tid >> f) << (f + 3)) | (e << f) | (tid & ((1 << f) - 1))
def rot_fwd(x): return [x[rotl3(p)] for p in range(8)]
def rot_bwd(x): return [x[rotr3(p)] for p in range(8)]

class Pass:
    def __init__(s, logn, tl, sb, k, tw, ct):
        s.logn, s.tl, s.sb, s.k, s.tw, s.ct = logn, tl, sb, k, tw, ct
        s.cols_log = tl - k; s.n = 1 << logn; s.ftop = tl - 3
    def run(s, vec, do_dif, do_scale, do_dit):
        tile = 1 << s.tl; nthreads = tile >> 3
        for blk in range(s.n >> s.tl):
            s.run_cta(vec, blk, nthreads, do_dif, do_scale, do_dit)
    def gidx(s, loc, blk): return tile_global_index(loc, blk, s.cols_log, s.sb, s.k)
    def rotate_to(s, X, K, want):
        for tid in range(len(X)):
            k = K[tid]
            d = (want - k) % 3
            if d == 1: X[tid] = rot_fwd(X[tid]); k += 1
            elif d == 2: X[tid] = rot_bwd(X[tid]); k += 2
            K[tid] = k % 3
    def exchange(s, X, f, nf):
        sm = {}
        for tid in range(len(X)):
            for p in range(8):
                a = sw(field_loc(tid, p, f)); assert a not in sm and a < (1 << s.tl); sm[a] = X[tid][p]
        for tid in range(len(X)):
            X[tid] = [sm[sw(field_loc(tid, p, nf))] for p in range(8)]
    def run_cta(s, vec, blk, nthreads, do_dif, do_scale, do_dit):
        cols_log, ftop, n, logn = s.cols_log, s.ftop, s.n, s.logn
        f = ftop if do_dif else min(cols_log, ftop)
        X = [[vec[s.gidx(field_loc(tid, p, f), blk)] for p in range(8)] for tid in range(nthreads)]
        K = [0] * nthreads
        if do_dif:
            qb = s.tl
            while qb > cols_log:
                qa = max(qb - 3, cols_log); nf = min(qa, ftop)
                if nf != f: s.exchange(X, f, nf); f = nf
                for q in range(qb - 1, qa - 1, -1):
                    eb = q - f; s.rotate_to(X, K, (eb + 1) % 3)
                    st = s.sb + (q - cols_log)
                    for tid in range(nthreads):
                        x = X[tid]; k = K[tid]
                        for p in range(4):
                            e = elem_of(p, k); assert (e >> eb) & 1 == 0 and elem_of(p + 4, k) == e | (1 << eb)
                            loc0 = field_loc(tid, e, f); gi = s.gidx(loc0, blk)
                            j = gi & ((1 << st) - 1); e2 = j << (logn - st)
                            u, v = x[p], x[p + 4]
                            x[p] = (u + v) % R
                            x[p + 4] = (u - v) % R if e2 == 0 else (v - u) * s.tw[n - e2] % R
                s.rotate_to(X, K, 0)
                qb = qa
        if do_scale:
            for tid in range(nthreads):
                for p in range(8):
                    g = s.gidx(field_loc(tid, p, f), blk)
                    X[tid][p] = X[tid][p] * s.ct[brev(g, logn)] % R
        if do_dit:
            qa = cols_log
            while qa < s.tl:
                qb = min(qa + 3, s.tl); nf = min(qa, ftop)
                if nf != f: s.exchange(X, f, nf); f = nf
                for q in range(qa, qb):
                    eb = q - f; s.rotate_to(X, K, (eb + 1) % 3)
                    st = s.sb + (q - cols_log)
                    for tid in range(nthreads):
                        x = X[tid]; k = K[tid]
                        for p in range(4):
                            e = elem_of(p, k); assert (e >> eb) & 1 == 0 and elem_of(p + 4, k) == e | (1 << eb)
                            loc0 = field_loc(tid, e, f); gi = s.gidx(loc0, blk)
                            j = gi & ((1 << st) - 1); e2 = j << (logn - st)
                            u, v = x[p], x[p + 4]
                            if e2: v = v * s.tw[e2] % R
                            x[p] = (u + v) % R; x[p + 4] = (u - v) % R
                s.rotate_to(X, K, 0)
                qa = qb
        for tid in range(nthreads):
            for p in range(8): vec[s.gidx(field_loc(tid, p, f), blk)] = X[tid][p]

def simple_ntt(x, w):
    # textbook recursive decimation-in-time transform, the independent reference above 256 points
    n = len(x)
    if n == 1: return x[:]
    e, o = simple_ntt(x[0::2], w * w % R), simple_ntt(x[1::2], w * w % R)
    out = [0] * n; t = 1
    for i in range(n // 2):
        u = o[i] * t % R
        out[i] = (e[i] + u) % R; out[i + n // 2] = (e[i] - u) % R
        t = t * w % R
    return out


def schedule(logn, tlmax=4, maxk=None):
    # mirrors ntt_domain_create (new): returns list of (sb, k, tl)
    k0 = min(logn, tlmax)
    passes = [(0, k0, k0)]
    rem = logn - k0
    if rem > 0:
        mk = maxk or tlmax
        np_ = (rem + mk - 1) // mk
        sb = k0
        for p in range(np_):
            k = rem // (np_ - p)
            tl = min(logn, max(k, tlmax))
            passes.append((sb, k, tl)); sb += k; rem -= k
    return passes

def check(logn, tlmax, maxk=None):
    import random
    n = 1 << logn; w2n = root(logn + 1); wn = w2n * w2n % R
    tw = [pow(w2n, i, R) for i in range(n)] + [None]
    ninv = pow(n, R - 2, R); ct = [ninv * t % R for t in tw[:n]]
    rnd = random.Random(logn * 100 + tlmax)
    x = [rnd.randrange(R) for _ in range(n)]
    P = [Pass(logn, tl, sb, k, tw, ct) for (sb, k, tl) in schedule(logn, tlmax, maxk)]
    # forward plain: bitrev then DIT passes
    v = [x[brev(i, logn)] for i in range(n)]
    for p in P: p.run(v, 0, 0, 1)
    ref = [sum(x[j] * pow(wn, i * j, R) for j in range(n)) % R for i in range(n)] if n <= 256 else simple_ntt(x, wn)
    assert v == ref, "forward"
    fwd = v[:]
    # inverse plain: DIF passes then bitrev + ninv
    v = fwd[:]
    for p in reversed(P): p.run(v, 1, 0, 0)
    back = [v[brev(i, logn)] * ninv % R for i in range(n)]
    assert back == x, "inverse"
    # fused: evaluations on H -> evaluations on the coset g*H
    v = x[:]
    for p in reversed(P[1:]): p.run(v, 1, 0, 0)
    P[0].run(v, 1, 1, 1)
    for p in P[1:]: p.run(v, 0, 0, 1)
    # reference: coef = iDFT(x); out_i = sum coef_j (g w^i)^j : compute via plain model pieces
    coef = back  # == x?? no: coefficients of the polynomial with evaluations x
    c = x[:]
    for p in reversed(P): p.run(c, 1, 0, 0)
    coef = [c[brev(i, logn)] * ninv % R for i in range(n)]
    sc = [coef[j] * tw[j] % R for j in range(n)]
    s2 = [sc[brev(i, logn)] for i in range(n)]
    for p in P: p.run(s2, 0, 0, 1)
    assert v == s2, "fused"
    print("ok", logn, tlmax, maxk, schedule(logn, tlmax, maxk))

if __name__ == "__main__":
    for logn, tlmax, maxk in [(3,3,None),(4,4,None),(5,5,None),(6,6,None),(7,7,None),(6,4,None),(7,4,None),(8,4,None),(8,5,None),(9,5,None),(10,5,None),
                              (9,4,None),(10,4,None),(8,5,2),(9,5,2),(10,6,3),(9,6,None),(11,6,None),(12,6,None),(11,4,2)]:
        check(logn, tlmax, maxk)
