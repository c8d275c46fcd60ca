/* cref.c - CPU restatement (plain C + OpenMP) of the ark-circom / ark-groth16 0.5 Groth16 prover hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the oracle and the CPU baseline.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` leg may load it; the product (circom_compat_b200/) never
 * links or calls it.
 *
 * The reference (/root/reference, Rust) cannot be compiled in this image (no cargo/rustc, arkworks crates not
 * vendored: Cargo.toml:24-32 pins ark-* ^0.5.0, Cargo.lock is git-ignored), so oracle/_ref does not exist and
 * this file is a "port": it follows
 *   - src/circom/qap.rs:23-88           CircomReduction::witness_map_from_matrices (same step order)
 *   - src/zkey.rs:320-368               Montgomery little-endian encodings of Fr/Fq/G1/G2
 *   - ark-groth16 0.5.0 prover.rs       create_proof_with_assignment (restated)
 *   - ark-ec 0.5.0 VariableBaseMSM      signed-digit Pippenger, c = floor(ceil(log2 n)*69/100)+2, threads over
 *                                       windows (plus point-range splits so that every host core is used)
 *   - ark-poly 0.5.0 Radix2EvaluationDomain  natural-order radix-2 (i)FFT, omega = 5^((r-1)/n), ifft * 1/n
 *   - ark-ff 0.5.0 (asm)                4x64 Montgomery arithmetic
 * It is pinned against oracle/pyref.py (independent big-int arithmetic + pairing check) and the golden vectors in
 * tests/golden/ by tests/test_oracle.py.
 *
 * Data conventions (identical to the product's C ABI, include/b2groth.h): field elements are 4 x u64 little-endian
 * limbs in Montgomery form unless a name says "canon"; G1 = x||y (64 B), G2 = x.c0||x.c1||y.c0||y.c1 (128 B),
 * infinity = all zero bytes.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

/* ------------------------------------------------------------------------------------------------ generic Fp */
static inline int fe_iszero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int fe_geq(const fe* a, const uint64_t p[4]) {
    for (int i = 3; i >= 0; i--) { if (a->l[i] > p[i]) return 1; if (a->l[i] < p[i]) return 0; }
    return 1;
}
static inline void fe_sub_p(fe* a, const uint64_t p[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a->l[i] - p[i] - (uint64_t)br; a->l[i] = (uint64_t)d; br = (d >> 64) & 1; }
}
static inline void mp_add(fe* r, const fe* a, const fe* b, const uint64_t p[4]) {
    u128 c = 0; fe t;
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; t.l[i] = (uint64_t)c; c >>= 64; }
    if (c || fe_geq(&t, p)) fe_sub_p(&t, p);
    *r = t;
}
static inline void mp_sub(fe* r, const fe* a, const fe* b, const uint64_t p[4]) {
    u128 br = 0; fe t;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)br; t.l[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)t.l[i] + p[i]; t.l[i] = (uint64_t)c; c >>= 64; } }
    *r = t;
}
static inline void mp_neg(fe* r, const fe* a, const uint64_t p[4]) {
    if (fe_iszero(a)) { *r = *a; return; }
    u128 br = 0; fe t;
    for (int i = 0; i < 4; i++) { u128 d = (u128)p[i] - a->l[i] - (uint64_t)br; t.l[i] = (uint64_t)d; br = (d >> 64) & 1; }
    *r = t;
}
/* CIOS Montgomery product, R = 2^256 */
static inline void mp_mul(fe* r, const fe* a, const fe* b, const uint64_t p[4], uint64_t inv) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->l[j] * b->l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * inv;
        c = (u128)m * p[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; c >>= 64;
        t[4] = t[5] + (uint64_t)c;
    }
    fe o = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fe_geq(&o, p)) fe_sub_p(&o, p);
    *r = o;
}

/* ------------------------------------------------------------------------------------------------ Fq and Fr */
static const uint64_t FQ_P[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FQ_INV = 0x87d20782e4866389ULL;
static const fe FQ_R1 = {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}};
static const fe FQ_R2 = {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};
static const uint64_t FR_P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FR_INV = 0xc2e1f593efffffffULL;
static const fe FR_R1 = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const fe FR_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};

static inline void fq_add(fe* r, const fe* a, const fe* b) { mp_add(r, a, b, FQ_P); }
static inline void fq_sub(fe* r, const fe* a, const fe* b) { mp_sub(r, a, b, FQ_P); }
static inline void fq_neg(fe* r, const fe* a) { mp_neg(r, a, FQ_P); }
static inline void fq_mul(fe* r, const fe* a, const fe* b) { mp_mul(r, a, b, FQ_P, FQ_INV); }
static inline void fq_sqr(fe* r, const fe* a) { mp_mul(r, a, a, FQ_P, FQ_INV); }
static inline void fq_dbl(fe* r, const fe* a) { mp_add(r, a, a, FQ_P); }
static inline void fq_setone(fe* r) { *r = FQ_R1; }
static inline void fq_setzero(fe* r) { memset(r, 0, sizeof *r); }

static inline void fr_add(fe* r, const fe* a, const fe* b) { mp_add(r, a, b, FR_P); }
static inline void fr_sub(fe* r, const fe* a, const fe* b) { mp_sub(r, a, b, FR_P); }
static inline void fr_mul(fe* r, const fe* a, const fe* b) { mp_mul(r, a, b, FR_P, FR_INV); }
static inline void fr_from_canon(fe* r, const fe* a) { mp_mul(r, a, &FR_R2, FR_P, FR_INV); }
static inline void fr_to_canon(fe* r, const fe* a) { fe one = {{1, 0, 0, 0}}; mp_mul(r, a, &one, FR_P, FR_INV); }
static inline void fq_to_canon(fe* r, const fe* a) { fe one = {{1, 0, 0, 0}}; mp_mul(r, a, &one, FQ_P, FQ_INV); }

static void mp_pow(fe* r, const fe* a, const uint64_t e[4], const uint64_t p[4], uint64_t inv, const fe* one) {
    fe acc = *one;
    for (int i = 255; i >= 0; i--) {
        mp_mul(&acc, &acc, &acc, p, inv);
        if ((e[i >> 6] >> (i & 63)) & 1) mp_mul(&acc, &acc, a, p, inv);
    }
    *r = acc;
}
static void fq_inv(fe* r, const fe* a) {
    uint64_t e[4] = {FQ_P[0] - 2, FQ_P[1], FQ_P[2], FQ_P[3]};
    mp_pow(r, a, e, FQ_P, FQ_INV, &FQ_R1);
}
static void fr_inv(fe* r, const fe* a) {
    uint64_t e[4] = {FR_P[0] - 2, FR_P[1], FR_P[2], FR_P[3]};
    mp_pow(r, a, e, FR_P, FR_INV, &FR_R1);
}

/* ------------------------------------------------------------------------------------------------ Fq2 = Fq[u]/(u^2+1) */
typedef struct { fe c0, c1; } fe2;
static inline void fq2_add(fe2* r, const fe2* a, const fe2* b) { fq_add(&r->c0, &a->c0, &b->c0); fq_add(&r->c1, &a->c1, &b->c1); }
static inline void fq2_sub(fe2* r, const fe2* a, const fe2* b) { fq_sub(&r->c0, &a->c0, &b->c0); fq_sub(&r->c1, &a->c1, &b->c1); }
static inline void fq2_dbl(fe2* r, const fe2* a) { fq_dbl(&r->c0, &a->c0); fq_dbl(&r->c1, &a->c1); }
static inline void fq2_neg(fe2* r, const fe2* a) { fq_neg(&r->c0, &a->c0); fq_neg(&r->c1, &a->c1); }
static inline void fq2_mul(fe2* r, const fe2* a, const fe2* b) {
    fe v0, v1, s, t, o0, o1;
    fq_mul(&v0, &a->c0, &b->c0); fq_mul(&v1, &a->c1, &b->c1);
    fq_add(&s, &a->c0, &a->c1); fq_add(&t, &b->c0, &b->c1);
    fq_mul(&o1, &s, &t); fq_sub(&o1, &o1, &v0); fq_sub(&o1, &o1, &v1);
    fq_sub(&o0, &v0, &v1);
    r->c0 = o0; r->c1 = o1;
}
static inline void fq2_sqr(fe2* r, const fe2* a) {
    fe s, d, m, o0;
    fq_add(&s, &a->c0, &a->c1); fq_sub(&d, &a->c0, &a->c1);
    fq_mul(&m, &a->c0, &a->c1);
    fq_mul(&o0, &s, &d);
    r->c0 = o0; fq_dbl(&r->c1, &m);
}
static void fq2_inv(fe2* r, const fe2* a) {
    fe n0, n1, d;
    fq_sqr(&n0, &a->c0); fq_sqr(&n1, &a->c1); fq_add(&d, &n0, &n1); fq_inv(&d, &d);
    fq_mul(&r->c0, &a->c0, &d); fq_mul(&n0, &a->c1, &d); fq_neg(&r->c1, &n0);
}
static inline int fq2_iszero(const fe2* a) { return fe_iszero(&a->c0) && fe_iszero(&a->c1); }
static inline int fq2_eq(const fe2* a, const fe2* b) { return fe_eq(&a->c0, &b->c0) && fe_eq(&a->c1, &b->c1); }
static inline void fq2_setone(fe2* r) { r->c0 = FQ_R1; fq_setzero(&r->c1); }
static inline void fq2_setzero(fe2* r) { memset(r, 0, sizeof *r); }

/* ------------------------------------------------------------------------------------------------ G1, G2 */
#define FE fe
#define FE_add fq_add
#define FE_sub fq_sub
#define FE_mul fq_mul
#define FE_sqr fq_sqr
#define FE_neg fq_neg
#define FE_dbl fq_dbl
#define FE_iszero fe_iszero
#define FE_eq fe_eq
#define FE_inv fq_inv
#define FE_setone fq_setone
#define FE_setzero fq_setzero
#define CN(x) g1_##x
#include "cref_curve.inc"
#undef FE
#undef FE_add
#undef FE_sub
#undef FE_mul
#undef FE_sqr
#undef FE_neg
#undef FE_dbl
#undef FE_iszero
#undef FE_eq
#undef FE_inv
#undef FE_setone
#undef FE_setzero
#undef CN

#define FE fe2
#define FE_add fq2_add
#define FE_sub fq2_sub
#define FE_mul fq2_mul
#define FE_sqr fq2_sqr
#define FE_neg fq2_neg
#define FE_dbl fq2_dbl
#define FE_iszero fq2_iszero
#define FE_eq fq2_eq
#define FE_inv fq2_inv
#define FE_setone fq2_setone
#define FE_setzero fq2_setzero
#define CN(x) g2_##x
#include "cref_curve.inc"

static int nthreads_or_default(int nthreads) { return nthreads > 0 ? nthreads : omp_get_max_threads(); }

/* ------------------------------------------------------------------------------------------------ radix-2 domain */
static const fe FR_ROOT_2_28_CANON = {{0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL, 0x00e0a7eb8ef62abcULL, 0x2a3c09f0a58a7e85ULL}};

static void fr_root_of_unity(fe* w, int log_n) {   /* Montgomery form */
    fe g; fr_from_canon(&g, &FR_ROOT_2_28_CANON);
    for (int i = 28; i > log_n; i--) fr_mul(&g, &g, &g);
    *w = g;
}

static inline size_t bitrev(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* natural-order in/out; inverse scales by 1/n (ark-poly Radix2EvaluationDomain::{fft,ifft}_in_place) */
static void fr_fft(fe* a, int log_n, int inverse, int nthreads) {
    size_t n = (size_t)1 << log_n;
    if (log_n == 0) return;
    fe w; fr_root_of_unity(&w, log_n);
    if (inverse) fr_inv(&w, &w);
    fe* tw = (fe*)malloc((n / 2) * sizeof(fe));
    /* twiddles: blocks of 1024 computed independently */
    size_t half = n / 2;
    const size_t BL = 1024;
    size_t nbl = (half + BL - 1) / BL;
    fe wbl; { uint64_t e[4] = {BL, 0, 0, 0}; mp_pow(&wbl, &w, e, FR_P, FR_INV, &FR_R1); }
    fe* starts = (fe*)malloc(nbl * sizeof(fe));
    starts[0] = FR_R1;
    for (size_t b = 1; b < nbl; b++) fr_mul(&starts[b], &starts[b - 1], &wbl);
    #pragma omp parallel for num_threads(nthreads)
    for (size_t b = 0; b < nbl; b++) {
        fe cur = starts[b];
        size_t hi = (b + 1) * BL > half ? half : (b + 1) * BL;
        for (size_t i = b * BL; i < hi; i++) { tw[i] = cur; fr_mul(&cur, &cur, &w); }
    }
    free(starts);
    #pragma omp parallel for num_threads(nthreads)
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, log_n);
        if (i < j) { fe t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int s = 1; s <= log_n; s++) {
        size_t len = (size_t)1 << s, hl = len >> 1, stride = n >> s;
        #pragma omp parallel for num_threads(nthreads)
        for (size_t idx = 0; idx < half; idx++) {
            size_t blk = idx / hl, k = idx % hl;
            fe* u = &a[blk * len + k];
            fe* v = u + hl;
            fe t; fr_mul(&t, v, &tw[k * stride]);
            fe x = *u;
            fr_add(u, &x, &t); fr_sub(v, &x, &t);
        }
    }
    if (inverse) {
        fe nc = {{n, 0, 0, 0}}, ninv; fr_from_canon(&ninv, &nc); fr_inv(&ninv, &ninv);
        #pragma omp parallel for num_threads(nthreads)
        for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &ninv);
    }
    free(tw);
}

/* a[i] *= g^i (distribute_powers_and_mul_by_const with c = 1) */
static void fr_distribute_powers(fe* a, size_t n, const fe* g, int nthreads) {
    const size_t BL = 4096;
    size_t nbl = (n + BL - 1) / BL;
    fe gbl; { uint64_t e[4] = {BL, 0, 0, 0}; mp_pow(&gbl, g, e, FR_P, FR_INV, &FR_R1); }
    fe* starts = (fe*)malloc(nbl * sizeof(fe));
    starts[0] = FR_R1;
    for (size_t b = 1; b < nbl; b++) fr_mul(&starts[b], &starts[b - 1], &gbl);
    #pragma omp parallel for num_threads(nthreads)
    for (size_t b = 0; b < nbl; b++) {
        fe cur = starts[b];
        size_t hi = (b + 1) * BL > n ? n : (b + 1) * BL;
        for (size_t i = b * BL; i < hi; i++) { fr_mul(&a[i], &a[i], &cur); fr_mul(&cur, &cur, g); }
    }
    free(starts);
}

/* ------------------------------------------------------------------------------------------------ exported API */
#define EXPORT __attribute__((visibility("default")))

EXPORT int cref_version(void) { return 1; }
EXPORT int cref_max_threads(void) { return omp_get_max_threads(); }

EXPORT void cref_fr_from_canon(uint64_t* out, const uint64_t* in, size_t n) {
    for (size_t i = 0; i < n; i++) fr_from_canon((fe*)(out + 4 * i), (const fe*)(in + 4 * i));
}
EXPORT void cref_fr_to_canon(uint64_t* out, const uint64_t* in, size_t n) {
    for (size_t i = 0; i < n; i++) fr_to_canon((fe*)(out + 4 * i), (const fe*)(in + 4 * i));
}
EXPORT void cref_fq_from_canon(uint64_t* out, const uint64_t* in, size_t n) {
    for (size_t i = 0; i < n; i++) mp_mul((fe*)(out + 4 * i), (const fe*)(in + 4 * i), &FQ_R2, FQ_P, FQ_INV);
}
EXPORT void cref_fq_to_canon(uint64_t* out, const uint64_t* in, size_t n) {
    for (size_t i = 0; i < n; i++) fq_to_canon((fe*)(out + 4 * i), (const fe*)(in + 4 * i));
}
EXPORT void cref_fr_mul(uint64_t* out, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) fr_mul((fe*)(out + 4 * i), (const fe*)(a + 4 * i), (const fe*)(b + 4 * i));
}
EXPORT void cref_fq_mul(uint64_t* out, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) fq_mul((fe*)(out + 4 * i), (const fe*)(a + 4 * i), (const fe*)(b + 4 * i));
}

/* plain (i)NTT, data Montgomery, in place */
EXPORT int cref_ntt(uint64_t* data, int log_n, int inverse, int nthreads) {
    if (log_n < 0 || log_n > 28) return -1;
    fr_fft((fe*)data, log_n, inverse, nthreads_or_default(nthreads));
    return 0;
}

static void eval_rows(fe* out, uint32_t m, const uint32_t* rowptr, const uint32_t* col, const fe* val,
                      const fe* w, int nthreads) {
    /* evaluate_constraint (ark-groth16 0.5.0 r1cs_to_qap.rs, called at qap.rs:42-43) */
    #pragma omp parallel for num_threads(nthreads)
    for (uint32_t i = 0; i < m; i++) {
        fe acc; memset(&acc, 0, sizeof acc);
        for (uint32_t k = rowptr[i]; k < rowptr[i + 1]; k++) {
            fe t; fr_mul(&t, &val[k], &w[col[k]]); fr_add(&acc, &acc, &t);
        }
        out[i] = acc;
    }
}

/* CircomReduction::witness_map_from_matrices, src/circom/qap.rs:23-88.  CSR matrices (values Montgomery),
 * w Montgomery, h_out = domain_size elements Montgomery.  Returns domain log2, <0 on error. */
EXPORT int cref_witness_map(uint32_t m, uint32_t num_inputs, uint32_t n_vars,
                            const uint32_t* a_rowptr, const uint32_t* a_col, const uint64_t* a_val,
                            const uint32_t* b_rowptr, const uint32_t* b_col, const uint64_t* b_val,
                            const uint64_t* w_mont, uint64_t* h_out, int nthreads) {
    (void)n_vars;
    nthreads = nthreads_or_default(nthreads);
    size_t need = (size_t)m + num_inputs, n = 1; int log_n = 0;
    while (n < need) { n <<= 1; log_n++; }
    if (log_n > 28) return -1;                                            /* qap.rs:31 PolynomialDegreeTooLarge */
    const fe* w = (const fe*)w_mont;
    fe* a = (fe*)calloc(n, sizeof(fe));
    fe* b = (fe*)calloc(n, sizeof(fe));
    fe* c = (fe*)calloc(n, sizeof(fe));
    eval_rows(a, m, a_rowptr, a_col, (const fe*)a_val, w, nthreads);      /* qap.rs:37-44 */
    eval_rows(b, m, b_rowptr, b_col, (const fe*)b_val, w, nthreads);
    for (uint32_t j = 0; j < num_inputs; j++) a[m + j] = w[j];            /* qap.rs:46-50 */
    #pragma omp parallel for num_threads(nthreads)
    for (uint32_t i = 0; i < m; i++) fr_mul(&c[i], &a[i], &b[i]);         /* qap.rs:52-58 */
    fr_fft(a, log_n, 1, nthreads); fr_fft(b, log_n, 1, nthreads);         /* qap.rs:60-61 */
    fe g; fr_root_of_unity(&g, log_n + 1);                                /* qap.rs:63-68 */
    fr_distribute_powers(a, n, &g, nthreads);                             /* qap.rs:69-70 */
    fr_distribute_powers(b, n, &g, nthreads);
    fr_fft(a, log_n, 0, nthreads); fr_fft(b, log_n, 0, nthreads);         /* qap.rs:72-73 */
    #pragma omp parallel for num_threads(nthreads)
    for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &b[i]);           /* qap.rs:75 */
    fr_fft(c, log_n, 1, nthreads);                                        /* qap.rs:79-81 */
    fr_distribute_powers(c, n, &g, nthreads);
    fr_fft(c, log_n, 0, nthreads);
    fe* h = (fe*)h_out;
    #pragma omp parallel for num_threads(nthreads)
    for (size_t i = 0; i < n; i++) fr_sub(&h[i], &a[i], &c[i]);           /* qap.rs:83-85 */
    free(a); free(b); free(c);
    return log_n;
}

/* LibsnarkReduction::witness_map_from_matrices (ark-groth16 0.5.0 r1cs_to_qap.rs; default QAP of Groth16<Bn254>,
 * /root/reference/tests/groth16.rs:9,25-35): coset offset g = Fr::GENERATOR = 5, c from the real C matrix,
 * h = coset_ifft((a*b - c) / Z(g)).  h_out = n coefficients (Montgomery).  Returns log2(n). */
EXPORT int cref_witness_map_libsnark(uint32_t m, uint32_t num_inputs,
                                     const uint32_t* a_rowptr, const uint32_t* a_col, const uint64_t* a_val,
                                     const uint32_t* b_rowptr, const uint32_t* b_col, const uint64_t* b_val,
                                     const uint32_t* c_rowptr, const uint32_t* c_col, const uint64_t* c_val,
                                     const uint64_t* w_mont, uint64_t* h_out, int nthreads) {
    nthreads = nthreads_or_default(nthreads);
    size_t need = (size_t)m + num_inputs, n = 1; int log_n = 0;
    while (n < need) { n <<= 1; log_n++; }
    if (log_n > 28) return -1;
    const fe* w = (const fe*)w_mont;
    fe* a = (fe*)calloc(n, sizeof(fe)); fe* b = (fe*)calloc(n, sizeof(fe)); fe* c = (fe*)calloc(n, sizeof(fe));
    eval_rows(a, m, a_rowptr, a_col, (const fe*)a_val, w, nthreads);
    eval_rows(b, m, b_rowptr, b_col, (const fe*)b_val, w, nthreads);
    eval_rows(c, m, c_rowptr, c_col, (const fe*)c_val, w, nthreads);
    for (uint32_t j = 0; j < num_inputs; j++) a[m + j] = w[j];
    fe five = {{5, 0, 0, 0}}, g, ginv; fr_from_canon(&g, &five); fr_inv(&ginv, &g);
    fe* vs[3] = {a, b, c};
    for (int k = 0; k < 3; k++) { fr_fft(vs[k], log_n, 1, nthreads); fr_distribute_powers(vs[k], n, &g, nthreads); fr_fft(vs[k], log_n, 0, nthreads); }
    fe gn = g; for (int i = 0; i < log_n; i++) fr_mul(&gn, &gn, &gn);
    fe z, zinv; fr_sub(&z, &gn, &FR_R1); fr_inv(&zinv, &z);
    fe* h = (fe*)h_out;
    #pragma omp parallel for num_threads(nthreads)
    for (size_t i = 0; i < n; i++) { fe t; fr_mul(&t, &a[i], &b[i]); fr_sub(&t, &t, &c[i]); fr_mul(&h[i], &t, &zinv); }
    fr_fft(h, log_n, 1, nthreads);
    fr_distribute_powers(h, n, &ginv, nthreads);
    free(a); free(b); free(c);
    return log_n;
}

/* MSM over G1: bases 64 B each (Montgomery), scalars canonical 4xu64.  out = affine Montgomery (zeros = infinity). */
EXPORT int cref_msm_g1(const uint64_t* bases, const uint64_t* scalars_canon, size_t n, uint64_t* out_xy, int nthreads) {
    g1_jac acc; g1_msm(&acc, (const g1_aff*)bases, scalars_canon, n, nthreads_or_default(nthreads));
    g1_jac_to_aff((g1_aff*)out_xy, &acc);
    return g1_jac_is_inf(&acc);
}
EXPORT int cref_msm_g2(const uint64_t* bases, const uint64_t* scalars_canon, size_t n, uint64_t* out_xy, int nthreads) {
    g2_jac acc; g2_msm(&acc, (const g2_aff*)bases, scalars_canon, n, nthreads_or_default(nthreads));
    g2_jac_to_aff((g2_aff*)out_xy, &acc);
    return g2_jac_is_inf(&acc);
}

/* k_i * G for the standard generators (synthetic trapdoor keys) */
static void g1_generator(g1_aff* g) {
    fe one = {{1, 0, 0, 0}}, two = {{2, 0, 0, 0}};
    mp_mul(&g->x, &one, &FQ_R2, FQ_P, FQ_INV); mp_mul(&g->y, &two, &FQ_R2, FQ_P, FQ_INV);
}
static void g2_generator(g2_aff* g) {   /* src/zkey.rs:443-463, canonical limbs */
    static const fe X0 = {{0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL}};
    static const fe X1 = {{0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL}};
    static const fe Y0 = {{0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL}};
    static const fe Y1 = {{0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL}};
    mp_mul(&g->x.c0, &X0, &FQ_R2, FQ_P, FQ_INV); mp_mul(&g->x.c1, &X1, &FQ_R2, FQ_P, FQ_INV);
    mp_mul(&g->y.c0, &Y0, &FQ_R2, FQ_P, FQ_INV); mp_mul(&g->y.c1, &Y1, &FQ_R2, FQ_P, FQ_INV);
}
EXPORT void cref_fixed_base_g1(const uint64_t* scalars_canon, size_t n, uint64_t* out, int nthreads) {
    g1_aff g; g1_generator(&g);
    g1_fixed_base((g1_aff*)out, &g, scalars_canon, n, nthreads_or_default(nthreads));
}
EXPORT void cref_fixed_base_g2(const uint64_t* scalars_canon, size_t n, uint64_t* out, int nthreads) {
    g2_aff g; g2_generator(&g);
    g2_fixed_base((g2_aff*)out, &g, scalars_canon, n, nthreads_or_default(nthreads));
}
/* k * P for one arbitrary point (affine Montgomery in/out) */
EXPORT void cref_mul_g1(const uint64_t* p, const uint64_t* k_canon, uint64_t* out) {
    g1_jac j; g1_jac_from_aff(&j, (const g1_aff*)p); g1_jac_mul(&j, &j, k_canon); g1_jac_to_aff((g1_aff*)out, &j);
}
EXPORT void cref_mul_g2(const uint64_t* p, const uint64_t* k_canon, uint64_t* out) {
    g2_jac j; g2_jac_from_aff(&j, (const g2_aff*)p); g2_jac_mul(&j, &j, k_canon); g2_jac_to_aff((g2_aff*)out, &j);
}
EXPORT void cref_add_g1(const uint64_t* p, const uint64_t* q, uint64_t* out) {
    g1_jac a, b; g1_jac_from_aff(&a, (const g1_aff*)p); g1_jac_from_aff(&b, (const g1_aff*)q);
    g1_jac_add(&a, &a, &b); g1_jac_to_aff((g1_aff*)out, &a);
}
EXPORT void cref_add_g2(const uint64_t* p, const uint64_t* q, uint64_t* out) {
    g2_jac a, b; g2_jac_from_aff(&a, (const g2_aff*)p); g2_jac_from_aff(&b, (const g2_aff*)q);
    g2_jac_add(&a, &a, &b); g2_jac_to_aff((g2_aff*)out, &a);
}

typedef struct {
    uint32_t n_vars, n_public, domain_size, num_constraints;
    const uint64_t *alpha_g1, *beta_g1, *delta_g1;      /* 64 B each */
    const uint64_t *beta_g2, *delta_g2;                 /* 128 B each */
    const uint64_t *a_query, *b_g1_query, *l_query, *h_query;   /* G1 arrays: n_vars, n_vars, n_vars-n_public-1, domain */
    const uint64_t *b_g2_query;                         /* G2 array: n_vars */
    const uint32_t *a_rowptr, *a_col; const uint64_t* a_val;
    const uint32_t *b_rowptr, *b_col; const uint64_t* b_val;
} cref_key;

/* phase seconds: [0] witness map, [1] H, [2] L, [3] A, [4] B1, [5] B2, [6] glue */
static double g_phase[8];
EXPORT void cref_last_phase_seconds(double* out) { memcpy(out, g_phase, sizeof g_phase); }

static void calc_coeff_g1(g1_jac* out, const uint64_t* query, const uint64_t* w_canon, size_t n_vars,
                          const uint64_t* vk_param, const uint64_t* delta, const uint64_t k[4], int nthreads) {
    /* calculate_coeff (ark-groth16 0.5.0 prover.rs): k*delta + query[0] + msm(query[1..], w[1..]) + vk_param */
    g1_jac acc, t;
    g1_msm(&acc, (const g1_aff*)query + 1, w_canon + 4, n_vars - 1, nthreads);
    g1_jac_from_aff(&t, (const g1_aff*)delta); g1_jac_mul(&t, &t, k);
    g1_jac_madd(&t, &t, (const g1_aff*)query);
    g1_jac_add(&t, &t, &acc);
    g1_jac_madd(&t, &t, (const g1_aff*)vk_param);
    *out = t;
}
static void calc_coeff_g2(g2_jac* out, const uint64_t* query, const uint64_t* w_canon, size_t n_vars,
                          const uint64_t* vk_param, const uint64_t* delta, const uint64_t k[4], int nthreads) {
    g2_jac acc, t;
    g2_msm(&acc, (const g2_aff*)query + 1, w_canon + 4, n_vars - 1, nthreads);
    g2_jac_from_aff(&t, (const g2_aff*)delta); g2_jac_mul(&t, &t, k);
    g2_jac_madd(&t, &t, (const g2_aff*)query);
    g2_jac_add(&t, &t, &acc);
    g2_jac_madd(&t, &t, (const g2_aff*)vk_param);
    *out = t;
}

/* Groth16::<Bn254, CircomReduction>::create_proof_with_reduction_and_matrices (src/zkey.rs:903-912,
 * benches/groth16.rs:52-61).  r, s canonical; w Montgomery (n_vars); proof_out = 256 B canonical LE
 * A.x A.y B.x.c0 B.x.c1 B.y.c0 B.y.c1 C.x C.y.  h_keep (optional) receives h (Montgomery). */
EXPORT int cref_prove(const cref_key* key, const uint64_t r[4], const uint64_t s[4], const uint64_t* w_mont,
                      uint8_t* proof_out, uint64_t* h_keep, int nthreads) {
    nthreads = nthreads_or_default(nthreads);
    size_t n = key->domain_size, nv = key->n_vars, li = (size_t)key->n_public + 1;
    double t0 = omp_get_wtime();
    fe* h = (fe*)malloc(n * sizeof(fe));
    int lg = cref_witness_map(key->num_constraints, (uint32_t)li, key->n_vars, key->a_rowptr, key->a_col, key->a_val,
                              key->b_rowptr, key->b_col, key->b_val, w_mont, (uint64_t*)h, nthreads);
    if (lg < 0 || ((size_t)1 << lg) != n) { free(h); return -1; }
    if (h_keep) memcpy(h_keep, h, n * sizeof(fe));
    double t1 = omp_get_wtime(); g_phase[0] = t1 - t0;
    /* into_bigint on every scalar */
    fe* hc = (fe*)malloc(n * sizeof(fe));
    fe* wc = (fe*)malloc(nv * sizeof(fe));
    #pragma omp parallel for num_threads(nthreads)
    for (size_t i = 0; i < n; i++) fr_to_canon(&hc[i], &h[i]);
    #pragma omp parallel for num_threads(nthreads)
    for (size_t i = 0; i < nv; i++) fr_to_canon(&wc[i], (const fe*)w_mont + i);
    g1_jac h_acc, l_acc, A, B1, C, t;
    g2_jac B2;
    t0 = omp_get_wtime();
    g1_msm(&h_acc, (const g1_aff*)key->h_query, (const uint64_t*)hc, n, nthreads);
    t1 = omp_get_wtime(); g_phase[1] = t1 - t0; t0 = t1;
    g1_msm(&l_acc, (const g1_aff*)key->l_query, (const uint64_t*)(wc + li), nv - li, nthreads);
    t1 = omp_get_wtime(); g_phase[2] = t1 - t0; t0 = t1;
    calc_coeff_g1(&A, key->a_query, (const uint64_t*)wc, nv, key->alpha_g1, key->delta_g1, r, nthreads);
    t1 = omp_get_wtime(); g_phase[3] = t1 - t0; t0 = t1;
    int r_zero = !(r[0] | r[1] | r[2] | r[3]);
    if (!r_zero) calc_coeff_g1(&B1, key->b_g1_query, (const uint64_t*)wc, nv, key->beta_g1, key->delta_g1, s, nthreads);
    else g1_jac_set_inf(&B1);
    t1 = omp_get_wtime(); g_phase[4] = t1 - t0; t0 = t1;
    calc_coeff_g2(&B2, key->b_g2_query, (const uint64_t*)wc, nv, key->beta_g2, key->delta_g2, s, nthreads);
    t1 = omp_get_wtime(); g_phase[5] = t1 - t0; t0 = t1;
    /* C = s*A + r*B1 - (r*s)*delta + l_acc + h_acc */
    fe rm, sm, rs; fr_from_canon(&rm, (const fe*)r); fr_from_canon(&sm, (const fe*)s);
    fr_mul(&rs, &rm, &sm); fr_to_canon(&rs, &rs);
    g1_aff Aaff; g1_jac_to_aff(&Aaff, &A);
    g1_jac Aj; g1_jac_from_aff(&Aj, &Aaff);
    g1_jac_mul(&C, &Aj, s);
    g1_jac_mul(&t, &B1, r); g1_jac_add(&C, &C, &t);
    g1_jac_from_aff(&t, (const g1_aff*)key->delta_g1); g1_jac_mul(&t, &t, rs.l); g1_jac_neg(&t, &t); g1_jac_add(&C, &C, &t);
    g1_jac_add(&C, &C, &l_acc); g1_jac_add(&C, &C, &h_acc);
    g1_aff Caff; g2_aff Baff;
    g1_jac_to_aff(&Caff, &C); g2_jac_to_aff(&Baff, &B2);
    fe o[8];
    fq_to_canon(&o[0], &Aaff.x); fq_to_canon(&o[1], &Aaff.y);
    fq_to_canon(&o[2], &Baff.x.c0); fq_to_canon(&o[3], &Baff.x.c1); fq_to_canon(&o[4], &Baff.y.c0); fq_to_canon(&o[5], &Baff.y.c1);
    fq_to_canon(&o[6], &Caff.x); fq_to_canon(&o[7], &Caff.y);
    memcpy(proof_out, o, 256);
    g_phase[6] = omp_get_wtime() - t0;
    free(h); free(hc); free(wc);
    return 0;
}
