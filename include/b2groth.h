/* b2groth.h - C ABI of libb2groth.so, the B200-native (sm_100a CUDA) Groth16/BN254 prover hot path that stands in
 * for what ark-circom 0.5 obtains from ark-groth16 / ark-ec / ark-poly on the CPU.
 *
 * The reference (arkworks-rs/circom-compat) has no FFI; its extension points are Rust traits and generic functions.
 * Each entry point below names the reference interface it replaces (paths relative to /root/reference):
 *
 *   b2g_pk_load            <- the ProvingKey<Bn254> half of read_zkey()                    src/zkey.rs:53-60, 103-133
 *   b2g_matrices_load      <- the ConstraintMatrices<Fr> half of read_zkey()               src/zkey.rs:151-196
 *   b2g_witness_map        <- CircomReduction::witness_map_from_matrices                   src/circom/qap.rs:23-88
 *   b2g_prove              <- Groth16::<Bn254, CircomReduction>::create_proof_with_reduction_and_matrices
 *                             (call sites src/zkey.rs:903-912, benches/groth16.rs:52-61, 72-80; body = ark-groth16
 *                             0.5.0 create_proof_with_assignment, restated in SURVEY.md 3.4)
 *   b2g_msm_g1 / b2g_msm_g2<- VariableBaseMSM::msm_bigint (ark-ec 0.5.0) as used by that function
 *   b2g_ntt                <- Radix2EvaluationDomain::{fft,ifft}_in_place (ark-poly 0.5.0) as used at qap.rs:60-81
 *   b2g_prove_partial / b2g_prove_finish : the same proof split for base-range sharding over several GPUs
 *   b2g_fixed_base_g1/g2   <- the batch fixed-base multiplications of generate_random_parameters_with_reduction
 *                             (tests/groth16.rs:25); used to manufacture synthetic proving keys
 *
 * Conventions
 *   - every function returns 0 (B2G_OK) or a negative error code; b2g_last_error() gives a thread-local message.
 *     No exception or unwinding ever crosses this boundary.
 *   - field elements are 32 bytes, little-endian.  "mont" = Montgomery form with R = 2^256 exactly as a .zkey stores
 *     points (src/zkey.rs:327-332) and as arkworks keeps Fp256 in memory; "canon" = the plain integer.
 *   - G1 affine = x||y (64 B, mont), G2 affine = x.c0||x.c1||y.c0||y.c1 (128 B, mont); all-zero bytes = infinity
 *     (src/zkey.rs:340-360).
 *   - host pointers are only read/written during the call; handles own all device memory.
 *   - one proof in flight per b2g_ctx (a ctx is not thread-safe); create one ctx per GPU.
 *   - there is no CPU fallback: without a CUDA device every call fails with B2G_E_DEVICE.
 */
#ifndef B2GROTH_H
#define B2GROTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B2G_API __attribute__((visibility("default")))
#else
#define B2G_API
#endif

#define B2G_OK 0
#define B2G_E_DOMAIN (-1) /* evaluation domain too large: SynthesisError::PolynomialDegreeTooLarge (qap.rs:31,66) */
#define B2G_E_SHAPE (-2)  /* inconsistent sizes / null pointers */
#define B2G_E_DEVICE (-3) /* CUDA failure, or no CUDA device */
#define B2G_E_INPUT (-4)  /* malformed input data */

typedef struct b2g_ctx b2g_ctx;
typedef struct b2g_pk b2g_pk;
typedef struct b2g_mat b2g_mat;

/* Proving key as read_zkey() produces it (src/zkey.rs:121-130); every pointer is a HOST pointer. */
typedef struct {
    uint32_t n_vars;          /* zkey header nVars  (src/zkey.rs:303) */
    uint32_t n_public;        /* zkey header nPublic */
    uint32_t domain_size;     /* zkey header domainSize = number of H bases */
    uint32_t reserved;
    const void* alpha_g1;     /* 64 B  */
    const void* beta_g1;      /* 64 B  */
    const void* delta_g1;     /* 64 B  */
    const void* beta_g2;      /* 128 B */
    const void* delta_g2;     /* 128 B */
    const void* a_query;      /* n_vars G1                 zkey section 5 */
    const void* b_g1_query;   /* n_vars G1                 zkey section 6 */
    const void* b_g2_query;   /* n_vars G2                 zkey section 7 */
    const void* l_query;      /* n_vars - n_public - 1 G1  zkey section 8 */
    const void* h_query;      /* domain_size G1            zkey section 9 */
} b2g_pk_desc;

/* ConstraintMatrices<Fr> a and b (src/zkey.rs:181-193) in CSR form; c is empty on the zkey route.  Values mont. */
typedef struct {
    uint32_t num_constraints; /* m: rows kept (src/zkey.rs:171-175) */
    uint32_t num_inputs;      /* num_instance_variables = n_public + 1 (src/zkey.rs:182) */
    uint32_t n_vars;          /* length of the full assignment */
    uint32_t reduction;       /* B2G_REDUCTION_CIRCOM (0) or B2G_REDUCTION_LIBSNARK (1): which R1CSToQAP the handle serves */
    const uint32_t* a_rowptr; /* m + 1 */
    const uint32_t* a_col;
    const void* a_val;        /* nnz x 32 B mont */
    const uint32_t* b_rowptr;
    const uint32_t* b_col;
    const void* b_val;
    const uint32_t* c_rowptr; /* LibsnarkReduction only (the zkey route has no C matrix, src/zkey.rs:188-192); else NULL */
    const uint32_t* c_col;
    const void* c_val;
} b2g_mat_desc;

/* CircomReduction: snarkjs keys, H query of domain_size bases, h = (ab - c)(g w^j), g = omega_2n   (src/circom/qap.rs:23-88).
 * LibsnarkReduction: arkworks-generated keys (the default QAP of Groth16<Bn254>, tests/groth16.rs:9,25-35), H query of
 * domain_size - 1 bases [tau^i Z(tau)/delta], h = coefficients of (ab - c)/Z (ark-groth16 0.5.0 r1cs_to_qap.rs). */
#define B2G_REDUCTION_CIRCOM 0
#define B2G_REDUCTION_LIBSNARK 1

B2G_API const char* b2g_last_error(void);
B2G_API int b2g_version(void);
B2G_API int b2g_device_count(int* count);

/* One context per GPU.  shard_rank / shard_count partition every query's base range (rank r of R owns the r-th
 * contiguous slice); use 0 / 1 for a whole-proof context. */
B2G_API int b2g_ctx_create(int device, int shard_rank, int shard_count, b2g_ctx** out);
B2G_API int b2g_ctx_destroy(b2g_ctx* ctx);

/* Optional: allocate this context's per-proof scratch for (pk, mat) now (otherwise the first proof does it, which
 * synchronises the device - do it up front when several contexts share a device). */
B2G_API int b2g_ctx_prepare(b2g_ctx* ctx, b2g_pk* pk, b2g_mat* mat);

B2G_API int b2g_pk_load(b2g_ctx* ctx, const b2g_pk_desc* desc, b2g_pk** out);
B2G_API int b2g_pk_free(b2g_pk* pk);
B2G_API int b2g_matrices_load(b2g_ctx* ctx, const b2g_mat_desc* desc, b2g_mat** out);
B2G_API int b2g_matrices_free(b2g_mat* mat);

/* h = witness map of the handle's reduction; w_mont = n_vars x 32 B (host); h_out = domain x 32 B mont, natural order (host). */
B2G_API int b2g_witness_map(b2g_ctx* ctx, b2g_mat* mat, const void* w_mont, void* h_out, uint32_t* domain_size_out);

/* 256-byte proof: A.x A.y B.x.c0 B.x.c1 B.y.c0 B.y.c1 C.x C.y, canon little-endian; infinity = zeros.
 * r, s canon (32 B each).  Requires shard_count == 1. */
B2G_API int b2g_prove(b2g_ctx* ctx, b2g_pk* pk, b2g_mat* mat, const void* r_canon, const void* s_canon, const void* w_mont,
              uint8_t proof_out[256]);

/* The same call split in two so that ONE host thread can keep several proofs in flight (one b2g_ctx each): submit
 * enqueues the upload, the captured proof graph and the read-back and returns without waiting for the device (w_mont
 * should be page-locked, or the upload is synchronous); wait blocks until that proof is done and fills the proof_out
 * given to submit, which must stay valid until then.  At most one proof may be pending per context.
 * b2g_prove == submit + wait.  (Reference shape: a rayon/thread pool calling the synchronous prove; here the pipeline is
 * a property of the device queue, not of host threads.) */
B2G_API int b2g_prove_submit(b2g_ctx* ctx, b2g_pk* pk, b2g_mat* mat, const void* r_canon, const void* s_canon, const void* w_mont,
                             uint8_t proof_out[256]);
B2G_API int b2g_prove_wait(b2g_ctx* ctx);
/* Page-lock / release a host buffer (cudaHostRegister): a witness vector owned by the caller (a Rust Vec<Fr>, a std::vector)
 * uploads asynchronously and at full PCIe speed once registered.  Registering twice / unregistering an unknown pointer is not
 * an error. */
B2G_API int b2g_host_register(const void* ptr, size_t bytes);
B2G_API int b2g_host_unregister(const void* ptr);

/* Sharded proof.  partial_out (768 B, host or device-accessible host memory) = this rank's partial MSM results
 * [H, L, A, B1] as G1 XYZZ (128 B each) followed by B2 as G2 XYZZ (256 B), mont.  b2g_prove_finish folds
 * shard_count partials in rank order and assembles the proof; every rank obtains identical bytes. */
#define B2G_PARTIAL_BYTES 768
/* r_canon / s_canon may be NULL; when given, the (r, s)-only part of the proof assembly (r*delta, s*delta, ...) is
 * started here on a side stream so that b2g_prove_finish with the same (r, s) finds it done. */
B2G_API int b2g_prove_partial(b2g_ctx* ctx, b2g_pk* pk, b2g_mat* mat, const void* r_canon, const void* s_canon, const void* w_mont,
                              void* partial_out);
B2G_API int b2g_prove_finish(b2g_ctx* ctx, b2g_pk* pk, const void* partials_all, int count, const void* r_canon,
                     const void* s_canon, uint8_t proof_out[256]);

/* Sharded proof with the exchange fused into the proof-assembly kernel (NVLink peer memory instead of a host-driven
 * collective).  Every rank owns an exchange arena in its HBM (two slots of a 1 KiB record - the 768-byte partial plus s*A_k and
 * r*B1_k - with an epoch word each, then room for one transformed vector of the split witness map); peers map it through CUDA
 * IPC.  b2g_prove_sharded_p2p runs the partial MSMs, publishes the record with a system-scope release of the epoch, and the
 * gather kernel acquires every peer's epoch and reads the records straight out of peer memory; the assembly folds them in rank
 * order.  The whole sharded proof, exchange included, is one captured CUDA graph (the epoch is a device-resident counter).
 * All ranks must call it for the same proof; every rank obtains identical bytes.
 *   b2g_p2p_export  : B2G_IPC_HANDLE_BYTES: the cudaIpcMemHandle_t of this context's exchange arena + its capacity
 *   b2g_p2p_import  : records of ALL ranks in rank order (count x B2G_IPC_HANDLE_BYTES; the own entry is ignored)
 * With >= 3 ranks the witness map is split as well (CircomReduction): ranks 0, 1, 2 each transform ONE of a, b, c into their
 * arena, and every rank forms its slice of h = a*b - c reading the three vectors from peer HBM inside the pointwise kernel.
 * The arena is sized when it is first needed (export / import / connect_local) for the largest domain the context has been
 * prepared for, so call b2g_ctx_prepare BEFORE wiring the peers; otherwise every rank computes the whole map itself. */
#define B2G_IPC_HANDLE_BYTES 80
B2G_API int b2g_p2p_export(b2g_ctx* ctx, void* handle_out);
B2G_API int b2g_p2p_import(b2g_ctx* ctx, const void* handles_all, int count);
/* same wiring for shard contexts that live in ONE process (IPC handles cannot be opened by their exporter): ctxs in rank order */
B2G_API int b2g_p2p_connect_local(b2g_ctx** ctxs, int count);
B2G_API int b2g_prove_sharded_p2p(b2g_ctx* ctx, b2g_pk* pk, b2g_mat* mat, const void* r_canon, const void* s_canon,
                                  const void* w_mont, uint8_t proof_out[256]);

/* Kernel-level entry points (parity tests, benchmarks). All pointers host. */
B2G_API int b2g_msm_g1(b2g_ctx* ctx, const void* bases, const void* scalars, size_t n, int scalars_mont, void* out_xy_mont);
B2G_API int b2g_msm_g2(b2g_ctx* ctx, const void* bases, const void* scalars, size_t n, int scalars_mont, void* out_xy_mont);
B2G_API int b2g_ntt(b2g_ctx* ctx, void* data_mont, int log_n, int inverse);
B2G_API int b2g_fixed_base_g1(b2g_ctx* ctx, const void* scalars_canon, size_t n, void* out_affine_mont);
B2G_API int b2g_fixed_base_g2(b2g_ctx* ctx, const void* scalars_canon, size_t n, void* out_affine_mont);

/* Element-wise device arithmetic, for unit parity tests of the field / group layers.
 * op: 0 fq_mul, 1 fq_add, 2 fq_sub, 3 fr_mul, 4 fr_add, 5 fr_sub, 6 fq_inv, 7 fr_inv (b ignored),
 *     8 g1_add (a, b, out = n x 64 B affine), 9 g2_add (n x 128 B), 10 g1_dbl, 11 g2_dbl (b ignored),
 *     12 g1 mixed add, 13 g2 mixed add, 14 fq_sqr, 15 fq a*b - b*b, 16 fq mul through the lazy-reduction blocks (32 B),
 *     17 fq2_mul, 18 fq2_sqr, 19 fq2 a*b - b*swap(a) (n x 64 B: c0 || c1). */
B2G_API int b2g_test_op(b2g_ctx* ctx, int op, const void* a, const void* b, size_t n, void* out);

/* Timing of the last b2g_prove / b2g_prove_partial on this ctx, CUDA-event milliseconds:
 * [0] h2d witness, [1] witness map, [2] msm H, [3] msm L, [4] msm A, [5] msm B1, [6] msm B2, [7] glue + d2h,
 * [8] whole call (first event to last event).  With the captured proof graph (default) [1]..[6] read 0: the graph has no
 * interior events (B2G_GRAPH=0 in the environment at b2g_ctx_create restores direct launches and the per-phase times).
 * Host wall-clock milliseconds of the same call: [9] entry -> upload enqueued, [10] entry -> everything enqueued,
 * [11] time blocked in the final wait. */
B2G_API int b2g_last_timings(b2g_ctx* ctx, float out_ms[16]);

/* Benchmark helper: the device-resident part of a proof (witness already in HBM from the last b2g_prove call):
 * runs witness map + 5 MSMs + glue `iters` times and returns the average CUDA-event milliseconds.  iters < 0: enqueue
 * |iters| proofs and return without waiting (*avg_ms = 0): the caller synchronises the device and times the window itself,
 * which is how several contexts are measured over ONE common window. */
B2G_API int b2g_bench_device(b2g_ctx* ctx, b2g_pk* pk, b2g_mat* mat, int iters, float* avg_ms);
/* One MSM alone on the main stream, `iters` times (query: 0 H, 1 L, 2 A, 3 B1, 4 B2): out_ms[0] = average CUDA-event
 * milliseconds of the whole MSM, out_ms[1] = of its bucket-accumulation kernel (the dominant kernel, for the roofline). */
B2G_API int b2g_bench_msm(b2g_ctx* ctx, b2g_pk* pk, b2g_mat* mat, int query, int iters, float out_ms[2]);
/* number of kernel launches issued by this library (process-wide, all contexts) so far; a replayed proof graph counts the kernel
 * nodes it contains */
B2G_API int b2g_launch_count(b2g_ctx* ctx, uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* B2GROTH_H */
