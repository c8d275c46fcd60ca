// msm.cuh - multi-scalar multiplication sum_i k_i * P_i over BN254 G1 / G2 for a FIXED base set (a proving-key query).
//
// Replaces ark-ec 0.5.0 VariableBaseMSM::msm_bigint as called five times by ark-groth16 0.5.0
// create_proof_with_assignment (call sites /root/reference/src/zkey.rs:903-912, benches/groth16.rs:52-61; restated in
// SURVEY.md 3.4 / App. C.3).  Same signed base-2^c digit decomposition, but organised for B200:
//
//   * the bases never change between proofs, and HBM is 180 GB, so at key-load time every base P_i is expanded into the
//     affine table T[w][i] = 2^(c*w) * P_i (w = 0..nwin-1).  All windows then share ONE bucket set of 2^(c-1) buckets:
//     no per-window bucket reduction and no Horner doubling chain on the critical path.
//   * per proof: (1) canonical scalars + digit histogram (one thread per (window, scalar)), (2) exclusive scan,
//     (3) scatter of (table row | sign) into bucket-sorted order - a counting sort with warp-aggregated atomics, no
//     library sort; the sorted list is shared by every query that pairs the same scalars (L, A, B1, B2);
//     (4) perfectly load-balanced accumulation: each thread owns a fixed-length run of the sorted list, mixed-adds its
//     gathered bases in registers (XYZZ), and emits complete buckets directly and at most two boundary fragments;
//     (5) fragments are folded per bucket (big buckets by a whole CTA); (6) the weighted bucket sum sum_b (b+1)*B_b by
//     chunked running sums, a small double-and-add and a shared-memory tree.
//   Nothing in (1)-(6) synchronises with the host.
#pragma once
#include "ec.cuh"

namespace b2g {

constexpr int MSM_MAX_WIN = 32;          // c >= 8
constexpr int MSM_BIG_FRAGS = 32;        // buckets with more fragments than this are folded by a whole CTA

struct MsmPlan {                         // static per query
    uint32_t n = 0;                      // number of bases
    int c = 0, nwin = 0;
    uint32_t nbuckets = 0;               // 2^(c-1)
    void* table = nullptr;               // affine [nwin][n]
    bool g2 = false;
};

// ---- batched-affine pre-reduction (msm.cu section 4a) --------------------------------------------------------------
// Before the XYZZ accumulation, the bucket-sorted list is halved R times by pairwise AFFINE additions inside each bucket
// (level k list: bucket b holds ceil(cnt_{k-1}[b] / 2) points).  An affine addition costs 1 inversion + 2M + 1S; the
// inversions of a whole level are shared by Montgomery's trick across the grid (three kernels per level: prefix products,
// batch inversion of the per-thread totals, back-substitution), i.e. 5M + 1S per addition instead of the 8M + 2S of an XYZZ
// mixed addition.  The level structure depends only on the sorted scalars, so it lives in the sorting scratch and is shared
// by every query that pairs the same scalars (L, A, B1, B2).
constexpr int MSM_AFF_MAX_ROUNDS = 6;
constexpr int MSM_AFF_M = 16;            // additions per thread and level (interleaved: thread t owns slots t + i * stride)
constexpr int MSM_AFF_S = 64;            // thread totals per inversion thread

struct MsmScratch {                      // one per in-flight MSM
    uint32_t cap_n = 0; int cap_nwin = 0; uint32_t cap_buckets = 0; uint32_t chunk = 64; uint32_t sorted_n = 0; uint32_t reduce_chunk = 8;
    uint32_t *counts = nullptr, *offsets = nullptr, *cursor = nullptr, *entries = nullptr;
    uint32_t *big_list = nullptr, *big_count = nullptr;
    void *frag_first = nullptr, *frag_last = nullptr, *buckets = nullptr, *partials = nullptr, *result = nullptr;
    fe* scalars_canon = nullptr;         // n canonical scalars (filled by the digit pass)
    bool g2 = false, result_owned = false;
    cudaEvent_t prof0 = nullptr, prof1 = nullptr;   // optional: bracket the accumulate kernel (b2g_bench_msm)
    cudaStream_t tail = nullptr;                     // high-priority stream for the low-parallelism fold / weighted-sum kernels
    cudaEvent_t ev_acc = nullptr, ev_tail = nullptr;
    // batched-affine pre-reduction.  Sorting scratch: level offsets / source maps; accumulating scratch: point lists.
    int aff_rounds = 0;                              // levels planned by the last msm_sort (0 = straight to XYZZ)
    int aff_cap_rounds = 0;                          // levels the buffers were sized for
    uint32_t aff_nmax[MSM_AFF_MAX_ROUNDS + 1] = {};  // host-side upper bounds of the level list lengths
    uint32_t* aff_off[MSM_AFF_MAX_ROUNDS + 1] = {};  // [k] = offsets of level k (nb + 1); [0] aliases `offsets`
    uint32_t** aff_off_dev = nullptr;                // device copy of aff_off[] for the level-planning kernel
    uint32_t* aff_src[MSM_AFF_MAX_ROUNDS + 1] = {};  // [k][q] = first input slot in level k-1 | (pair << 31)
    void* aff_list[2] = {nullptr, nullptr};          // ping-pong affine point lists (levels 1, 3, 5 / 2, 4, 6)
    void *aff_pref = nullptr, *aff_totals = nullptr, *aff_tscratch = nullptr;
};

// Window size by base count, from whole-proof measurements on B200 (round 2: gpurun_out/r2_b25_*, r2_b26_*, r2_shard_sweep*.log;
// round 1 for >= 0.8 M).  Larger windows than the textbook log2(n) - 3 pay off below 2^19 because the table removes the
// per-window reductions: 2^14 keys 740 -> 970 proofs/s (c 11 -> 13), 2^16 457 -> 474 (13 -> 15), one rank of an 8-way sharded 2^20
// proof 8.8 -> 7.9 ms (14 -> 15).  B2G_MSM_C overrides (8..22) for tuning.
inline int msm_pick_c(uint32_t n) {
    if (n >= 3u << 18) return 17;        // 15 windows instead of 16 pay for the doubled bucket set from ~0.8 M bases up (18 already loses)
    if (n >= 1u << 19) return 16;
    if (n >= 1u << 15) return 15;
    if (n >= 1u << 13) return 13;
    if (n >= 1u << 11) return 11;
    int lg = 0; while ((1ull << (lg + 1)) <= n) lg++;
    return lg - 3 < 8 ? 8 : lg - 3;
}
inline int msm_nwin(int c) { return (255 + c - 1) / c; }

}  // namespace b2g
