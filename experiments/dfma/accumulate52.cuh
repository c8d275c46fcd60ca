// accumulate52.cuh - ROUND-2 CANDIDATE: the G1 bucket-accumulation kernel of csrc/msm.cu on the FP64 pipe.
//
// Same contract as msm_accumulate_kernel<G1, Fq> (each thread owns `chunk` consecutive sorted entries; complete buckets are
// written directly, at most two boundary fragments per thread) restricted to threads t in [t_begin, t_end), so that the
// integer kernel and this one can split the run list and execute concurrently on two streams: they contend for issue slots
// but not for a multiplier pipe (IMAD.WIDE there, DFMA + integer adds here).  Points are written in the product's layout
// (X, Y, ZZ, ZZZ as 8 x u32 residues in the 2^256 Montgomery domain; ZZ = 0 for infinity), so fold / reduce are unchanged.
// A run that meets the one case madd52 does not implement (equal x-coordinates) is appended to `redo_list`; the caller
// replays those runs with the integer kernel afterwards (it overwrites whatever this kernel had stored for them).
#pragma once
#include "ec52.cuh"
#include "fq2_52.cuh"

namespace b2g52 {

template <class ReduceOnce /* fe -> fe, subtracts p once if >= p */>
__device__ __forceinline__ void store_point52(const Pt52& acc, bool inf, uint4* dst, ReduceOnce reduce_once) {
    uint32_t w[32];
    if (inf) {
        #pragma unroll
        for (int i = 0; i < 32; i++) w[i] = 0;
    } else {
        store52(acc, w, [&](uint32_t* v) { reduce_once(v); reduce_once(v); reduce_once(v); });
    }
    #pragma unroll
    for (int i = 0; i < 8; i++) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

template <class ReduceOnce>
__device__ __forceinline__ void accumulate52_run(uint32_t t, const void* __restrict__ table, const uint32_t* __restrict__ entries,
                                                 const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t chunk,
                                                 void* __restrict__ buckets, void* __restrict__ frag_first, void* __restrict__ frag_last,
                                                 uint32_t* __restrict__ redo_list, uint32_t* __restrict__ redo_count, ReduceOnce reduce_once) {
    const uint32_t total = offsets[nb];
    const uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= total) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (uint32_t)min((uint64_t)total, start64 + chunk);
    uint32_t lo = 0, hi = nb;                       // invariant: offsets[lo] <= start < offsets[hi]
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (offsets[mid] <= start) lo = mid; else hi = mid; }
    uint32_t b = lo;
    uint32_t bucket_end = offsets[b + 1];
    while (bucket_end <= start) { b++; bucket_end = offsets[b + 1]; }
    Pt52 acc; bool inf = true;
    uint32_t seg_start = start;
    for (uint32_t pos = start; pos < end;) {
        const uint32_t e = entries[pos];
        const uint4* row = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(table) + (size_t)(e & 0x7fffffffu) * 64);
        const uint4 r0 = __ldg(row), r1 = __ldg(row + 1), r2 = __ldg(row + 2), r3 = __ldg(row + 3);
        const uint32_t xw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w}, yw[8] = {r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
        const bool p_inf = ((r0.x | r0.y | r0.z | r0.w | r1.x | r1.y | r1.z | r1.w | r2.x | r2.y | r2.z | r2.w | r3.x | r3.y | r3.z | r3.w) == 0u);
        if (!p_inf) {
            const fe52 x2 = from_u32(xw);
            fe52 y2 = from_u32(yw);
            if (e >> 31) y2 = neg(y2);
            if (inf) { from_affine52(acc, x2, y2); inf = false; }
            else if (!madd52(acc, x2, y2)) { redo_list[atomicAdd(redo_count, 1u)] = t; return; }
        }
        pos++;
        if (pos == bucket_end || pos == end) {
            const uint32_t bucket_start = offsets[b];
            uint4* dst = (bucket_start >= start && bucket_end <= end) ? reinterpret_cast<uint4*>(static_cast<uint8_t*>(buckets) + (size_t)b * 128)
                       : (seg_start == start)                         ? reinterpret_cast<uint4*>(static_cast<uint8_t*>(frag_first) + (size_t)t * 128)
                                                                      : reinterpret_cast<uint4*>(static_cast<uint8_t*>(frag_last) + (size_t)t * 128);
            store_point52(acc, inf, dst, reduce_once);
            inf = true;
            seg_start = pos;
            if (pos == bucket_end && pos < end) {
                do { b++; bucket_end = offsets[b + 1]; } while (bucket_end <= pos);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- G2
template <class ReduceOnce>
__device__ __forceinline__ void store_point52_g2(const Pt52x2& acc, bool inf, uint4* dst, ReduceOnce reduce_once) {
    uint32_t w[64];
    if (inf) {
        #pragma unroll
        for (int i = 0; i < 64; i++) w[i] = 0;
    } else {
        store52_g2(acc, w, [&](uint32_t* v) { reduce_once(v); reduce_once(v); reduce_once(v); });
    }
    #pragma unroll
    for (int i = 0; i < 16; i++) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

// G2 counterpart of accumulate52_run: table rows are 128 B (x.c0, x.c1, y.c0, y.c1), points 256 B
template <class ReduceOnce>
__device__ __forceinline__ void accumulate52_g2_run(uint32_t t, const void* __restrict__ table, const uint32_t* __restrict__ entries,
                                                    const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t chunk,
                                                    void* __restrict__ buckets, void* __restrict__ frag_first, void* __restrict__ frag_last,
                                                    uint32_t* __restrict__ redo_list, uint32_t* __restrict__ redo_count, ReduceOnce reduce_once) {
    const uint32_t total = offsets[nb];
    const uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= total) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (uint32_t)min((uint64_t)total, start64 + chunk);
    uint32_t lo = 0, hi = nb;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (offsets[mid] <= start) lo = mid; else hi = mid; }
    uint32_t b = lo;
    uint32_t bucket_end = offsets[b + 1];
    while (bucket_end <= start) { b++; bucket_end = offsets[b + 1]; }
    Pt52x2 acc; bool inf = true;
    uint32_t seg_start = start;
    for (uint32_t pos = start; pos < end;) {
        const uint32_t e = entries[pos];
        const uint4* row = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(table) + (size_t)(e & 0x7fffffffu) * 128);
        uint32_t w[32]; uint32_t any = 0;
        #pragma unroll
        for (int i = 0; i < 8; i++) { const uint4 r = __ldg(row + i); w[4 * i] = r.x; w[4 * i + 1] = r.y; w[4 * i + 2] = r.z; w[4 * i + 3] = r.w; any |= r.x | r.y | r.z | r.w; }
        if (any) {
            fe52x2 x2, y2;
            x2.c0 = from_u32(w); x2.c1 = from_u32(w + 8); y2.c0 = from_u32(w + 16); y2.c1 = from_u32(w + 24);
            if (e >> 31) y2 = neg2(y2);
            if (inf) { from_affine52_g2(acc, x2, y2); inf = false; }
            else if (!madd52_g2(acc, x2, y2)) { redo_list[atomicAdd(redo_count, 1u)] = t; return; }
        }
        pos++;
        if (pos == bucket_end || pos == end) {
            const uint32_t bucket_start = offsets[b];
            uint4* dst = (bucket_start >= start && bucket_end <= end) ? reinterpret_cast<uint4*>(static_cast<uint8_t*>(buckets) + (size_t)b * 256)
                       : (seg_start == start)                         ? reinterpret_cast<uint4*>(static_cast<uint8_t*>(frag_first) + (size_t)t * 256)
                                                                      : reinterpret_cast<uint4*>(static_cast<uint8_t*>(frag_last) + (size_t)t * 256);
            store_point52_g2(acc, inf, dst, reduce_once);
            inf = true;
            seg_start = pos;
            if (pos == bucket_end && pos < end) {
                do { b++; bucket_end = offsets[b + 1]; } while (bucket_end <= pos);
            }
        }
    }
}

}  // namespace b2g52
