// dfma_probe.cu - ROUND-2 CANDIDATE harness (standalone; not part of the product, its tests or its bench).
//   1. checks fp52 mont_mul / mont_sqr (FP64 pipe) against Fp<FqParams>::mul (IMAD pipe) on random residues, on the GPU;
//   2. measures field multiplications per second for: all warps integer, all warps FP64, and warps alternating between the
//      two (the question that decides whether a mixed accumulation kernel is worth building: do the pipes overlap?).
// Build: make -C experiments/dfma ; run on a B200: experiments/dfma/dfma_probe [n_threads_log2=20] [iters=2000]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../circom_compat_b200/csrc/fp.cuh"
#include "../../circom_compat_b200/csrc/ec.cuh"
#include "accumulate52.cuh"

using namespace b2g;
using b2g52::fe52;

__device__ fe52 to52(const fe& x) { return b2g52::from_u32(x.l); }
// five signed limbs with |value| < 2 p  ->  canonical residue in [0, p)
__device__ void reduce3(uint32_t* w) {
    fe r;
    for (int i = 0; i < 8; i++) r.l[i] = w[i];
    r = Fq::reduce_once(r); r = Fq::reduce_once(r); r = Fq::reduce_once(r);
    for (int i = 0; i < 8; i++) w[i] = r.l[i];
}
__device__ fe from52(const fe52& a) {
    fe r;
    b2g52::to_u32_plus_2p(a, r.l);
    reduce3(r.l);
    return r;
}

using G1c = Curve<Fq>;

// thread i: points P_k = [(i * 8 + k) * 2654435761 + 1] G (affine, product representation), k < 8, alternating signs;
// integer path: G1 XYZZ madd chain; FP64 path: madd52 chain + store52.  out[8 i ..] = X, Y, ZZ, ZZZ of both.
__global__ void ec_check_kernel(fe* __restrict__ out, uint32_t n, uint32_t* __restrict__ redo_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1c::Aff g; g.x = Fq::one(); g.y = Fq::dbl(Fq::one());                   // generator (1, 2)
    G1c::Pt acc = G1c::infinity();
    b2g52::Pt52 acc52; bool first = true, ok = true;
    for (int k = 0; k < 8; k++) {
        fe sc = fe_zero(); sc.l[0] = (i * 8u + k) * 2654435761u + 1u; sc.l[1] = i ^ (0x9e3779b9u * k);
        G1c::Aff p = G1c::to_affine(G1c::mul_scalar(G1c::from_affine(g), sc.l));
        if (k & 1) p.y = Fq::neg(p.y);
        G1c::madd(acc, p);
        fe52 x2 = to52(p.x), y2 = to52(p.y);
        if (first) { b2g52::from_affine52(acc52, x2, y2); first = false; }
        else ok = ok && b2g52::madd52(acc52, x2, y2);
    }
    if (!ok) atomicAdd(redo_count, 1u);
    out[8 * i + 0] = acc.x; out[8 * i + 1] = acc.y; out[8 * i + 2] = acc.zz; out[8 * i + 3] = acc.zzz;
    uint32_t w[32];
    b2g52::store52(acc52, w, reduce3);
    for (int c = 0; c < 4; c++) { fe r; for (int j = 0; j < 8; j++) r.l[j] = w[8 * c + j]; out[8 * i + 4 + c] = r; }
}

using G2c = Curve<Fq2>;
using b2g52::fe52x2;
__device__ fe52x2 to52x2(const fe2& x) { fe52x2 r; r.c0 = to52(x.c0); r.c1 = to52(x.c1); return r; }
__device__ G2c::Aff g2_gen() {       // the standard BN254 G2 generator (same words as csrc/prover.cu: g2_generator)
    auto w = [](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7) {
        fe r; r.l[0] = a0; r.l[1] = a1; r.l[2] = a2; r.l[3] = a3; r.l[4] = a4; r.l[5] = a5; r.l[6] = a6; r.l[7] = a7; return Fq::from_canonical(r); };
    G2c::Aff g;
    g.x.c0 = w(0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu);
    g.x.c1 = w(0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u);
    g.y.c0 = w(0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u);
    g.y.c1 = w(0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u);
    return g;
}
// G2 counterpart of ec_check_kernel: 6 additions per thread; out[16 i ..] = 8 residues of the integer result, then of the FP64 one
__global__ void g2_check_kernel(fe* __restrict__ out, uint32_t n, uint32_t* __restrict__ redo_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const G2c::Aff g = g2_gen();
    G2c::Pt acc = G2c::infinity();
    b2g52::Pt52x2 acc52; bool first = true, ok = true;
    for (int k = 0; k < 6; k++) {
        uint32_t sc[8] = {(i * 6u + k) * 2654435761u + 1u, i ^ (0x9e3779b9u * k), 0, 0, 0, 0, 0, 0};
        G2c::Aff p = G2c::to_affine(G2c::mul_scalar(G2c::from_affine(g), sc));
        if (k & 1) p.y = Fq2::neg(p.y);
        G2c::madd(acc, p);
        fe52x2 x2 = to52x2(p.x), y2 = to52x2(p.y);
        if (first) { b2g52::from_affine52_g2(acc52, x2, y2); first = false; }
        else ok = ok && b2g52::madd52_g2(acc52, x2, y2);
    }
    if (!ok) atomicAdd(redo_count, 1u);
    fe* o = out + 16 * (size_t)i;
    o[0] = acc.x.c0; o[1] = acc.x.c1; o[2] = acc.y.c0; o[3] = acc.y.c1; o[4] = acc.zz.c0; o[5] = acc.zz.c1; o[6] = acc.zzz.c0; o[7] = acc.zzz.c1;
    uint32_t w[64];
    b2g52::store52_g2(acc52, w, reduce3);
    for (int c = 0; c < 8; c++) { fe r; for (int j = 0; j < 8; j++) r.l[j] = w[8 * c + j]; o[8 + c] = r; }
}
// G2 madd throughput, same scheme as madd_probe_kernel (points = pairs of residues from the input buffer)
__global__ void __launch_bounds__(128) madd_g2_probe_kernel(int mode, int iters, const fe* __restrict__ pts, fe* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool use_fp64 = mode == 1 || (mode == 2 && ((threadIdx.x >> 5) & 1));
    const fe* my = pts + 4 * (size_t)((i * 37u) & 511u);
    if (!use_fp64) {
        G2c::Pt acc = G2c::infinity();
        #pragma unroll 1
        for (int k = 0; k < iters; k++) {
            const fe* q = my + 4 * (k & 31);
            G2c::Aff p; p.x.c0 = fe_load_nc(q); p.x.c1 = fe_load_nc(q + 1); p.y.c0 = fe_load_nc(q + 2); p.y.c1 = fe_load_nc(q + 3);
            G2c::madd(acc, p);
        }
        out[i] = Fq::add(acc.x.c0, acc.zz.c1);
    } else {
        b2g52::Pt52x2 acc;
        { fe2 x, y; x.c0 = fe_load_nc(my); x.c1 = fe_load_nc(my + 1); y.c0 = fe_load_nc(my + 2); y.c1 = fe_load_nc(my + 3);
          b2g52::from_affine52_g2(acc, to52x2(x), to52x2(y)); }
        bool ok = true;
        #pragma unroll 1
        for (int k = 1; k < iters; k++) {
            const fe* q = my + 4 * (k & 31);
            fe2 x, y; x.c0 = fe_load_nc(q); x.c1 = fe_load_nc(q + 1); y.c0 = fe_load_nc(q + 2); y.c1 = fe_load_nc(q + 3);
            ok = b2g52::madd52_g2(acc, to52x2(x), to52x2(y)) && ok;
        }
        fe r = from52(b2g52::add(acc.X.c0, acc.ZZ.c1));
        if (!ok) r.l[0] ^= 1u;
        out[i] = r;
    }
}

// compile check of the accumulation kernel (launched by round 2's integration, not by this probe)
__global__ void __launch_bounds__(128, 4) msm_accumulate52_kernel(const void* __restrict__ table, const uint32_t* __restrict__ entries,
                                      const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t chunk, void* __restrict__ buckets,
                                      void* __restrict__ frag_first, void* __restrict__ frag_last, uint32_t t_begin, uint32_t t_end,
                                      uint32_t* __restrict__ redo_list, uint32_t* __restrict__ redo_count) {
    const uint32_t t = t_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_end) return;
    b2g52::accumulate52_run(t, table, entries, offsets, nb, chunk, buckets, frag_first, frag_last, redo_list, redo_count,
                            [](uint32_t* v) { fe r; for (int i = 0; i < 8; i++) r.l[i] = v[i]; r = Fq::reduce_once(r); for (int i = 0; i < 8; i++) v[i] = r.l[i]; });
}

// madd throughput: every thread walks `iters` times over 32 affine points that sit in global memory (L2-resident), as the
// accumulation kernel does with table rows.  mode as in probe_kernel.
__global__ void __launch_bounds__(128) madd_probe_kernel(int mode, int iters, const fe* __restrict__ pts, fe* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool use_fp64 = mode == 1 || (mode == 2 && ((threadIdx.x >> 5) & 1));
    const fe* my = pts + 2 * (size_t)((i * 37u) & 1023u);
    if (!use_fp64) {
        G1c::Pt acc = G1c::infinity();
        #pragma unroll 1
        for (int k = 0; k < iters; k++) {
            G1c::Aff p; p.x = fe_load_nc(my + 2 * (k & 31)); p.y = fe_load_nc(my + 2 * (k & 31) + 1);
            G1c::madd(acc, p);
        }
        out[i] = Fq::add(acc.x, acc.zz);
    } else {
        b2g52::Pt52 acc;
        { fe x = fe_load_nc(my), y = fe_load_nc(my + 1); b2g52::from_affine52(acc, to52(x), to52(y)); }
        bool ok = true;
        #pragma unroll 1
        for (int k = 1; k < iters; k++) {
            fe x = fe_load_nc(my + 2 * (k & 31)), y = fe_load_nc(my + 2 * (k & 31) + 1);
            ok = b2g52::madd52(acc, to52(x), to52(y)) && ok;
        }
        fe r = from52(b2g52::add(acc.X, acc.ZZ));
        if (!ok) r.l[0] ^= 1u;
        out[i] = r;
    }
}

// out[3 i + 0] = x * y * 2^-256 (integer path), out[3 i + 1] = 16 * (x * y * 2^-260) (FP64 path), out[3 i + 2] = same for x^2
__global__ void check_kernel(const fe* __restrict__ in, fe* __restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x = in[2 * i], y = in[2 * i + 1];
    out[4 * i + 0] = Fq::mul(x, y);
    out[4 * i + 2] = Fq::mul(x, x);
    fe52 a = to52(x), b = to52(y);
    fe r = from52(b2g52::mont_mul(a, b));
    for (int k = 0; k < 4; k++) r = Fq::dbl(r);
    out[4 * i + 1] = r;
    fe s = from52(b2g52::mont_sqr(a));
    for (int k = 0; k < 4; k++) s = Fq::dbl(s);
    out[4 * i + 3] = s;
}

// mode 0: every warp on the integer pipe; 1: every warp on the FP64 pipe; 2: even warps integer, odd warps FP64.
// Two independent dependency chains per thread, `iters` products on each.
__global__ void __launch_bounds__(128) probe_kernel(int mode, int iters, const fe* __restrict__ in, fe* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool use_fp64 = mode == 1 || (mode == 2 && ((threadIdx.x >> 5) & 1));
    fe x = in[2 * i], y = in[2 * i + 1];
    if (!use_fp64) {
        fe u = x, v = y;
        #pragma unroll 1
        for (int k = 0; k < iters; k++) { u = Fq::mul(u, y); v = Fq::mul(v, x); }
        out[i] = Fq::add(u, v);
    } else {
        fe52 a = to52(x), b = to52(y), u = a, v = b;
        #pragma unroll 1
        for (int k = 0; k < iters; k++) { u = b2g52::mont_mul(u, b); v = b2g52::mont_mul(v, a); }
        out[i] = from52(b2g52::normalize(b2g52::add(u, v)));
    }
}


// ---- pipe micro-benchmarks: 8 independent chains per thread of one instruction kind, to read off issue rates
//   kind 0 DFMA (rn)   1 DFMA.RM (round down)   2 DADD   3 IMAD.WIDE.U32   4 64-bit integer add (IADD3 + IADD3.X)
//   kind 5: even warps IMAD.WIDE, odd warps DFMA (rn)     kind 6: even warps IMAD.WIDE, odd warps 64-bit adds
__global__ void __launch_bounds__(256) pipe_probe_kernel(int kind, int iters, double seed, double* __restrict__ out, unsigned long long* __restrict__ clk) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int odd = (threadIdx.x >> 5) & 1;
    int k = kind;
    if (kind == 5) k = odd ? 0 : 3;
    if (kind == 6) k = odd ? 4 : 3;
    double d[8]; unsigned long long u[8]; uint32_t m = (uint32_t)seed | 1u;
    #pragma unroll
    for (int j = 0; j < 8; j++) { d[j] = seed + j + i * 1e-9; u[j] = (unsigned long long)(i + j) * 0x9e3779b97f4a7c15ull; }
    const double a = 1.0000001, b = seed * 1e-7;
    unsigned long long t0 = 0, g0 = 0;
    if (i == 0) { t0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0)); }
    #pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (k == 0) {
            #pragma unroll
            for (int j = 0; j < 8; j++) d[j] = __fma_rn(d[j], a, b);
        } else if (k == 1) {
            #pragma unroll
            for (int j = 0; j < 8; j++) d[j] = __fma_rd(d[j], a, b);
        } else if (k == 2) {
            #pragma unroll
            for (int j = 0; j < 8; j++) d[j] = __dadd_rn(d[j], b);
        } else if (k == 3) {
            #pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(u[j]) : "r"((uint32_t)u[j]), "r"(m));
        } else {
            #pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t lo = (uint32_t)u[j], hi = (uint32_t)(u[j] >> 32);
                asm volatile("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(m), "r"((uint32_t)it));
                u[j] = (unsigned long long)lo | ((unsigned long long)hi << 32);
            }
        }
    }
    if (i == 0) { unsigned long long t1 = clock64(), g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1)); clk[0] = t1 - t0; clk[1] = g1 - g0; }
    double r = 0; unsigned long long q = 0;
    #pragma unroll
    for (int j = 0; j < 8; j++) { r += d[j]; q ^= u[j]; }
    out[i] = r + (double)q;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    const int logn = argc > 1 ? atoi(argv[1]) : 20, iters = argc > 2 ? atoi(argv[2]) : 2000;
    const uint32_t n = 1u << logn;
    std::vector<uint32_t> h((size_t)n * 16);
    uint64_t s = 0x9e3779b97f4a7c15ull;
    for (auto& w : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w = (uint32_t)(s >> 16); }
    for (size_t e = 0; e < (size_t)n * 2; e++) h[e * 8 + 7] &= 0x1fffffffu;      // < 2^253 < p
    fe *d_in, *d_out;
    CK(cudaMalloc(&d_in, (size_t)n * 64)); CK(cudaMalloc(&d_out, (size_t)n * 128));
    CK(cudaMemcpy(d_in, h.data(), (size_t)n * 64, cudaMemcpyHostToDevice));
    // 1. parity of the two multipliers
    const uint32_t nc = n < 65536 ? n : 65536;
    check_kernel<<<(nc + 127) / 128, 128>>>(d_in, d_out, nc);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> o((size_t)nc * 32);
    CK(cudaMemcpy(o.data(), d_out, (size_t)nc * 128, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (uint32_t i = 0; i < nc; i++)
        bad += (memcmp(&o[(size_t)i * 32], &o[(size_t)i * 32 + 8], 32) != 0) + (memcmp(&o[(size_t)i * 32 + 16], &o[(size_t)i * 32 + 24], 32) != 0);
    printf("{\"check\": {\"pairs\": %u, \"mismatches\": %zu}}\n", nc, bad);
    // 1b. parity of the two mixed additions
    {
        const uint32_t ne = 8192;
        uint32_t* d_redo; CK(cudaMalloc(&d_redo, 4)); CK(cudaMemset(d_redo, 0, 4));
        ec_check_kernel<<<ne / 64, 64>>>(d_out, ne, d_redo);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> e((size_t)ne * 64); uint32_t redo = 0;
        CK(cudaMemcpy(e.data(), d_out, (size_t)ne * 256, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&redo, d_redo, 4, cudaMemcpyDeviceToHost));
        size_t ebad = 0;
        for (uint32_t i = 0; i < ne; i++) ebad += memcmp(&e[(size_t)i * 64], &e[(size_t)i * 64 + 32], 128) != 0;
        printf("{\"ec_check\": {\"chains\": %u, \"mismatches\": %zu, \"redo_flags\": %u}}\n", ne, ebad, redo);
        bad += ebad;
        // points for the madd probe: reuse the affine points implied by the first 2048 outputs?  simpler: X, Y of the integer
        // results are not affine; build 1024 affine points = first 1024 (x, y) pairs of a fixed-base walk on the host side is
        // overkill - take the check kernel's inputs instead: d_in holds residues < p that are not curve points, which is fine
        // for timing (no exceptional case can trigger: the formulas never test curve membership).
    }
    // 1c. parity of the two G2 mixed additions
    {
        const uint32_t ne = 2048;
        uint32_t* d_redo; CK(cudaMalloc(&d_redo, 4)); CK(cudaMemset(d_redo, 0, 4));
        g2_check_kernel<<<ne / 64, 64>>>(d_out, ne, d_redo);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> e((size_t)ne * 128); uint32_t redo = 0;
        CK(cudaMemcpy(e.data(), d_out, (size_t)ne * 512, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&redo, d_redo, 4, cudaMemcpyDeviceToHost));
        size_t ebad = 0;
        for (uint32_t i = 0; i < ne; i++) ebad += memcmp(&e[(size_t)i * 128], &e[(size_t)i * 128 + 64], 256) != 0;
        printf("{\"g2_check\": {\"chains\": %u, \"mismatches\": %zu, \"redo_flags\": %u}}\n", ne, ebad, redo);
        bad += ebad;
    }
    // 0. pipe micro-benchmarks (and the SM clock under each load: clock64 / globaltimer of one thread)
    {
        double* d_o; unsigned long long* d_clk; CK(cudaMalloc(&d_o, (size_t)n * 8)); CK(cudaMalloc(&d_clk, 16));
        cudaEvent_t a0, a1; CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1));
        const char* names[7] = {"DFMA rn", "DFMA rd (.RM)", "DADD", "IMAD.WIDE.U32", "64-bit add", "even IMAD.WIDE / odd DFMA", "even IMAD.WIDE / odd 64-bit add"};
        int dev = 0, sms = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const uint32_t np = (uint32_t)sms * 2048u;                        // full occupancy: 8 CTAs of 256 threads per SM
        for (int kind = 0; kind < 7; kind++) {
            const int it = 4096;
            pipe_probe_kernel<<<np / 256, 256>>>(kind, 64, 3.0, d_o, d_clk);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a0));
            pipe_probe_kernel<<<np / 256, 256>>>(kind, it, 3.0, d_o, d_clk);
            CK(cudaEventRecord(a1)); CK(cudaDeviceSynchronize());
            float ms = 0; CK(cudaEventElapsedTime(&ms, a0, a1));
            unsigned long long c[2]; CK(cudaMemcpy(c, d_clk, 16, cudaMemcpyDeviceToHost));
            const double ops = 8.0 * it * np, mhz = c[1] ? 1e3 * (double)c[0] / (double)c[1] : 0.0;
            printf("{\"pipe\": \"%s\", \"ms\": %.3f, \"thread_ops_per_s\": %.4g, \"lanes_per_clk_per_sm\": %.1f, \"sm_mhz\": %.0f}\n",
                   names[kind], ms, ops / (ms * 1e-3), mhz > 0 ? ops / (ms * 1e-3) / (mhz * 1e6) / sms : 0.0, mhz);
        }
    }
    // 2. throughput
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 3; mode++) {
        probe_kernel<<<n / 128, 128>>>(mode, 16, d_in, d_out);                  // warm-up
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        probe_kernel<<<n / 128, 128>>>(mode, iters, d_in, d_out);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double muls = 2.0 * iters * n;
        printf("{\"mode\": %d, \"what\": \"%s\", \"ms\": %.3f, \"field_muls_per_s\": %.4g}\n", mode,
               mode == 0 ? "all warps IMAD" : mode == 1 ? "all warps DFMA" : "even warps IMAD, odd warps DFMA", ms, muls / (ms * 1e-3));
    }
    for (int mode = 0; mode < 3; mode++) {
        const int it = iters / 8 > 32 ? iters / 8 : 32;
        madd_probe_kernel<<<n / 128, 128>>>(mode, 32, d_in, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        madd_probe_kernel<<<n / 128, 128>>>(mode, it, d_in, d_out);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"madd_mode\": %d, \"what\": \"%s\", \"ms\": %.3f, \"mixed_adds_per_s\": %.4g}\n", mode,
               mode == 0 ? "all warps IMAD" : mode == 1 ? "all warps DFMA" : "even warps IMAD, odd warps DFMA", ms, (double)it * n / (ms * 1e-3));
    }
    for (int mode = 0; mode < 3; mode++) {
        const int it = iters / 24 > 32 ? iters / 24 : 32;
        madd_g2_probe_kernel<<<n / 128, 128>>>(mode, 32, d_in, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        madd_g2_probe_kernel<<<n / 128, 128>>>(mode, it, d_in, d_out);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"madd_g2_mode\": %d, \"what\": \"%s\", \"ms\": %.3f, \"mixed_adds_per_s\": %.4g}\n", mode,
               mode == 0 ? "all warps IMAD" : mode == 1 ? "all warps DFMA" : "even warps IMAD, odd warps DFMA", ms, (double)it * n / (ms * 1e-3));
    }
    return bad ? 2 : 0;
}
