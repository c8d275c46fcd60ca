// ntt.cuh - host-visible interface of ntt.cu (radix-2 domain tables, witness-map transforms, sparse mat-vec).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "fp.cuh"

namespace b2g {

struct NttDomain {
    int logn = -1, tl = 0, npass = 0;
    int pass_sb[4] = {0, 0, 0, 0}, pass_k[4] = {0, 0, 0, 0}, pass_tl[4] = {0, 0, 0, 0};
    bool radix8 = false;                                      // ntt_pass8_kernel (register radix-8 rounds) instead of ntt_pass_kernel
    fe *tw = nullptr, *ct = nullptr, *pw = nullptr;
    // LibsnarkReduction only: coset by the field generator g = 5 (ark-poly get_coset(F::GENERATOR))
    fe *cg = nullptr, *cginv = nullptr, *zinv = nullptr;      // n^-1 g^k, n^-1 g^-k (k < n), (g^n - 1)^-1
};

void ntt_domain_create(NttDomain& d, int logn, cudaStream_t st, bool libsnark = false);
void ntt_domain_destroy(NttDomain& d);
void ntt_witness_transform(const NttDomain& d, fe* a, fe* b, fe* c, fe* out, cudaStream_t st);
void ntt_transform_single(const NttDomain& d, fe* v, cudaStream_t st);
void ntt_witness_transform_libsnark(const NttDomain& d, fe* a, fe* b, fe* c, fe* scratch, fe* out, cudaStream_t st);
void ntt_plain(const NttDomain& d, fe* data, fe* tmp, bool inverse, cudaStream_t st);
void spmv_launch(uint32_t n, uint32_t m, uint32_t num_inputs, const uint32_t* a_rowptr, const uint32_t* a_col, const fe* a_val,
                 const uint32_t* b_rowptr, const uint32_t* b_col, const fe* b_val, const fe* w, fe* a, fe* b, fe* c, cudaStream_t st,
                 const uint32_t* c_rowptr = nullptr, const uint32_t* c_col = nullptr, const fe* c_val = nullptr);

}  // namespace b2g
