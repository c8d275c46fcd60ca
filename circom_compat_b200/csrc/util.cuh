// util.cuh - error plumbing shared by the host side of the library.  Exceptions never cross the C ABI: every
// extern "C" entry point catches B2gError and returns its code (include/b2groth.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <atomic>
#include <cstdint>

#define B2G_OK 0
#define B2G_E_DOMAIN (-1)   /* domain larger than 2^28: SynthesisError::PolynomialDegreeTooLarge, qap.rs:31 */
#define B2G_E_SHAPE (-2)    /* inconsistent sizes / null pointers */
#define B2G_E_DEVICE (-3)   /* CUDA or NCCL failure */
#define B2G_E_INPUT (-4)    /* malformed input data (e.g. off-curve point, bad zkey) */

namespace b2g {

struct B2gError : public std::runtime_error {
    int code;
    B2gError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

extern std::atomic<uint64_t> g_launch_count;   // kernels launched by this library (defined in msm.cu)

[[noreturn]] inline void throw_error(int code, const std::string& msg) { throw B2gError(code, msg); }

#define CUDA_CHECK(expr)                                                                                   \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            ::b2g::throw_error(B2G_E_DEVICE, std::string("CUDA error ") + cudaGetErrorString(_e) + " at " + \
                                                 __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")"); \
    } while (0)

}  // namespace b2g
