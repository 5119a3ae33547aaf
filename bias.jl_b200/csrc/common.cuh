// common.cuh — shared device/host helpers for libbias_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <cstdio>
#include <cstdarg>
#include "../../include/bias_cuda.h"

namespace bias {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char *fmt, ...);
int  fail(int code, const char *fmt, ...);

#define BIAS_CUDA_TRY(expr)                                                                 \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::bias::fail(_e == cudaErrorMemoryAllocation ? BIAS_ENOMEM : BIAS_ECUDA, \
                                "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                                __FILE__, __LINE__);                                        \
    } while (0)

#define BIAS_TRY(expr)              \
    do {                            \
        int _rc = (expr);           \
        if (_rc != BIAS_OK) return _rc; \
    } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- Philox4x32-10 (counter-based; generated on chip, no RNG state in HBM) --------------
struct Philox {
    uint32_t k0, k1;
    __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#ifdef __CUDA_ARCH__
        lo = a * b;
        hi = __umulhi(a, b);
#else
        uint64_t p = (uint64_t)a * b;
        lo = (uint32_t)p;
        hi = (uint32_t)(p >> 32);
#endif
    }
    // counter = (c0, c1, c2, c3) -> four 32-bit outputs
    __host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t out[4]) const {
        uint32_t key0 = k0, key1 = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ key0, n1 = lo1, n2 = hi0 ^ c3 ^ key1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            key0 += 0x9E3779B9u;
            key1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// 53-bit uniform in (0, 1) from two 32-bit words
__host__ __device__ inline double u01_double(uint32_t a, uint32_t b) {
    uint64_t v = (((uint64_t)a << 32) | b) >> 11;           // 53 bits
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}
// 24-bit uniform in (0, 1)
__host__ __device__ inline float u01_float(uint32_t a) {
    return ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// Stream ids keep the Philox streams of different consumers disjoint.
enum : uint32_t { kStreamItemDraw = 1, kStreamTable = 2, kStreamHyper = 3, kStreamBirth = 4 };

#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// x / y as one MUFU.RCP and one FMUL (no denormal rescaling: the callers' divisors are >= alpha or >= V beta)
__device__ __forceinline__ float fast_div(float x, float y) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    return x * r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float y = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += y;
    }
    return v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double y = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += y;
    }
    return v;
}
// fire-and-forget reductions (no value returns to the SM)
__device__ __forceinline__ void red_add(int32_t *p, int32_t v) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add(unsigned *p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
#endif

}  // namespace bias
