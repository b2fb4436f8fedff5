// common.cuh -- shared host/device helpers of libswt (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/swt.h"

#define SWT_API extern "C" __attribute__((visibility("default")))

namespace swt {

void set_error(const std::string &msg);

#define SWT_CUDA_OK(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::swt::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
            return SWT_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define SWT_REQUIRE(cond, msg)                                                                    \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            ::swt::set_error(std::string("invalid argument: ") + msg);                            \
            return SWT_ERR_ARG;                                                                   \
        }                                                                                         \
    } while (0)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static inline uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve aligned regions out of one caller-provided workspace
struct Carver {
    uint8_t *base; size_t off = 0;
    explicit Carver(void *p) : base((uint8_t *)p) {}
    template <typename T> T *take(size_t n) {
        off = align_up(off, 256);
        T *r = base ? (T *)(base + off) : nullptr;
        off += n * sizeof(T);
        return r;
    }
    size_t used() const { return align_up(off, 256); }
};

#ifdef __CUDACC__
// ---- UTF-8 -------------------------------------------------------------------------------------
// decodes the sequence starting at p (no validation: the host encoder produced it); adv = bytes used
__device__ __forceinline__ uint32_t utf8_decode(const uint8_t *p, uint32_t avail, uint32_t &adv) {
    uint32_t c = p[0];
    if (c < 0x80u) { adv = 1; return c; }
    if (c < 0xE0u && avail >= 2) { adv = 2; return ((c & 0x1Fu) << 6) | (p[1] & 0x3Fu); }
    if (c < 0xF0u && avail >= 3) { adv = 3; return ((c & 0x0Fu) << 12) | ((p[1] & 0x3Fu) << 6) | (p[2] & 0x3Fu); }
    if (avail >= 4) { adv = 4; return ((c & 0x07u) << 18) | ((p[1] & 0x3Fu) << 12) | ((p[2] & 0x3Fu) << 6) | (p[3] & 0x3Fu); }
    adv = 1; return c;
}

// block-wide exclusive scan of one u32 per thread (blockDim.x <= 1024, multiple of 32);
// returns the exclusive prefix, *total = block sum.  `warp_sums` is 33 words of shared memory.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < nw ? warp_sums[lane] : 0, si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, si, d);
            if (lane >= d) si += t;
        }
        if (lane < nw) warp_sums[lane] = si - s;
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    uint32_t r = warp_sums[wid] + incl - v;
    *total = warp_sums[32];
    __syncthreads();
    return r;
}
#endif  // __CUDACC__

}  // namespace swt
