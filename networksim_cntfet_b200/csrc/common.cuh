// Shared helpers for the cntnet_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/cntnet_b200.h"

namespace cnt {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define CNT_CHECK_ARG(cond)                                                       \
    do {                                                                          \
        if (!(cond)) {                                                            \
            cnt::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond); \
            return CNT_EINVAL;                                                    \
        }                                                                         \
    } while (0)

#define CNT_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            cnt::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return CNT_ECUDA;                                                     \
        }                                                                         \
    } while (0)

// launch check: configuration errors surface here; asynchronous faults at the next sync
#define CNT_LAUNCHED()                                                            \
    do {                                                                          \
        cnt::count_launch();                                                      \
        cudaError_t e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) {                                                  \
            cnt::set_error("%s:%d: kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return CNT_ECUDA;                                                     \
        }                                                                         \
    } while (0)

// segmented stable LSD radix sort of (u64 key, u32 value) pairs (radix.cu); segments are
// [h_off[s], h_off[s+1]) with h_off[0] == 0; passes cover key bits [0, 8*npasses); buffers
// ping-pong between A and B and the result pointers are returned
int64_t segsort_workspace_bytes(int64_t n, int nseg);
int segsort_pairs(int nseg, const int32_t *h_off, unsigned long long *keyA, uint32_t *valA, unsigned long long *keyB,
                  uint32_t *valB, int npasses, void *ws, int64_t ws_bytes, cudaStream_t st,
                  unsigned long long **key_out, uint32_t **val_out);
int segsort_pairs32(int nseg, const int32_t *h_off, uint32_t *keyA, uint32_t *valA, uint32_t *keyB, uint32_t *valB, int npasses,
                    void *ws, int64_t ws_bytes, cudaStream_t st, uint32_t **key_out, uint32_t **val_out);

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// bump allocator over a caller-provided workspace
struct Workspace {
    char *base;
    int64_t size, used;
    Workspace(void *p, int64_t n) : base((char *)p), size(n), used(0) {}
    template <typename T>
    T *take(int64_t count) {
        int64_t bytes = align_up(count * (int64_t)sizeof(T), 256);
        if (used + bytes > size) return nullptr;
        T *r = (T *)(base + used);
        used += bytes;
        return r;
    }
};
static inline int64_t ws_bytes_for(int64_t count, int64_t elem) { return align_up(count * elem, 256); }

// largest b with off[b] <= i  (off has nb+1 entries, off[0]=0)
__device__ __forceinline__ int find_segment(const int32_t *__restrict__ off, int nb, int i) {
    int lo = 0, hi = nb;  // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (fixed tree); result valid in thread 0; sm must hold 32 doubles
__device__ __forceinline__ double block_sum(double v, double *sm) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? sm[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

}  // namespace cnt
