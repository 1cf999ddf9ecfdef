// Error plumbing, launch counter and the device-wide exclusive scan used by every two-pass
// (count -> scan -> fill) stage of the pipeline.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace cnt {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }
}  // namespace cnt

extern "C" int cnt_abi_version(void) { return CNT_ABI_VERSION; }
extern "C" const char *cnt_last_error(void) { return cnt::g_err; }
extern "C" int64_t cnt_launch_count(void) { return cnt::g_launches.load(); }

namespace {
constexpr int SCAN_T = 1024;      // threads per block
constexpr int SCAN_ITEMS = 4;     // items per thread
constexpr int SCAN_TILE = SCAN_T * SCAN_ITEMS;

// exclusive scan of one int per thread across the block; returns the exclusive prefix and
// writes the block total to *total (all threads)
__device__ __forceinline__ int block_excl_scan(int v, int *sm /*33 ints*/, int *total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm[w] = incl;
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        int s = lane < nw ? sm[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        sm[lane] = si - s;          // exclusive warp offsets
        if (lane == 31) sm[32] = si;  // block total
    }
    __syncthreads();
    int r = sm[w] + incl - v;
    *total = sm[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_blocksum(const int32_t *__restrict__ in, int64_t n,
                                                           int32_t *__restrict__ bsum) {
    __shared__ int sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) s += in[base + k];
    int total;
    block_excl_scan(s, sm, &total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single block: exclusive scan of bsum[nb] in place
__global__ void __launch_bounds__(SCAN_T) k_scan_top(int32_t *__restrict__ bsum, int64_t nb) {
    __shared__ int sm[33];
    int64_t per = (nb + SCAN_T - 1) / SCAN_T;
    int64_t lo = (int64_t)threadIdx.x * per, hi = lo + per < nb ? lo + per : nb;
    int s = 0;
    for (int64_t i = lo; i < hi; i++) s += bsum[i];
    int total;
    int run = block_excl_scan(s, sm, &total);
    for (int64_t i = lo; i < hi; i++) {
        int v = bsum[i];
        bsum[i] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(SCAN_T) k_scan_final(const int32_t *in, int32_t *out, int64_t n,
                                                        const int32_t *__restrict__ bsum, int write_total) {
    __shared__ int sm[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int run = block_excl_scan(s, sm, &total) + bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    // the thread owning element n-1 also publishes the grand total
    if (write_total && base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = run;
}

__global__ void k_zero_total(int32_t *out) { out[0] = 0; }
}  // namespace

extern "C" int64_t cnt_scan_workspace_bytes(int64_t n) {
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    return cnt::ws_bytes_for(nb > 0 ? nb : 1, 4);
}

extern "C" int cnt_exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, int write_total, void *ws,
                                      int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(n >= 0);
    if (n == 0) {
        if (write_total) {
            k_zero_total<<<1, 1, 0, st>>>(out);
            CNT_LAUNCHED();
        }
        return CNT_OK;
    }
    CNT_CHECK_ARG(in && out && ws);
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ws_bytes < cnt_scan_workspace_bytes(n)) {
        cnt::set_error("scan workspace too small");
        return CNT_ECAPACITY;
    }
    int32_t *bsum = (int32_t *)ws;
    k_scan_blocksum<<<(unsigned)nb, SCAN_T, 0, st>>>(in, n, bsum);
    CNT_LAUNCHED();
    k_scan_top<<<1, SCAN_T, 0, st>>>(bsum, nb);
    CNT_LAUNCHED();
    k_scan_final<<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, n, bsum, write_total);
    CNT_LAUNCHED();
    return CNT_OK;
}

// ---------------------------------------------------------------------------------------
// FP64 FMA peak probe: the junction kernel's roofline denominator (MEASURED_PEAKS.json has no
// FP64 figure).  Every thread runs 8 independent DFMA chains; returns through *tflops the
// best of `reps` timed launches.  Synchronises.
// ---------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace

extern "C" int cnt_fp64_peak(double *scratch /* >= 148*16*256 doubles */, int reps, double *h_tflops, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(scratch && h_tflops && reps > 0);
    const int blocks = 148 * 16, iters = 20000;
    cudaEvent_t e0, e1;
    CNT_CUDA(cudaEventCreate(&e0));
    CNT_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int r = 0; r < reps + 1; r++) {
        CNT_CUDA(cudaEventRecord(e0, st));
        k_fp64_peak<<<blocks, 256, 0, st>>>(scratch, iters, 0.999999, 1e-9);
        CNT_LAUNCHED();
        CNT_CUDA(cudaEventRecord(e1, st));
        CNT_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CNT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 8 * (double)iters * blocks * 256 / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;     // first launch is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *h_tflops = best;
    return CNT_OK;
}
