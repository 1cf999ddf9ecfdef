// Streaming CSR SpMV for short rows (the reduced MNA matrices have ~4.7 off-diagonal entries per row).
//
// A thread-per-row loop reads val / col with a stride of one row length between lanes: every load
// instruction touches ~20 sectors for 32 useful words.  Here a CTA of 256 threads owns 256 consecutive
// rows, i.e. ONE contiguous range of entries.  It streams that range with coalesced 16-byte loads (int4 of
// col, 2 x double2 of val per thread and trip), multiplies every entry with its gathered x[col] on the fly
// and parks the products in shared memory; after one barrier thread t adds up the products of row t in
// CSR order (fixed order: deterministic bits).  Ranges longer than the staging buffer take several trips.
// DRAM sees val, col, rowptr, diag, y once and x roughly once (neighbouring rows are neighbouring sticks --
// Z-order numbering -- so the gathers are served by L1 / L2).
#pragma once
#include "common.cuh"

namespace cnt {
namespace spmv {

constexpr int TB = 256;            // threads = rows per CTA
constexpr int CAP = 2048;          // products staged per trip (16 KB)

// rows [r0, r0 + TB) of y = diag .* x + A x (diag may be null); returns y of row r0 + threadIdx.x (0 beyond
// n_rows).  s_prod: CAP doubles of shared memory.  Must be called by all TB threads of the CTA.
__device__ __forceinline__ double stream_rows(int64_t r0, int64_t n_rows, int64_t nnz, const int32_t *__restrict__ rowptr,
                                              const int32_t *__restrict__ col, const double *__restrict__ val,
                                              const double *__restrict__ diag, const double *__restrict__ x,
                                              double *s_prod, const double *__restrict__ xd = nullptr) {
    const int t = threadIdx.x;
    const int64_t r = r0 + t, rend = r0 + TB < n_rows ? r0 + TB : n_rows;
    const int kb = __ldg(rowptr + r0), ke = __ldg(rowptr + rend);
    int rs = 0, re = 0;
    double acc = 0.0;
    if (r < n_rows) {
        rs = __ldg(rowptr + r);
        re = __ldg(rowptr + r + 1);
        if (diag) acc = __ldg(diag + r) * __ldg((xd ? xd : x) + r);     // xd: x indexed like the rows when the columns
                                                                        // are relative to another base (batched systems)
    }
    const bool vec = ((reinterpret_cast<uintptr_t>(col) & 15) == 0) && ((reinterpret_cast<uintptr_t>(val) & 15) == 0);
    for (int kc = kb & ~3; kc < ke; kc += CAP) {
        const int kend = kc + CAP < ke ? kc + CAP : ke;
        for (int k = kc + 4 * t; k < kend; k += 4 * TB) {
            if (vec && (int64_t)k + 4 <= nnz) {
                const int4 c4 = __ldcs(reinterpret_cast<const int4 *>(col + k));
                const double2 v01 = __ldcs(reinterpret_cast<const double2 *>(val + k));
                const double2 v23 = __ldcs(reinterpret_cast<const double2 *>(val + k + 2));
                const double x0 = __ldg(x + c4.x), x1 = __ldg(x + c4.y), x2 = __ldg(x + c4.z), x3 = __ldg(x + c4.w);
                double2 p01, p23;
                p01.x = v01.x * x0; p01.y = v01.y * x1;
                p23.x = v23.x * x2; p23.y = v23.y * x3;
                *reinterpret_cast<double2 *>(s_prod + (k - kc)) = p01;
                *reinterpret_cast<double2 *>(s_prod + (k - kc) + 2) = p23;
            } else {
                for (int j = k; j < k + 4 && j < kend; j++) s_prod[j - kc] = __ldg(val + j) * __ldg(x + __ldg(col + j));
            }
        }
        __syncthreads();
        const int lo = rs > kc ? rs : kc, hi = re < kend ? re : kend;
        for (int k = lo; k < hi; k++) acc += s_prod[k - kc];
        __syncthreads();
    }
    return acc;
}

}  // namespace spmv
}  // namespace cnt
