// y = diag .* x + A x for NS value sets over one CSR pattern (off-diagonal entries, global columns):
// the SpMV of the PCG solvers (cnet.py:147-149 is replaced by them) as a stand-alone C-ABI entry point.
// Kernel: csrc/spmv_stream.cuh (coalesced 16-byte streaming of val / col through shared memory).
// The row-sharded solver (pcg_slab.cu) runs the same device function inside cnt_slab_spmv.
#include "spmv_stream.cuh"

namespace {
__global__ void __launch_bounds__(cnt::spmv::TB) k_spmv_stream(int64_t n_rows, int64_t nnz, const int32_t *__restrict__ rowptr,
                                                               const int32_t *__restrict__ col, const double *__restrict__ val,
                                                               int64_t val_stride, const double *__restrict__ diag,
                                                               int64_t diag_stride, const double *__restrict__ x,
                                                               int64_t x_stride, double *__restrict__ y, int64_t y_stride) {
    __shared__ __align__(16) double s_prod[cnt::spmv::CAP];
    const int q = blockIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * cnt::spmv::TB, r = r0 + threadIdx.x;
    const double acc = cnt::spmv::stream_rows(r0, n_rows, nnz, rowptr, col, val + (int64_t)q * val_stride,
                                              diag ? diag + (int64_t)q * diag_stride : nullptr, x + (int64_t)q * x_stride,
                                              s_prod);
    if (r < n_rows) y[(int64_t)q * y_stride + r] = acc;
}
}  // namespace

extern "C" int cnt_spmv_csr(int NS, int64_t n_rows, int64_t nnz, const int32_t *rowptr, const int32_t *col,
                            const double *val, int64_t val_stride, const double *diag, int64_t diag_stride,
                            const double *x, int64_t x_stride, double *y, int64_t y_stride, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(NS > 0 && NS <= 65535 && n_rows >= 0 && nnz >= 0 && nnz < (1ll << 31) && rowptr && x && y);
    CNT_CHECK_ARG(nnz == 0 || (col && val));
    if (n_rows == 0) return CNT_OK;
    k_spmv_stream<<<dim3(cnt::grid_for(n_rows, cnt::spmv::TB), (unsigned)NS, 1), cnt::spmv::TB, 0, st>>>(
        n_rows, nnz, rowptr, col, val, val_stride, diag, diag_stride, x, x_stride, y, y_stride);
    CNT_LAUNCHED();
    return CNT_OK;
}
