// Row-slab sharded two-level PCG for ONE very large network spread over several GPUs
// (BASELINE config 5; the reference has no counterpart -- it extends solve_mna, cnet.py:147-149).
//
// Every rank owns a contiguous range of the Z-ordered unknown rows (a compact spatial patch) of
// the same replicated system and keeps only its slice of the matrix and of the vectors.  The NS
// gate configurations of measure_fullnet are advanced together, so every exchange carries NS
// systems.  This file holds the per-rank phases of one iteration; the exchanges between them
// (torch.distributed over NCCL / NVLink, networksim_cntfet_b200/slab.py) are
//
//   spmv       q = A p (own rows; p of remote rows sits behind the own part of p)     -> all-reduce p.q
//   update     x += alpha p, r -= alpha q, aggregate sums of r                       -> all-reduce [r.r/d + local coarse
//                                                                                        part, r.r, sums of the
//                                                                                        aggregates that span ranks]
//   direction  z = r/d + einv[a] S[a], p = z + beta p, exported rows of p packed      -> all-gather of the exports
//
// r.z is never reduced on its own: r.z = sum r^2/d + sum_a einv_a S_a^2, the second sum splits into
// the aggregates that live on one rank (added to the local partial) and the shared ones, which every
// rank evaluates identically from the reduced S.  All partial sums are reduced in a fixed order
// (block partials, then one block per configuration), so a rank's arithmetic is deterministic.
// Stopping rule as in pcg.cu: recurrence residual ||r|| <= rtol ||b||.
#include "spmv_stream.cuh"

namespace {
constexpr int TB = 256;
static_assert(TB == cnt::spmv::TB, "k_slab_spmv: one row per thread of the streaming SpMV");

__device__ __forceinline__ const int32_t *row_i32(const int32_t *base, int q, int64_t stride) {
    return base + (int64_t)q * stride;
}

// r = rhs, x = 0, dinv = 1/diag; partials of r^2/d and r^2
__global__ void __launch_bounds__(TB) k_slab_init(cnt_slab_ctx c) {
    __shared__ double sm[32];
    const int q = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x, o = (int64_t)q * c.n_own + i;
    double a = 0.0, b = 0.0;
    if (i < c.n_own) {
        double d = c.diag[o], r = c.rhs[o], di = 1.0 / d;
        c.dinv[o] = di;
        c.r[o] = r;
        c.x[o] = 0.0;
        a = r * r * di;
        b = r * r;
    }
    a = cnt::block_sum(a, sm);
    b = cnt::block_sum(b, sm);
    if (threadIdx.x == 0) {
        c.part[((int64_t)q * 3 + 0) * c.nblk + blockIdx.x] = a;
        c.part[((int64_t)q * 3 + 1) * c.nblk + blockIdx.x] = b;
    }
}

// halo values of p from the gathered export buffers
__global__ void k_slab_unpack(cnt_slab_ctx c) {
    const int q = blockIdx.y;
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= c.n_halo) return;
    c.p[(int64_t)q * (c.n_own + c.n_halo) + c.n_own + h] = c.recvbuf[c.halo_src[h] + (int64_t)q * c.emax];
}

// q = A p on the own rows: the CTA streams the contiguous entry range of its 256 rows with coalesced 16-byte
// loads through shared memory (spmv_stream.cuh), thread t then sums row t in CSR order; block partial of p.q
__global__ void __launch_bounds__(TB) k_slab_spmv(cnt_slab_ctx c) {
    __shared__ double sm[32];
    __shared__ __align__(16) double s_prod[cnt::spmv::CAP];
    const int q = blockIdx.y;
    if (c.done[q]) {                                         // CTA-uniform
        if (threadIdx.x == 0) c.part[((int64_t)q * 3 + 2) * c.nblk + blockIdx.x] = 0.0;
        return;
    }
    const int64_t r0 = (int64_t)blockIdx.x * TB, i = r0 + threadIdx.x;
    const double *p = c.p + (int64_t)q * (c.n_own + c.n_halo);
    const double acc = cnt::spmv::stream_rows(r0, c.n_own, (int64_t)c.rowptr[c.n_own], c.rowptr, c.col,
                                              c.val + (int64_t)q * c.val_stride, c.diag + (int64_t)q * c.n_own, p, s_prod);
    double pq = 0.0;
    if (i < c.n_own) {
        c.q[(int64_t)q * c.n_own + i] = acc;
        pq = p[i] * acc;
    }
    pq = cnt::block_sum(pq, sm);
    if (threadIdx.x == 0) c.part[((int64_t)q * 3 + 2) * c.nblk + blockIdx.x] = pq;
}

// one block per configuration: fixed-order sum of the block partials of p.q
__global__ void __launch_bounds__(1024) k_slab_reduce_pq(cnt_slab_ctx c) {
    __shared__ double sm[32];
    const int q = blockIdx.x;
    const double *part = c.part + ((int64_t)q * 3 + 2) * c.nblk;
    double a = 0.0;
    for (int k = threadIdx.x; k < c.nblk; k += blockDim.x) a += part[k];
    a = cnt::block_sum(a, sm);
    if (threadIdx.x == 0) c.red_pq[q] = a;
}

// x += alpha p, r -= alpha q; partials of r^2/d and r^2
__global__ void __launch_bounds__(TB) k_slab_update(cnt_slab_ctx c) {
    __shared__ double sm[32];
    const int q = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x, o = (int64_t)q * c.n_own + i;
    double a = 0.0, b = 0.0;
    if (i < c.n_own) {
        double r = c.r[o];
        if (!c.done[q]) {
            const double pq = c.red_pq[q];
            const double alpha = pq > 0.0 ? c.scal[q * CNT_SLAB_NSCAL + CNT_SLAB_RZ] / pq : 0.0;
            c.x[o] += alpha * c.p[(int64_t)q * (c.n_own + c.n_halo) + i];
            r -= alpha * c.q[o];
            c.r[o] = r;
        }
        a = r * r * c.dinv[o];
        b = r * r;
    }
    a = cnt::block_sum(a, sm);
    b = cnt::block_sum(b, sm);
    if (threadIdx.x == 0) {
        c.part[((int64_t)q * 3 + 0) * c.nblk + blockIdx.x] = a;
        c.part[((int64_t)q * 3 + 1) * c.nblk + blockIdx.x] = b;
    }
}

// sums of r over the local members of every local aggregate: 8 lanes per aggregate (most have
// 2..10 members), fixed order
__global__ void __launch_bounds__(TB) k_slab_aggsum(cnt_slab_ctx c) {
    const int q = blockIdx.y;
    const int64_t a = ((int64_t)blockIdx.x * TB + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const int nl = c.nloc[q];
    double acc = 0.0;
    int k0 = 0, k1 = 0;
    if (a < nl) {
        const int32_t *ap = row_i32(c.aggptr, q, c.maxL + 1);
        k0 = ap[a];
        k1 = ap[a + 1];
        if (k1 - k0 > CNT_SLAB_BIGAGG) k1 = k0;          // left to k_slab_aggsum_big
    }
    const int32_t *mem = row_i32(c.aggmem, q, c.n_own);
    const double *r = c.r + (int64_t)q * c.n_own;
    for (int k = k0 + sub; k < k1; k += 8) acc += r[mem[k]];
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (a < nl && sub == 0 && k1 > k0) c.S[(int64_t)q * c.maxL + a] = acc;
}
// aggregates with more than CNT_SLAB_BIGAGG local members: one block each
__global__ void __launch_bounds__(TB) k_slab_aggsum_big(cnt_slab_ctx c) {
    __shared__ double sm[32];
    const int q = blockIdx.y;
    if ((int)blockIdx.x >= c.nbig[q]) return;
    const int a = row_i32(c.bigagg, q, c.maxBig)[blockIdx.x];
    const int32_t *ap = row_i32(c.aggptr, q, c.maxL + 1);
    const int32_t *mem = row_i32(c.aggmem, q, c.n_own);
    const double *r = c.r + (int64_t)q * c.n_own;
    double acc = 0.0;
    for (int k = ap[a] + threadIdx.x; k < ap[a + 1]; k += TB) acc += r[mem[k]];
    acc = cnt::block_sum(acc, sm);
    if (threadIdx.x == 0) c.S[(int64_t)q * c.maxL + a] = acc;
}

// one block per configuration: red2[q] = [sum r^2/d + sum over the aggregates that live on this
// rank only of einv S^2, sum r^2, partial sums of the shared aggregates (0 where the rank holds none)]
__global__ void __launch_bounds__(1024) k_slab_reduce2(cnt_slab_ctx c) {
    __shared__ double sm[32];
    const int q = blockIdx.x;
    const double *pa = c.part + ((int64_t)q * 3 + 0) * c.nblk, *pb = c.part + ((int64_t)q * 3 + 1) * c.nblk;
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < c.nblk; k += blockDim.x) {
        a += pa[k];
        b += pb[k];
    }
    const int nl = c.nloc[q];
    const double *S = c.S + (int64_t)q * c.maxL, *einv = c.einv + (int64_t)q * c.maxL;
    const int32_t *ash = row_i32(c.agg_shared, q, c.maxL);
    double e = 0.0;
    for (int k = threadIdx.x; k < nl; k += blockDim.x)
        if (ash[k] < 0) e += einv[k] * S[k] * S[k];
    a = cnt::block_sum(a, sm);
    b = cnt::block_sum(b, sm);
    e = cnt::block_sum(e, sm);
    double *out = c.red2 + (int64_t)q * (2 + c.maxSh);
    if (threadIdx.x == 0) {
        out[0] = a + e;
        out[1] = b;
    }
    const int32_t *sl = row_i32(c.shared_local, q, c.maxSh);
    for (int s = threadIdx.x; s < c.maxSh; s += blockDim.x) out[2 + s] = (s < c.nsh[q] && sl[s] >= 0) ? S[sl[s]] : 0.0;
}

// after the all-reduce of red2: r.z, beta, convergence; the reduced sums of the shared aggregates
// go back into S.  first = 1: start of the solve (b.b = r.r, no beta)
__global__ void __launch_bounds__(1024) k_slab_scal(cnt_slab_ctx c, int first) {
    __shared__ double sm[32];
    const int q = blockIdx.x;
    const double *in = c.red2 + (int64_t)q * (2 + c.maxSh);
    const int ns = c.nsh[q];
    const double *es = c.einv_sh + (int64_t)q * c.maxSh;
    const int32_t *sl = row_i32(c.shared_local, q, c.maxSh);
    double *S = c.S + (int64_t)q * c.maxL;
    double e = 0.0;
    for (int s = threadIdx.x; s < ns; s += blockDim.x) {
        double v = in[2 + s];
        e += es[s] * v * v;
        if (sl[s] >= 0) S[sl[s]] = v;
    }
    e = cnt::block_sum(e, sm);
    if (threadIdx.x != 0) return;
    double *sc = c.scal + q * CNT_SLAB_NSCAL;
    const double rz = in[0] + e, rr = in[1];
    if (first) {
        sc[CNT_SLAB_RZ] = rz;
        sc[CNT_SLAB_BB] = rr;
        sc[CNT_SLAB_RR] = rr;
        sc[CNT_SLAB_BETA] = 0.0;
        sc[CNT_SLAB_ITERS] = 0.0;
        sc[CNT_SLAB_STATUS] = 0.0;
        c.done[q] = rr == 0.0 ? 1 : 0;          // zero right-hand side: x = 0
        return;
    }
    if (c.done[q]) {
        sc[CNT_SLAB_BETA] = 0.0;
        return;
    }
    if (!(c.red_pq[q] > 0.0)) {       // breakdown (matrix not SPD or NaN): x, r were left alone
        c.done[q] = 1;
        sc[CNT_SLAB_STATUS] = 2.0;
        sc[CNT_SLAB_BETA] = 0.0;
        return;
    }
    sc[CNT_SLAB_BETA] = rz / sc[CNT_SLAB_RZ];
    sc[CNT_SLAB_RZ] = rz;
    sc[CNT_SLAB_RR] = rr;
    sc[CNT_SLAB_ITERS] += 1.0;
    if (rr <= c.rtol * c.rtol * sc[CNT_SLAB_BB]) {
        c.done[q] = 1;
        sc[CNT_SLAB_STATUS] = 0.0;
    } else if (sc[CNT_SLAB_ITERS] >= (double)c.maxit) {
        c.done[q] = 1;
        sc[CNT_SLAB_STATUS] = 1.0;
    } else if (!(rr == rr)) {
        c.done[q] = 1;
        sc[CNT_SLAB_STATUS] = 2.0;
    }
}

// z = r/d + einv[a] S[a]; p = z + beta p (own rows)
__global__ void __launch_bounds__(TB) k_slab_direction(cnt_slab_ctx c) {
    const int q = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x, o = (int64_t)q * c.n_own + i;
    if (i >= c.n_own) return;
    const int a = row_i32(c.row_agg, q, c.n_own)[i];
    double z = c.r[o] * c.dinv[o];
    if (a >= 0) z += c.einv[(int64_t)q * c.maxL + a] * c.S[(int64_t)q * c.maxL + a];
    double *p = c.p + (int64_t)q * (c.n_own + c.n_halo) + i;
    const double beta = c.scal[q * CNT_SLAB_NSCAL + CNT_SLAB_BETA];
    *p = beta != 0.0 ? z + beta * *p : z;
}

// exported rows of p -> send buffer
__global__ void k_slab_pack(cnt_slab_ctx c) {
    const int q = blockIdx.y;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= c.n_exp) return;
    c.sendbuf[(int64_t)q * c.emax + e] = c.p[(int64_t)q * (c.n_own + c.n_halo) + c.exp_rows[e]];
}

int check_ctx(const cnt_slab_ctx *c) {
    CNT_CHECK_ARG(c && c->NS > 0 && c->n_own >= 0 && c->n_halo >= 0 && c->nblk == (int)cnt::grid_for(c->n_own > 0 ? c->n_own : 1, TB));
    CNT_CHECK_ARG(c->rowptr && c->diag && c->rhs && c->x && c->r && c->q && c->dinv && c->p && c->part && c->red_pq &&
                  c->red2 && c->scal && c->done && c->S && c->nloc && c->nsh && c->nbig);
    CNT_CHECK_ARG(c->maxL >= 1 && c->maxSh >= 1 && c->maxBig >= 1 && c->emax >= 1 && c->rtol > 0);
    return CNT_OK;
}
dim3 grid_rows(const cnt_slab_ctx *c) { return dim3((unsigned)c->nblk, (unsigned)c->NS, 1); }

int launch_aggsums(const cnt_slab_ctx *c, cudaStream_t st) {
    k_slab_aggsum<<<dim3(cnt::grid_for(c->maxL * 8, TB), c->NS, 1), TB, 0, st>>>(*c);
    CNT_LAUNCHED();
    if (c->max_nbig > 0) {
        k_slab_aggsum_big<<<dim3((unsigned)c->max_nbig, c->NS, 1), TB, 0, st>>>(*c);
        CNT_LAUNCHED();
    }
    k_slab_reduce2<<<c->NS, 1024, 0, st>>>(*c);
    CNT_LAUNCHED();
    return CNT_OK;
}
int launch_direction(const cnt_slab_ctx *c, int first, cudaStream_t st) {
    k_slab_scal<<<c->NS, 1024, 0, st>>>(*c, first);
    CNT_LAUNCHED();
    k_slab_direction<<<grid_rows(c), TB, 0, st>>>(*c);
    CNT_LAUNCHED();
    if (c->n_exp > 0) {
        k_slab_pack<<<dim3(cnt::grid_for(c->n_exp, TB), c->NS, 1), TB, 0, st>>>(*c);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}
}  // namespace

extern "C" int cnt_slab_block_rows(void) { return TB; }

// start: r = b, x = 0, aggregate sums -> red2 (to be all-reduced)
extern "C" int cnt_slab_init(const cnt_slab_ctx *c, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_ctx(c);
    if (rc) return rc;
    CNT_CUDA(cudaMemsetAsync(c->done, 0, sizeof(int32_t) * c->NS, st));
    CNT_CUDA(cudaMemsetAsync(c->p, 0, sizeof(double) * c->NS * (c->n_own + c->n_halo), st));
    k_slab_init<<<grid_rows(c), TB, 0, st>>>(*c);
    CNT_LAUNCHED();
    return launch_aggsums(c, st);
}
// after the all-reduce of red2 (first = 1 right after cnt_slab_init): scalars, z, p, exports
// packed into sendbuf (to be all-gathered)
extern "C" int cnt_slab_direction(const cnt_slab_ctx *c, int first, void *stream) {
    int rc = check_ctx(c);
    if (rc) return rc;
    return launch_direction(c, first, (cudaStream_t)stream);
}
// after the all-gather: halo of p, q = A p, red_pq (to be all-reduced)
extern "C" int cnt_slab_spmv(const cnt_slab_ctx *c, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_ctx(c);
    if (rc) return rc;
    if (c->n_halo > 0) {
        CNT_CHECK_ARG(c->recvbuf && c->halo_src);
        k_slab_unpack<<<dim3(cnt::grid_for(c->n_halo, TB), c->NS, 1), TB, 0, st>>>(*c);
        CNT_LAUNCHED();
    }
    k_slab_spmv<<<grid_rows(c), TB, 0, st>>>(*c);
    CNT_LAUNCHED();
    k_slab_reduce_pq<<<c->NS, 1024, 0, st>>>(*c);
    CNT_LAUNCHED();
    return CNT_OK;
}
// after the all-reduce of red_pq: x, r, aggregate sums -> red2 (to be all-reduced)
extern "C" int cnt_slab_update(const cnt_slab_ctx *c, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_ctx(c);
    if (rc) return rc;
    k_slab_update<<<grid_rows(c), TB, 0, st>>>(*c);
    CNT_LAUNCHED();
    return launch_aggsums(c, st);
}
