// Numeric phase of the star-mesh direct solver (networksim_cntfet_b200/starmesh.py) -- replaces
// solve_mna (cnet.py:147-149) by sparse Cholesky written on the conductances:
//
//   eliminate v (neighbours u_i, conductances g_i, electrode coupling c_v):
//     d_v = c_v + sum_i g_i;   g(u_i,u_j) += g_i g_j / d_v;   c_ui += g_i c_v / d_v;   b_ui += g_i b_v / d_v
//   back-substitution:  x_v = (b_v + sum_i g_i x_ui) / d_v
//
// Every quantity is a sum of positive terms (no cancellation -> componentwise accurate).  The
// symbolic phase eliminates an independent set of unknowns per round and hands over index lists;
// each kernel below is a fixed-order segmented sum over such a list (deterministic), for all NS
// gate configurations at once (grid.y).
#include "common.cuh"

namespace {
constexpr int TB = 128;

// edges of round 0: sums of the CSR entries (-val) of every stick pair
__global__ void k_sm_edges0(int64_t nE, const int32_t *__restrict__ ptr, const int32_t *__restrict__ src,
                            const double *__restrict__ val, int64_t val_stride, double *__restrict__ g, int64_t g_stride) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    const double *v = val + (int64_t)blockIdx.y * val_stride;
    double s = 0.0;
    for (int k = ptr[e]; k < ptr[e + 1]; k++) s += -v[src[k]];
    g[(int64_t)blockIdx.y * g_stride + e] = s;
}

// electrode coupling of every unknown: c = gs + gd; with leak != 0 the rounding error of the
// assembled FP64 diagonal (the same sequential sum as k_values: entries in CSR order, + gs, + gd)
// is added, so that the direct solve answers the SAME floating-point matrix the iterative solvers
// and the reference see; the errors of the additions are exact (TwoSum)
__global__ void k_sm_coupling(int64_t R, const int32_t *__restrict__ rowptr, const double *__restrict__ val,
                              int64_t val_stride, const double *__restrict__ gs, const double *__restrict__ gd,
                              int leak, double *__restrict__ c) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int q = blockIdx.y;
    const double *v = val + (int64_t)q * val_stride;
    const double s_ = gs[(int64_t)q * R + r], d_ = gd[(int64_t)q * R + r];
    double out = __dadd_rn(s_, d_);
    if (leak) {
        double sum = 0.0, err = 0.0;
        auto add = [&](double x) {
            double t = __dadd_rn(sum, x);
            double bb = __dsub_rn(t, sum);
            err = __dadd_rn(err, __dadd_rn(__dsub_rn(sum, __dsub_rn(t, bb)), __dsub_rn(x, bb)));
            sum = t;
        };
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++) add(-v[k]);
        add(s_);
        add(d_);
        out = __dsub_rn(out, err);        // exact sum = sum + err  ->  fl - exact = -err
    }
    c[(int64_t)q * R + r] = out;
}

struct RoundArgs {
    int64_t nelim, nE_new, ntarget, R, ebase, gbase;
    int64_t g_stride, pool_e_stride, pool_g_stride;
    const int32_t *E, *nptr, *nbr, *eslot, *copy_src, *cptr, *cm, *ca, *cb, *tn, *tptr, *tm, *tslot;
    const double *g_old;
    double *g_new, *c, *b, *x, *dinv, *gE;
};

__global__ void k_sm_pivot(RoundArgs a) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.nelim) return;
    const int q = blockIdx.y;
    const double *g = a.g_old + (int64_t)q * a.g_stride;
    double *gE = a.gE + (int64_t)q * a.pool_g_stride + a.gbase;
    double d = a.c[(int64_t)q * a.R + a.E[m]];
    for (int k = a.nptr[m]; k < a.nptr[m + 1]; k++) {
        double ge = g[a.eslot[k]];
        gE[k] = ge;
        d += ge;
    }
    a.dinv[(int64_t)q * a.pool_e_stride + a.ebase + m] = 1.0 / d;
}
__global__ void k_sm_newedges(RoundArgs a) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.nE_new) return;
    const int q = blockIdx.y;
    const double *g = a.g_old + (int64_t)q * a.g_stride;
    const double *dinv = a.dinv + (int64_t)q * a.pool_e_stride + a.ebase;
    const int cs = a.copy_src[e];
    double v = cs >= 0 ? g[cs] : 0.0;
    for (int k = a.cptr[e]; k < a.cptr[e + 1]; k++) v += g[a.ca[k]] * g[a.cb[k]] * dinv[a.cm[k]];
    a.g_new[(int64_t)q * a.g_stride + e] = v;
}
__global__ void k_sm_targets(RoundArgs a) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.ntarget) return;
    const int q = blockIdx.y;
    const double *g = a.g_old + (int64_t)q * a.g_stride;
    const double *dinv = a.dinv + (int64_t)q * a.pool_e_stride + a.ebase;
    double *c = a.c + (int64_t)q * a.R, *b = a.b + (int64_t)q * a.R;
    double sc = 0.0, sb = 0.0;
    for (int k = a.tptr[t]; k < a.tptr[t + 1]; k++) {
        const int m = a.tm[k];
        const double w = g[a.tslot[k]] * dinv[m];
        const int v = a.E[m];
        sc += w * c[v];
        sb += w * b[v];
    }
    const int u = a.tn[t];
    c[u] += sc;
    b[u] += sb;
}
__global__ void k_sm_back(RoundArgs a) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.nelim) return;
    const int q = blockIdx.y;
    const double *gE = a.gE + (int64_t)q * a.pool_g_stride + a.gbase;
    double *x = a.x + (int64_t)q * a.R;
    const int v = a.E[m];
    double s = a.b[(int64_t)q * a.R + v];
    for (int k = a.nptr[m]; k < a.nptr[m + 1]; k++) s += gE[k] * x[a.nbr[k]];
    x[v] = s * a.dinv[(int64_t)q * a.pool_e_stride + a.ebase + m];
}

// ---- dense tail: once every network is down to a few hundred unknowns (the dense core that the
// rounds would take one handful at a time), one CTA per (network, configuration) finishes the same
// star-mesh elimination on a dense upper-triangular conductance matrix G [n x n] in global memory
// (L2 resident), in ascending row order, and back-substitutes.  A warp takes one neighbour row of the
// pivot at a time; rows without a junction to the pivot are skipped, so the cost follows the actual
// degrees.  Same formulas as the sparse rounds: sums of positive terms only.
constexpr int DENSE_TB = 256;
__global__ void k_sm_dense_fill(int64_t nE, const long long *__restrict__ epos, const double *__restrict__ g,
                                int64_t g_stride, double *__restrict__ G, int64_t gpool) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    G[(int64_t)blockIdx.y * gpool + epos[e]] = g[(int64_t)blockIdx.y * g_stride + e];
}
__global__ void __launch_bounds__(DENSE_TB) k_sm_dense(int64_t R, const long long *__restrict__ doff,
                                                       const int32_t *__restrict__ dptr, const int32_t *__restrict__ dnodes,
                                                       double *__restrict__ Gall, int64_t gpool, const double *__restrict__ c_in,
                                                       const double *__restrict__ b_in, double *__restrict__ x_out, int nmax) {
    extern __shared__ double smd[];
    double *rowk = smd, *cc = smd + nmax, *bb = smd + 2 * nmax, *dinv = smd + 3 * nmax, *red = smd + 4 * nmax;
    const int net = blockIdx.x, q = blockIdx.y;
    const int n = dptr[net + 1] - dptr[net];
    if (n <= 0) return;
    const int32_t *nodes = dnodes + dptr[net];
    double *G = Gall + (int64_t)q * gpool + doff[net];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = DENSE_TB >> 5;
    for (int i = t; i < n; i += DENSE_TB) {
        cc[i] = c_in[(int64_t)q * R + nodes[i]];
        bb[i] = b_in[(int64_t)q * R + nodes[i]];
    }
    __syncthreads();
    for (int k = 0; k < n; k++) {
        // pivot row (upper part) into shared memory, d_k = c_k + its sum (fixed order)
        double part = 0.0;
        for (int j = k + 1 + t; j < n; j += DENSE_TB) {
            double v = G[(int64_t)k * n + j];
            rowk[j] = v;
            part += v;
        }
        part = cnt::block_sum(part, red);
        if (t == 0) {
            double d = cc[k] + part;
            dinv[k] = 1.0 / d;
        }
        __syncthreads();
        const double di = dinv[k], ck = cc[k], bk = bb[k];
        for (int i = k + 1 + w; i < n; i += nw) {          // one warp per neighbour row
            const double gi = rowk[i];
            if (gi == 0.0) continue;
            const double f = gi * di;
            for (int j = i + 1 + lane; j < n; j += 32) {
                const double gj = rowk[j];
                if (gj != 0.0) G[(int64_t)i * n + j] += f * gj;
            }
            if (lane == 0) {
                cc[i] += f * ck;
                bb[i] += f * bk;
            }
        }
        __syncthreads();
    }
    // back-substitution: x_k = (b_k + sum_{j>k} G[k][j] x_j) / d_k; x kept in rowk
    for (int k = n - 1; k >= 0; k--) {
        double part = 0.0;
        for (int j = k + 1 + t; j < n; j += DENSE_TB) part += G[(int64_t)k * n + j] * rowk[j];
        part = cnt::block_sum(part, red);
        if (t == 0) rowk[k] = (bb[k] + part) * dinv[k];
        __syncthreads();
    }
    for (int i = t; i < n; i += DENSE_TB) x_out[(int64_t)q * R + nodes[i]] = rowk[i];
}

RoundArgs make_args(const cnt_sm_round *r, int64_t R, int64_t g_stride, int64_t pool_e_stride, int64_t pool_g_stride,
                    const double *g_old, double *g_new, double *c, double *b, double *x, double *dinv, double *gE) {
    RoundArgs a;
    a.nelim = r->nelim; a.nE_new = r->nE_new; a.ntarget = r->ntarget; a.R = R; a.ebase = r->ebase; a.gbase = r->gbase;
    a.g_stride = g_stride; a.pool_e_stride = pool_e_stride; a.pool_g_stride = pool_g_stride;
    a.E = r->E; a.nptr = r->nptr; a.nbr = r->nbr; a.eslot = r->eslot; a.copy_src = r->copy_src; a.cptr = r->cptr;
    a.cm = r->cm; a.ca = r->ca; a.cb = r->cb; a.tn = r->tn; a.tptr = r->tptr; a.tm = r->tm; a.tslot = r->tslot;
    a.g_old = g_old; a.g_new = g_new; a.c = c; a.b = b; a.x = x; a.dinv = dinv; a.gE = gE;
    return a;
}
}  // namespace

extern "C" int cnt_sm_edges0(int NS, int64_t nE, const int32_t *ptr, const int32_t *src, const double *val,
                             int64_t val_stride, double *g, int64_t g_stride, void *stream) {
    CNT_CHECK_ARG(NS > 0 && nE >= 0 && g);
    if (nE > 0) {
        CNT_CHECK_ARG(ptr && src && val);
        k_sm_edges0<<<dim3(cnt::grid_for(nE, TB), NS, 1), TB, 0, (cudaStream_t)stream>>>(nE, ptr, src, val, val_stride, g,
                                                                                      g_stride);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}

extern "C" int cnt_sm_coupling(int NS, int64_t R, const int32_t *rowptr, const double *val, int64_t val_stride,
                               const double *gs, const double *gd, int leak, double *c, void *stream) {
    CNT_CHECK_ARG(NS > 0 && R >= 0 && c);
    if (R > 0) {
        CNT_CHECK_ARG(rowptr && gs && gd && (val || !leak));
        k_sm_coupling<<<dim3(cnt::grid_for(R, TB), NS, 1), TB, 0, (cudaStream_t)stream>>>(R, rowptr, val, val_stride, gs, gd,
                                                                                       leak, c);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}

// forward elimination of all rounds: g0 [NS, g_stride] holds the round-0 edge values and is used as
// one half of a ping-pong pair with g1; c, b [NS, R] are updated in place; dinv [NS, pool_e_stride]
// and gE [NS, pool_g_stride] keep what the back-substitution needs
extern "C" int cnt_sm_forward(int NS, int nrounds, const cnt_sm_round *rounds, int64_t R, double *g0, double *g1,
                              int64_t g_stride, double *c, double *b, double *dinv, int64_t pool_e_stride, double *gE,
                              int64_t pool_g_stride, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(NS > 0 && nrounds >= 0 && (nrounds == 0 || rounds) && g0 && g1 && c && b && dinv && gE);
    double *cur = g0, *nxt = g1;
    for (int k = 0; k < nrounds; k++) {
        const cnt_sm_round *r = rounds + k;
        CNT_CHECK_ARG(r->nelim > 0 && r->nE_new <= g_stride && r->ebase + r->nelim <= pool_e_stride);
        RoundArgs a = make_args(r, R, g_stride, pool_e_stride, pool_g_stride, cur, nxt, c, b, nullptr, dinv, gE);
        k_sm_pivot<<<dim3(cnt::grid_for(r->nelim, TB), NS, 1), TB, 0, st>>>(a);
        CNT_LAUNCHED();
        if (r->nE_new > 0) {
            k_sm_newedges<<<dim3(cnt::grid_for(r->nE_new, TB), NS, 1), TB, 0, st>>>(a);
            CNT_LAUNCHED();
        }
        if (r->ntarget > 0) {
            k_sm_targets<<<dim3(cnt::grid_for(r->ntarget, TB), NS, 1), TB, 0, st>>>(a);
            CNT_LAUNCHED();
        }
        double *t = cur; cur = nxt; nxt = t;
    }
    return CNT_OK;
}

extern "C" int cnt_sm_backward(int NS, int nrounds, const cnt_sm_round *rounds, int64_t R, const double *b,
                               const double *dinv, int64_t pool_e_stride, const double *gE, int64_t pool_g_stride,
                               double *x, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(NS > 0 && nrounds >= 0 && (nrounds == 0 || rounds) && b && dinv && gE && x);
    for (int k = nrounds - 1; k >= 0; k--) {
        const cnt_sm_round *r = rounds + k;
        RoundArgs a = make_args(r, R, 0, pool_e_stride, pool_g_stride, nullptr, nullptr, nullptr, const_cast<double *>(b), x,
                                const_cast<double *>(dinv), const_cast<double *>(gE));
        k_sm_back<<<dim3(cnt::grid_for(r->nelim, TB), NS, 1), TB, 0, st>>>(a);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}

// dense tail (see k_sm_dense): g = edge values after the last sparse round; epos[e] = position of edge e in
// the pool of dense matrices; c, b as left by the sparse rounds; writes x of the remaining unknowns
extern "C" int cnt_sm_dense(int NS, int64_t R, int nnet, const int64_t *doff, const int32_t *dptr, const int32_t *dnodes,
                            int64_t nE, const int64_t *epos, const double *g, int64_t g_stride, double *G, int64_t gpool,
                            const double *c, const double *b, double *x, int nmax, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(NS > 0 && nnet > 0 && doff && dptr && dnodes && G && c && b && x && nmax > 0 && nmax <= 1400 && gpool > 0);
    CNT_CUDA(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)NS * (size_t)gpool, st));
    if (nE > 0) {
        CNT_CHECK_ARG(epos && g);
        k_sm_dense_fill<<<dim3(cnt::grid_for(nE, TB), NS, 1), TB, 0, st>>>(nE, (const long long *)epos, g, g_stride, G, gpool);
        CNT_LAUNCHED();
    }
    size_t smem = sizeof(double) * (4 * (size_t)nmax + 32);
    k_sm_dense<<<dim3(nnet, NS, 1), DENSE_TB, smem, st>>>(R, (const long long *)doff, dptr, dnodes, G, gpool, c, b, x, nmax);
    CNT_LAUNCHED();
    return CNT_OK;
}
