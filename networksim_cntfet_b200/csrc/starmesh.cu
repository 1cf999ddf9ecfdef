// Numeric phase of the star-mesh direct solver -- replaces solve_mna (cnet.py:147-149) by sparse
// Cholesky written on the conductances:
//
//   eliminate v (neighbours u_i, conductances g_i, electrode coupling c_v):
//     d_v = c_v + sum_i g_i;   g(u_i,u_j) += g_i g_j / d_v;   c_ui += g_i c_v / d_v;   b_ui += g_i b_v / d_v
//   back-substitution:  x_v = (b_v + sum_i g_i x_ui) / d_v
//
// Every quantity is a sum of positive terms (no cancellation -> componentwise accurate).  The
// symbolic phase (starmesh_sym.cu) eliminates an independent set of unknowns per round and hands
// over index lists over edge slots that never move: the value of an edge lives in g[slot] from its
// creation to the end (a dead edge keeps the value the back-substitution needs, so there is no
// copy of the factor), fill edges start at zero.  Each kernel below is a fixed-order segmented sum
// over such a list (deterministic), for all NS gate configurations at once (grid.y).
#include "starmesh_plan.cuh"

namespace {
using cnt::sm::Plan;
constexpr int TB = 128;

// values of the round-0 edges: sums of the CSR entries (-val) of every stick pair; fill slots = 0
__global__ void k_sm_edges0(int nE0, int nslots, const int32_t *__restrict__ ptr, const int32_t *__restrict__ src,
                            const double *__restrict__ val, int64_t val_stride, double *__restrict__ g,
                            int64_t g_stride) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nslots) return;
    double s = 0.0;
    if (e < nE0) {
        const double *v = val + (int64_t)blockIdx.y * val_stride;
        for (int k = ptr[e]; k < ptr[e + 1]; k++) s += -v[src[k]];
    }
    g[(int64_t)blockIdx.y * g_stride + e] = s;
}

// electrode coupling of every unknown: c = gs + gd (+ extra); with leak != 0 the rounding error of the
// assembled FP64 diagonal (the same sequential sum as k_values: entries in CSR order, + gs, + gd)
// is added, so that the direct solve answers the SAME floating-point matrix the iterative solvers
// and the reference see; the errors of the additions are exact (TwoSum).  Also b = rhs.
__global__ void k_sm_coupling(int R, const int32_t *__restrict__ rowptr, const double *__restrict__ val,
                              int64_t val_stride, const int32_t *__restrict__ src_eid,
                              const int32_t *__restrict__ drn_eid, const double *__restrict__ gj, int64_t g_stride,
                              const double *__restrict__ extra, const double *__restrict__ rhs, int64_t rhs_stride,
                              int leak, double *__restrict__ c, double *__restrict__ b) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int q = blockIdx.y;
    const double *v = val + (int64_t)q * val_stride;
    const double *gq = gj ? gj + (int64_t)q * g_stride : nullptr;
    const int se = src_eid ? src_eid[r] : -1, de = drn_eid ? drn_eid[r] : -1;
    const double s_ = se >= 0 ? gq[se] : 0.0, d_ = de >= 0 ? gq[de] : 0.0;
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
    if (extra) out += extra[(int64_t)q * R + r];
    c[(int64_t)q * R + r] = out;
    b[(int64_t)q * R + r] = rhs[(int64_t)q * rhs_stride + r];
}

struct Num {
    int R;
    int64_t g_stride;
    const int32_t *piv, *nptr, *nbr, *eslot, *tgt, *cptr, *cm, *ca, *cb, *tn, *tptr, *tm, *tslot;
    double *g, *c, *b, *x, *dinv;
};

// pivots of a round: d = c + sum of the incident conductances (neighbours ascending)
__global__ void k_sm_pivot(Num a, int ebase, int nelim) {
    int ml = blockIdx.x * blockDim.x + threadIdx.x;
    if (ml >= nelim) return;
    const int q = blockIdx.y, m = ebase + ml;
    const double *g = a.g + (int64_t)q * a.g_stride;
    double d = a.c[(int64_t)q * a.R + a.piv[m]];
    for (int k = a.nptr[m]; k < a.nptr[m + 1]; k++) d += g[a.eslot[k]];
    a.dinv[(int64_t)q * a.R + m] = 1.0 / d;
}
// targets of a round: first the edge groups (g[tgt] += sum g[ca] g[cb] / d[cm]), then the node groups
// (c, b of a neighbour += sum g[tslot] / d[tm] * {c, b}[piv[tm]]); contributions ordered by pivot
__global__ void k_sm_update(Num a, int gebase, int nge, int gnbase, int ngn) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nge + ngn) return;
    const int q = blockIdx.y;
    double *g = a.g + (int64_t)q * a.g_stride;
    const double *dinv = a.dinv + (int64_t)q * a.R;
    if (t < nge) {
        const int gi = gebase + t;
        double v = 0.0;
        for (int k = a.cptr[gi]; k < a.cptr[gi + 1]; k++) v += g[a.ca[k]] * g[a.cb[k]] * dinv[a.cm[k]];
        g[a.tgt[gi]] += v;
    } else {
        const int gi = gnbase + (t - nge);
        double *c = a.c + (int64_t)q * a.R, *b = a.b + (int64_t)q * a.R;
        double sc = 0.0, sb = 0.0;
        for (int k = a.tptr[gi]; k < a.tptr[gi + 1]; k++) {
            const int m = a.tm[k];
            const double w = g[a.tslot[k]] * dinv[m];
            const int v = a.piv[m];
            sc += w * c[v];
            sb += w * b[v];
        }
        const int u = a.tn[gi];
        c[u] += sc;
        b[u] += sb;
    }
}
__global__ void k_sm_back(Num a, int ebase, int nelim) {
    int ml = blockIdx.x * blockDim.x + threadIdx.x;
    if (ml >= nelim) return;
    const int q = blockIdx.y, m = ebase + ml;
    const double *g = a.g + (int64_t)q * a.g_stride;
    double *x = a.x + (int64_t)q * a.R;
    const int v = a.piv[m];
    double s = a.b[(int64_t)q * a.R + v];
    for (int k = a.nptr[m]; k < a.nptr[m + 1]; k++) s += g[a.eslot[k]] * x[a.nbr[k]];
    x[v] = s * a.dinv[(int64_t)q * a.R + m];
}

// ---- dense tail: once a network is down to a few hundred unknowns (the dense core that the
// rounds would take one handful at a time), one CTA per (network, configuration) finishes the same
// star-mesh elimination on a dense upper-triangular conductance matrix G [n x n] in global memory
// (L2 resident), in ascending row order, and back-substitutes.  A warp takes one neighbour row of the
// pivot at a time; rows without a junction to the pivot are skipped, so the cost follows the actual
// degrees.  Same formulas as the sparse rounds: sums of positive terms only.
constexpr int DENSE_TB = 256;
__global__ void k_sm_dense_fill(int nE, const int32_t *__restrict__ dslot, const long long *__restrict__ depos,
                                const double *__restrict__ g, int64_t g_stride, double *__restrict__ G, int64_t gpool) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    G[(int64_t)blockIdx.y * gpool + depos[e]] = g[(int64_t)blockIdx.y * g_stride + dslot[e]];
}
__global__ void __launch_bounds__(DENSE_TB) k_sm_dense(int R, const long long *__restrict__ doff,
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

struct NumWs {
    int64_t g, c, b, dinv, G, total;
};
NumWs num_layout(int NS, const cnt_sm_info *h) {
    NumWs L;
    int64_t pos = 0;
    auto take = [&](int64_t count) {
        int64_t at = pos;
        pos += cnt::align_up((count > 0 ? count : 1) * 8, 256);
        return at;
    };
    L.g = take((int64_t)NS * (h->nslots > 0 ? h->nslots : 1));
    L.c = take((int64_t)NS * h->R);
    L.b = take((int64_t)NS * h->R);
    L.dinv = take((int64_t)NS * h->R);
    L.G = take((int64_t)NS * (h->dense_gpool > 0 ? h->dense_gpool : 1));
    L.total = pos;
    return L;
}
}  // namespace

extern "C" int64_t cnt_sm_solve_workspace_bytes(int NS, const cnt_sm_info *h_info) {
    if (!h_info || NS <= 0) return 0;
    return num_layout(NS, h_info).total;
}

extern "C" int cnt_sm_solve(int NS, int B, int64_t R, int64_t NNZ, const cnt_sm_info *h, const void *plan,
                            int64_t cap_slots, int64_t cap_contrib, const int32_t *rowptr, const double *val,
                            int64_t val_stride, const int32_t *src_eid, const int32_t *drn_eid, const double *gj,
                            int64_t g_stride, const double *coupling, const double *rhs, int64_t rhs_stride, int leak,
                            double *x, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(NS > 0 && NS <= 65535 && B > 0 && R > 0 && h && plan && rowptr && val && rhs && x && ws);
    CNT_CHECK_ARG(h->done && !h->error && h->R == R && h->nrounds >= 0 && h->nrounds < CNT_SM_MAX_ROUNDS);
    CNT_CHECK_ARG((!src_eid && !drn_eid) || gj);
    CNT_CHECK_ARG(h->dense_nmax <= 1400);
    int64_t off[CNT_SM_NARR + 1];
    cnt::sm::plan_layout(R, B, NNZ, cap_slots, cap_contrib, off);
    const Plan p = cnt::sm::plan_view(const_cast<void *>(plan), off);
    const NumWs L = num_layout(NS, h);
    if (L.total > ws_bytes) {
        cnt::set_error("star-mesh numeric workspace too small");
        return CNT_ECAPACITY;
    }
    char *w = (char *)ws;
    const int64_t gs = h->nslots > 0 ? h->nslots : 1;
    Num a;
    a.R = (int)R;
    a.g_stride = gs;
    a.piv = p.piv; a.nptr = p.nptr; a.nbr = p.nbr; a.eslot = p.eslot; a.tgt = p.tgt; a.cptr = p.cptr; a.cm = p.cm;
    a.ca = p.ca; a.cb = p.cb; a.tn = p.tn; a.tptr = p.tptr; a.tm = p.tm; a.tslot = p.tslot;
    a.g = (double *)(w + L.g); a.c = (double *)(w + L.c); a.b = (double *)(w + L.b); a.x = x;
    a.dinv = (double *)(w + L.dinv);
    double *G = (double *)(w + L.G);

    if (h->nslots > 0) {
        k_sm_edges0<<<dim3(cnt::grid_for(h->nslots, 256), NS, 1), 256, 0, st>>>(h->nE0, h->nslots, p.e0_ptr, p.e0_src, val,
                                                                              val_stride, a.g, gs);
        CNT_LAUNCHED();
    }
    k_sm_coupling<<<dim3(cnt::grid_for(R, TB), NS, 1), TB, 0, st>>>((int)R, rowptr, val, val_stride, src_eid, drn_eid, gj,
                                                                  g_stride, coupling, rhs, rhs_stride, leak, a.c, a.b);
    CNT_LAUNCHED();
    for (int k = 0; k < h->nrounds; k++) {
        const cnt_sm_round *r = &h->rounds[k];
        k_sm_pivot<<<dim3(cnt::grid_for(r->nelim, TB), NS, 1), TB, 0, st>>>(a, r->ebase, r->nelim);
        CNT_LAUNCHED();
        if (r->nge + r->ngn > 0) {
            k_sm_update<<<dim3(cnt::grid_for(r->nge + r->ngn, TB), NS, 1), TB, 0, st>>>(a, r->gebase, r->nge, r->gnbase,
                                                                                      r->ngn);
            CNT_LAUNCHED();
        }
    }
    if (h->dense_nnodes > 0) {
        CNT_CUDA(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)NS * (size_t)h->dense_gpool, st));
        if (h->dense_nE > 0) {
            k_sm_dense_fill<<<dim3(cnt::grid_for(h->dense_nE, 256), NS, 1), 256, 0, st>>>(h->dense_nE, p.dslot, p.depos, a.g,
                                                                                        gs, G, h->dense_gpool);
            CNT_LAUNCHED();
        }
        size_t smem = sizeof(double) * (4 * (size_t)h->dense_nmax + 32);
        k_sm_dense<<<dim3(h->dense_nnet, NS, 1), DENSE_TB, smem, st>>>((int)R, p.doff, p.dptr, p.dnodes, G, h->dense_gpool,
                                                                       a.c, a.b, x, h->dense_nmax);
        CNT_LAUNCHED();
    }
    for (int k = h->nrounds - 1; k >= 0; k--) {
        const cnt_sm_round *r = &h->rounds[k];
        k_sm_back<<<dim3(cnt::grid_for(r->nelim, TB), NS, 1), TB, 0, st>>>(a, r->ebase, r->nelim);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}
