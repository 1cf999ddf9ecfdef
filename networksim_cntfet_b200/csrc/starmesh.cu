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
constexpr int TB = 256;

// Internal layout of the numeric phase: CONFIGURATION-MINOR.  g[slot * NS + q], c / b / x[row * NS + q],
// dinv[pivot * NS + q]: the values of one edge (row) for all NS gate configurations are adjacent, so
// one random access serves every configuration (NS = 4 doubles = one 32-byte sector) and the index
// lists are read once per warp, not once per configuration.  Thread = (item, q), q fastest.

// values of the round-0 edges: sums of the CSR entries (-val) of every stick pair; fill slots = 0
__global__ void k_sm_edges0(int NS, int nE0, int nslots, const int32_t *__restrict__ ptr, const int32_t *__restrict__ src,
                            const double *__restrict__ val, int64_t val_stride, double *__restrict__ g) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nslots * NS) return;
    const int e = (int)(t / NS), q = (int)(t - (int64_t)e * NS);
    double s = 0.0;
    if (e < nE0) {
        const double *v = val + (int64_t)q * val_stride;
        for (int k = ptr[e]; k < ptr[e + 1]; k++) s += -v[src[k]];
    }
    g[t] = s;
}

// electrode coupling of every unknown: c = gs + gd (+ extra); with leak != 0 the rounding error of the
// assembled FP64 diagonal (the same sequential sum as k_values: entries in CSR order, + gs, + gd)
// is added, so that the direct solve answers the SAME floating-point matrix the iterative solvers
// and the reference see; the errors of the additions are exact (TwoSum).  Also b = rhs.
__global__ void k_sm_coupling(int NS, int R, const int32_t *__restrict__ rowptr, const double *__restrict__ val,
                              int64_t val_stride, const int32_t *__restrict__ src_eid,
                              const int32_t *__restrict__ drn_eid, const double *__restrict__ gj, int64_t g_stride,
                              const double *__restrict__ extra, const double *__restrict__ rhs, int64_t rhs_stride,
                              int leak, double2 *__restrict__ cb) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)R * NS) return;
    const int r = (int)(t / NS), q = (int)(t - (int64_t)r * NS);
    const double *v = val + (int64_t)q * val_stride;
    const double *gq = gj ? gj + (int64_t)q * g_stride : nullptr;
    const int se = src_eid ? src_eid[r] : -1, de = drn_eid ? drn_eid[r] : -1;
    const double s_ = se >= 0 ? gq[se] : 0.0, d_ = de >= 0 ? gq[de] : 0.0;
    double out = __dadd_rn(s_, d_);
    if (leak) {
        double sum = 0.0, err = 0.0;
        auto add = [&](double x) {
            double tt = __dadd_rn(sum, x);
            double bb = __dsub_rn(tt, sum);
            err = __dadd_rn(err, __dadd_rn(__dsub_rn(sum, __dsub_rn(tt, bb)), __dsub_rn(x, bb)));
            sum = tt;
        };
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++) add(-v[k]);
        add(s_);
        add(d_);
        out = __dsub_rn(out, err);        // exact sum = sum + err  ->  fl - exact = -err
    }
    if (extra) out += extra[(int64_t)q * R + r];
    cb[t] = make_double2(out, rhs[(int64_t)q * rhs_stride + r]);
}

struct Num {
    int R, NS;
    const int32_t *piv, *nptr, *nbr, *eslot, *tgt, *cptr, *cm, *ca, *cb, *tn, *tptr, *tm, *tslot;
    double *g, *x, *dinv;
    double2 *nv;           // [R, NS] (c, b) of every unknown side by side: the node updates touch both, and DRAM moves 64 bytes
                           // per access -- with separate arrays every 32-byte gather paid for a line and used half of it
    double *gl;            // [nlist, NS] values of the dead edges in the order of the incident lists: final once their
                           // pivot is eliminated, copied by k_sm_pivot, so that the contributions of a round (ca, cb, tslot
                           // = list positions) and the back-substitution read them in order instead of gathering g[slot]
    double *xout;          // [NS, R] configuration-major copy of the solution for the caller
};

// pivots of a round: d = c + sum of the incident conductances (neighbours ascending)
__global__ void k_sm_pivot(Num a, int ebase, int nelim) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nelim * a.NS) return;
    const int ml = (int)(t / a.NS), q = (int)(t - (int64_t)ml * a.NS), m = ebase + ml, NS = a.NS;
    double d = a.nv[(int64_t)a.piv[m] * NS + q].x;
    for (int k = a.nptr[m]; k < a.nptr[m + 1]; k++) {
        const double v = a.g[(int64_t)a.eslot[k] * NS + q];
        a.gl[(int64_t)k * NS + q] = v;
        d += v;
    }
    a.dinv[(int64_t)m * NS + q] = 1.0 / d;
}
// targets of a round: first the edge groups (g[tgt] += sum gl[ca] gl[cb] / d[cm]), then the node groups
// (c, b of a neighbour += sum gl[tslot] / d[tm] * {c, b}[piv[tm]]); contributions ordered by pivot
__global__ void k_sm_update(Num a, int gebase, int nge, int gnbase, int ngn) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int NS = a.NS;
    if (t >= (int64_t)(nge + ngn) * NS) return;
    const int it = (int)(t / NS), q = (int)(t - (int64_t)it * NS);
    const double *g = a.gl;
    if (it < nge) {
        const int gi = gebase + it;
        double v = 0.0;
        for (int k = a.cptr[gi]; k < a.cptr[gi + 1]; k++)
            v += g[(int64_t)a.ca[k] * NS + q] * g[(int64_t)a.cb[k] * NS + q] * a.dinv[(int64_t)a.cm[k] * NS + q];
        a.g[(int64_t)a.tgt[gi] * NS + q] += v;
    } else {
        const int gi = gnbase + (it - nge);
        double sc = 0.0, sb = 0.0;
        for (int k = a.tptr[gi]; k < a.tptr[gi + 1]; k++) {
            const int m = a.tm[k];
            const double w = g[(int64_t)a.tslot[k] * NS + q] * a.dinv[(int64_t)m * NS + q];
            const int64_t v = (int64_t)a.piv[m] * NS + q;
            const double2 cbv = a.nv[v];
            sc += w * cbv.x;
            sb += w * cbv.y;
        }
        const int64_t u = (int64_t)a.tn[gi] * NS + q;
        double2 t2 = a.nv[u];
        t2.x += sc;
        t2.y += sb;
        a.nv[u] = t2;
    }
}
__global__ void k_sm_back(Num a, int ebase, int nelim) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nelim * a.NS) return;
    const int ml = (int)(t / a.NS), q = (int)(t - (int64_t)ml * a.NS), m = ebase + ml, NS = a.NS;
    const int v = a.piv[m];
    double s = a.nv[(int64_t)v * NS + q].y;
    for (int k = a.nptr[m]; k < a.nptr[m + 1]; k++) s += a.gl[(int64_t)k * NS + q] * a.x[(int64_t)a.nbr[k] * NS + q];
    s *= a.dinv[(int64_t)m * NS + q];
    a.x[(int64_t)v * NS + q] = s;
    a.xout[(int64_t)q * a.R + v] = s;
}

// ---- dense tail: once a network is down to a couple of hundred unknowns (the dense core that the
// rounds would take one handful at a time), one CTA per (network, configuration) finishes the same
// star-mesh elimination on a dense upper-triangular conductance matrix, in ascending row order, and
// back-substitutes.  A warp takes one neighbour row of the pivot at a time; rows without a junction
// to the pivot are skipped, so the cost follows the actual degrees.  Same formulas as the sparse
// rounds: sums of positive terms only.  Two variants: k_sm_dense_smem keeps the packed strict upper
// triangle in shared memory (n <= DENSE_SMEM_N); k_sm_dense works on n x n matrices in global memory.
constexpr int DENSE_TB = 256;
constexpr int DENSE_SMEM_N = 224;      // 224 * 223 / 2 doubles = 195 KB + 4 vectors
__global__ void k_sm_dense_fill(int NS, int nE, const int32_t *__restrict__ dslot, const long long *__restrict__ depos,
                                const double *__restrict__ g, double *__restrict__ G, int64_t gpool) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nE * NS) return;
    const int e = (int)(t / NS), q = (int)(t - (int64_t)e * NS);
    G[(int64_t)q * gpool + depos[e]] = g[(int64_t)dslot[e] * NS + q];
}
__global__ void __launch_bounds__(DENSE_TB) k_sm_dense(int NS, int R, const long long *__restrict__ doff,
                                                       const int32_t *__restrict__ dptr, const int32_t *__restrict__ dnodes,
                                                       double *__restrict__ Gall, int64_t gpool, const double2 *__restrict__ cb_in,
                                                       double *__restrict__ x_int,
                                                       double *__restrict__ x_out, int nmax) {
    extern __shared__ double smd[];
    double *rowk = smd, *cc = smd + nmax, *bb = smd + 2 * nmax, *dinv = smd + 3 * nmax, *red = smd + 4 * nmax;
    const int net = blockIdx.x, q = blockIdx.y;
    const int n = dptr[net + 1] - dptr[net];
    if (n <= 0) return;
    const int32_t *nodes = dnodes + dptr[net];
    double *G = Gall + (int64_t)q * gpool + doff[net];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = DENSE_TB >> 5;
    for (int i = t; i < n; i += DENSE_TB) {
        const double2 cbv = cb_in[(int64_t)nodes[i] * NS + q];
        cc[i] = cbv.x;
        bb[i] = cbv.y;
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
    for (int i = t; i < n; i += DENSE_TB) {
        x_int[(int64_t)nodes[i] * NS + q] = rowk[i];
        x_out[(int64_t)q * R + nodes[i]] = rowk[i];
    }
}

// the same elimination with the packed strict upper triangle in shared memory: row i holds the
// entries j > i at T[i * (2n - i - 1) / 2 + (j - i - 1)]; filled straight from the edge values
__global__ void __launch_bounds__(DENSE_TB) k_sm_dense_smem(int NS, int R, int nE, const int32_t *__restrict__ dslot,
                                                            const long long *__restrict__ depos,
                                                            const long long *__restrict__ doff,
                                                            const int32_t *__restrict__ dptr,
                                                            const int32_t *__restrict__ dnodes,
                                                            const int32_t *__restrict__ dedge /* [nnet+1] */,
                                                            const double *__restrict__ g, const double2 *__restrict__ cb_in,
                                                            double *__restrict__ x_int,
                                                            double *__restrict__ x_out) {
    extern __shared__ double smd[];
    const int net = blockIdx.x, q = blockIdx.y;
    const int n = dptr[net + 1] - dptr[net];
    if (n <= 0) return;
    double *cc = smd, *bb = smd + n, *dinv = smd + 2 * n, *xs = smd + 3 * n, *red = smd + 4 * n, *T = smd + 4 * n + 32;
    const int32_t *nodes = dnodes + dptr[net];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = DENSE_TB >> 5;
    const int ntri = n * (n - 1) / 2;
    for (int i = t; i < ntri; i += DENSE_TB) T[i] = 0.0;
    for (int i = t; i < n; i += DENSE_TB) {
        const double2 cbv = cb_in[(int64_t)nodes[i] * NS + q];
        cc[i] = cbv.x;
        bb[i] = cbv.y;
    }
    __syncthreads();
    const long long base = doff[net];
    for (int e = dedge[net] + t; e < dedge[net + 1]; e += DENSE_TB) {
        const long long pos = depos[e] - base;
        const int i = (int)(pos / n), j = (int)(pos - (long long)i * n);
        T[i * (2 * n - i - 1) / 2 + (j - i - 1)] = g[(int64_t)dslot[e] * NS + q];
    }
    __syncthreads();
    for (int k = 0; k < n; k++) {
        const double *rowk = T + k * (2 * n - k - 1) / 2 - (k + 1);      // rowk[j], j > k
        double part = 0.0;
        for (int j = k + 1 + t; j < n; j += DENSE_TB) part += rowk[j];
        part = cnt::block_sum(part, red);
        if (t == 0) dinv[k] = 1.0 / (cc[k] + part);
        __syncthreads();
        const double di = dinv[k], ck = cc[k], bk = bb[k];
        for (int i = k + 1 + w; i < n; i += nw) {
            const double gi = rowk[i];
            if (gi == 0.0) continue;
            const double f = gi * di;
            double *rowi = T + i * (2 * n - i - 1) / 2 - (i + 1);
            for (int j = i + 1 + lane; j < n; j += 32) {
                const double gj = rowk[j];
                if (gj != 0.0) rowi[j] += f * gj;
            }
            if (lane == 0) {
                cc[i] += f * ck;
                bb[i] += f * bk;
            }
        }
        __syncthreads();
    }
    for (int k = n - 1; k >= 0; k--) {
        const double *rowk = T + k * (2 * n - k - 1) / 2 - (k + 1);
        double part = 0.0;
        for (int j = k + 1 + t; j < n; j += DENSE_TB) part += rowk[j] * xs[j];
        part = cnt::block_sum(part, red);
        if (t == 0) xs[k] = (bb[k] + part) * dinv[k];
        __syncthreads();
    }
    for (int i = t; i < n; i += DENSE_TB) {
        x_int[(int64_t)nodes[i] * NS + q] = xs[i];
        x_out[(int64_t)q * R + nodes[i]] = xs[i];
    }
}

struct NumWs {
    int64_t g, gl, cb, dinv, x, G, total;
};
bool dense_in_smem(const cnt_sm_info *h) { return h->dense_nmax <= DENSE_SMEM_N; }
NumWs num_layout(int NS, const cnt_sm_info *h) {
    NumWs L;
    int64_t pos = 0;
    auto take = [&](int64_t count) {
        int64_t at = pos;
        pos += cnt::align_up((count > 0 ? count : 1) * 8, 256);
        return at;
    };
    L.g = take((int64_t)NS * (h->nslots > 0 ? h->nslots : 1));
    L.gl = take((int64_t)NS * (h->nlist > 0 ? h->nlist : 1));
    L.cb = take(2 * (int64_t)NS * h->R);
    L.dinv = take((int64_t)NS * h->R);
    L.x = take((int64_t)NS * h->R);
    L.G = take(dense_in_smem(h) ? 1 : (int64_t)NS * (h->dense_gpool > 0 ? h->dense_gpool : 1));
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
    Num a;
    a.R = (int)R;
    a.NS = NS;
    a.piv = p.piv; a.nptr = p.nptr; a.nbr = p.nbr; a.eslot = p.eslot; a.tgt = p.tgt; a.cptr = p.cptr; a.cm = p.cm;
    a.ca = p.ca; a.cb = p.cb; a.tn = p.tn; a.tptr = p.tptr; a.tm = p.tm; a.tslot = p.tslot;
    a.g = (double *)(w + L.g); a.nv = (double2 *)(w + L.cb); a.x = (double *)(w + L.x);
    a.dinv = (double *)(w + L.dinv);
    a.gl = (double *)(w + L.gl);
    a.xout = x;
    double *G = (double *)(w + L.G);
    auto grid = [&](int64_t items) { return cnt::grid_for(items * NS, TB); };

    if (h->nslots > 0) {
        k_sm_edges0<<<grid(h->nslots), TB, 0, st>>>(NS, h->nE0, h->nslots, p.e0_ptr, p.e0_src, val, val_stride, a.g);
        CNT_LAUNCHED();
    }
    k_sm_coupling<<<grid(R), TB, 0, st>>>(NS, (int)R, rowptr, val, val_stride, src_eid, drn_eid, gj, g_stride, coupling, rhs,
                                          rhs_stride, leak, a.nv);
    CNT_LAUNCHED();
    for (int k = 0; k < h->nrounds; k++) {
        const cnt_sm_round *r = &h->rounds[k];
        k_sm_pivot<<<grid(r->nelim), TB, 0, st>>>(a, r->ebase, r->nelim);
        CNT_LAUNCHED();
        if (r->nge + r->ngn > 0) {
            k_sm_update<<<grid(r->nge + r->ngn), TB, 0, st>>>(a, r->gebase, r->nge, r->gnbase, r->ngn);
            CNT_LAUNCHED();
        }
    }
    if (h->dense_nnodes > 0) {
        if (dense_in_smem(h)) {
            const int n = h->dense_nmax;
            const size_t smem = sizeof(double) * (4 * (size_t)n + 32 + (size_t)n * (n - 1) / 2);
            CNT_CUDA(cudaFuncSetAttribute(k_sm_dense_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_sm_dense_smem<<<dim3(h->dense_nnet, NS, 1), DENSE_TB, smem, st>>>(NS, (int)R, h->dense_nE, p.dslot, p.depos, p.doff,
                                                                                p.dptr, p.dnodes, p.dedge, a.g, a.nv, a.x,
                                                                                x);
            CNT_LAUNCHED();
        } else {
            CNT_CUDA(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)NS * (size_t)h->dense_gpool, st));
            if (h->dense_nE > 0) {
                k_sm_dense_fill<<<grid(h->dense_nE), TB, 0, st>>>(NS, h->dense_nE, p.dslot, p.depos, a.g, G, h->dense_gpool);
                CNT_LAUNCHED();
            }
            size_t smem = sizeof(double) * (4 * (size_t)h->dense_nmax + 32);
            k_sm_dense<<<dim3(h->dense_nnet, NS, 1), DENSE_TB, smem, st>>>(NS, (int)R, p.doff, p.dptr, p.dnodes, G,
                                                                           h->dense_gpool, a.nv, a.x, x, h->dense_nmax);
            CNT_LAUNCHED();
        }
    }
    for (int k = h->nrounds - 1; k >= 0; k--) {
        const cnt_sm_round *r = &h->rounds[k];
        k_sm_back<<<grid(r->nelim), TB, 0, st>>>(a, r->ebase, r->nelim);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}
