// MNA assembly in reduced SPD form  L_UU V_U = 0.1 g_{U,source}.
// Replaces populate_graph (netsim.py:192-195), update_conductivity / make_G / make_A /
// make_z (cnet.py:99-131) and the gate setters (cnet.py:159-183).
//
// Unknowns = sticks of the percolating component except stick 0 (source, 0.1 V) and stick 1
// (ground), in ascending stick index (canonical node order).  The CSR pattern
// (without diagonal) is built once per batch; one gate configuration is a value pass:
// val = -g[junction of the entry], diag = ordered row sum incl. electrode contacts.
#include "common.cuh"

namespace {

__global__ void k_row_flags(int B, const int32_t *__restrict__ soff, int64_t S, const int32_t *__restrict__ label,
                            const int32_t *__restrict__ stats, const uint8_t *__restrict__ alive,
                            const int32_t *__restrict__ elim_vid, int32_t *__restrict__ flag) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    int i = (int)s - soff[b];
    bool in = stats[(int64_t)b * CNT_NSTAT + CNT_STAT_PERCOLATING] && i >= 2 && label[s] == 0;
    if (alive) in = in && alive[s];     // dangling sticks pruned by cnt_prune_leaves
    if (elim_vid) in = in && elim_vid[s] < 0;     // series sticks eliminated by cnt_series_select
    flag[s] = in ? 1 : 0;
}

__global__ void k_rows_finish(int B, const int32_t *__restrict__ soff, int64_t S, const int32_t *__restrict__ flag,
                              const int32_t *__restrict__ scan, int32_t *__restrict__ urow,
                              int32_t *__restrict__ roff) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) urow[s] = flag[s] ? scan[s] : -1;
    if (s <= B) roff[s] = scan[soff[s]];   // scan has S+1 entries (total at S)
}

template <bool FILL>
__global__ void k_pattern(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ joff, int64_t J,
                          const int32_t *__restrict__ stick1, const int32_t *__restrict__ stick2,
                          const int32_t *__restrict__ urow, const int32_t *__restrict__ roff,
                          const int32_t *__restrict__ rowptr, int32_t *__restrict__ rowcnt /* count or cursor */,
                          int32_t *__restrict__ col, int32_t *__restrict__ nz_eid, int32_t *__restrict__ src_eid,
                          int32_t *__restrict__ drn_eid, int eid_base) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= J) return;
    int b = cnt::find_segment(joff, B, (int)e);
    int s0 = soff[b];
    int a = stick1[e], c = stick2[e];
    int ua = urow[s0 + a], uc = urow[s0 + c];
    if (ua >= 0 && uc >= 0) {
        int ka = atomicAdd(rowcnt + ua, 1), kc = atomicAdd(rowcnt + uc, 1);
        if (FILL) {
            int r0 = roff[b];
            col[rowptr[ua] + ka] = uc - r0;
            nz_eid[rowptr[ua] + ka] = eid_base + (int)e;
            col[rowptr[uc] + kc] = ua - r0;
            nz_eid[rowptr[uc] + kc] = eid_base + (int)e;
        }
    } else if (FILL && uc >= 0) {   // a is an electrode of the percolating component
        if (a == 0) src_eid[uc] = (int)e;
        else if (a == 1) drn_eid[uc] = (int)e;
    }
}

// sort every row by (column, junction) (rows are short) so the pattern is deterministic even
// with parallel entries (a virtual series junction next to a real one)
__global__ void k_sort_rows(int64_t R, const int32_t *__restrict__ rowptr, int32_t *__restrict__ col,
                            int32_t *__restrict__ nz_eid) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int k0 = rowptr[r], k1 = rowptr[r + 1];
    constexpr int NREG = 8;
    if (k1 - k0 <= NREG) {
        // short rows (3.7 entries on average): one read, an odd-even transposition network on (column, junction) keys
        // in registers, one write -- instead of a chain of dependent global loads and stores
        if (k1 - k0 < 2) return;
        unsigned long long key[NREG];
#pragma unroll
        for (int a = 0; a < NREG; a++)
            key[a] = k0 + a < k1 ? ((unsigned long long)(unsigned)col[k0 + a] << 32) | (unsigned)nz_eid[k0 + a] : ~0ull;
#pragma unroll
        for (int rr = 0; rr < NREG; rr++) {
#pragma unroll
            for (int a = rr & 1; a + 1 < NREG; a += 2) {
                const unsigned long long lo = key[a] < key[a + 1] ? key[a] : key[a + 1];
                const unsigned long long hi = key[a] < key[a + 1] ? key[a + 1] : key[a];
                key[a] = lo;
                key[a + 1] = hi;
            }
        }
#pragma unroll
        for (int a = 0; a < NREG; a++) {
            if (k0 + a < k1) {
                col[k0 + a] = (int)(key[a] >> 32);
                nz_eid[k0 + a] = (int)(key[a] & 0xFFFFFFFFull);
            }
        }
        return;
    }
    for (int a = k0 + 1; a < k1; a++) {
        int cc = col[a], ee = nz_eid[a];
        int c = a - 1;
        while (c >= k0 && (col[c] > cc || (col[c] == cc && nz_eid[c] > ee))) {
            col[c + 1] = col[c];
            nz_eid[c + 1] = nz_eid[c];
            c--;
        }
        col[c + 1] = cc;
        nz_eid[c + 1] = ee;
    }
}

__global__ void k_row_stick(int B, const int32_t *__restrict__ soff, int64_t S, const int32_t *__restrict__ urow,
                            int32_t *__restrict__ row_stick) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int u = urow[s];
    if (u >= 0) row_stick[u] = (int)s - soff[cnt::find_segment(soff, B, (int)s)];
}

__global__ void k_gate_levels(int64_t J, const double *__restrict__ jx, const double *__restrict__ jy,
                              uint8_t *__restrict__ level, int global_set, uint8_t global_level, int rect_set,
                              double left, double right, double bottom, double top, uint8_t rect_level) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= J) return;
    uint8_t l = level[e];
    if (global_set) l = global_level;
    if (rect_set) {
        double x = jx[e], y = jy[e];
        if (left <= x && x <= right && bottom <= y && y <= top) l = rect_level;   // cnet.py:166-175
    }
    level[e] = l;
}

__global__ void k_conductance(int64_t J, const uint8_t *__restrict__ kind2, const uint8_t *__restrict__ level,
                              const double *__restrict__ gtable, int nlevels, double *__restrict__ g) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < J) g[e] = gtable[(int)kind2[e] * nlevels + (int)level[e]];
}

// conductance of two junctions in series, with fixed roundings (the host builds the same table)
__device__ __forceinline__ double series_g(double g1, double g2) {
    return __ddiv_rn(__dmul_rn(g1, g2), __dadd_rn(g1, g2));
}
// one thread per row: fixed summation order (columns ascending, then source, then drain).
// Entries with junction id >= J are virtual series junctions: conductance g1 g2 / (g1 + g2) of the
// eliminated stick's two junctions; their class is the pair class nb + tri(c1, c2) (c1 <= c2) of
// the two base classes, nb = 9 * nlevels.
__global__ void k_values(int64_t R, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ nz_eid,
                         const int32_t *__restrict__ src_eid, const int32_t *__restrict__ drn_eid,
                         const double *__restrict__ g, double *__restrict__ val, double *__restrict__ diag,
                         double *__restrict__ rhs, const uint8_t *__restrict__ kind2,
                         const uint8_t *__restrict__ level, int nlevels, uint8_t *__restrict__ cls, int64_t J,
                         const int32_t *__restrict__ ve1, const int32_t *__restrict__ ve2) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int k0 = rowptr[r], k1 = rowptr[r + 1];
    const int nb = 9 * nlevels;
    double sum = 0.0;
    for (int k = k0; k < k1; k++) {
        int e = nz_eid[k];
        double ge;
        int c;
        if (e < J) {
            ge = g[e];
            c = cls ? (int)kind2[e] * nlevels + (int)level[e] : 0;
        } else {
            int e1 = ve1[e - J], e2 = ve2[e - J];
            ge = series_g(g[e1], g[e2]);
            c = 0;
            if (cls) {
                int c1 = (int)kind2[e1] * nlevels + (int)level[e1], c2 = (int)kind2[e2] * nlevels + (int)level[e2];
                if (c1 > c2) { int t = c1; c1 = c2; c2 = t; }
                c = nb + c1 * nb - c1 * (c1 - 1) / 2 + (c2 - c1);
            }
        }
        val[k] = -ge;
        if (cls) cls[k] = (uint8_t)c;                    // index into the value table
        sum += ge;
    }
    int se = src_eid[r], de = drn_eid[r];
    double gs = se >= 0 ? g[se] : 0.0;
    double gd = de >= 0 ? g[de] : 0.0;
    sum += gs;
    sum += gd;
    diag[r] = sum;
    rhs[r] = 0.1 * gs;
}
}  // namespace

extern "C" int64_t cnt_assemble_workspace_bytes(int64_t S) {
    return 2 * cnt::ws_bytes_for(S + 1, 4) + cnt_scan_workspace_bytes(S + 1);
}

extern "C" int cnt_assemble_rows(int B, const int32_t *soff, int64_t S, const int32_t *label, const int32_t *stats,
                                 const uint8_t *alive, const int32_t *elim_vid, int32_t *urow, int32_t *roff, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && label && stats && urow && roff && ws && S >= B);
    cnt::Workspace w(ws, ws_bytes);
    int32_t *flag = w.take<int32_t>(S + 1);
    int32_t *scan = w.take<int32_t>(S + 1);
    int64_t sb = cnt_scan_workspace_bytes(S + 1);
    void *sws = w.take<char>(sb);
    if (!sws) {
        cnt::set_error("assemble workspace too small");
        return CNT_ECAPACITY;
    }
    k_row_flags<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, label, stats, alive, elim_vid, flag);
    CNT_LAUNCHED();
    int rc = cnt_exclusive_scan_i32(flag, scan, S, 1, sws, sb, stream);
    if (rc) return rc;
    k_rows_finish<<<cnt::grid_for(S + 1, 256), 256, 0, st>>>(B, soff, S, flag, scan, urow, roff);
    CNT_LAUNCHED();
    return CNT_OK;
}

namespace {
// 16-bit x 16-bit Morton (Z-order) code of a point of the unit box
__device__ __forceinline__ uint32_t spread16(uint32_t v) {
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
__global__ void k_morton_keys(int64_t S, const int32_t *__restrict__ urow, const double *__restrict__ xc,
                              const double *__restrict__ yc, uint32_t *__restrict__ key,
                              uint32_t *__restrict__ val) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int r = urow[s];
    if (r < 0) return;
    double x = xc[s] * 65536.0, y = yc[s] * 65536.0;
    uint32_t ix = x < 0 ? 0u : (x >= 65535.0 ? 65535u : (uint32_t)x);
    uint32_t iy = y < 0 ? 0u : (y >= 65535.0 ? 65535u : (uint32_t)y);
    if (!(x == x)) ix = 0;
    if (!(y == y)) iy = 0;
    key[r] = spread16(ix) | (spread16(iy) << 1);
    val[r] = (uint32_t)s;
}
__global__ void k_apply_order(int64_t R, const uint32_t *__restrict__ order, int32_t *__restrict__ urow) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) urow[order[r]] = (int)r;
}
}  // namespace

extern "C" int64_t cnt_assemble_reorder_workspace_bytes(int64_t R, int B) {
    return 2 * cnt::ws_bytes_for(R, 8) + 2 * cnt::ws_bytes_for(R, 4) + cnt::segsort_workspace_bytes(R, B);
}

// Renumber the unknown rows of every network in Z-order of the stick centres (stable: ties keep
// ascending stick index).  Neighbouring sticks get neighbouring rows, so the SpMV gathers of
// p[col] hit the same sectors and a contiguous row block is a compact patch of the device --
// which is what lets the cluster PCG kernel serve most gathers from its own shared memory.
// The canonical (ascending stick index) order of the reference is only a reporting convention;
// results are written back per stick through urow.
extern "C" int cnt_assemble_reorder(int B, const int32_t *h_roff, int64_t S, int64_t R, const double *xc,
                                    const double *yc, int32_t *urow, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && h_roff && urow && ws && S >= 0 && R >= 0 && h_roff[B] == R);
    if (R == 0) return CNT_OK;
    CNT_CHECK_ARG(xc && yc);
    cnt::Workspace w(ws, ws_bytes);
    uint32_t *keyA = (uint32_t *)w.take<unsigned long long>(R), *keyB = (uint32_t *)w.take<unsigned long long>(R);   // 32-bit keys
    uint32_t *valA = w.take<uint32_t>(R), *valB = w.take<uint32_t>(R);
    int64_t sb = cnt::segsort_workspace_bytes(R, B);
    void *sws = w.take<char>(sb);
    if (!sws) {
        cnt::set_error("assemble_reorder workspace too small");
        return CNT_ECAPACITY;
    }
    k_morton_keys<<<cnt::grid_for(S, 256), 256, 0, st>>>(S, urow, xc, yc, keyA, valA);
    CNT_LAUNCHED();
    uint32_t *ks, *order;
    int rc = cnt::segsort_pairs32(B, h_roff, keyA, valA, keyB, valB, 4, sws, sb, st, &ks, &order);
    if (rc) return rc;
    k_apply_order<<<cnt::grid_for(R, 256), 256, 0, st>>>(R, order, urow);
    CNT_LAUNCHED();
    return CNT_OK;
}

extern "C" int cnt_assemble_pattern_count(int B, const int32_t *soff, const int32_t *joff, int64_t J,
                                          const int32_t *stick1, const int32_t *stick2, const int32_t *urow, int64_t R,
                                          int32_t *rowcnt, int64_t NV, const int32_t *voff, const int32_t *va,
                                          const int32_t *vc, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && joff && urow && rowcnt && R >= 0 && J >= 0);
    CNT_CUDA(cudaMemsetAsync(rowcnt, 0, 4 * (R + 1), st));
    if (J > 0) {
        k_pattern<false><<<cnt::grid_for(J, 256), 256, 0, st>>>(B, soff, joff, J, stick1, stick2, urow, nullptr,
                                                                nullptr, rowcnt, nullptr, nullptr, nullptr, nullptr, 0);
        CNT_LAUNCHED();
    }
    if (NV > 0) {     // virtual junctions of the series elimination: a second junction table
        CNT_CHECK_ARG(voff && va && vc);
        k_pattern<false><<<cnt::grid_for(NV, 256), 256, 0, st>>>(B, soff, voff, NV, va, vc, urow, nullptr, nullptr,
                                                                 rowcnt, nullptr, nullptr, nullptr, nullptr, 0);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}

extern "C" int cnt_assemble_pattern_fill(int B, const int32_t *soff, const int32_t *joff, int64_t J,
                                         const int32_t *stick1, const int32_t *stick2, const int32_t *urow,
                                         const int32_t *roff, int64_t R, const int32_t *rowptr, int32_t *col,
                                         int32_t *nz_eid, int32_t *src_eid, int32_t *drn_eid, int64_t NV,
                                         const int32_t *voff, const int32_t *va, const int32_t *vc, void *ws,
                                         int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && joff && urow && roff && rowptr && R >= 0 && J >= 0 && ws);
    if (R == 0) return CNT_OK;
    CNT_CHECK_ARG(col && nz_eid && src_eid && drn_eid);
    cnt::Workspace w(ws, ws_bytes);
    int32_t *cursor = w.take<int32_t>(R);
    if (!cursor) {
        cnt::set_error("assemble workspace too small");
        return CNT_ECAPACITY;
    }
    CNT_CUDA(cudaMemsetAsync(cursor, 0, 4 * R, st));
    CNT_CUDA(cudaMemsetAsync(src_eid, 0xFF, 4 * R, st));
    CNT_CUDA(cudaMemsetAsync(drn_eid, 0xFF, 4 * R, st));
    if (J > 0) {
        k_pattern<true><<<cnt::grid_for(J, 256), 256, 0, st>>>(B, soff, joff, J, stick1, stick2, urow, roff, rowptr,
                                                               cursor, col, nz_eid, src_eid, drn_eid, 0);
        CNT_LAUNCHED();
    }
    if (NV > 0) {     // entries of the virtual junctions carry the ids J, J+1, ...
        CNT_CHECK_ARG(voff && va && vc);
        k_pattern<true><<<cnt::grid_for(NV, 256), 256, 0, st>>>(B, soff, voff, NV, va, vc, urow, roff, rowptr, cursor, col,
                                                                nz_eid, src_eid, drn_eid, (int)J);
        CNT_LAUNCHED();
    }
    k_sort_rows<<<cnt::grid_for(R, 128), 128, 0, st>>>(R, rowptr, col, nz_eid);
    CNT_LAUNCHED();
    return CNT_OK;
}

extern "C" int cnt_assemble_row_stick(int B, const int32_t *soff, int64_t S, const int32_t *urow,
                                      int32_t *row_stick, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && urow && row_stick && S >= 0);
    if (S > 0) {
        k_row_stick<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, urow, row_stick);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}

extern "C" int cnt_gate_levels(int64_t J, const double *jx, const double *jy, uint8_t *level, int global_set,
                               uint8_t global_level, int rect_set, double cx, double cy, double w, double h,
                               uint8_t rect_level, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(J >= 0);
    if (J == 0) return CNT_OK;
    CNT_CHECK_ARG(jx && jy && level);
    // rectangle edges exactly as cnet.py:168-171 computes them
    double right = cx + w / 2, left = cx - w / 2, top = cy + h / 2, bottom = cy - h / 2;
    k_gate_levels<<<cnt::grid_for(J, 256), 256, 0, st>>>(J, jx, jy, level, global_set, global_level, rect_set, left,
                                                         right, bottom, top, rect_level);
    CNT_LAUNCHED();
    return CNT_OK;
}

extern "C" int cnt_assemble_values(int64_t J, const uint8_t *kind2, const uint8_t *level, const double *gtable,
                                   int nlevels, int64_t R, const int32_t *rowptr, const int32_t *nz_eid,
                                   const int32_t *src_eid, const int32_t *drn_eid, double *g, double *val,
                                   double *diag, double *rhs, uint8_t *cls, const int32_t *ve1, const int32_t *ve2,
                                   void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(J >= 0 && R >= 0 && nlevels > 0);
    if (J > 0 && gtable) {
        CNT_CHECK_ARG(kind2 && level && g);
        k_conductance<<<cnt::grid_for(J, 256), 256, 0, st>>>(J, kind2, level, gtable, nlevels, g);
        CNT_LAUNCHED();
    } else if (J > 0) {
        CNT_CHECK_ARG(g);   // gtable == NULL: g[] is an input (conductances evaluated by the caller)
    }
    if (R > 0) {
        CNT_CHECK_ARG(rowptr && nz_eid && src_eid && drn_eid && val && diag && rhs);
        const int nb = 9 * nlevels;
        if (cls) CNT_CHECK_ARG(gtable && kind2 && level && (ve1 ? nb + nb * (nb + 1) / 2 : nb) <= 256);
        k_values<<<cnt::grid_for(R, 128), 128, 0, st>>>(R, rowptr, nz_eid, src_eid, drn_eid, g, val, diag, rhs, kind2,
                                                        level, nlevels, cls, J, ve1, ve2);
        CNT_LAUNCHED();
    }
    return CNT_OK;
}
