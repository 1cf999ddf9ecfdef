// Aggregates of strongly coupled unknowns for the additive two-level preconditioner
//        M^-1 = D^-1 + Z diag(Z^T A Z)^-1 Z^T            (Z = aggregate indicator vectors)
// used by the PCG solvers (pcg.cu, pcg_cluster.cu).
//
// Why: junction conductances are bimodal (~1 for mm/ss/vm contacts, e^-5 ... e^-10 for gated
// ms/vs contacts, cnet.py:23-41).  Sticks joined by strong junctions form nearly equipotential
// clusters that float on a few weak links; each contributes one tiny eigenvalue that plain
// Jacobi cannot see (217k iterations at Vg=+10 on a 100x100 um network).  Adding Jacobi on the
// aggregate level (one unknown per cluster, diagonal = total conductance leaving the cluster)
// brings that to ~4k iterations.  With no contrast there is one aggregate and the method
// degenerates gracefully to Jacobi plus a constant-mode correction.
//
// Everything is deterministic: aggregates = connected components of the strong sub-graph by
// union-find with minimum-index roots, members ordered by a stable radix sort, per-aggregate
// sums taken by one warp per segment (<= SEGW members) in a fixed order.
#include <algorithm>
#include <vector>
#include <chrono>
#include "common.cuh"

namespace {
constexpr int SEGW = 256;       // members per segment (one warp sums a segment)
constexpr int TILE = 256;

struct ATile {
    int32_t sys, row0, nrows, slab;   // row0 = global pattern row
};
struct ASys {
    int32_t net_row0, nrows, slab, pad;
};

__device__ __forceinline__ int uf_find(int *parent, int x) {
    int p = ((volatile int *)parent)[x];
    while (p != x) {
        int gp = ((volatile int *)parent)[p];
        if (gp != p) ((volatile int *)parent)[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}

__global__ void k_iota(int64_t n, int32_t *parent) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) parent[v] = (int)v;
}

__global__ void __launch_bounds__(TILE) k_gmax(const ATile *__restrict__ tiles, int64_t NNZ,
                                                const int32_t *__restrict__ rowptr, const double *__restrict__ val,
                                                unsigned long long *__restrict__ gmax) {
    ATile t = tiles[blockIdx.x];
    double m = 0.0;
    if (threadIdx.x < t.nrows) {
        int row = t.row0 + threadIdx.x;
        const double *v = val + (int64_t)t.slab * NNZ;
        for (int e = rowptr[row]; e < rowptr[row + 1]; e++) m = fmax(m, fabs(v[e]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(gmax + t.sys, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(TILE) k_strong_union(const ATile *__restrict__ tiles, const ASys *__restrict__ sys,
                                                        int64_t R, int64_t NNZ, const int32_t *__restrict__ rowptr,
                                                        const int32_t *__restrict__ col, const double *__restrict__ val,
                                                        const unsigned long long *__restrict__ gmax, double theta,
                                                        int32_t *parent) {
    ATile t = tiles[blockIdx.x];
    if (threadIdx.x >= t.nrows) return;
    ASys s = sys[t.sys];
    double thr = theta * __longlong_as_double((long long)gmax[t.sys]);
    int row = t.row0 + threadIdx.x;
    int li = row - s.net_row0;
    int64_t vb = (int64_t)t.slab * R + s.net_row0;
    const double *v = val + (int64_t)t.slab * NNZ;
    for (int e = rowptr[row]; e < rowptr[row + 1]; e++) {
        int lj = col[e];
        if (lj <= li || !(fabs(v[e]) > thr)) continue;
        int a = (int)(vb + li), c = (int)(vb + lj);
        while (true) {
            a = uf_find(parent, a);
            c = uf_find(parent, c);
            if (a == c) break;
            int hi = a > c ? a : c, lo = a > c ? c : a;
            if (atomicCAS(parent + hi, hi, lo) == hi) break;
            a = hi;
            c = lo;
        }
    }
}

__global__ void k_roots(int64_t n, int32_t *parent, unsigned long long *__restrict__ key, uint32_t *__restrict__ v32) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    key[v] = (unsigned long long)uf_find(parent, (int)v);
    v32[v] = (uint32_t)v;
}

__global__ void k_fill_i32(int64_t n, int32_t *p, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// one warp per segment: mark the members' aggregate
__global__ void k_mark(int64_t nseg, const int32_t *__restrict__ seg_start, const int32_t *__restrict__ seg_count,
                       const int32_t *__restrict__ seg_agg, const uint32_t *__restrict__ mem,
                       int32_t *__restrict__ agg_of_row) {
    int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nseg) return;
    int lane = threadIdx.x & 31;
    int s0 = seg_start[s], n = seg_count[s], a = seg_agg[s];
    for (int k = lane; k < n; k += 32) agg_of_row[mem[s0 + k]] = a;
}

// one warp per segment: sum over members of (diag + entries that stay inside the aggregate)
__global__ void k_seg_E(int64_t nseg, const int32_t *__restrict__ seg_start, const int32_t *__restrict__ seg_count,
                        const int32_t *__restrict__ seg_agg, const uint32_t *__restrict__ mem,
                        const int32_t *__restrict__ agg_of_row, int B, const int32_t *__restrict__ roff, int64_t R,
                        int64_t NNZ, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                        const double *__restrict__ val, const double *__restrict__ diag, double *__restrict__ segE) {
    int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nseg) return;
    int lane = threadIdx.x & 31;
    int s0 = seg_start[s], n = seg_count[s], a = seg_agg[s];
    double acc = 0.0;
    for (int k = lane; k < n; k += 32) {
        int64_t v = mem[s0 + k];
        int slab = (int)(v / R), row = (int)(v - (int64_t)slab * R);
        int net0 = roff[cnt::find_segment(roff, B, row)];
        int64_t vb = (int64_t)slab * R + net0;
        const double *vv = val + (int64_t)slab * NNZ;
        double e = diag[v];
        for (int q = rowptr[row]; q < rowptr[row + 1]; q++)
            if (agg_of_row[vb + col[q]] == a) e += vv[q];
        acc += e;
    }
    acc = cnt::warp_sum(acc);
    if (lane == 0) segE[s] = acc;
}

// per aggregated row: (first segment, number of segments) of its aggregate and 1/E -- the
// coarse term then needs one coalesced load + the segment sums instead of a 4-deep chain
__global__ void k_row_tables(int64_t nseg, const int32_t *__restrict__ seg_start, const int32_t *__restrict__ seg_count,
                             const int32_t *__restrict__ seg_agg, const uint32_t *__restrict__ mem,
                             const int32_t *__restrict__ agg_seg0, const double *__restrict__ einv,
                             int32_t *__restrict__ row_q, double *__restrict__ row_einv) {
    int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nseg) return;
    int lane = threadIdx.x & 31;
    int s0 = seg_start[s], n = seg_count[s], a = seg_agg[s];
    int q0 = agg_seg0[a], qn = agg_seg0[a + 1] - q0;
    double e = einv[a];
    for (int k = lane; k < n; k += 32) {
        int64_t v = mem[s0 + k];
        row_q[2 * v] = q0;
        row_q[2 * v + 1] = qn;
        row_einv[v] = e;
    }
}

__global__ void k_agg_einv(int64_t nagg, const int32_t *__restrict__ agg_seg0, const double *__restrict__ segE,
                           double *__restrict__ einv) {
    int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nagg) return;
    double e = 0.0;
    for (int s = agg_seg0[a]; s < agg_seg0[a + 1]; s++) e += segE[s];
    // E is a difference of large numbers (internal strong links cancel): accuracy is irrelevant
    // for a preconditioner, positivity is not
    einv[a] = (e > 0.0 && e == e) ? 1.0 / e : 0.0;
}

// ---------------------------------------------------------------------------------------
// Segment / CTA tables on the device (the host builder further down is the reference the tests
// compare against: same tables, same order).  Input: positions 0..n-1 sorted by (system, root,
// member); a run of equal roots with >= 2 members is an aggregate; it is cut into segments
// wherever the owning CTA (row chunk of the cluster solver) changes and every SEGW members.
// Flag arrays + exclusive scans give run / segment / aggregate numbers; nothing is atomic.
// ---------------------------------------------------------------------------------------
struct ASysD {
    int64_t first;                     // first position (= slab * R + net_row0), ascending
    int32_t net_row0, nrows, sidx, pad;
};
__device__ __forceinline__ int sys_lookup(const ASysD *__restrict__ ord, int nsys, int64_t pos) {
    int lo = 0, hi = nsys - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (ord[mid].first <= pos) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int cta_of(const ASysD &sy, int C, uint32_t member) {
    if (C <= 0) return 0;
    int chunk = (sy.nrows + C - 1) / C;
    return (int)(((int64_t)member - sy.first) / chunk);
}
__device__ __forceinline__ bool in_aggregate(const unsigned long long *__restrict__ ks, int64_t n, int64_t i) {
    unsigned long long r = ks[i];
    return (i > 0 && ks[i - 1] == r) || (i + 1 < n && ks[i + 1] == r);
}
// head of a (root, CTA) run
__global__ void k_run_heads(int64_t n, const unsigned long long *__restrict__ ks, const uint32_t *__restrict__ vs,
                            const ASysD *__restrict__ ord, int nsys, int C, int32_t *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = 0;
    if (in_aggregate(ks, n, i)) {
        if (i == 0 || ks[i - 1] != ks[i]) {
            f = 1;
        } else if (C > 0) {
            ASysD sy = ord[sys_lookup(ord, nsys, (int64_t)ks[i])];
            f = cta_of(sy, C, vs[i]) != cta_of(sy, C, vs[i - 1]);
        }
    }
    flag[i] = f;
}
__global__ void k_run_start(int64_t n, const int32_t *__restrict__ scan, int32_t *__restrict__ start) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && scan[i + 1] != scan[i]) start[scan[i]] = (int32_t)i;
}
// segment heads (every SEGW members of a run) and, written at the head's position, the member count
__global__ void k_seg_heads(int64_t n, const unsigned long long *__restrict__ ks, const int32_t *__restrict__ runscan,
                            const int32_t *__restrict__ start, int32_t *__restrict__ flag,
                            int32_t *__restrict__ count_at) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!in_aggregate(ks, n, i)) {
        flag[i] = 0;
        return;
    }
    int st = start[runscan[i + 1] - 1];
    int off = (int)i - st;
    flag[i] = (off % SEGW) == 0;
    bool last = (i + 1 == n) || !in_aggregate(ks, n, i + 1) || runscan[i + 2] != runscan[i + 1] || ((off + 1) % SEGW) == 0;
    if (last) {
        int hp = st + (off / SEGW) * SEGW;
        count_at[hp] = (int)i - hp + 1;
    }
}
__global__ void k_agg_heads(int64_t n, const unsigned long long *__restrict__ ks, int32_t *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || ks[i - 1] != ks[i]) && (i + 1 < n && ks[i + 1] == ks[i]);
}
__global__ void k_emit(int64_t n, const unsigned long long *__restrict__ ks, const uint32_t *__restrict__ vs,
                       const ASysD *__restrict__ ord, int nsys, int C, const int32_t *__restrict__ segscan,
                       const int32_t *__restrict__ aggscan, const int32_t *__restrict__ count_at,
                       int32_t *__restrict__ seg_start, int32_t *__restrict__ seg_count, int32_t *__restrict__ seg_agg,
                       int32_t *__restrict__ seg_sys, int32_t *__restrict__ seg_cta, int32_t *__restrict__ agg_seg0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) agg_seg0[aggscan[n]] = segscan[n];
    int q = segscan[i];
    if (segscan[i + 1] == q) return;
    int a = aggscan[i + 1] - 1;
    ASysD sy = ord[sys_lookup(ord, nsys, (int64_t)ks[i])];
    seg_start[q] = (int32_t)i;
    seg_count[q] = count_at[i];
    seg_agg[q] = a;
    seg_sys[q] = sy.sidx;
    seg_cta[q] = cta_of(sy, C, vs[i]);
    if (aggscan[i + 1] != aggscan[i]) agg_seg0[a] = q;
}
__global__ void k_sys_seg(int64_t nseg, const int32_t *__restrict__ seg_sys, int32_t *__restrict__ sys_seg) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nseg) return;
    int s = seg_sys[q];
    if (q == 0 || seg_sys[q - 1] != s) sys_seg[2 * s] = (int32_t)q;
    if (q + 1 == nseg || seg_sys[q + 1] != s) sys_seg[2 * s + 1] = (int32_t)q + 1;
}
__global__ void k_cta_keys(int64_t nseg, const int32_t *__restrict__ seg_sys, const int32_t *__restrict__ seg_cta, int C,
                           unsigned long long *__restrict__ key, uint32_t *__restrict__ val) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nseg) return;
    key[q] = (unsigned long long)seg_sys[q] * (unsigned)C + (unsigned)seg_cta[q];
    val[q] = (uint32_t)q;
}
__global__ void k_cta_off(int64_t nkeys, int64_t nseg, const unsigned long long *__restrict__ skey,
                          int32_t *__restrict__ cta_seg_off) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nkeys) return;
    int64_t lo = 0, hi = nseg;                 // first position with key >= c
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (skey[mid] < (unsigned long long)c) lo = mid + 1; else hi = mid;
    }
    cta_seg_off[c] = (int32_t)lo;
}
__global__ void k_cta_fill(int64_t nseg, const unsigned long long *__restrict__ skey, const uint32_t *__restrict__ sval,
                           int C, const int32_t *__restrict__ cta_seg_off, const int32_t *__restrict__ seg_start,
                           const int32_t *__restrict__ seg_count, const int32_t *__restrict__ seg_agg,
                           int32_t *__restrict__ cta_seg, int32_t *__restrict__ seg_home, int32_t *__restrict__ flag) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= nseg) return;
    int q = (int)sval[pos];
    int64_t c = (int64_t)skey[pos];
    int first = cta_seg_off[c];
    cta_seg[4 * pos + 0] = q;
    cta_seg[4 * pos + 1] = seg_start[q];
    cta_seg[4 * pos + 2] = seg_count[q];
    cta_seg[4 * pos + 3] = 0;
    seg_home[q] = (int32_t)(((c % C) << 16) | (pos - first));
    flag[pos] = (pos == first) || seg_agg[q] != seg_agg[sval[pos - 1]];
}
__global__ void k_cta_agg(int64_t nseg, int64_t nkeys, const uint32_t *__restrict__ sval,
                          const int32_t *__restrict__ seg_agg, const int32_t *__restrict__ fscan,
                          const int32_t *__restrict__ cta_seg_off, int32_t *__restrict__ cta_agg,
                          int32_t *__restrict__ cta_agg_off) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg && fscan[i + 1] != fscan[i]) cta_agg[fscan[i]] = seg_agg[sval[i]];
    if (i <= nkeys) cta_agg_off[i] = fscan[cta_seg_off[i]];
}
// one warp per list position: slot of the segment's aggregate in its CTA's local aggregate list
__global__ void k_row_slot(int64_t nseg, const unsigned long long *__restrict__ skey, const uint32_t *__restrict__ sval,
                           const int32_t *__restrict__ fscan, const int32_t *__restrict__ cta_agg_off,
                           const int32_t *__restrict__ seg_start, const int32_t *__restrict__ seg_count,
                           const uint32_t *__restrict__ mem, int32_t *__restrict__ row_slot) {
    int64_t pos = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pos >= nseg) return;
    int lane = threadIdx.x & 31;
    int q = (int)sval[pos];
    int slot = fscan[pos + 1] - 1 - cta_agg_off[skey[pos]];
    int s0 = seg_start[q], cnt = seg_count[q];
    for (int k = lane; k < cnt; k += 32) row_slot[mem[s0 + k]] = slot;
}
}  // namespace

extern "C" int64_t cnt_pcg_aggregate_workspace_bytes(int NSYS, int NSLAB, int64_t R) {
    using cnt::ws_bytes_for;
    int64_t n = (int64_t)NSLAB * R;
    int64_t nt = (int64_t)NSYS * ((R + TILE - 1) / TILE + 1);
    return ws_bytes_for(n, 4) + ws_bytes_for(NSYS, 8) + 2 * ws_bytes_for(n, 8) + 2 * ws_bytes_for(n, 4) +
           ws_bytes_for(nt, sizeof(ATile)) + ws_bytes_for(NSYS, sizeof(ASys)) + ws_bytes_for(n + 2, 8) +
           cnt::segsort_workspace_bytes(n, 2 * NSYS + 2) + 5 * ws_bytes_for(n + 1, 4) + ws_bytes_for(NSYS, sizeof(ASysD)) +
           cnt_scan_workspace_bytes(n + 1);
}

// Capacity the caller must provide: agg_of_row, mem: NSLAB*R; seg_* (seg_capacity entries; checked): NSLAB*R/2 +
// NSLAB*R/256 + 2 for chunk_C == 0, NSLAB*R + 2 for the cluster tables (segments also end at CTA chunk boundaries);
// agg_seg0: NSLAB*R/2 + 2; agg_einv: NSLAB*R/2 + 1.  Synchronises the stream (the run structure
// of the sorted roots is read back once to build the segment tables).
extern "C" int cnt_pcg_aggregates(int B, int NSYS, int NSLAB, const int32_t *h_sys_net, const int32_t *h_sys_slab,
                                  const int32_t *h_roff, const int32_t *roff, int64_t R, int64_t NNZ,
                                  const int32_t *rowptr, const int32_t *col, const double *val, const double *diag,
                                  double theta, int32_t *agg_of_row, int32_t *mem, int32_t *seg_start,
                                  int32_t *seg_count, int32_t *seg_agg, int32_t *seg_sys, int32_t *agg_seg0,
                                  double *agg_einv, int32_t *sys_seg, int chunk_C, int32_t *row_q, double *row_einv,
                                  int32_t *cta_seg_off, int32_t *cta_seg, int32_t *cta_agg_off, int32_t *cta_agg,
                                  int32_t *row_slot, int32_t *seg_home, int64_t seg_capacity,
                                  int64_t *h_nagg, int64_t *h_nseg, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && NSYS > 0 && NSLAB > 0 && h_sys_net && h_sys_slab && h_roff && roff && R >= 0 && h_nagg &&
                  h_nseg && ws && theta > 0);
    *h_nagg = 0;
    *h_nseg = 0;
    int64_t n = (int64_t)NSLAB * R;
    if (n == 0) return CNT_OK;
    CNT_CHECK_ARG(n < ((int64_t)1 << 31));
    CNT_CHECK_ARG(rowptr && diag && agg_of_row && mem && seg_start && seg_count && seg_agg && seg_sys && agg_seg0 &&
                  agg_einv && sys_seg && chunk_C >= 0 && (chunk_C == 0 || (cta_seg_off && cta_seg && cta_agg_off && cta_agg && row_slot && seg_home)));
    // systems sorted by (slab, network) are the sort segments; tiles for the row kernels
    std::vector<ASys> sys(NSYS);
    std::vector<ATile> tiles;
    std::vector<std::pair<int64_t, int>> order(NSYS);
    for (int s = 0; s < NSYS; s++) {
        int b = h_sys_net[s];
        CNT_CHECK_ARG(b >= 0 && b < B && h_sys_slab[s] >= 0 && h_sys_slab[s] < NSLAB);
        sys[s].net_row0 = h_roff[b];
        sys[s].nrows = h_roff[b + 1] - h_roff[b];
        sys[s].slab = h_sys_slab[s];
        sys[s].pad = 0;
        order[s] = {(int64_t)h_sys_slab[s] * R + h_roff[b], s};
        for (int t = 0; t * TILE < sys[s].nrows; t++) {
            ATile tl;
            tl.sys = s;
            tl.row0 = sys[s].net_row0 + t * TILE;
            tl.nrows = sys[s].nrows - t * TILE < TILE ? sys[s].nrows - t * TILE : TILE;
            tl.slab = sys[s].slab;
            tiles.push_back(tl);
        }
    }
    std::sort(order.begin(), order.end());
    // sort segments must tile [0, n): systems plus the gaps between them (rows of pairs that are
    // not being solved stay singletons)
    std::vector<int32_t> h_off;
    h_off.push_back(0);
    for (int k = 0; k < NSYS; k++) {
        int64_t lo = order[k].first, hi = lo + sys[order[k].second].nrows;
        if (lo < h_off.back()) {
            cnt::set_error("duplicate (network, slab) system");
            return CNT_EINVAL;
        }
        if (lo > h_off.back()) h_off.push_back((int32_t)lo);
        if (hi > h_off.back()) h_off.push_back((int32_t)hi);
    }
    if (h_off.back() < n) h_off.push_back((int32_t)n);
    int nsortseg = (int)h_off.size() - 1;
    int64_t nt = (int64_t)tiles.size();

    cnt::Workspace w(ws, ws_bytes);
    int32_t *parent = w.take<int32_t>(n);
    unsigned long long *gmax = w.take<unsigned long long>(NSYS);
    unsigned long long *keyA = w.take<unsigned long long>(n), *keyB = w.take<unsigned long long>(n);
    uint32_t *valA = w.take<uint32_t>(n), *valB = w.take<uint32_t>(n);
    ATile *d_tiles = w.take<ATile>((int64_t)NSYS * ((R + TILE - 1) / TILE + 1));
    ASys *d_sys = w.take<ASys>(NSYS);
    double *segE = w.take<double>(n + 2);          // nseg <= n (segments also end at CTA chunk boundaries)
    int64_t sb = cnt::segsort_workspace_bytes(n, 2 * NSYS + 2);
    void *sws = w.take<char>(sb);
    int32_t *tA = w.take<int32_t>(n + 1), *tB = w.take<int32_t>(n + 1), *tC = w.take<int32_t>(n + 1);
    int32_t *tD = w.take<int32_t>(n + 1), *tE = w.take<int32_t>(n + 1);
    ASysD *d_ord = w.take<ASysD>(NSYS);
    int64_t scb = cnt_scan_workspace_bytes(n + 1);
    void *scws = w.take<char>(scb);
    if (!sws || !scws || nsortseg > 2 * NSYS + 1) {
        cnt::set_error("aggregate workspace too small");
        return CNT_ECAPACITY;
    }
    const bool timing = getenv("CNT_AGG_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = tnow(), t1 = 0, t2 = 0, t3 = 0;
    CNT_CUDA(cudaMemcpyAsync(d_sys, sys.data(), sizeof(ASys) * NSYS, cudaMemcpyHostToDevice, st));
    if (nt) CNT_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(ATile) * nt, cudaMemcpyHostToDevice, st));
    CNT_CUDA(cudaMemsetAsync(gmax, 0, 8 * NSYS, st));
    k_iota<<<cnt::grid_for(n, 256), 256, 0, st>>>(n, parent);
    CNT_LAUNCHED();
    if (nt && NNZ > 0) {
        k_gmax<<<(unsigned)nt, TILE, 0, st>>>(d_tiles, NNZ, rowptr, val, gmax);
        CNT_LAUNCHED();
        k_strong_union<<<(unsigned)nt, TILE, 0, st>>>(d_tiles, d_sys, R, NNZ, rowptr, col, val, gmax, theta, parent);
        CNT_LAUNCHED();
    }
    k_roots<<<cnt::grid_for(n, 256), 256, 0, st>>>(n, parent, keyA, valA);
    CNT_LAUNCHED();
    int bits = 1;
    while (((int64_t)1 << bits) < n) bits++;
    unsigned long long *ks;
    uint32_t *vs;
    int rc = cnt::segsort_pairs(nsortseg, h_off.data(), keyA, valA, keyB, valB, (bits + 7) / 8, sws, sb, st, &ks, &vs);
    if (rc) return rc;
    // members in aggregate order -> caller's `mem`; sorted roots -> host
    CNT_CUDA(cudaMemcpyAsync(mem, vs, 4 * n, cudaMemcpyDeviceToDevice, st));
    const bool host_tables = getenv("CNT_AGG_HOST") != nullptr;      // reference builder (tests)
    int64_t nagg = 0, nseg = 0;
    std::vector<unsigned long long> roots;
    std::vector<uint32_t> members;
    std::vector<int32_t> h_seg_start, h_seg_count, h_seg_agg, h_seg_sys, h_seg_cta, h_agg_seg0;
    std::vector<int32_t> h_cta_off, h_cta_seg, h_cta_agg_off, h_cta_agg, h_row_slot, h_seg_home, h_sys_seg;
    if (!host_tables) {
        // ---- device: flags + scans (see the kernels above) ----
        std::vector<ASysD> ordv(NSYS);
        for (int k = 0; k < NSYS; k++) {
            const ASys &sy = sys[order[k].second];
            ordv[k] = ASysD{order[k].first, sy.net_row0, sy.nrows, order[k].second, 0};
        }
        CNT_CUDA(cudaMemcpyAsync(d_ord, ordv.data(), sizeof(ASysD) * NSYS, cudaMemcpyHostToDevice, st));
        const unsigned gn = cnt::grid_for(n, 256);
        k_run_heads<<<gn, 256, 0, st>>>(n, ks, vs, d_ord, NSYS, chunk_C, tA);
        CNT_LAUNCHED();
        if ((rc = cnt_exclusive_scan_i32(tA, tB, n, 1, scws, scb, stream))) return rc;          // tB: run number
        k_run_start<<<gn, 256, 0, st>>>(n, tB, tC);                                             // tC: run start
        CNT_LAUNCHED();
        k_seg_heads<<<gn, 256, 0, st>>>(n, ks, tB, tC, tA, tD);                                 // tD: count at head
        CNT_LAUNCHED();
        if ((rc = cnt_exclusive_scan_i32(tA, tE, n, 1, scws, scb, stream))) return rc;          // tE: segment number
        k_agg_heads<<<gn, 256, 0, st>>>(n, ks, tA);
        CNT_LAUNCHED();
        if ((rc = cnt_exclusive_scan_i32(tA, tB, n, 1, scws, scb, stream))) return rc;          // tB: aggregate number
        int32_t h_tot[2] = {0, 0};
        CNT_CUDA(cudaMemcpyAsync(&h_tot[0], tE + n, 4, cudaMemcpyDeviceToHost, st));
        CNT_CUDA(cudaMemcpyAsync(&h_tot[1], tB + n, 4, cudaMemcpyDeviceToHost, st));
        CNT_CUDA(cudaStreamSynchronize(st));          // the ordv upload is complete as well
        nseg = h_tot[0];
        nagg = h_tot[1];
        if (nseg > seg_capacity) {
            cnt::set_error("aggregate tables: %lld segments exceed the caller's seg_* capacity %lld", (long long)nseg,
                           (long long)seg_capacity);
            return CNT_ECAPACITY;
        }
        t1 = t2 = tnow();
        k_emit<<<gn, 256, 0, st>>>(n, ks, vs, d_ord, NSYS, chunk_C, tE, tB, tD, seg_start, seg_count, seg_agg, seg_sys,
                                   tC, agg_seg0);                                               // tC: CTA of segment
        CNT_LAUNCHED();
        CNT_CUDA(cudaMemsetAsync(sys_seg, 0, 4 * 2 * (size_t)NSYS, st));
        if (nseg > 0) {
            k_sys_seg<<<cnt::grid_for(nseg, 256), 256, 0, st>>>(nseg, seg_sys, sys_seg);
            CNT_LAUNCHED();
        }
    } else {
        roots.resize(n);
        members.resize(chunk_C > 0 ? n : 0);
        CNT_CUDA(cudaMemcpyAsync(roots.data(), ks, 8 * n, cudaMemcpyDeviceToHost, st));
        if (chunk_C > 0) CNT_CUDA(cudaMemcpyAsync(members.data(), vs, 4 * n, cudaMemcpyDeviceToHost, st));
        CNT_CUDA(cudaStreamSynchronize(st));
        t1 = tnow();
        // position -> system: `order` holds the systems sorted by first position
        auto sys_of = [&](int64_t pos) {
            int lo = 0, hi = NSYS - 1;
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (order[mid].first <= pos) lo = mid; else hi = mid - 1;
            }
            return order[lo].second;
        };
        int64_t i = 0;
        while (i < n) {
            int64_t j = i + 1;
            while (j < n && roots[j] == roots[i]) j++;
            if (j - i >= 2) {
                int a = (int)h_agg_seg0.size();
                int sidx = sys_of((int64_t)roots[i]);
                h_agg_seg0.push_back((int32_t)h_seg_start.size());
                // segments: <= SEGW members, and (cluster solver) never across a CTA chunk boundary
                const ASys &sy = sys[sidx];
                const int64_t vb0 = (int64_t)sy.slab * R + sy.net_row0;
                const int chunk = chunk_C > 0 ? (sy.nrows + chunk_C - 1) / chunk_C : 0;
                int64_t p0 = i;
                while (p0 < j) {
                    int64_t p1 = p0 + 1;
                    int cta = chunk ? (int)((members[p0] - vb0) / chunk) : 0;
                    while (p1 < j && p1 - p0 < SEGW && (!chunk || (int)((members[p1] - vb0) / chunk) == cta)) p1++;
                    h_seg_start.push_back((int32_t)p0);
                    h_seg_count.push_back((int32_t)(p1 - p0));
                    h_seg_agg.push_back(a);
                    h_seg_sys.push_back(sidx);
                    h_seg_cta.push_back(cta);
                    p0 = p1;
                }
            }
            i = j;
        }

        nagg = (int64_t)h_agg_seg0.size();
        nseg = (int64_t)h_seg_start.size();
        if (nseg > seg_capacity) {
            cnt::set_error("aggregate tables: %lld segments exceed the caller's seg_* capacity %lld", (long long)nseg,
                           (long long)seg_capacity);
            return CNT_ECAPACITY;
        }
        t2 = tnow();
        // segments of one system are contiguous (the member array is ordered by system)
        h_sys_seg.assign(2 * (size_t)NSYS, 0);
        for (int64_t q = 0; q < nseg; q++) {
            int sidx = h_seg_sys[q];
            if (h_sys_seg[2 * sidx + 1] == 0) h_sys_seg[2 * sidx] = (int32_t)q;
            h_sys_seg[2 * sidx + 1] = (int32_t)q + 1;
        }
        CNT_CUDA(cudaMemcpyAsync(sys_seg, h_sys_seg.data(), 4 * 2 * (size_t)NSYS, cudaMemcpyHostToDevice, st));
        h_agg_seg0.push_back((int32_t)nseg);
        CNT_CUDA(cudaMemcpyAsync(agg_seg0, h_agg_seg0.data(), 4 * (nagg + 1), cudaMemcpyHostToDevice, st));
        if (nseg > 0) {
            CNT_CUDA(cudaMemcpyAsync(seg_start, h_seg_start.data(), 4 * nseg, cudaMemcpyHostToDevice, st));
            CNT_CUDA(cudaMemcpyAsync(seg_count, h_seg_count.data(), 4 * nseg, cudaMemcpyHostToDevice, st));
            CNT_CUDA(cudaMemcpyAsync(seg_agg, h_seg_agg.data(), 4 * nseg, cudaMemcpyHostToDevice, st));
            CNT_CUDA(cudaMemcpyAsync(seg_sys, h_seg_sys.data(), 4 * nseg, cudaMemcpyHostToDevice, st));
        }
    }
    *h_nagg = nagg;
    *h_nseg = nseg;
    k_fill_i32<<<cnt::grid_for(n, 256), 256, 0, st>>>(n, agg_of_row, -1);
    CNT_LAUNCHED();
    if (nseg > 0) {
        unsigned gwarp = cnt::grid_for(nseg * 32, 256);
        k_mark<<<gwarp, 256, 0, st>>>(nseg, seg_start, seg_count, seg_agg, (const uint32_t *)mem, agg_of_row);
        CNT_LAUNCHED();
        k_seg_E<<<gwarp, 256, 0, st>>>(nseg, seg_start, seg_count, seg_agg, (const uint32_t *)mem, agg_of_row, B, roff, R,
                                       NNZ, rowptr, col, val, diag, segE);
        CNT_LAUNCHED();
        k_agg_einv<<<cnt::grid_for(nagg, 256), 256, 0, st>>>(nagg, agg_seg0, segE, agg_einv);
        CNT_LAUNCHED();
    }
    if (row_q && row_einv) {
        // per-row (first segment, segment count, 1/E): short load chain for the coarse term
        CNT_CUDA(cudaMemsetAsync(row_q, 0, 8 * n, st));
        CNT_CUDA(cudaMemsetAsync(row_einv, 0, 8 * n, st));
        if (nseg > 0) {
            k_row_tables<<<cnt::grid_for(nseg * 32, 256), 256, 0, st>>>(nseg, seg_start, seg_count, seg_agg,
                                                                        (const uint32_t *)mem, agg_seg0, agg_einv,
                                                                        row_q, row_einv);
            CNT_LAUNCHED();
        }
    }
    if (chunk_C > 0 && !host_tables) {
        // per (system, CTA): segment list (stable sort of the segments by system * C + CTA), local
        // aggregate list (adjacent dedup of the list's aggregates) and the rows' aggregate slots
        const int64_t nkeys = (int64_t)NSYS * chunk_C;
        unsigned long long *sk = keyA;
        uint32_t *sv = valA;
        if (nseg > 0) {
            k_cta_keys<<<cnt::grid_for(nseg, 256), 256, 0, st>>>(nseg, seg_sys, tC, chunk_C, keyA, valA);
            CNT_LAUNCHED();
            int kb = 1;
            while (((int64_t)1 << kb) < nkeys) kb++;
            int32_t h_off2[2] = {0, (int32_t)nseg};
            if ((rc = cnt::segsort_pairs(1, h_off2, keyA, valA, keyB, valB, (kb + 7) / 8, sws, sb, st, &sk, &sv))) return rc;
        }
        k_cta_off<<<cnt::grid_for(nkeys + 1, 256), 256, 0, st>>>(nkeys, nseg, sk, cta_seg_off);
        CNT_LAUNCHED();
        if (nseg > 0) {
            k_cta_fill<<<cnt::grid_for(nseg, 256), 256, 0, st>>>(nseg, sk, sv, chunk_C, cta_seg_off, seg_start, seg_count,
                                                                seg_agg, cta_seg, seg_home, tA);
            CNT_LAUNCHED();
        }
        if ((rc = cnt_exclusive_scan_i32(tA, tD, nseg, 1, scws, scb, stream))) return rc;
        const int64_t m = nseg > nkeys + 1 ? nseg : nkeys + 1;
        k_cta_agg<<<cnt::grid_for(m, 256), 256, 0, st>>>(nseg, nkeys, sv, seg_agg, tD, cta_seg_off, cta_agg, cta_agg_off);
        CNT_LAUNCHED();
        k_fill_i32<<<cnt::grid_for(n, 256), 256, 0, st>>>(n, row_slot, -1);
        CNT_LAUNCHED();
        if (nseg > 0) {
            k_row_slot<<<cnt::grid_for(nseg * 32, 256), 256, 0, st>>>(nseg, sk, sv, tD, cta_agg_off, seg_start, seg_count,
                                                                     (const uint32_t *)mem, row_slot);
            CNT_LAUNCHED();
        }
    }
    if (chunk_C > 0 && host_tables) {
        // per (system, CTA): its segments as (segment id, first member, count, 0) quadruples
        h_cta_off.assign((size_t)NSYS * chunk_C + 1, 0);
        for (int64_t q = 0; q < nseg; q++) h_cta_off[(size_t)h_seg_sys[q] * chunk_C + h_seg_cta[q] + 1]++;
        for (size_t k = 1; k < h_cta_off.size(); k++) h_cta_off[k] += h_cta_off[k - 1];
        std::vector<int32_t> cur(h_cta_off.begin(), h_cta_off.end() - 1);
        h_cta_seg.assign(4 * (size_t)(nseg > 0 ? nseg : 1), 0);
        h_seg_home.assign((size_t)(nseg > 0 ? nseg : 1), 0);
        for (int64_t q = 0; q < nseg; q++) {
            int32_t pos = cur[(size_t)h_seg_sys[q] * chunk_C + h_seg_cta[q]]++;
            h_cta_seg[4 * (size_t)pos + 0] = (int32_t)q;
            h_cta_seg[4 * (size_t)pos + 1] = h_seg_start[q];
            h_cta_seg[4 * (size_t)pos + 2] = h_seg_count[q];
            // home of the segment: owning CTA and its index in that CTA's list (DSMEM exchange)
            h_seg_home[(size_t)q] = (h_seg_cta[q] << 16) | (pos - h_cta_off[(size_t)h_seg_sys[q] * chunk_C + h_seg_cta[q]]);
        }
        // per (system, CTA): the aggregates it touches (dedup of its segments' aggregates, which are
        // adjacent because segment ids are contiguous per aggregate) and, per row, the slot of its
        // aggregate in that CTA-local list
        h_cta_agg_off.assign((size_t)NSYS * chunk_C + 1, 0);
        h_row_slot.assign((size_t)n, -1);
        for (size_t c = 0; c + 1 < h_cta_off.size(); c++) {
            int prev = -1, nloc = 0;
            for (int32_t q = h_cta_off[c]; q < h_cta_off[c + 1]; q++) {
                int sg = h_cta_seg[4 * (size_t)q];
                int ag = h_seg_agg[sg];
                if (ag != prev) {
                    h_cta_agg.push_back(ag);
                    prev = ag;
                    nloc++;
                }
                for (int k = 0; k < h_seg_count[sg]; k++) h_row_slot[members[h_seg_start[sg] + k]] = nloc - 1;
            }
            h_cta_agg_off[c + 1] = (int32_t)h_cta_agg.size();
        }
        if (h_cta_agg.empty()) h_cta_agg.push_back(0);
        CNT_CUDA(cudaMemcpyAsync(cta_agg_off, h_cta_agg_off.data(), 4 * h_cta_agg_off.size(), cudaMemcpyHostToDevice, st));
        CNT_CUDA(cudaMemcpyAsync(cta_agg, h_cta_agg.data(), 4 * h_cta_agg.size(), cudaMemcpyHostToDevice, st));
        CNT_CUDA(cudaMemcpyAsync(row_slot, h_row_slot.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CNT_CUDA(cudaMemcpyAsync(seg_home, h_seg_home.data(), 4 * h_seg_home.size(), cudaMemcpyHostToDevice, st));
        CNT_CUDA(cudaMemcpyAsync(cta_seg_off, h_cta_off.data(), 4 * h_cta_off.size(), cudaMemcpyHostToDevice, st));
        CNT_CUDA(cudaMemcpyAsync(cta_seg, h_cta_seg.data(), 4 * h_cta_seg.size(), cudaMemcpyHostToDevice, st));
    }
    // host vectors must outlive the async uploads
    t3 = tnow();
    CNT_CUDA(cudaStreamSynchronize(st));
    if (timing)
        fprintf(stderr, "[aggregates] n=%lld: device+D2H %.2f ms, run detection %.2f ms, CTA lists %.2f ms, uploads+kernels %.2f ms\n",
                (long long)n, t1 - t0, t2 - t1, t3 - t2, tnow() - t3);
    return CNT_OK;
}
