// Junction finder: cell-list binning (counting sort by cell) + FP64 intersection kernels.
// Replaces netsim.py:75-94 (check_intersect) and netsim.py:135-150 (KD-tree pair loop).
//
// Layout.  Sticks of all B networks are concatenated (soff).  Network b gets a G_b x G_b
// uniform grid over the unit box, G_b = clamp(isqrt(n_b/2), 1, 2048) (~2 sticks per cell).
// A stick whose length exceeds RMAX cells is "long" (the two electrodes, length 100, always
// are): long sticks form a prefix of the length-sorted table and are handled by a dense
// row kernel (they test every later stick, like the reference's radius-100 KD query).
// All other sticks are binned by the cell of their centre; the cell kernel gives each
// stick i (in cell order, so neighbouring threads share neighbourhoods in L1) the cells
// overlapping [c_i - l_i, c_i + l_i]^2 and tests the sticks j > i found there.
// Rows (junctions with stick1 == i) are produced by two passes (count, scan, fill) so the
// output is compact, deterministic and sorted by (network, stick1, stick2).
#include <vector>
#include "common.cuh"
#include "predicate.cuh"

namespace {

constexpr int RMAX = 6;          // max search radius in cells for binned sticks
constexpr int GMAX = 2048;
constexpr int DENSE_T = 512;     // threads of the dense-row kernel
constexpr int DENSE_ROWS = 8;    // grid.x of the dense-row kernel (rows strided)

struct __align__(16) StickRec {  // 64 bytes, one per binned stick, in cell order
    double m, b, xmin, xmax;     // CntLine
    double xc, yc, len;
    int32_t idx;                 // local post-sort index
    int32_t gid;                 // global stick id
};

struct Geo {                     // per-network grid geometry (host computed, uploaded)
    int32_t G;                   // cells per side
    int32_t coff;                // offset of the network's first cell in the global cell array
};

static int isqrt_i(int64_t v) {
    int64_t r = 0;
    while ((r + 1) * (r + 1) <= v) r++;
    return (int)r;
}
static int grid_dim_for(int64_t n) {
    int g = isqrt_i(n / 2);
    if (g < 1) g = 1;
    if (g > GMAX) g = GMAX;
    return g;
}

__device__ __forceinline__ int cell_of(double v, int G) {
    int c = (int)floor(v * (double)G);
    return c < 0 ? 0 : (c >= G ? G - 1 : c);
}

// ---- pass A: per-stick line records, long flags, cell histogram -----------------------
__global__ void k_prep(int B, const int32_t *__restrict__ soff, const Geo *__restrict__ geo, int64_t S,
                       const double *__restrict__ xa, const double *__restrict__ ya,
                       const double *__restrict__ xb, const double *__restrict__ yb,
                       const double *__restrict__ xc, const double *__restrict__ yc,
                       const double *__restrict__ len, CntLine *__restrict__ line, int32_t *__restrict__ cellkey,
                       int32_t *__restrict__ cell_count, int32_t *__restrict__ nlong, int32_t *__restrict__ errflag) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    int i = (int)s - soff[b], n = soff[b + 1] - soff[b];
    line[s] = cnt_make_line(xa[s], ya[s], xb[s], yb[s]);
    double l = len[s];
    if (i + 1 < n && l < len[s + 1]) atomicOr(errflag, 1);   // table must be sorted by length desc
    Geo g = geo[b];
    if (l * (double)g.G > (double)RMAX || !(l == l)) {
        cellkey[s] = -1;
        atomicAdd(nlong + b, 1);
    } else {
        int key = g.coff + cell_of(yc[s], g.G) * g.G + cell_of(xc[s], g.G);
        cellkey[s] = key;
        atomicAdd(cell_count + key, 1);
    }
}

// ---- pass B: scatter records into cell order -------------------------------------------
__global__ void k_bin(int64_t S, const int32_t *__restrict__ soff, int B, const int32_t *__restrict__ cellkey,
                      const int32_t *__restrict__ cell_start, int32_t *__restrict__ cell_fill,
                      const CntLine *__restrict__ line, const double *__restrict__ xc,
                      const double *__restrict__ yc, const double *__restrict__ len,
                      StickRec *__restrict__ rec) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int key = cellkey[s];
    if (key < 0) return;
    int b = cnt::find_segment(soff, B, (int)s);
    int pos = cell_start[key] + atomicAdd(cell_fill + key, 1);
    CntLine l = line[s];
    StickRec r;
    r.m = l.m; r.b = l.b; r.xmin = l.xmin; r.xmax = l.xmax;
    r.xc = xc[s]; r.yc = yc[s]; r.len = len[s];
    r.idx = (int)s - soff[b];
    r.gid = (int)s;
    rec[pos] = r;
}

// ---- cell kernel: one thread per binned stick (home stick i, candidates j > i) ---------
template <bool FILL>
__global__ void __launch_bounds__(128) k_cells(int B, const int32_t *__restrict__ soff, const Geo *__restrict__ geo,
                                               const int32_t *__restrict__ cell_start, int32_t total_cells,
                                               const StickRec *__restrict__ rec,
                                               int32_t *__restrict__ jcount, unsigned long long *__restrict__ ntests,
                                               const int32_t *__restrict__ jstart, const uint8_t *__restrict__ kind,
                                               int32_t *__restrict__ stick1, int32_t *__restrict__ stick2,
                                               double *__restrict__ jx, double *__restrict__ jy,
                                               uint8_t *__restrict__ kind2) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int nbinned = __ldg(cell_start + total_cells);
    bool active = p < nbinned;
    int cnt = 0;
    unsigned long long tests = 0;
    int b = -1;
    if (active) {
        StickRec me = rec[p];
        b = cnt::find_segment(soff, B, me.gid);
        Geo g = geo[b];
        CntLine l1{me.m, me.b, me.xmin, me.xmax};
        double G = (double)g.G;
        // cells overlapping the bounding square of the ball; 1e-6 cell of slack covers the
        // rounding of the products (cells are >= 1/2048 wide, the ball test is exact)
        int cx0 = (int)floor((me.xc - me.len) * G - 1e-6), cx1 = (int)floor((me.xc + me.len) * G + 1e-6);
        int cy0 = (int)floor((me.yc - me.len) * G - 1e-6), cy1 = (int)floor((me.yc + me.len) * G + 1e-6);
        cx0 = max(cx0, 0); cy0 = max(cy0, 0);
        cx1 = min(cx1, g.G - 1); cy1 = min(cy1, g.G - 1);
        int base = FILL ? jstart[me.gid] : 0;
        int s0 = soff[b];
        for (int cy = cy0; cy <= cy1; cy++) {
            int q0 = __ldg(cell_start + g.coff + cy * g.G + cx0);
            int q1 = __ldg(cell_start + g.coff + cy * g.G + cx1 + 1);
            for (int q = q0; q < q1; q++) {
                const StickRec *o = rec + q;
                int j = __ldg(&o->idx);
                if (j <= me.idx) continue;
                double oxc = __ldg(&o->xc), oyc = __ldg(&o->yc);
                if (!cnt_in_ball(me.xc, me.yc, me.len, oxc, oyc)) continue;
                tests++;
                CntLine l2{__ldg(&o->m), __ldg(&o->b), __ldg(&o->xmin), __ldg(&o->xmax)};
                double xi, yi;
                if (cnt_junction(l1, l2, &xi, &yi)) {
                    if (FILL) {
                        int w = base + cnt;
                        stick1[w] = me.idx;
                        stick2[w] = j;
                        jx[w] = xi;
                        jy[w] = yi;
                        kind2[w] = (uint8_t)(3 * kind[me.gid] + kind[s0 + j]);
                    }
                    cnt++;
                }
            }
        }
        if (FILL) {
            // order the row by stick2 (rows are short: insertion sort in place)
            for (int a = 1; a < cnt; a++) {
                int w = base + a;
                int kj = stick2[w];
                double kx = jx[w], ky = jy[w];
                uint8_t kk = kind2[w];
                int c = w - 1;
                while (c >= base && stick2[c] > kj) {
                    stick2[c + 1] = stick2[c]; jx[c + 1] = jx[c]; jy[c + 1] = jy[c]; kind2[c + 1] = kind2[c];
                    c--;
                }
                stick2[c + 1] = kj; jx[c + 1] = kx; jy[c + 1] = ky; kind2[c + 1] = kk;
            }
        } else {
            jcount[me.gid] = cnt;
        }
    }
    if (!FILL) {
        // per-network test counter: warp-aggregate when the warp sits in one network
        int b0 = __shfl_sync(0xffffffffu, b, 0);
        if (__all_sync(0xffffffffu, b == b0 || !active)) {
            unsigned long long t = tests;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if ((threadIdx.x & 31) == 0 && t && b0 >= 0) atomicAdd(ntests + b0, t);
        } else if (active && tests) {
            atomicAdd(ntests + b, tests);
        }
    }
}

// ---- dense rows: long stick i against every later stick of its network ------------------
template <bool FILL>
__global__ void __launch_bounds__(DENSE_T) k_dense(int B, const int32_t *__restrict__ soff,
                                                   const int32_t *__restrict__ nlong,
                                                   const CntLine *__restrict__ line, const double *__restrict__ xc,
                                                   const double *__restrict__ yc, const double *__restrict__ len,
                                                   int32_t *__restrict__ jcount, unsigned long long *__restrict__ ntests,
                                                   const int32_t *__restrict__ jstart, const uint8_t *__restrict__ kind,
                                                   int32_t *__restrict__ stick1, int32_t *__restrict__ stick2,
                                                   double *__restrict__ jx, double *__restrict__ jy,
                                                   uint8_t *__restrict__ kind2) {
    __shared__ int warp_cnt[DENSE_T / 32];
    __shared__ double red[32];
    int b = blockIdx.y;
    int s0 = soff[b], n = soff[b + 1] - s0;
    int nl = nlong[b];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = blockIdx.x; i < nl; i += gridDim.x) {
        CntLine l1 = line[s0 + i];
        double xci = xc[s0 + i], yci = yc[s0 + i], li = len[s0 + i];
        int base = FILL ? jstart[s0 + i] : 0;
        int written = 0;                       // identical in every thread
        unsigned long long tests = 0;
        for (int j0 = i + 1; j0 < n; j0 += DENSE_T) {
            int j = j0 + threadIdx.x;
            bool hit = false;
            double xi = 0, yi = 0;
            if (j < n && cnt_in_ball(xci, yci, li, xc[s0 + j], yc[s0 + j])) {
                tests++;
                CntLine l2 = line[s0 + j];
                hit = cnt_junction(l1, l2, &xi, &yi);
            }
            // ordered compaction: ballot + warp prefix + prefix over warps
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) warp_cnt[w] = __popc(m);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int k = 0; k < DENSE_T / 32; k++) {
                int c = warp_cnt[k];
                if (k < w) before += c;
                total += c;
            }
            if (FILL && hit) {
                int o = base + written + before + __popc(m & ((1u << lane) - 1u));
                stick1[o] = i;
                stick2[o] = j;
                jx[o] = xi;
                jy[o] = yi;
                kind2[o] = (uint8_t)(3 * kind[s0 + i] + kind[s0 + j]);
            }
            written += total;
            __syncthreads();
        }
        if (!FILL) {
            if (threadIdx.x == 0) jcount[s0 + i] = written;
            double t = cnt::block_sum((double)tests, red);   // exact: counts < 2^53
            if (threadIdx.x == 0 && t > 0) atomicAdd(ntests + b, (unsigned long long)t);
        }
    }
}

__global__ void k_joff(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ jstart,
                       int32_t *__restrict__ joff) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= B) joff[b] = jstart[soff[b]];
}

struct Plan {
    int64_t S, total_cells;
    std::vector<Geo> geo;
    // workspace carve-up
    Geo *d_geo;
    CntLine *line;
    StickRec *rec;
    int32_t *cellkey, *cell_count /* total_cells+1, scanned in place */, *cell_fill, *nlong, *errflag;
    void *scan_ws;
    int64_t scan_ws_bytes;
};

static int64_t plan_bytes(int B, int64_t S, int64_t total_cells) {
    using cnt::ws_bytes_for;
    return ws_bytes_for(B, sizeof(Geo)) + ws_bytes_for(S, sizeof(CntLine)) + ws_bytes_for(S, sizeof(StickRec)) +
           ws_bytes_for(S, 4) + ws_bytes_for(total_cells + 1, 4) + ws_bytes_for(total_cells + 1, 4) +
           ws_bytes_for(B, 4) + ws_bytes_for(1, 4) + cnt_scan_workspace_bytes(total_cells + 1);
}

static void make_geo(int B, const int32_t *h_soff, std::vector<Geo> &geo, int64_t &total_cells) {
    geo.resize(B);
    total_cells = 0;
    for (int b = 0; b < B; b++) {
        int g = grid_dim_for(h_soff[b + 1] - h_soff[b]);
        geo[b].G = g;
        geo[b].coff = (int32_t)total_cells;
        total_cells += (int64_t)g * g;
    }
}

static int carve(Plan &p, int B, const int32_t *h_soff, void *ws, int64_t ws_bytes) {
    p.S = h_soff[B];
    make_geo(B, h_soff, p.geo, p.total_cells);
    if (p.total_cells >= (int64_t)1 << 31) {
        cnt::set_error("too many cells");
        return CNT_EINVAL;
    }
    cnt::Workspace w(ws, ws_bytes);
    p.d_geo = w.take<Geo>(B);
    p.line = w.take<CntLine>(p.S);
    p.rec = w.take<StickRec>(p.S);
    p.cellkey = w.take<int32_t>(p.S);
    p.cell_count = w.take<int32_t>(p.total_cells + 1);
    p.cell_fill = w.take<int32_t>(p.total_cells + 1);
    p.nlong = w.take<int32_t>(B);
    p.errflag = w.take<int32_t>(1);
    p.scan_ws_bytes = cnt_scan_workspace_bytes(p.total_cells + 1);
    p.scan_ws = w.take<char>(p.scan_ws_bytes);
    if (!p.scan_ws) {
        cnt::set_error("junction workspace too small: have %lld need %lld", (long long)ws_bytes,
                       (long long)plan_bytes(B, p.S, p.total_cells));
        return CNT_ECAPACITY;
    }
    return CNT_OK;
}
}  // namespace

extern "C" int64_t cnt_junction_workspace_bytes(int B, const int32_t *h_soff) {
    if (B <= 0 || !h_soff) return 0;
    std::vector<Geo> geo;
    int64_t tc;
    make_geo(B, h_soff, geo, tc);
    return plan_bytes(B, h_soff[B], tc);
}

extern "C" int cnt_junction_count(int B, const int32_t *soff, const int32_t *h_soff, const double *xa,
                                  const double *ya, const double *xb, const double *yb, const double *xc,
                                  const double *yc, const double *length, int32_t *jcount,
                                  unsigned long long *ntests, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && h_soff && xa && ya && xb && yb && xc && yc && length && jcount && ntests && ws);
    CNT_CHECK_ARG(h_soff[0] == 0 && B <= 65535);
    for (int b = 0; b < B; b++) CNT_CHECK_ARG(h_soff[b + 1] >= h_soff[b]);
    Plan p;
    int rc = carve(p, B, h_soff, ws, ws_bytes);
    if (rc) return rc;
    if (p.S == 0) {
        CNT_CUDA(cudaMemsetAsync(ntests, 0, sizeof(unsigned long long) * B, st));
        return CNT_OK;
    }
    CNT_CUDA(cudaMemcpyAsync(p.d_geo, p.geo.data(), sizeof(Geo) * B, cudaMemcpyHostToDevice, st));
    CNT_CUDA(cudaMemsetAsync(p.cell_count, 0, 4 * (p.total_cells + 1), st));
    CNT_CUDA(cudaMemsetAsync(p.cell_fill, 0, 4 * (p.total_cells + 1), st));
    CNT_CUDA(cudaMemsetAsync(p.nlong, 0, 4 * B, st));
    CNT_CUDA(cudaMemsetAsync(p.errflag, 0, 4, st));
    CNT_CUDA(cudaMemsetAsync(ntests, 0, sizeof(unsigned long long) * B, st));
    CNT_CUDA(cudaMemsetAsync(jcount, 0, 4 * (p.S + 1), st));
    k_prep<<<cnt::grid_for(p.S, 256), 256, 0, st>>>(B, soff, p.d_geo, p.S, xa, ya, xb, yb, xc, yc, length, p.line,
                                                    p.cellkey, p.cell_count, p.nlong, p.errflag);
    CNT_LAUNCHED();
    rc = cnt_exclusive_scan_i32(p.cell_count, p.cell_count, p.total_cells, 1, p.scan_ws, p.scan_ws_bytes, stream);
    if (rc) return rc;
    k_bin<<<cnt::grid_for(p.S, 256), 256, 0, st>>>(p.S, soff, B, p.cellkey, p.cell_count, p.cell_fill, p.line, xc, yc,
                                                   length, p.rec);
    CNT_LAUNCHED();
    k_cells<false><<<cnt::grid_for(p.S, 128), 128, 0, st>>>(B, soff, p.d_geo, p.cell_count, (int32_t)p.total_cells,
                                                            p.rec, jcount, ntests, nullptr, nullptr, nullptr, nullptr,
                                                            nullptr, nullptr, nullptr);
    CNT_LAUNCHED();
    k_dense<false><<<dim3(DENSE_ROWS, B), DENSE_T, 0, st>>>(B, soff, p.nlong, p.line, xc, yc, length, jcount, ntests,
                                                            nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                            nullptr);
    CNT_LAUNCHED();
    return CNT_OK;
}

extern "C" int cnt_junction_fill(int B, const int32_t *soff, const int32_t *h_soff, const double *xa,
                                 const double *ya, const double *xb, const double *yb, const double *xc,
                                 const double *yc, const double *length, const uint8_t *kind,
                                 const int32_t *jstart, int64_t J, int32_t *stick1, int32_t *stick2, double *jx,
                                 double *jy, uint8_t *kind2, int32_t *joff, void *ws, int64_t ws_bytes,
                                 void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && h_soff && kind && jstart && joff && ws && J >= 0);
    CNT_CHECK_ARG(J == 0 || (stick1 && stick2 && jx && jy && kind2));
    Plan p;   // same carve-up as _count: the binned records built there are reused
    int rc = carve(p, B, h_soff, ws, ws_bytes);
    if (rc) return rc;
    k_joff<<<cnt::grid_for(B + 1, 128), 128, 0, st>>>(B, soff, jstart, joff);
    CNT_LAUNCHED();
    if (p.S == 0 || J == 0) return CNT_OK;
    k_cells<true><<<cnt::grid_for(p.S, 128), 128, 0, st>>>(B, soff, p.d_geo, p.cell_count, (int32_t)p.total_cells,
                                                           p.rec, nullptr, nullptr, jstart, kind, stick1, stick2, jx,
                                                           jy, kind2);
    CNT_LAUNCHED();
    k_dense<true><<<dim3(DENSE_ROWS, B), DENSE_T, 0, st>>>(B, soff, p.nlong, p.line, xc, yc, length, nullptr, nullptr,
                                                           jstart, kind, stick1, stick2, jx, jy, kind2);
    CNT_LAUNCHED();
    return CNT_OK;
}

// sortedness flag of the last _count on this workspace (1 = a network was not sorted by
// length descending); device pointer so the caller can fetch it with its own D2H copy.
extern "C" int cnt_junction_errflag_ptr(int B, const int32_t *h_soff, void *ws, int64_t ws_bytes, int32_t **out) {
    Plan p;
    int rc = carve(p, B, h_soff, ws, ws_bytes);
    if (rc) return rc;
    *out = p.errflag;
    return CNT_OK;
}
