// Junction finder: cell-list binning (counting sort by cell) + FP64 intersection kernels.
// Replaces netsim.py:75-94 (check_intersect) and netsim.py:135-150 (KD-tree pair loop).
//
// Layout.  Sticks of all B networks are concatenated (soff).  Network b gets a G_b x G_b
// uniform grid over the unit box, G_b = clamp(isqrt(n_b/2), 1, 2048) (~2 sticks per cell).
// Three length classes (is_dense / is_wide below): the electrodes (longer than a quarter of the box) are DENSE rows
// -- they test every later stick, like the reference's radius-100 KD query -- cut into segments (k_dense); everything
// else is binned by the cell of its centre; of those, the few sticks longer than RMAX cells are WIDE: candidates
// like everybody else, but their own row is walked by a warp over their cell range (k_wide).  All other sticks are
// home sticks of the tile kernel (k_tiles): one CTA per tile of 8 x 8 cells stages the records (centre, length, ids)
// of the tile and of the halo its sticks reach into SHARED MEMORY once, flattens the (home stick, candidate) pairs
// of the tile over all threads (cheap part: j > i and the KD-ball test), compacts the survivors with warp ballot /
// prefix into a queue, runs the FP64 predicate -- the two IEEE divisions -- on the queue with full warps (line
// records of the survivors through L1 / L2), and writes the tile's junctions grouped by home stick to a temporary
// table.  The predicate is evaluated ONCE: after the scan of the row lengths k_place copies every row to its place
// and orders it by stick2, so the output is compact, deterministic and sorted by (network, stick1, stick2).  A batch
// of home sticks that runs out of hit buffer (more than CAP_HIT junctions) or of temporary table space (more than ~8
// junctions per stick overall) only counts; exactly its rows are then recomputed in place by the thread-per-row kernel
// k_cells_fill.  cnt_junction_*_slab shard the search by spatial slab over ranks (slab_owner below).
#include <vector>
#include "common.cuh"
#include "predicate.cuh"

namespace {

constexpr int RMAX = 6;          // max search radius in cells for binned sticks
constexpr int GMAX = 2048;
constexpr int DENSE_T = 512;     // threads of the dense-row kernel
constexpr int DENSE_ROWS = 8;    // grid.x of the dense-row kernel (rows strided)

struct __align__(16) RecA {      // 32 bytes: what the cheap tests need (one sector)
    double xc, yc, len;
    int32_t idx, gid;
};

#ifndef CNT_TILE_C
#define CNT_TILE_C 8
#endif
#ifndef CNT_TILE_HOME
#define CNT_TILE_HOME 256
#endif
constexpr int TILE_C = CNT_TILE_C;     // home cells per tile side
// shape of the tile kernel (A/B on the GPU box, tools/gpu_tile_ab.sh; junction count pass per 56 networks):
// 512 threads / 2048 pairs / 1024 hits + staged line records (105 KB, 2 CTAs per SM) 4.18 ms; without staged line
// records 4.49 (512 threads, 2 CTAs per SM by registers), 4.03 (384 threads), 3.86 ms (256 threads / 1024 / 512: 53 KB,
// 4 CTAs per SM -- four independent tiles per SM hide each other's barriers)
#ifndef CNT_TILE_CT
#define CNT_TILE_CT 256
#endif
#ifndef CNT_TILE_Q
#define CNT_TILE_Q 1024
#endif
#ifndef CNT_TILE_HIT
#define CNT_TILE_HIT 512
#endif
constexpr int CT = CNT_TILE_CT;  // threads of the tile kernel
constexpr int CAP_REC = 1024;    // staged records per tile (32 KB)
constexpr int CAP_HOME = CNT_TILE_HOME;   // home sticks per batch of a tile
constexpr int MAXRUN = 1024;           // runs (home stick, cell row) per batch of home sticks
constexpr int WQ = 64;                 // per-warp survivor queue (ring buffer: < 32 left over + 32 new)
constexpr int CAP_Q = (CNT_TILE_CT / 32) * WQ;     // all queues of a CTA
constexpr int CAP_HIT = CNT_TILE_HIT;  // junctions per batch of home sticks
static_assert(CAP_HOME <= CT, "the scans over a batch of home sticks use one thread per stick");
constexpr int MAX_RH = TILE_C + 2 * (RMAX + 2);      // rows / columns of the staged region

struct Geo {                     // per-network grid geometry (host computed, uploaded)
    int32_t G;                   // cells per side
    int32_t coff;                // offset of the network's first cell in the global cell array
    int32_t toff;                // offset of the network's first tile in the global tile list
    int32_t TG;                  // tiles per side
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

// Length classes of a stick in a grid of G cells per side: "dense" (the electrodes: longer than a quarter of the box,
// or than RMAX cells in a small grid) -> rows of k_dense, not binned; "wide" (longer than RMAX cells: a handful of
// body sticks out of millions, the tail of the length distribution) -> binned like everybody else, so shorter home
// sticks find them, but their own row is walked by a warp over their (large) cell range in k_wide instead of a
// tile; everything else is a tile home stick.
__device__ __forceinline__ bool is_dense(double len, int G) {
    const double lg = len * (double)G, lim = (double)(G / 4 > RMAX ? G / 4 : RMAX);
    return lg > lim || !(len == len);
}
__device__ __forceinline__ bool is_wide(double len, int G) { return len * (double)G > (double)RMAX && !is_dense(len, G); }

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
                       int32_t *__restrict__ cell_count, int32_t *__restrict__ nlong, int32_t *__restrict__ errflag,
                       int32_t *__restrict__ nwide, int32_t *__restrict__ widelist) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    int i = (int)s - soff[b], n = soff[b + 1] - soff[b];
    line[s] = cnt_make_line(xa[s], ya[s], xb[s], yb[s]);
    double l = len[s];
    if (i + 1 < n && l < len[s + 1]) atomicOr(errflag, 1);   // table must be sorted by length desc
    Geo g = geo[b];
    if (is_dense(l, g.G)) {
        cellkey[s] = -1;
        atomicAdd(nlong + b, 1);
    } else {
        if (is_wide(l, g.G)) widelist[atomicAdd(nwide, 1)] = (int)s;      // order irrelevant: every row stands alone
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
                      RecA *__restrict__ recA) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int key = cellkey[s];
    if (key < 0) return;
    int b = cnt::find_segment(soff, B, (int)s);
    int pos = cell_start[key] + atomicAdd(cell_fill + key, 1);
    RecA r;
    r.xc = xc[s]; r.yc = yc[s]; r.len = len[s];
    r.idx = (int)s - soff[b];
    r.gid = (int)s;
    recA[pos] = r;                // the line records stay in stick order: only survivors of the ball test read them
}

// cells overlapping the bounding square of the ball of a stick; 1e-6 cell of slack covers the rounding
// of the products (cells are >= 1/2048 wide, the ball test is exact)
__device__ __forceinline__ void cell_range(const RecA &me, int G, int &cx0, int &cx1, int &cy0, int &cy1) {
    const double g = (double)G;
    cx0 = (int)floor((me.xc - me.len) * g - 1e-6); cx1 = (int)floor((me.xc + me.len) * g + 1e-6);
    cy0 = (int)floor((me.yc - me.len) * g - 1e-6); cy1 = (int)floor((me.yc + me.len) * g + 1e-6);
    cx0 = max(cx0, 0); cy0 = max(cy0, 0);
    cx1 = min(cx1, G - 1); cy1 = min(cy1, G - 1);
}

// ---- tile kernel: one CTA per tile of TILE_C x TILE_C cells ---------------------------------
// Slab sharding of ONE junction search over several ranks (config 5: a single huge network on several GPUs).  The unit
// box is cut into `world` vertical slabs of whole tile columns; a rank evaluates the tiles -- i.e. the home sticks
// whose centre lies -- in its slab, reading the halo sticks around it from the (replicated) binned table, and leaves
// the rows of all other home sticks empty.  Rank 0 also takes the long sticks (electrodes).  The caller adds the
// per-rank jcount / ntests and the per-rank tables (disjoint supports, everything else zero) over the ranks.
__host__ __device__ __forceinline__ int slab_owner(int tile_col, int TG, int world) {
    return (int)(((long long)tile_col * world) / TG);
}

struct Hit {
    int32_t hr;                  // home stick (batch-local) << 16 | arrival rank inside its row
    int32_t j;
    double x, y;
};

__global__ void __launch_bounds__(CT) k_tiles(int B, const Geo *__restrict__ geo,
                                              const int32_t *__restrict__ cell_start, const RecA *__restrict__ gA,
                                              const CntLine *__restrict__ line, int32_t *__restrict__ jcount,
                                              int32_t *__restrict__ jtmp, unsigned long long *__restrict__ ntests,
                                              int32_t *__restrict__ tj, double *__restrict__ tx, double *__restrict__ ty,
                                              int64_t cap_tmp, int32_t *__restrict__ ntmp, int32_t *__restrict__ errflag,
                                              int shard_rank, int shard_world) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_xc = (double *)smem_raw;                            // [CAP_REC] staged records of the region, SoA: centre,
    double *s_yc = s_xc + CAP_REC;                                // [CAP_REC]
    double *s_hlen = s_yc + CAP_REC;                              // [CAP_HOME] length of every home stick of the batch
    Hit *s_hit = (Hit *)(s_hlen + CAP_HOME);                      // [CAP_HIT]
    int32_t *s_idx = (int32_t *)(s_hit + CAP_HIT);                // [CAP_REC] ... and network-local stick index
    int32_t *s_q = s_idx + CAP_REC;                               // [CT / 32][WQ] survivor queues: home << 23 | region index
    int32_t *s_cs = s_q + CAP_Q;                                  // [MAX_RH * (MAX_RH + 1)] cell starts (region index)
    int32_t *s_gq0 = s_cs + MAX_RH * (MAX_RH + 1);                // [MAX_RH] first global record of every region row
    int32_t *s_rb = s_gq0 + MAX_RH;                               // [MAX_RH + 1] first region index of every region row
    int32_t *s_rpre = s_rb + MAX_RH + 1;                          // [MAXRUN + 1] first pair of every run of the batch
    int32_t *s_rpack = s_rpre + MAXRUN + 1;                       // [MAXRUN] run: home << 23 | first region index
    int32_t *s_cnt = s_rpack + MAXRUN;                            // [CAP_HOME + 1] junctions per home stick / row offsets
    int32_t *s_hq = s_cnt + CAP_HOME + 1;                         // [CAP_HOME] region index of every home stick
    __shared__ int s_red[8];
    __shared__ int s_nhit, s_base, s_warp[CT / 32], s_warp2[CT / 32], s_P;
    __shared__ unsigned long long s_tests;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int b = 0;
    {
        int lo = 0, hi = B;                                       // geo[lo].toff <= blockIdx.x < geo[hi].toff
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (geo[mid].toff <= (int)blockIdx.x) lo = mid; else hi = mid;
        }
        b = lo;
    }
    const Geo g = geo[b];
    const int tloc = blockIdx.x - g.toff, tyi = tloc / g.TG, txi = tloc - tyi * g.TG;
    if (shard_world > 1 && slab_owner(txi, g.TG, shard_world) != shard_rank) return;     // another rank's slab
    const int hx0 = txi * TILE_C, hy0 = tyi * TILE_C;
    const int hx1 = min(hx0 + TILE_C, g.G) - 1, hy1 = min(hy0 + TILE_C, g.G) - 1;
    const int32_t *cs = cell_start + g.coff;
    // home sticks: the records of the home cells, row by row
    __shared__ int s_nhrow[TILE_C + 1], s_hq0[TILE_C];
    if (wid == 0) {                                               // one lane per home row: loads side by side, warp scan
        int q0 = 0, cnt_r = 0;
        if (lane < TILE_C && hy0 + lane <= hy1) {
            q0 = __ldg(cs + (hy0 + lane) * g.G + hx0);
            cnt_r = __ldg(cs + (hy0 + lane) * g.G + hx1 + 1) - q0;
        }
        int inc = cnt_r;
#pragma unroll
        for (int o = 1; o < TILE_C; o <<= 1) {
            const int t = __shfl_up_sync(~0u, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane < TILE_C) {
            s_hq0[lane] = q0;
            s_nhrow[lane] = inc - cnt_r;
        }
        if (lane == TILE_C - 1) s_nhrow[TILE_C] = inc;
    }
    __syncthreads();
    const int nhome = s_nhrow[TILE_C];
    if (nhome == 0) return;
    auto home_row = [&](int h) {
        int r = 0;
#pragma unroll
        for (int k = 1; k < TILE_C; k++) r += h >= s_nhrow[k];
        return r;
    };
    auto home_q = [&](int h) {                                    // global record index of home stick h of the tile
        const int r = home_row(h);
        return s_hq0[r] + (h - s_nhrow[r]);
    };
    // region = union of the cell ranges of the home sticks
    int rx0 = g.G, rx1 = -1, ry0 = g.G, ry1 = -1;
    for (int h = tid; h < nhome; h += CT) {
        const RecA me = gA[home_q(h)];
        if (is_wide(me.len, g.G)) continue;                       // its row is k_wide's
        int cx0, cx1, cy0, cy1;
        cell_range(me, g.G, cx0, cx1, cy0, cy1);
        rx0 = min(rx0, cx0); rx1 = max(rx1, cx1); ry0 = min(ry0, cy0); ry1 = max(ry1, cy1);
    }
    if (tid == 0) {
        s_red[0] = g.G; s_red[1] = -1; s_red[2] = g.G; s_red[3] = -1;
        s_tests = 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rx0 = min(rx0, __shfl_xor_sync(~0u, rx0, o)); rx1 = max(rx1, __shfl_xor_sync(~0u, rx1, o));
        ry0 = min(ry0, __shfl_xor_sync(~0u, ry0, o)); ry1 = max(ry1, __shfl_xor_sync(~0u, ry1, o));
    }
    if (lane == 0) {
        atomicMin(&s_red[0], rx0); atomicMax(&s_red[1], rx1); atomicMin(&s_red[2], ry0); atomicMax(&s_red[3], ry1);
    }
    __syncthreads();
    rx0 = s_red[0]; rx1 = s_red[1]; ry0 = s_red[2]; ry1 = s_red[3];
    if (rx1 < rx0) return;                                        // only wide home sticks in this tile (uniform)
    const int RW = rx1 - rx0 + 1, RH = ry1 - ry0 + 1;            // <= MAX_RH by construction (len * G <= RMAX)
    // region rows: global start, region-index start; cell starts as region indices
    static_assert(MAX_RH <= 32, "region rows are scanned by one warp");
    if (wid == 0) {                                               // one lane per region row
        int q0 = 0, cnt_r = 0;
        if (lane < RH) {
            q0 = __ldg(cs + (ry0 + lane) * g.G + rx0);
            cnt_r = __ldg(cs + (ry0 + lane) * g.G + rx1 + 1) - q0;
            s_gq0[lane] = q0;
        }
        int inc = cnt_r;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(~0u, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane < RH) s_rb[lane] = inc - cnt_r;
        if (lane == RH - 1) s_rb[RH] = inc;
    }
    __syncthreads();
    const int nrec = s_rb[RH];
    for (int k = tid; k < RH * (RW + 1); k += CT) {
        const int r = k / (RW + 1), c = k - r * (RW + 1);
        s_cs[k] = s_rb[r] + (__ldg(cs + (ry0 + r) * g.G + rx0 + c) - s_gq0[r]);
    }
    const bool staged = nrec <= CAP_REC;
    int s0net;                                                    // first stick of the network: gid = s0net + idx
    {
        const RecA r0 = gA[home_q(0)];
        s0net = r0.gid - r0.idx;
    }
    if (staged) {
        for (int k = tid; k < nrec; k += CT) {
            int r = 0;
            while (k >= s_rb[r + 1]) r++;
            const RecA rec = gA[s_gq0[r] + (k - s_rb[r])];
            s_xc[k] = rec.xc;
            s_yc[k] = rec.yc;
            s_idx[k] = rec.idx;
        }
    }
    __syncthreads();
    // record accessor: region index -> centre and index (shared memory when staged, global memory otherwise)
    auto cand = [&](int v, double &xc, double &yc, int &idx) {
        if (staged) {
            xc = s_xc[v];
            yc = s_yc[v];
            idx = s_idx[v];
            return;
        }
        int r = 0;
        while (v >= s_rb[r + 1]) r++;
        const RecA rec = gA[s_gq0[r] + (v - s_rb[r])];
        xc = rec.xc;
        yc = rec.yc;
        idx = rec.idx;
    };
    unsigned long long tests = 0;
    for (int hb = 0; hb < nhome;) {                              // batches of home sticks
        const int nb0 = min(CAP_HOME, nhome - hb);
        // candidate pairs and runs (one per cell row of its range) of every home stick
        int ncand = 0, nrows = 0, rng = 0;
        double mylen = 0.0;
        if (tid < nb0) {
            const int q = home_q(hb + tid);
            s_cnt[tid] = 0;
            s_hq[tid] = -1;                                       // wide stick: no pairs here, its row is k_wide's
            const RecA me = gA[q];
            mylen = me.len;
            if (!is_wide(me.len, g.G)) {
                // region index of the home stick: its cell row inside the region
                const int hr = home_row(hb + tid);
                s_hq[tid] = s_rb[hy0 + hr - ry0] + (q - s_gq0[hy0 + hr - ry0]);
                int cx0, cx1, cy0, cy1;
                cell_range(me, g.G, cx0, cx1, cy0, cy1);
                for (int cy = cy0; cy <= cy1; cy++)
                    ncand += s_cs[(cy - ry0) * (RW + 1) + (cx1 + 1 - rx0)] - s_cs[(cy - ry0) * (RW + 1) + (cx0 - rx0)];
                nrows = cy1 - cy0 + 1;
                rng = (cx0 - rx0) | ((cx1 - rx0) << 8) | ((cy0 - ry0) << 16) | ((cy1 - ry0) << 24);
            }
        }
        // exclusive scans of both counts over the batch (nb0 <= CAP_HOME <= CT)
        int inc = ncand, incr = nrows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(~0u, inc, o), tr = __shfl_up_sync(~0u, incr, o);
            if (lane >= o) {
                inc += t;
                incr += tr;
            }
        }
        if (lane == 31) {
            s_warp[wid] = inc;
            s_warp2[wid] = incr;
        }
        if (tid == 0) s_nhit = 0;
        __syncthreads();
        if (wid == 0) {
            int w = lane < CT / 32 ? s_warp[lane] : 0, wi = w, w2 = lane < CT / 32 ? s_warp2[lane] : 0, wi2 = w2;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(~0u, wi, o), t2 = __shfl_up_sync(~0u, wi2, o);
                if (lane >= o) {
                    wi += t;
                    wi2 += t2;
                }
            }
            if (lane < CT / 32) {
                s_warp[lane] = wi - w;
                s_warp2[lane] = wi2 - w2;
            }
        }
        __syncthreads();
        const int pre = s_warp[wid] + inc - ncand, rpre = s_warp2[wid] + incr - nrows;     // first pair / run of this home
        // the batch ends where the run table is full (prefixes are monotone: the home sticks that fit are a prefix)
        const int nb = __syncthreads_count(tid < nb0 && rpre + nrows <= MAXRUN);
        if (tid < nb) {
            s_hlen[tid] = mylen;
            const int cx0 = rng & 255, cx1 = (rng >> 8) & 255, cy0 = (rng >> 16) & 255;
            int pp = pre;
            for (int k = 0; k < nrows; k++) {
                const int a0 = s_cs[(cy0 + k) * (RW + 1) + cx0], a1 = s_cs[(cy0 + k) * (RW + 1) + cx1 + 1];
                s_rpre[rpre + k] = pp;
                s_rpack[rpre + k] = (tid << 23) | a0;
                pp += a1 - a0;
            }
            if (tid == nb - 1) {
                s_rpre[rpre + nrows] = pp;                        // sentinel: total pairs of the batch
                s_P = pp;
                s_warp2[0] = rpre + nrows;                        // runs of the batch
            }
        }
        __syncthreads();
        // The P candidate pairs of the batch are cut into one contiguous range per warp.  A warp walks its range 32
        // pairs at a time; pair p lies in run r (s_rpre[r] <= p < s_rpre[r + 1]: a short forward search from the run of
        // the warp's first pair) and is candidate a0 + (p - s_rpre[r]) of that run's home stick -- runs are contiguous in
        // the region.  Cheap tests (j > i, KD ball), survivors compacted by ballot into the warp's OWN queue; as soon as
        // the queue holds 32 entries the warp evaluates the FP64 predicate on them with all lanes.  No CTA-wide barrier
        // between the two phases.
        const int P = s_P, nruns = s_warp2[0];
        {
            constexpr int NW = CT / 32;
            const int per = ((P + NW * 32 - 1) / (NW * 32)) * 32;            // pairs per warp, a multiple of 32
            const int pbeg = wid * per, pend = min(P, pbeg + per);
            int32_t *wq = s_q + wid * WQ;
            int qh = 0, qn = 0;                                              // ring buffer: head, entries (warp-uniform)
            auto emit = [&](int entry) {
                const int h = entry >> 23, v = entry & ((1 << 23) - 1);
                double xh, yh, xo, yo;
                int ih, io;
                cand(s_hq[h], xh, yh, ih);
                cand(v, xo, yo, io);
                const CntLine l1 = line[s0net + ih], l2 = line[s0net + io];
                double xi, yi;
                if (cnt_junction(l1, l2, &xi, &yi)) {
                    const int rank = atomicAdd(&s_cnt[h], 1);
                    const int at = atomicAdd(&s_nhit, 1);
                    if (at < CAP_HIT) {
                        Hit hh;
                        hh.hr = (h << 16) | (rank & 0xFFFF);
                        hh.j = io;
                        hh.x = xi;
                        hh.y = yi;
                        s_hit[at] = hh;
                    }
                }
            };
            int rcur = 0;
            if (pbeg < pend) {                                               // run of the warp's first pair
                int lo = 0, hi = nruns - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_rpre[mid + 1] <= pbeg) lo = mid + 1; else hi = mid;
                }
                rcur = lo;
            }
            for (int pp0 = pbeg; pp0 < pend; pp0 += 32) {
                while (s_rpre[rcur + 1] <= pp0) rcur++;                      // uniform
                const int pp = pp0 + lane;
                bool pass = false;
                int entry = 0;
                if (pp < pend) {
                    int r = rcur;
                    while (s_rpre[r + 1] <= pp) r++;                         // a few steps further for this lane
                    const int pack = s_rpack[r];
                    const int h = pack >> 23, v = (pack & ((1 << 23) - 1)) + (pp - s_rpre[r]);
                    double xh, yh, xo, yo;
                    int ih, io;
                    cand(s_hq[h], xh, yh, ih);
                    cand(v, xo, yo, io);
                    pass = io > ih && cnt_in_ball(xh, yh, s_hlen[h], xo, yo);
                    entry = (h << 23) | v;
                }
                const unsigned m = __ballot_sync(~0u, pass);
                if (pass) {
                    wq[(qh + qn + __popc(m & ((1u << lane) - 1u))) & (WQ - 1)] = entry;
                    tests++;
                }
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    emit(wq[(qh + lane) & (WQ - 1)]);
                    qh = (qh + 32) & (WQ - 1);
                    qn -= 32;
                    __syncwarp();
                }
            }
            if (lane < qn) emit(wq[(qh + lane) & (WQ - 1)]);
        }
        __syncthreads();
        // rows of the batch: offsets (exclusive scan of s_cnt), space in the temporary table, write-out
        const int nhit = s_nhit;
        int cnt_h = tid < nb ? s_cnt[tid] : 0;
        int inc2 = cnt_h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(~0u, inc2, o);
            if (lane >= o) inc2 += t;
        }
        __syncthreads();
        if (lane == 31) s_warp[wid] = inc2;
        __syncthreads();
        if (wid == 0) {
            int w = lane < CT / 32 ? s_warp[lane] : 0, wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(~0u, wi, o);
                if (lane >= o) wi += t;
            }
            if (lane < CT / 32) s_warp[lane] = wi - w;
        }
        __syncthreads();
        const int roff = s_warp[wid] + inc2 - cnt_h;
        if (tid == 0) {
            bool ok = nhit <= CAP_HIT;
            int base = -1;
            if (ok && nhit > 0) {
                base = atomicAdd(ntmp, nhit);
                if ((int64_t)base + nhit > cap_tmp) ok = false;
            }
            if (!ok) atomicOr(errflag, 2);                        // the rows of this batch will be recomputed by k_cells_fill
            s_base = ok ? base : -1;
        }
        __syncthreads();
        const int base = s_base;
        if (tid < nb) {
            if (s_hq[tid] >= 0) {
                double xh, yh;
                int ih;
                cand(s_hq[tid], xh, yh, ih);
                const int gid = s0net + ih;
                jcount[gid] = cnt_h;
                jtmp[gid] = base >= 0 ? base + roff : -1;
            }
            s_cnt[tid] = roff;                                    // row offset from now on
        }
        __syncthreads();
        if (base >= 0) {
            for (int k = tid; k < nhit; k += CT) {
                const Hit hh = s_hit[k];
                const int at = base + s_cnt[hh.hr >> 16] + (hh.hr & 0xFFFF);
                tj[at] = hh.j;
                tx[at] = hh.x;
                ty[at] = hh.y;
            }
        }
        __syncthreads();
        hb += nb;
    }
    // candidate pairs inside the ball (the reference's KD-tree query count), one atomic per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tests += __shfl_down_sync(~0u, tests, o);
    if (lane == 0 && tests) atomicAdd(&s_tests, tests);
    __syncthreads();
    if (tid == 0 && s_tests) atomicAdd(ntests + b, s_tests);
}

// every binned stick's row from the temporary table to its place, ordered by stick2
__global__ void k_place(int B, const int32_t *__restrict__ soff, int64_t S, const int32_t *__restrict__ jtmp,
                        const int32_t *__restrict__ jstart /* [S + 1] */,
                        const int32_t *__restrict__ tj, const double *__restrict__ tx, const double *__restrict__ ty,
                        const uint8_t *__restrict__ kind, const int32_t *__restrict__ errflag,
                        int32_t *__restrict__ stick1, int32_t *__restrict__ stick2, double *__restrict__ jx,
                        double *__restrict__ jy, uint8_t *__restrict__ kind2) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int t0 = jtmp[s], base = jstart[s], cnt = jstart[s + 1] - base;
    if (t0 < 0 || cnt == 0) return;                               // long stick (k_dense), empty row, or a row whose tile
                                                                  // ran out of buffer space (k_cells_fill recomputes it)
    const int b = cnt::find_segment(soff, B, (int)s);
    const int s0 = soff[b], i = (int)s - s0;
    const uint8_t ki = kind[s];
    constexpr int NREG = 8;
    if (cnt <= NREG) {
        // the common case (1.3 junctions per row on average): the row is read once, ordered in registers by an
        // odd-even transposition network (unused places hold the largest key) and written once -- the in-place
        // insertion below is a chain of dependent global loads and stores
        int kj[NREG];
        double kx[NREG], ky[NREG];
#pragma unroll
        for (int a = 0; a < NREG; a++) {
            kj[a] = 0x7fffffff;
            kx[a] = ky[a] = 0.0;
            if (a < cnt) {
                kj[a] = tj[t0 + a];
                kx[a] = tx[t0 + a];
                ky[a] = ty[t0 + a];
            }
        }
#pragma unroll
        for (int r = 0; r < NREG; r++) {
#pragma unroll
            for (int a = r & 1; a + 1 < NREG; a += 2) {
                if (kj[a] > kj[a + 1]) {
                    const int tjj = kj[a]; kj[a] = kj[a + 1]; kj[a + 1] = tjj;
                    const double txx = kx[a]; kx[a] = kx[a + 1]; kx[a + 1] = txx;
                    const double tyy = ky[a]; ky[a] = ky[a + 1]; ky[a + 1] = tyy;
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NREG; a++) {
            if (a < cnt) {
                stick1[base + a] = i;
                stick2[base + a] = kj[a];
                jx[base + a] = kx[a];
                jy[base + a] = ky[a];
                kind2[base + a] = (uint8_t)(3 * ki + kind[s0 + kj[a]]);
            }
        }
        return;
    }
    for (int a = 0; a < cnt; a++) {                               // insertion by stick2 (longer rows)
        const int kj = tj[t0 + a];
        const double kx = tx[t0 + a], ky = ty[t0 + a];
        int c = base + a - 1;
        while (c >= base && stick2[c] > kj) {
            stick2[c + 1] = stick2[c]; jx[c + 1] = jx[c]; jy[c + 1] = jy[c];
            c--;
        }
        stick2[c + 1] = kj; jx[c + 1] = kx; jy[c + 1] = ky;
    }
    for (int a = 0; a < cnt; a++) {
        stick1[base + a] = i;
        kind2[base + a] = (uint8_t)(3 * ki + kind[s0 + stick2[base + a]]);
    }
}

// ---- fallback: the rows of home-stick batches that ran out of hit buffer / temporary table space (jtmp < 0) are
// recomputed in place, one thread per row; errflag bit 1 says whether there is any such row ----
__global__ void __launch_bounds__(128) k_cells_fill(int B, const int32_t *__restrict__ soff, const Geo *__restrict__ geo,
                                                    const int32_t *__restrict__ cell_start, int32_t total_cells,
                                                    const RecA *__restrict__ recA, const CntLine *__restrict__ line,
                                                    const int32_t *__restrict__ errflag, const int32_t *__restrict__ jtmp,
                                                    const int32_t *__restrict__ jstart, const uint8_t *__restrict__ kind,
                                                    int32_t *__restrict__ stick1, int32_t *__restrict__ stick2,
                                                    double *__restrict__ jx, double *__restrict__ jy,
                                                    uint8_t *__restrict__ kind2, int shard_rank, int shard_world) {
    if (!(*errflag & 2)) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int nbinned = __ldg(cell_start + total_cells);
    if (p >= nbinned) return;
    const RecA me = recA[p];
    if (jtmp[me.gid] >= 0 || jstart[me.gid + 1] == jstart[me.gid]) return;       // placed by k_place / nothing to place
    const int b = cnt::find_segment(soff, B, me.gid);
    const Geo g = geo[b];
    if (is_wide(me.len, g.G)) return;                             // k_wide's row
    if (shard_world > 1 && slab_owner(cell_of(me.xc, g.G) / TILE_C, g.TG, shard_world) != shard_rank) return;
    const CntLine l1 = line[me.gid];
    int cx0, cx1, cy0, cy1;
    cell_range(me, g.G, cx0, cx1, cy0, cy1);
    const int base = jstart[me.gid], s0 = soff[b];
    int cnt = 0;
    for (int cy = cy0; cy <= cy1; cy++) {
        const int q0 = __ldg(cell_start + g.coff + cy * g.G + cx0);
        const int q1 = __ldg(cell_start + g.coff + cy * g.G + cx1 + 1);
        for (int q = q0; q < q1; q++) {
            const RecA o = recA[q];
            if (o.idx <= me.idx || !cnt_in_ball(me.xc, me.yc, me.len, o.xc, o.yc)) continue;
            double xi, yi;
            if (cnt_junction(l1, line[o.gid], &xi, &yi)) {
                // insertion by stick2 (rows are short)
                int c = base + cnt - 1;
                while (c >= base && stick2[c] > o.idx) {
                    stick2[c + 1] = stick2[c]; jx[c + 1] = jx[c]; jy[c + 1] = jy[c];
                    c--;
                }
                stick2[c + 1] = o.idx; jx[c + 1] = xi; jy[c + 1] = yi;
                cnt++;
            }
        }
    }
    for (int a = 0; a < cnt; a++) {
        stick1[base + a] = me.idx;
        kind2[base + a] = (uint8_t)(3 * kind[me.gid] + kind[s0 + stick2[base + a]]);
    }
}

// ---- wide rows: one warp per wide stick walks the records of its cell range (any radius) ----------------
// COUNT: junctions and in-ball tests of the row; FILL: the row in walk order (deterministic: cells in order, lanes
// in order), then lane 0 orders it by stick2 (a wide stick crosses a few dozen others).
template <bool FILL>
__global__ void __launch_bounds__(256) k_wide(int B, const int32_t *__restrict__ soff, const Geo *__restrict__ geo,
                                              const int32_t *__restrict__ cell_start, const RecA *__restrict__ recA,
                                              const int32_t *__restrict__ nwide,
                                              const int32_t *__restrict__ widelist, const CntLine *__restrict__ line,
                                              const double *__restrict__ xc, const double *__restrict__ yc,
                                              const double *__restrict__ len, int32_t *__restrict__ jcount,
                                              unsigned long long *__restrict__ ntests, const int32_t *__restrict__ jstart,
                                              const uint8_t *__restrict__ kind, int32_t *__restrict__ stick1,
                                              int32_t *__restrict__ stick2, double *__restrict__ jx, double *__restrict__ jy,
                                              uint8_t *__restrict__ kind2, int shard_rank, int shard_world) {
    const int lane = threadIdx.x & 31;
    const int nw = *nwide;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nw; w += (gridDim.x * blockDim.x) >> 5) {
        const int s = widelist[w];
        const int b = cnt::find_segment(soff, B, s);
        const Geo g = geo[b];
        const int s0 = soff[b];
        RecA me;
        me.xc = xc[s]; me.yc = yc[s]; me.len = len[s]; me.idx = s - s0; me.gid = s;
        if (shard_world > 1 && slab_owner(cell_of(me.xc, g.G) / TILE_C, g.TG, shard_world) != shard_rank) continue;
        const CntLine l1 = line[s];
        int cx0, cx1, cy0, cy1;
        cell_range(me, g.G, cx0, cx1, cy0, cy1);
        const int base = FILL ? jstart[s] : 0;
        int written = 0;                                          // identical in every lane
        unsigned long long tests = 0;
        for (int cy = cy0; cy <= cy1; cy++) {
            const int q0 = __ldg(cell_start + g.coff + cy * g.G + cx0);
            const int q1 = __ldg(cell_start + g.coff + cy * g.G + cx1 + 1);
            for (int qq = q0; qq < q1; qq += 32) {
                const int q = qq + lane;
                bool hit = false;
                double xi = 0, yi = 0;
                int oj = 0;
                if (q < q1) {
                    const RecA o = recA[q];
                    if (o.idx > me.idx && cnt_in_ball(me.xc, me.yc, me.len, o.xc, o.yc)) {
                        tests++;
                        hit = cnt_junction(l1, line[o.gid], &xi, &yi);
                        oj = o.idx;
                    }
                }
                const unsigned m = __ballot_sync(~0u, hit);
                if (FILL && hit) {
                    const int o = base + written + __popc(m & ((1u << lane) - 1u));
                    stick2[o] = oj;
                    jx[o] = xi;
                    jy[o] = yi;
                }
                written += __popc(m);
            }
        }
        if (!FILL) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tests += __shfl_down_sync(~0u, tests, o);
            if (lane == 0) {
                jcount[s] = written;
                if (tests) atomicAdd(ntests + b, tests);
            }
        } else {
            __syncwarp();
            if (lane == 0) {
                for (int a = 1; a < written; a++) {               // insertion by stick2
                    const int kj = stick2[base + a];
                    const double kx = jx[base + a], ky = jy[base + a];
                    int c = base + a - 1;
                    while (c >= base && stick2[c] > kj) {
                        stick2[c + 1] = stick2[c]; jx[c + 1] = jx[c]; jy[c + 1] = jy[c];
                        c--;
                    }
                    stick2[c + 1] = kj; jx[c + 1] = kx; jy[c + 1] = ky;
                }
                for (int a = 0; a < written; a++) {
                    stick1[base + a] = me.idx;
                    kind2[base + a] = (uint8_t)(3 * kind[s] + kind[s0 + stick2[base + a]]);
                }
            }
        }
    }
}

// ---- dense rows: long stick i against every later stick of its network ------------------
// Rows 0 and 1 of a network (the electrodes: ~100k candidates each) are cut into DENSE_SEGS segments handled by
// different CTAs -- 16 CTAs per network instead of 2; the count pass leaves the junctions per segment in segcnt
// (and their sum in jcount), the fill pass starts every segment behind its predecessors, so the row comes out in
// stick2 order.  Further long rows (small networks, where a body stick can span more than RMAX cells) are
// walked by one CTA each (segment 0).
constexpr int DENSE_SEGS = 8;    // segments of an electrode row
constexpr int DENSE_SPLIT = 2;   // rows that are split
template <bool FILL>
__global__ void __launch_bounds__(DENSE_T) k_dense(int B, const int32_t *__restrict__ soff,
                                                   const int32_t *__restrict__ nlong,
                                                   const CntLine *__restrict__ line, const double *__restrict__ xc,
                                                   const double *__restrict__ yc, const double *__restrict__ len,
                                                   int32_t *__restrict__ jcount, unsigned long long *__restrict__ ntests,
                                                   int32_t *__restrict__ segcnt /* [B][DENSE_SPLIT][DENSE_SEGS] */,
                                                   const int32_t *__restrict__ jstart, const uint8_t *__restrict__ kind,
                                                   int32_t *__restrict__ stick1, int32_t *__restrict__ stick2,
                                                   double *__restrict__ jx, double *__restrict__ jy,
                                                   uint8_t *__restrict__ kind2) {
    __shared__ int warp_cnt[DENSE_T / 32];
    __shared__ double red[32];
    const int b = blockIdx.y;
    const int s0 = soff[b], n = soff[b + 1] - s0;
    const int nl = nlong[b];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rowslot = blockIdx.x / DENSE_SEGS, seg = blockIdx.x - rowslot * DENSE_SEGS;
    for (int i = rowslot; i < nl; i += DENSE_ROWS) {
        const bool split = i < DENSE_SPLIT;
        if (!split && seg != 0) continue;                        // CTA-uniform
        int jb = i + 1, je = n;
        int32_t *sc = split ? segcnt + ((int64_t)b * DENSE_SPLIT + i) * DENSE_SEGS : nullptr;
        int before_seg = 0;
        if (split) {
            const int seglen = ((n - i - 1 + DENSE_SEGS - 1) / DENSE_SEGS + DENSE_T - 1) / DENSE_T * DENSE_T;
            jb = i + 1 + seg * seglen;
            je = min(n, jb + seglen);
            if (FILL)
                for (int k = 0; k < seg; k++) before_seg += sc[k];
        }
        CntLine l1 = line[s0 + i];
        double xci = xc[s0 + i], yci = yc[s0 + i], li = len[s0 + i];
        int base = FILL ? jstart[s0 + i] + before_seg : 0;
        int written = 0;                       // identical in every thread
        unsigned long long tests = 0;
        for (int j0 = jb; j0 < je; j0 += DENSE_T) {
            int j = j0 + threadIdx.x;
            bool hit = false;
            double xi = 0, yi = 0;
            if (j < je && cnt_in_ball(xci, yci, li, xc[s0 + j], yc[s0 + j])) {
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
            if (threadIdx.x == 0) {
                if (split) {
                    sc[seg] = written;
                    if (written) atomicAdd(jcount + s0 + i, written);     // jcount is cleared before the pass
                } else {
                    jcount[s0 + i] = written;
                }
            }
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
    int64_t S, total_cells, total_tiles, cap_tmp;
    std::vector<Geo> geo;
    // workspace carve-up
    Geo *d_geo;
    CntLine *line;
    RecA *recA;
    int32_t *cellkey, *cell_count /* total_cells+1, scanned in place */, *cell_fill, *nlong, *errflag, *ntmp, *jtmp, *tj, *segcnt, *nwide, *widelist;
    double *tx, *ty;
    void *scan_ws;
    int64_t scan_ws_bytes;
};

static int64_t tmp_capacity(int64_t S) { return 8 * S + 65536; }

static int64_t plan_bytes(int B, int64_t S, int64_t total_cells) {
    using cnt::ws_bytes_for;
    const int64_t ct = tmp_capacity(S);
    return ws_bytes_for(B + 1, sizeof(Geo)) + ws_bytes_for(S, sizeof(CntLine)) + ws_bytes_for(S, sizeof(RecA)) +
           ws_bytes_for(S, 4) + ws_bytes_for(total_cells + 1, 4) +
           ws_bytes_for(total_cells + 1, 4) + ws_bytes_for(B, 4) + ws_bytes_for(2, 4) + ws_bytes_for(S + 1, 4) +
           ws_bytes_for((int64_t)B * DENSE_SPLIT * DENSE_SEGS, 4) + ws_bytes_for(S + 1, 4) +
           ws_bytes_for(ct, 4) + 2 * ws_bytes_for(ct, 8) + cnt_scan_workspace_bytes(total_cells + 1);
}

static void make_geo(int B, const int32_t *h_soff, std::vector<Geo> &geo, int64_t &total_cells, int64_t &total_tiles) {
    geo.resize(B + 1);
    total_cells = 0;
    total_tiles = 0;
    for (int b = 0; b < B; b++) {
        int g = grid_dim_for(h_soff[b + 1] - h_soff[b]);
        geo[b].G = g;
        geo[b].coff = (int32_t)total_cells;
        geo[b].TG = (g + TILE_C - 1) / TILE_C;
        geo[b].toff = (int32_t)total_tiles;
        total_cells += (int64_t)g * g;
        total_tiles += (int64_t)geo[b].TG * geo[b].TG;
    }
    geo[B].G = 0;
    geo[B].coff = (int32_t)total_cells;
    geo[B].TG = 0;
    geo[B].toff = (int32_t)total_tiles;
}

static int carve(Plan &p, int B, const int32_t *h_soff, void *ws, int64_t ws_bytes) {
    p.S = h_soff[B];
    make_geo(B, h_soff, p.geo, p.total_cells, p.total_tiles);
    if (p.total_cells >= (int64_t)1 << 31 || p.total_tiles >= (int64_t)1 << 31) {
        cnt::set_error("too many cells");
        return CNT_EINVAL;
    }
    p.cap_tmp = tmp_capacity(p.S);
    cnt::Workspace w(ws, ws_bytes);
    p.d_geo = w.take<Geo>(B + 1);
    p.line = w.take<CntLine>(p.S);
    p.recA = w.take<RecA>(p.S);
    p.cellkey = w.take<int32_t>(p.S);
    p.cell_count = w.take<int32_t>(p.total_cells + 1);
    p.cell_fill = w.take<int32_t>(p.total_cells + 1);
    p.nlong = w.take<int32_t>(B);
    p.errflag = w.take<int32_t>(2);
    p.ntmp = p.errflag ? p.errflag + 1 : nullptr;
    p.jtmp = w.take<int32_t>(p.S + 1);
    p.segcnt = w.take<int32_t>((int64_t)B * DENSE_SPLIT * DENSE_SEGS);
    p.widelist = w.take<int32_t>(p.S + 1);                        // [0]: number of wide sticks, then their ids
    p.nwide = p.widelist;
    if (p.widelist) p.widelist += 1;
    p.tj = w.take<int32_t>(p.cap_tmp);
    p.tx = w.take<double>(p.cap_tmp);
    p.ty = w.take<double>(p.cap_tmp);
    p.scan_ws_bytes = cnt_scan_workspace_bytes(p.total_cells + 1);
    p.scan_ws = w.take<char>(p.scan_ws_bytes);
    if (!p.scan_ws) {
        cnt::set_error("junction workspace too small: have %lld need %lld", (long long)ws_bytes,
                       (long long)plan_bytes(B, p.S, p.total_cells));
        return CNT_ECAPACITY;
    }
    return CNT_OK;
}

static size_t tiles_smem() {
    return 8 * ((size_t)2 * CAP_REC + CAP_HOME) + (size_t)CAP_HIT * sizeof(Hit) +
           4 * ((size_t)CAP_REC + CAP_Q + MAX_RH * (MAX_RH + 1) + MAX_RH + (MAX_RH + 1) + (MAXRUN + 1) + MAXRUN + (CAP_HOME + 1) +
                CAP_HOME);
}
}  // namespace

extern "C" int64_t cnt_junction_workspace_bytes(int B, const int32_t *h_soff) {
    if (B <= 0 || !h_soff) return 0;
    std::vector<Geo> geo;
    int64_t tc, tt;
    make_geo(B, h_soff, geo, tc, tt);
    return plan_bytes(B, h_soff[B], tc);
}

static int junction_count(int B, const int32_t *soff, const int32_t *h_soff, const double *xa,
                          const double *ya, const double *xb, const double *yb, const double *xc,
                          const double *yc, const double *length, int32_t *jcount,
                          unsigned long long *ntests, void *ws, int64_t ws_bytes, void *stream, int shard_rank,
                          int shard_world) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world);
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
    CNT_CUDA(cudaMemcpyAsync(p.d_geo, p.geo.data(), sizeof(Geo) * (B + 1), cudaMemcpyHostToDevice, st));
    CNT_CUDA(cudaMemsetAsync(p.cell_count, 0, 4 * (p.total_cells + 1), st));
    CNT_CUDA(cudaMemsetAsync(p.cell_fill, 0, 4 * (p.total_cells + 1), st));
    CNT_CUDA(cudaMemsetAsync(p.nlong, 0, 4 * B, st));
    CNT_CUDA(cudaMemsetAsync(p.nwide, 0, 4, st));
    CNT_CUDA(cudaMemsetAsync(p.errflag, 0, 8, st));
    CNT_CUDA(cudaMemsetAsync(ntests, 0, sizeof(unsigned long long) * B, st));
    CNT_CUDA(cudaMemsetAsync(jcount, 0, 4 * (p.S + 1), st));
    CNT_CUDA(cudaMemsetAsync(p.jtmp, 0xFF, 4 * (p.S + 1), st));
    k_prep<<<cnt::grid_for(p.S, 256), 256, 0, st>>>(B, soff, p.d_geo, p.S, xa, ya, xb, yb, xc, yc, length, p.line,
                                                    p.cellkey, p.cell_count, p.nlong, p.errflag, p.nwide, p.widelist);
    CNT_LAUNCHED();
    rc = cnt_exclusive_scan_i32(p.cell_count, p.cell_count, p.total_cells, 1, p.scan_ws, p.scan_ws_bytes, stream);
    if (rc) return rc;
    k_bin<<<cnt::grid_for(p.S, 256), 256, 0, st>>>(p.S, soff, B, p.cellkey, p.cell_count, p.cell_fill, p.line, xc, yc,
                                                   length, p.recA);
    CNT_LAUNCHED();
    static bool attr_set = false;
    if (!attr_set) {
        CNT_CUDA(cudaFuncSetAttribute(k_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiles_smem()));
        attr_set = true;
    }
    k_tiles<<<(unsigned)p.total_tiles, CT, tiles_smem(), st>>>(B, p.d_geo, p.cell_count, p.recA, p.line, jcount, p.jtmp,
                                                               ntests, p.tj, p.tx, p.ty, p.cap_tmp, p.ntmp, p.errflag,
                                                               shard_rank, shard_world);
    CNT_LAUNCHED();
    k_wide<false><<<148 * 2, 256, 0, st>>>(B, soff, p.d_geo, p.cell_count, p.recA, p.nwide, p.widelist, p.line, xc, yc,
                                           length, jcount, ntests, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                           nullptr, shard_rank, shard_world);
    CNT_LAUNCHED();
    if (shard_rank != 0) return CNT_OK;                           // the long sticks (electrodes) are rank 0's
    k_dense<false><<<dim3(DENSE_ROWS * DENSE_SEGS, B), DENSE_T, 0, st>>>(B, soff, p.nlong, p.line, xc, yc, length, jcount,
                                                                         ntests, p.segcnt, nullptr, nullptr, nullptr, nullptr,
                                                                         nullptr, nullptr, nullptr);
    CNT_LAUNCHED();
    return CNT_OK;
}

extern "C" int cnt_junction_count(int B, const int32_t *soff, const int32_t *h_soff, const double *xa,
                                  const double *ya, const double *xb, const double *yb, const double *xc,
                                  const double *yc, const double *length, int32_t *jcount,
                                  unsigned long long *ntests, void *ws, int64_t ws_bytes, void *stream) {
    return junction_count(B, soff, h_soff, xa, ya, xb, yb, xc, yc, length, jcount, ntests, ws, ws_bytes, stream, 0, 1);
}
extern "C" int cnt_junction_count_slab(int B, const int32_t *soff, const int32_t *h_soff, const double *xa,
                                       const double *ya, const double *xb, const double *yb, const double *xc,
                                       const double *yc, const double *length, int32_t *jcount,
                                       unsigned long long *ntests, void *ws, int64_t ws_bytes, int shard_rank,
                                       int shard_world, void *stream) {
    return junction_count(B, soff, h_soff, xa, ya, xb, yb, xc, yc, length, jcount, ntests, ws, ws_bytes, stream, shard_rank,
                          shard_world);
}

static int junction_fill(int B, const int32_t *soff, const int32_t *h_soff, const double *xa,
                         const double *ya, const double *xb, const double *yb, const double *xc,
                         const double *yc, const double *length, const uint8_t *kind,
                         const int32_t *jstart, int64_t J, int32_t *stick1, int32_t *stick2, double *jx,
                         double *jy, uint8_t *kind2, int32_t *joff, void *ws, int64_t ws_bytes,
                         void *stream, int shard_rank, int shard_world) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world);
    CNT_CHECK_ARG(B > 0 && soff && h_soff && kind && jstart && joff && ws && J >= 0);
    CNT_CHECK_ARG(J == 0 || (stick1 && stick2 && jx && jy && kind2));
    Plan p;   // same carve-up as _count: the binned records built there are reused
    int rc = carve(p, B, h_soff, ws, ws_bytes);
    if (rc) return rc;
    k_joff<<<cnt::grid_for(B + 1, 128), 128, 0, st>>>(B, soff, jstart, joff);
    CNT_LAUNCHED();
    if (p.S == 0 || J == 0) return CNT_OK;
    k_place<<<cnt::grid_for(p.S, 256), 256, 0, st>>>(B, soff, p.S, p.jtmp, jstart, p.tj, p.tx, p.ty, kind, p.errflag, stick1,
                                                     stick2, jx, jy, kind2);
    CNT_LAUNCHED();
    k_cells_fill<<<cnt::grid_for(p.S, 128), 128, 0, st>>>(B, soff, p.d_geo, p.cell_count, (int32_t)p.total_cells, p.recA,
                                                          p.line, p.errflag, p.jtmp, jstart, kind, stick1, stick2, jx, jy, kind2,
                                                          shard_rank, shard_world);
    CNT_LAUNCHED();
    k_wide<true><<<148 * 2, 256, 0, st>>>(B, soff, p.d_geo, p.cell_count, p.recA, p.nwide, p.widelist, p.line, xc, yc,
                                          length, nullptr, nullptr, jstart, kind, stick1, stick2, jx, jy, kind2, shard_rank,
                                          shard_world);
    CNT_LAUNCHED();
    if (shard_rank != 0) return CNT_OK;
    k_dense<true><<<dim3(DENSE_ROWS * DENSE_SEGS, B), DENSE_T, 0, st>>>(B, soff, p.nlong, p.line, xc, yc, length, nullptr,
                                                                        nullptr, p.segcnt, jstart, kind, stick1, stick2, jx, jy,
                                                                        kind2);
    CNT_LAUNCHED();
    return CNT_OK;
}

extern "C" int cnt_junction_fill(int B, const int32_t *soff, const int32_t *h_soff, const double *xa,
                                 const double *ya, const double *xb, const double *yb, const double *xc,
                                 const double *yc, const double *length, const uint8_t *kind,
                                 const int32_t *jstart, int64_t J, int32_t *stick1, int32_t *stick2, double *jx,
                                 double *jy, uint8_t *kind2, int32_t *joff, void *ws, int64_t ws_bytes,
                                 void *stream) {
    return junction_fill(B, soff, h_soff, xa, ya, xb, yb, xc, yc, length, kind, jstart, J, stick1, stick2, jx, jy, kind2, joff,
                         ws, ws_bytes, stream, 0, 1);
}
extern "C" int cnt_junction_fill_slab(int B, const int32_t *soff, const int32_t *h_soff, const double *xa,
                                      const double *ya, const double *xb, const double *yb, const double *xc,
                                      const double *yc, const double *length, const uint8_t *kind,
                                      const int32_t *jstart, int64_t J, int32_t *stick1, int32_t *stick2, double *jx,
                                      double *jy, uint8_t *kind2, int32_t *joff, void *ws, int64_t ws_bytes,
                                      int shard_rank, int shard_world, void *stream) {
    return junction_fill(B, soff, h_soff, xa, ya, xb, yb, xc, yc, length, kind, jstart, J, stick1, stick2, jx, jy, kind2, joff,
                         ws, ws_bytes, stream, shard_rank, shard_world);
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
