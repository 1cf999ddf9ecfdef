// Batched Jacobi-preconditioned conjugate gradient, multi-launch ("flat") variant.
// Replaces solve_mna (cnet.py:147-149): the reference's SuperLU solve of the indefinite
// MNA matrix becomes PCG on the equivalent reduced SPD system.
//
// NSYS systems (network x gate configuration) are advanced in lock step by five launches
// per iteration; every system owns row tiles of TILE rows, per-tile partial dot products
// are written to memory and reduced in a fixed order by a per-system scalar kernel, so a
// system's iterates are bit-reproducible and independent of what else is in the batch.
// Converged systems drop out (their tiles return immediately).  The iteration body is
// captured once into a CUDA graph of `check_every` iterations; the host polls the per-system
// done flags between graph launches.
//
// Stopping rule: recurrence residual ||r||_2 <= rtol ||b||_2.  No residual replacement on
// purpose: for these Laplacians the TRUE residual stagnates near 1e-12 ||b|| while the error
// keeps falling with the recurrence residual ("PCG accuracy").
#include <vector>
#include "spmv_stream.cuh"

namespace {
constexpr int TILE = 256;
static_assert(TILE == cnt::spmv::TB, "k_spmv: one row per thread of the streaming SpMV");

struct Tile {
    int32_t sys;    // system index
    int32_t row0;   // first row (global pattern row)
    int32_t nrows;
    int32_t slab;   // value / vector slab of the system
};
struct Sys {
    int32_t net, row0, nrows, tile0, ntiles, slab;
};
struct Scal {      // per-system scalars
    double rz, bb, alpha, beta, rr;
    int32_t done, status, iters, pad;
};

struct Ctx {
    const Tile *tiles;
    const Sys *sys;
    Scal *scal;
    int64_t R, NNZ;
    const int32_t *rowptr, *col;
    const double *val, *diag, *rhs;
    double *x, *r, *p, *q;
    double *part0, *part1;   // per-tile partials
    // two-level preconditioner (agg.nseg == 0: plain Jacobi)
    cnt_aggregates agg;
    double *segS;            // per-segment sums of r
};

__device__ __forceinline__ double spmv_row(const Ctx &c, int sysi, int net_row0, int row, const double *v) {
    // sysi = slab; row = global pattern row; slab vectors live at slab*R + row; columns are local
    const double *vs = v + (int64_t)sysi * c.R + net_row0;
    const double *vals = c.val + (int64_t)sysi * c.NNZ;
    int k0 = __ldg(c.rowptr + row), k1 = __ldg(c.rowptr + row + 1);
    double acc = __ldg(c.diag + (int64_t)sysi * c.R + row) * v[(int64_t)sysi * c.R + row];
    for (int k = k0; k < k1; k++) acc += __ldg(vals + k) * vs[__ldg(c.col + k)];
    return acc;
}

__global__ void __launch_bounds__(TILE) k_init(Ctx c) {
    __shared__ double sm[32];
    Tile t = c.tiles[blockIdx.x];
    Sys s = c.sys[t.sys];
    int i = threadIdx.x;
    double rz = 0, bb = 0, rr = 0;
    if (i < t.nrows) {
        int row = t.row0 + i;
        int64_t v = (int64_t)t.slab * c.R + row;
        double b = c.rhs[v];
        double ri = b - spmv_row(c, t.slab, s.row0, row, c.x);
        double zi = ri / c.diag[v];
        c.r[v] = ri;
        c.p[v] = zi;
        rz = ri * zi;
        bb = b * b;
        rr = ri * ri;
    }
    rz = cnt::block_sum(rz, sm);
    bb = cnt::block_sum(bb, sm);
    rr = cnt::block_sum(rr, sm);
    if (i == 0) {
        c.part0[blockIdx.x] = rz;
        c.part1[blockIdx.x] = bb;
        c.part1[gridDim.x + blockIdx.x] = rr;   // third partial array lives in part1's second half
    }
}

// ---- per-system scalar kernels: ONE WARP per system.  Lane l sums tiles l, l+32, ... in order,
// then a fixed shuffle tree combines the lanes: the result depends only on the system's own
// tile count, never on the batch, so iterates stay bit-reproducible.
__device__ __forceinline__ double warp_tiles_sum(const double *__restrict__ part, int t0, int nt) {
    int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int t = lane; t < nt; t += 32) s += part[t0 + t];
    return __shfl_sync(0xffffffffu, cnt::warp_sum(s), 0);
}

__global__ void k_scal_init(Ctx c, int ntiles_total, double rtol, int nsys) {
    int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nsys) return;
    Sys sy = c.sys[s];
    double rz = warp_tiles_sum(c.part0, sy.tile0, sy.ntiles);
    double bb = warp_tiles_sum(c.part1, sy.tile0, sy.ntiles);
    double rr = warp_tiles_sum(c.part1 + ntiles_total, sy.tile0, sy.ntiles);
    if ((threadIdx.x & 31) != 0) return;
    Scal sc;
    sc.rz = rz; sc.bb = bb; sc.rr = rr; sc.alpha = 0; sc.beta = 0;
    sc.iters = 0; sc.status = 0; sc.pad = 0;
    sc.done = (sy.nrows == 0 || rr <= rtol * rtol * bb) ? 1 : 0;
    c.scal[s] = sc;
}

// q = A p of one tile of 256 rows of one system: the streaming SpMV of spmv_stream.cuh (the tile's entries are one
// contiguous range, read with coalesced 16-byte loads, products through shared memory, rows summed in CSR order)
__global__ void __launch_bounds__(TILE) k_spmv(Ctx c) {
    __shared__ double sm[32];
    __shared__ __align__(16) double s_prod[cnt::spmv::CAP];
    Tile t = c.tiles[blockIdx.x];
    if (c.scal[t.sys].done) return;                    // CTA-uniform
    Sys s = c.sys[t.sys];
    int i = threadIdx.x;
    const double *pslab = c.p + (int64_t)t.slab * c.R;
    const double qi = cnt::spmv::stream_rows(t.row0, (int64_t)t.row0 + t.nrows, c.NNZ, c.rowptr, c.col,
                                             c.val + (int64_t)t.slab * c.NNZ, c.diag + (int64_t)t.slab * c.R,
                                             pslab + s.row0, s_prod, pslab);
    double pq = 0;
    if (i < t.nrows) {
        int64_t v = (int64_t)t.slab * c.R + t.row0 + i;
        c.q[v] = qi;
        pq = c.p[v] * qi;
    }
    pq = cnt::block_sum(pq, sm);
    if (i == 0) c.part0[blockIdx.x] = pq;
}

__global__ void k_scal_alpha(Ctx c, int nsys) {
    int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nsys) return;
    Scal *sc = c.scal + s;
    if (sc->done) return;
    Sys sy = c.sys[s];
    double pq = warp_tiles_sum(c.part0, sy.tile0, sy.ntiles);
    if ((threadIdx.x & 31) != 0) return;
    if (!(pq > 0.0)) {   // breakdown (matrix not SPD or NaN)
        sc->done = 1;
        sc->status = 2;
        return;
    }
    sc->alpha = sc->rz / pq;
}

__global__ void __launch_bounds__(TILE) k_update(Ctx c) {
    __shared__ double sm[32];
    Tile t = c.tiles[blockIdx.x];
    const Scal *sc = c.scal + t.sys;
    if (sc->done) return;
    double alpha = sc->alpha;
    int i = threadIdx.x;
    double rz = 0, rr = 0;
    if (i < t.nrows) {
        int64_t v = (int64_t)t.slab * c.R + t.row0 + i;
        c.x[v] += alpha * c.p[v];
        double ri = c.r[v] - alpha * c.q[v];
        c.r[v] = ri;
        rz = ri * (ri / c.diag[v]);
        rr = ri * ri;
    }
    rz = cnt::block_sum(rz, sm);
    rr = cnt::block_sum(rr, sm);
    if (i == 0) {
        c.part0[blockIdx.x] = rz;
        c.part1[blockIdx.x] = rr;
    }
}

__global__ void k_scal_beta(Ctx c, int nsys, double rtol, int maxit) {
    int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nsys) return;
    Scal *sc = c.scal + s;
    if (sc->done) return;
    Sys sy = c.sys[s];
    double rz = warp_tiles_sum(c.part0, sy.tile0, sy.ntiles);
    double rr = warp_tiles_sum(c.part1, sy.tile0, sy.ntiles);
    if ((threadIdx.x & 31) != 0) return;
    sc->beta = rz / sc->rz;
    sc->rz = rz;
    sc->rr = rr;
    sc->iters += 1;
    if (rr <= rtol * rtol * sc->bb) {
        sc->done = 1;
        sc->status = 0;
    } else if (sc->iters >= maxit) {
        sc->done = 1;
        sc->status = 1;
    } else if (!(rr == rr)) {
        sc->done = 1;
        sc->status = 2;
    }
}

__global__ void __launch_bounds__(TILE) k_pupdate(Ctx c) {
    Tile t = c.tiles[blockIdx.x];
    const Scal *sc = c.scal + t.sys;
    if (sc->done) return;
    double beta = sc->beta;
    int i = threadIdx.x;
    if (i < t.nrows) {
        int64_t v = (int64_t)t.slab * c.R + t.row0 + i;
        c.p[v] = c.r[v] / c.diag[v] + beta * c.p[v];
    }
}

// ---------------- two-level preconditioner: z = r/d + (sum of r over the row's aggregate) * einv
// segment sums of r: one THREAD per segment of <= 8 rows (the vast majority), one warp per larger
// segment -- members are added in list order / lane-strided + fixed tree, so the result is
// deterministic either way
constexpr int SEG_SMALL = 8;
__global__ void k_seg_sum_small(Ctx c) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= c.agg.nseg) return;
    int n = c.agg.seg_count[s];
    if (n > SEG_SMALL || c.scal[c.agg.seg_sys[s]].done) return;
    int s0 = c.agg.seg_start[s];
    double acc = 0.0;
    for (int k = 0; k < n; k++) acc += c.r[c.agg.mem[s0 + k]];
    c.segS[s] = acc;
}
__global__ void k_seg_sum(Ctx c) {
    int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= c.agg.nseg) return;
    int n = c.agg.seg_count[s];
    if (n <= SEG_SMALL || c.scal[c.agg.seg_sys[s]].done) return;
    int lane = threadIdx.x & 31;
    int s0 = c.agg.seg_start[s];
    double acc = 0.0;
    for (int k = lane; k < n; k += 32) acc += c.r[c.agg.mem[s0 + k]];
    acc = cnt::warp_sum(acc);
    if (lane == 0) c.segS[s] = acc;
}

__device__ __forceinline__ double coarse_term(const Ctx &c, int64_t v) {
    if (c.agg.row_q) {      // precomputed (first segment, count, 1/E) per row
        int2 rq = __ldg(reinterpret_cast<const int2 *>(c.agg.row_q) + v);
        if (rq.y == 0) return 0.0;
        double S = c.segS[rq.x];
        for (int q = 1; q < rq.y; q++) S += c.segS[rq.x + q];
        return S * __ldg(c.agg.row_einv + v);
    }
    int a = c.agg.agg_of_row[v];
    if (a < 0) return 0.0;
    double S = 0.0;
    for (int s = c.agg.agg_seg0[a]; s < c.agg.agg_seg0[a + 1]; s++) S += c.segS[s];
    return S * c.agg.agg_einv[a];
}

__global__ void __launch_bounds__(TILE) k_init_r(Ctx c) {
    __shared__ double sm[32];
    Tile t = c.tiles[blockIdx.x];
    Sys s = c.sys[t.sys];
    int i = threadIdx.x;
    double bb = 0, rr = 0;
    if (i < t.nrows) {
        int row = t.row0 + i;
        int64_t v = (int64_t)t.slab * c.R + row;
        double b = c.rhs[v];
        double ri = b - spmv_row(c, t.slab, s.row0, row, c.x);
        c.r[v] = ri;
        bb = b * b;
        rr = ri * ri;
    }
    bb = cnt::block_sum(bb, sm);
    rr = cnt::block_sum(rr, sm);
    if (i == 0) {
        c.part1[blockIdx.x] = bb;
        c.part1[gridDim.x + blockIdx.x] = rr;
    }
}

__global__ void __launch_bounds__(TILE) k_init_z(Ctx c) {
    __shared__ double sm[32];
    Tile t = c.tiles[blockIdx.x];
    int i = threadIdx.x;
    double rz = 0;
    if (i < t.nrows) {
        int64_t v = (int64_t)t.slab * c.R + t.row0 + i;
        double ri = c.r[v];
        double zi = ri / c.diag[v] + coarse_term(c, v);
        c.p[v] = zi;
        rz = ri * zi;
    }
    rz = cnt::block_sum(rz, sm);
    if (i == 0) c.part0[blockIdx.x] = rz;
}

__global__ void __launch_bounds__(TILE) k_update_r(Ctx c) {
    __shared__ double sm[32];
    Tile t = c.tiles[blockIdx.x];
    const Scal *sc = c.scal + t.sys;
    if (sc->done) return;
    double alpha = sc->alpha;
    int i = threadIdx.x;
    double rr = 0;
    if (i < t.nrows) {
        int64_t v = (int64_t)t.slab * c.R + t.row0 + i;
        c.x[v] += alpha * c.p[v];
        double ri = c.r[v] - alpha * c.q[v];
        c.r[v] = ri;
        rr = ri * ri;
    }
    rr = cnt::block_sum(rr, sm);
    if (i == 0) c.part1[blockIdx.x] = rr;
}

__global__ void __launch_bounds__(TILE) k_zdot(Ctx c) {
    __shared__ double sm[32];
    Tile t = c.tiles[blockIdx.x];
    if (c.scal[t.sys].done) return;
    int i = threadIdx.x;
    double rz = 0;
    if (i < t.nrows) {
        int64_t v = (int64_t)t.slab * c.R + t.row0 + i;
        double ri = c.r[v];
        double zi = ri / c.diag[v] + coarse_term(c, v);
        c.q[v] = zi;               // q is free after k_update_r: reuse it for z
        rz = ri * zi;
    }
    rz = cnt::block_sum(rz, sm);
    if (i == 0) c.part0[blockIdx.x] = rz;
}

__global__ void __launch_bounds__(TILE) k_pupdate_z(Ctx c) {
    Tile t = c.tiles[blockIdx.x];
    const Scal *sc = c.scal + t.sys;
    if (sc->done) return;
    double beta = sc->beta;
    int i = threadIdx.x;
    if (i < t.nrows) {
        int64_t v = (int64_t)t.slab * c.R + t.row0 + i;
        c.p[v] = c.q[v] + beta * c.p[v];
    }
}

__global__ void k_export(Ctx c, int nsys, int32_t *iters, double *relres, int32_t *status, int32_t *alldone) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsys) return;
    Scal sc = c.scal[s];
    iters[s] = sc.iters;
    relres[s] = sc.bb > 0 ? sqrt(sc.rr / sc.bb) : 0.0;
    status[s] = sc.done ? sc.status : 1;
    if (!sc.done) atomicAnd(alldone, 0);
}
__global__ void k_set1(int32_t *p) { *p = 1; }

struct Layout {
    std::vector<Tile> tiles;
    std::vector<Sys> sys;
};
static void make_layout(int NSYS, const int32_t *h_sys_net, const int32_t *h_sys_slab, const int32_t *h_roff,
                        Layout &L) {
    L.sys.resize(NSYS);
    L.tiles.clear();
    for (int s = 0; s < NSYS; s++) {
        int b = h_sys_net[s];
        Sys sy;
        sy.net = b;
        sy.row0 = h_roff[b];
        sy.nrows = h_roff[b + 1] - h_roff[b];
        sy.tile0 = (int)L.tiles.size();
        sy.ntiles = (sy.nrows + TILE - 1) / TILE;
        sy.slab = h_sys_slab[s];
        for (int t = 0; t < sy.ntiles; t++) {
            Tile tl;
            tl.sys = s;
            tl.row0 = sy.row0 + t * TILE;
            tl.nrows = sy.nrows - t * TILE < TILE ? sy.nrows - t * TILE : TILE;
            tl.slab = sy.slab;
            L.tiles.push_back(tl);
        }
        L.sys[s] = sy;
    }
}
// upper bound on the number of tiles: every system has at most R rows
static int64_t count_tiles(int NSYS, int64_t R) { return (int64_t)NSYS * ((R + TILE - 1) / TILE + 1); }
}  // namespace

extern "C" int64_t cnt_pcg_workspace_bytes(int NSYS, int NSLAB, int64_t R) {
    using cnt::ws_bytes_for;
    int64_t nt = count_tiles(NSYS, R);
    return ws_bytes_for(nt, sizeof(Tile)) + ws_bytes_for(NSYS, sizeof(Sys)) + ws_bytes_for(NSYS, sizeof(Scal)) +
           3 * ws_bytes_for((int64_t)NSLAB * R + 1, 8) + ws_bytes_for(nt, 8) + ws_bytes_for(2 * nt, 8) +
           ws_bytes_for(1, 4) + ws_bytes_for((int64_t)NSLAB * R / 2 + (int64_t)NSLAB * R / 256 + 2, 8);
}

extern "C" int cnt_pcg_solve(int B, int NSYS, int NSLAB, const int32_t *h_sys_net, const int32_t *h_sys_slab,
                             const int32_t *h_roff, int64_t R, int64_t NNZ, const int32_t *rowptr, const int32_t *col, const double *val, const double *diag,
                             const double *rhs, double *x, double rtol, int maxit, int check_every, int32_t *iters,
                             double *relres, int32_t *status, const cnt_aggregates *agg, void *ws, int64_t ws_bytes,
                             void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && NSYS > 0 && NSLAB > 0 && h_sys_net && h_sys_slab && h_roff && R >= 0 && NNZ >= 0 && iters && relres && status && ws);
    CNT_CHECK_ARG(rtol > 0 && maxit >= 0 && check_every > 0);
    for (int s = 0; s < NSYS; s++)
        CNT_CHECK_ARG(h_sys_net[s] >= 0 && h_sys_net[s] < B && h_sys_slab[s] >= 0 && h_sys_slab[s] < NSLAB);
    CNT_CHECK_ARG(h_roff[B] == R);
    if (R > 0) CNT_CHECK_ARG(rowptr && diag && rhs && x && (NNZ == 0 || (col && val)));
    Layout L;
    make_layout(NSYS, h_sys_net, h_sys_slab, h_roff, L);
    int64_t nt = (int64_t)L.tiles.size();
    int64_t nt_alloc = count_tiles(NSYS, R);
    cnt::Workspace w(ws, ws_bytes);
    Tile *d_tiles = w.take<Tile>(nt_alloc);
    Sys *d_sys = w.take<Sys>(NSYS);
    Scal *d_scal = w.take<Scal>(NSYS);
    double *r = w.take<double>((int64_t)NSLAB * R + 1);
    double *p = w.take<double>((int64_t)NSLAB * R + 1);
    double *q = w.take<double>((int64_t)NSLAB * R + 1);
    double *part0 = w.take<double>(nt_alloc);
    double *part1 = w.take<double>(2 * nt_alloc);
    int32_t *alldone = w.take<int32_t>(1);
    const bool two_level = agg && agg->nseg > 0;
    double *segS = w.take<double>(two_level ? agg->nseg : 1);
    if (!alldone || !segS) {
        cnt::set_error("pcg workspace too small: have %lld need %lld", (long long)ws_bytes,
                       (long long)cnt_pcg_workspace_bytes(NSYS, NSLAB, R));
        return CNT_ECAPACITY;
    }
    CNT_CUDA(cudaMemcpyAsync(d_sys, L.sys.data(), sizeof(Sys) * NSYS, cudaMemcpyHostToDevice, st));
    if (nt > 0) CNT_CUDA(cudaMemcpyAsync(d_tiles, L.tiles.data(), sizeof(Tile) * nt, cudaMemcpyHostToDevice, st));
    Ctx c;
    c.tiles = d_tiles; c.sys = d_sys; c.scal = d_scal; c.R = R; c.NNZ = NNZ;
    c.rowptr = rowptr; c.col = col; c.val = val; c.diag = diag; c.rhs = rhs;
    c.x = x; c.r = r; c.p = p; c.q = q; c.part0 = part0; c.part1 = part1;
    c.segS = segS;
    if (two_level) c.agg = *agg; else { c.agg = cnt_aggregates{}; }
    unsigned gseg = two_level ? cnt::grid_for(agg->nseg * 32, 256) : 0;
    unsigned gseg_small = two_level ? cnt::grid_for(agg->nseg, 256) : 0;
    unsigned gs = cnt::grid_for(NSYS, 128);            // one thread per system (export)
    unsigned gw = cnt::grid_for((int64_t)NSYS * 32, 128);   // one warp per system (scalar kernels)
    if (nt > 0 && !two_level) {
        k_init<<<(unsigned)nt, TILE, 0, st>>>(c);
        CNT_LAUNCHED();
    } else if (nt > 0) {
        k_init_r<<<(unsigned)nt, TILE, 0, st>>>(c);
        CNT_LAUNCHED();
        // every system is live here: Scal is not initialised yet, so zero the done flags first
        CNT_CUDA(cudaMemsetAsync(d_scal, 0, sizeof(Scal) * NSYS, st));
        k_seg_sum_small<<<gseg_small, 256, 0, st>>>(c);
        CNT_LAUNCHED();
        k_seg_sum<<<gseg, 256, 0, st>>>(c);
        CNT_LAUNCHED();
        k_init_z<<<(unsigned)nt, TILE, 0, st>>>(c);
        CNT_LAUNCHED();
    }
    k_scal_init<<<gw, 128, 0, st>>>(c, (int)nt, rtol, NSYS);
    CNT_LAUNCHED();

    auto enqueue_iter = [&]() -> int {
        k_spmv<<<(unsigned)nt, TILE, 0, st>>>(c);
        CNT_LAUNCHED();
        k_scal_alpha<<<gw, 128, 0, st>>>(c, NSYS);
        CNT_LAUNCHED();
        if (!two_level) {
            k_update<<<(unsigned)nt, TILE, 0, st>>>(c);
            CNT_LAUNCHED();
            k_scal_beta<<<gw, 128, 0, st>>>(c, NSYS, rtol, maxit);
            CNT_LAUNCHED();
            k_pupdate<<<(unsigned)nt, TILE, 0, st>>>(c);
            CNT_LAUNCHED();
        } else {
            k_update_r<<<(unsigned)nt, TILE, 0, st>>>(c);
            CNT_LAUNCHED();
            k_seg_sum_small<<<gseg_small, 256, 0, st>>>(c);
            CNT_LAUNCHED();
            k_seg_sum<<<gseg, 256, 0, st>>>(c);
            CNT_LAUNCHED();
            k_zdot<<<(unsigned)nt, TILE, 0, st>>>(c);
            CNT_LAUNCHED();
            k_scal_beta<<<gw, 128, 0, st>>>(c, NSYS, rtol, maxit);
            CNT_LAUNCHED();
            k_pupdate_z<<<(unsigned)nt, TILE, 0, st>>>(c);
            CNT_LAUNCHED();
        }
        return CNT_OK;
    };

    int32_t h_alldone = (nt == 0) ? 1 : 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    bool use_graph = (st != nullptr) && nt > 0 && maxit > 0 && !getenv("CNT_NO_GRAPH");
    int rc = CNT_OK;
    if (use_graph) {
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) {
            use_graph = false;
            cudaGetLastError();
        } else {
            for (int k = 0; k < check_every && rc == CNT_OK; k++) rc = enqueue_iter();
            e = cudaStreamEndCapture(st, &graph);
            if (rc != CNT_OK) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            if (e != cudaSuccess || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
                cnt::set_error("CUDA graph capture/instantiate failed: %s", cudaGetErrorString(cudaGetLastError()));
                if (graph) cudaGraphDestroy(graph);
                return CNT_ECUDA;
            }
        }
    }
    int done_iters = 0;
    while (!h_alldone && done_iters < maxit) {
        if (use_graph) {
            cudaError_t e = cudaGraphLaunch(gexec, st);
            cnt::count_launch((two_level ? 8 : 5) * check_every);
            if (e != cudaSuccess) {
                cnt::set_error("cudaGraphLaunch: %s", cudaGetErrorString(e));
                rc = CNT_ECUDA;
                break;
            }
        } else {
            for (int k = 0; k < check_every && rc == CNT_OK; k++) rc = enqueue_iter();
            if (rc) break;
        }
        done_iters += check_every;
        k_set1<<<1, 1, 0, st>>>(alldone);
        k_export<<<gs, 128, 0, st>>>(c, NSYS, iters, relres, status, alldone);
        cnt::count_launch(2);
        cudaError_t e = cudaMemcpyAsync(&h_alldone, alldone, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cnt::set_error("pcg poll: %s", cudaGetErrorString(e));
            rc = CNT_ECUDA;
            break;
        }
    }
    if (rc == CNT_OK) {
        k_set1<<<1, 1, 0, st>>>(alldone);
        k_export<<<gs, 128, 0, st>>>(c, NSYS, iters, relres, status, alldone);
        cnt::count_launch(2);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            cnt::set_error("pcg export: %s", cudaGetErrorString(e));
            rc = CNT_ECUDA;
        }
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    return rc;
}
