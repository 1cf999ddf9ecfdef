// Batched Jacobi-PCG, persistent thread-block-cluster variant ("one network per block-group").
// Replaces solve_mna (cnet.py:147-149).  Same arithmetic contract as pcg.cu.
//
// One launch solves a whole batch.  The grid is as many clusters of C CTAs as fit on the chip;
// each cluster pulls systems from a global work queue until it is empty (continuous batching:
// no lock-step, no host polling, no tail of finished systems).  While a cluster owns a system
//   * the rows are split into C contiguous chunks -- with the Z-ordered numbering of
//     cnt_assemble_reorder each chunk is a compact patch of the device, so ~90 % of the SpMV
//     gathers p[col] are served from the CTA's OWN shared memory; only halo columns go to L2;
//   * p, r and diag of the chunk live in shared memory, x and q/z in registers for the whole
//     solve: per iteration the only memory traffic is the matrix stream (val, col, rowptr) from
//     L2, the halo gathers, and one 8-byte store per row to publish p for the other CTAs;
//   * the two dot products are block-reduced, exchanged through distributed shared memory and
//     combined in a fixed order (bit-reproducible, identical in every CTA); the three phase
//     boundaries per iteration are hardware cluster barriers (barrier.cluster), no global spin.
// Systems whose chunk does not fit (rows > C_max * CAP) are reported with CNT_ECAPACITY so the
// caller can route them to the multi-launch solver in pcg.cu.
#include <cooperative_groups.h>
#include <vector>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {
constexpr int CT = 1024;             // threads per CTA
constexpr int MAXR = 6;              // rows per thread
constexpr int CAP = CT * MAXR;       // rows per CTA (6144)
constexpr int MAXC = 16;             // largest cluster (non-portable size, opt-in)

struct CSys {
    int32_t row0;      // first row of the network in the batch pattern
    int32_t nrows;
    int32_t slab;
    int32_t out;       // index into iters/relres/status
};

struct CArgs {
    const CSys *sys;
    int nsys;
    int *counter;
    int64_t R, NNZ;
    const int32_t *rowptr, *col;
    const double *val, *diag, *rhs;
    double *x, *pglob;
    double rtol;
    int maxit;
    int32_t *iters, *status;
    double *relres;
};

struct Shared {
    double red[2][4];      // [phase parity][value]: this CTA's partial sums, read by cluster peers
    double bcast[4];       // cluster-wide sums, broadcast inside the CTA
    double warp[32];       // block reduction scratch
    int next;              // next system (valid in CTA rank 0)
};

// deterministic block sums of up to 3 values at once; results valid in thread 0
template <int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double *scratch /* 32*N */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++) v[k] = cnt::warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; k++) scratch[k * 32 + w] = v[k];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < N; k++) {
            double t = lane < (CT >> 5) ? scratch[k * 32 + lane] : 0.0;
            v[k] = cnt::warp_sum(t);
        }
    }
}

// cluster-wide sum of N values: every CTA publishes its partials in sh->red[slot], the cluster
// barrier makes them visible, warp 0 of every CTA reads all C partials through DSMEM and adds
// them in rank order.  Every CTA obtains bit-identical results.
template <int N>
__device__ __forceinline__ void cluster_sum_n(cg::cluster_group &cluster, Shared *sh, int slot, double (&v)[N],
                                              double *scratch, int C) {
    block_sum_n<N>(v, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; k++) sh->red[slot][k] = v[k];
    }
    cluster.sync();
    if (threadIdx.x < 32) {
        int lane = threadIdx.x;
#pragma unroll
        for (int k = 0; k < N; k++) {
            double t = 0.0;
            if (lane < C) {
                const Shared *peer = cluster.map_shared_rank(sh, lane);
                t = peer->red[slot][k];
            }
            // fixed-order tree over ranks 0..C-1 (C <= 16)
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if (lane == 0) sh->bcast[k] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; k++) v[k] = sh->bcast[k];
}

__global__ void __launch_bounds__(CT, 1) k_pcg_cluster(CArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    Shared *sh = reinterpret_cast<Shared *>(smem_raw);
    double *scratch = reinterpret_cast<double *>(smem_raw + sizeof(Shared));   // 3*32 doubles
    double *p_s = scratch + 96;
    double *r_s = p_s + CAP;
    double *d_s = r_s + CAP;
    const int t = threadIdx.x;

    while (true) {
        if (crank == 0 && t == 0) sh->next = atomicAdd(a.counter, 1);
        cluster.sync();
        const int sidx = cluster.map_shared_rank(sh, 0)->next;
        cluster.sync();               // everybody has read `next` before rank 0 overwrites it
        if (sidx >= a.nsys) break;
        const CSys sy = a.sys[sidx];
        const int chunk = (sy.nrows + C - 1) / C;
        const int own0 = crank * chunk;                                  // local (network) row index
        const int nown = max(0, min(chunk, sy.nrows - own0));
        const int64_t vbase = (int64_t)sy.slab * a.R + sy.row0;          // slab vectors, network-local index
        const double *val = a.val + (int64_t)sy.slab * a.NNZ;
        const double *diag = a.diag + vbase;
        const double *rhs = a.rhs + vbase;
        double *x = a.x + vbase;
        double *pg = a.pglob + vbase;
        const int32_t *rowptr = a.rowptr + sy.row0;

        double x_r[MAXR], qz_r[MAXR];
        // ---- r = b - A x0, z = r / d, p = z
        double acc[3] = {0.0, 0.0, 0.0};    // rz, bb, rr
#pragma unroll
        for (int k = 0; k < MAXR; k++) {
            int li = t + k * CT;
            x_r[k] = 0.0;
            qz_r[k] = 0.0;
            if (li < nown) {
                int gi = own0 + li;
                double di = diag[gi], xi = x[gi], bi = rhs[gi];
                double ax = di * xi;
                int k0 = __ldg(rowptr + gi), k1 = __ldg(rowptr + gi + 1);
                for (int e = k0; e < k1; e++) ax += __ldg(val + e) * x[__ldg(a.col + e)];
                double ri = bi - ax, zi = ri / di;
                x_r[k] = xi;
                d_s[li] = di;
                r_s[li] = ri;
                p_s[li] = zi;
                pg[gi] = zi;
                acc[0] += ri * zi;
                acc[1] += bi * bi;
                acc[2] += ri * ri;
            }
        }
        cluster_sum_n<3>(cluster, sh, 0, acc, scratch, C);     // barrier also publishes p
        double rz = acc[0];
        const double bb = acc[1];
        double rr = acc[2];
        const double tol2 = a.rtol * a.rtol * bb;
        int it = 0, status = 0;
        bool done = (sy.nrows == 0) || (rr <= tol2);
        if (!done && a.maxit <= 0) { done = true; status = 1; }

        while (!done) {
            // ---- phase 1: q = A p (own rows), pq = p.q
            double s1[1] = {0.0};
#pragma unroll
            for (int k = 0; k < MAXR; k++) {
                int li = t + k * CT;
                if (li < nown) {
                    int gi = own0 + li;
                    double pi = p_s[li];
                    double q = d_s[li] * pi;
                    int k0 = __ldg(rowptr + gi), k1 = __ldg(rowptr + gi + 1);
                    for (int e = k0; e < k1; e++) {
                        int c = __ldg(a.col + e);
                        unsigned lc = (unsigned)(c - own0);
                        double pv = lc < (unsigned)nown ? p_s[lc] : __ldcg(pg + c);
                        q += __ldg(val + e) * pv;
                    }
                    qz_r[k] = q;
                    s1[0] += pi * q;
                }
            }
            cluster_sum_n<1>(cluster, sh, 1, s1, scratch, C);
            const double pq = s1[0];
            if (!(pq > 0.0)) { status = 2; break; }             // breakdown; identical in every CTA
            const double alpha = rz / pq;
            // ---- phase 2: x += alpha p, r -= alpha q, z = r / d
            double s2[2] = {0.0, 0.0};
#pragma unroll
            for (int k = 0; k < MAXR; k++) {
                int li = t + k * CT;
                if (li < nown) {
                    x_r[k] += alpha * p_s[li];
                    double ri = r_s[li] - alpha * qz_r[k];
                    double zi = ri / d_s[li];
                    r_s[li] = ri;
                    qz_r[k] = zi;
                    s2[0] += ri * zi;
                    s2[1] += ri * ri;
                }
            }
            cluster_sum_n<2>(cluster, sh, 0, s2, scratch, C);
            const double beta = s2[0] / rz;
            rz = s2[0];
            rr = s2[1];
            it++;
            if (rr <= tol2) { status = 0; break; }
            if (it >= a.maxit) { status = 1; break; }
            if (!(rr == rr)) { status = 2; break; }
            // ---- phase 3: p = z + beta p, publish for the halo gathers of the other CTAs
#pragma unroll
            for (int k = 0; k < MAXR; k++) {
                int li = t + k * CT;
                if (li < nown) {
                    double pn = qz_r[k] + beta * p_s[li];
                    p_s[li] = pn;
                    pg[own0 + li] = pn;
                }
            }
            cluster.sync();
        }
        // ---- write back
#pragma unroll
        for (int k = 0; k < MAXR; k++) {
            int li = t + k * CT;
            if (li < nown) x[own0 + li] = x_r[k];
        }
        if (crank == 0 && t == 0) {
            a.iters[sy.out] = it;
            a.status[sy.out] = status;
            a.relres[sy.out] = bb > 0 ? sqrt(rr / bb) : 0.0;
        }
        cluster.sync();     // nobody starts the next system (overwriting smem / pglob) early
    }
}

constexpr size_t SMEM_BYTES = sizeof(Shared) + 96 * sizeof(double) + 3 * (size_t)CAP * sizeof(double);

static int pick_cluster(int max_rows) {
    int c = 1;
    while (c < MAXC && (int64_t)c * CAP < max_rows) c <<= 1;
    return c;
}
}  // namespace

extern "C" int64_t cnt_pcg_cluster_workspace_bytes(int NSYS, int NSLAB, int64_t R) {
    return cnt::ws_bytes_for(NSYS, sizeof(CSys)) + cnt::ws_bytes_for((int64_t)NSLAB * R + 1, 8) + cnt::ws_bytes_for(1, 4);
}

// largest system (rows) the cluster kernel can take
extern "C" int64_t cnt_pcg_cluster_max_rows(void) { return (int64_t)MAXC * CAP; }

static int launch_config(int C, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr, cudaStream_t st, int *nclusters) {
    CNT_CUDA(cudaFuncSetAttribute(k_pcg_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    if (C > 8) CNT_CUDA(cudaFuncSetAttribute(k_pcg_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    *cfg = cudaLaunchConfig_t{};
    cfg->blockDim = dim3(CT, 1, 1);
    cfg->dynamicSmemBytes = SMEM_BYTES;
    cfg->stream = st;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
    cfg->gridDim = dim3((unsigned)C, 1, 1);
    cudaError_t e = cudaOccupancyMaxActiveClusters(nclusters, k_pcg_cluster, cfg);
    if (e != cudaSuccess || *nclusters <= 0) {
        cnt::set_error("cluster size %d not launchable: %s", C, cudaGetErrorString(e));
        cudaGetLastError();
        return CNT_ECUDA;
    }
    return CNT_OK;
}

// number of clusters of C CTAs that can be resident at once on the current device (0 on error)
extern "C" int cnt_pcg_cluster_active(int C) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int n = 0;
    if (C < 1 || C > MAXC || (C & (C - 1))) return 0;
    if (launch_config(C, &cfg, attr, nullptr, &n) != CNT_OK) return 0;
    return n;
}

extern "C" int cnt_pcg_cluster_solve(int B, int NSYS, int NSLAB, const int32_t *h_sys_net, const int32_t *h_sys_slab,
                                     const int32_t *h_roff, int64_t R, int64_t NNZ, const int32_t *rowptr,
                                     const int32_t *col, const double *val, const double *diag, const double *rhs,
                                     double *x, double rtol, int maxit, int32_t *iters, double *relres,
                                     int32_t *status, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && NSYS > 0 && NSLAB > 0 && h_sys_net && h_sys_slab && h_roff && R >= 0 && NNZ >= 0 && iters &&
                  relres && status && ws);
    CNT_CHECK_ARG(rtol > 0 && maxit >= 0 && h_roff[B] == R);
    if (R > 0) CNT_CHECK_ARG(rowptr && diag && rhs && x && (NNZ == 0 || (col && val)));
    std::vector<CSys> sys(NSYS);
    int max_rows = 0;
    for (int s = 0; s < NSYS; s++) {
        int b = h_sys_net[s];
        CNT_CHECK_ARG(b >= 0 && b < B && h_sys_slab[s] >= 0 && h_sys_slab[s] < NSLAB);
        sys[s].row0 = h_roff[b];
        sys[s].nrows = h_roff[b + 1] - h_roff[b];
        sys[s].slab = h_sys_slab[s];
        sys[s].out = s;
        if (sys[s].nrows > max_rows) max_rows = sys[s].nrows;
    }
    if (max_rows > MAXC * CAP) {
        cnt::set_error("system with %d rows exceeds the cluster kernel's on-chip capacity (%d)", max_rows, MAXC * CAP);
        return CNT_ECAPACITY;
    }
    cnt::Workspace w(ws, ws_bytes);
    CSys *d_sys = w.take<CSys>(NSYS);
    double *pglob = w.take<double>((int64_t)NSLAB * R + 1);
    int *counter = w.take<int>(1);
    if (!counter) {
        cnt::set_error("pcg_cluster workspace too small");
        return CNT_ECAPACITY;
    }
    CNT_CUDA(cudaMemcpyAsync(d_sys, sys.data(), sizeof(CSys) * NSYS, cudaMemcpyHostToDevice, st));
    CNT_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));

    int C = pick_cluster(max_rows);
    if (const char *e = getenv("CNT_CLUSTER_SIZE")) {
        int c = atoi(e);
        if (c >= C && c <= MAXC && (c & (c - 1)) == 0) C = c;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int nclusters = 0;
    int rc = launch_config(C, &cfg, attr, st, &nclusters);
    if (rc) return rc;
    if (nclusters > NSYS) nclusters = NSYS;
    cfg.gridDim = dim3((unsigned)(nclusters * C), 1, 1);

    CArgs a;
    a.sys = d_sys; a.nsys = NSYS; a.counter = counter; a.R = R; a.NNZ = NNZ;
    a.rowptr = rowptr; a.col = col; a.val = val; a.diag = diag; a.rhs = rhs;
    a.x = x; a.pglob = pglob; a.rtol = rtol; a.maxit = maxit;
    a.iters = iters; a.status = status; a.relres = relres;
    CNT_CUDA(cudaLaunchKernelEx(&cfg, k_pcg_cluster, a));
    CNT_LAUNCHED();
    return CNT_OK;
}
