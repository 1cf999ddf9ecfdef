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
#ifndef CNT_CLUSTER_THREADS
#define CNT_CLUSTER_THREADS 512
#endif
constexpr int CT = CNT_CLUSTER_THREADS;   // threads per CTA
constexpr int MAXR = 6144 / CT;       // rows per thread
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
    const uint8_t *cls;      // v2: class code per entry (val == gneg[cls])
    const double *gneg;      // v2: [NSLAB][256]
    double *x, *pglob;
    // two-level preconditioner (TWO kernels only)
    cnt_aggregates agg;
    double *rglob;           // [NSLAB*R] r published once per iteration for the segment sums
    double *segS;            // [nseg]
    long long *dbg;          // [gridDim.x][16] per-phase cycle sums (NULL = off)
    int bankopt;             // 1: bank-aware entry order (setup)
    int mb;                  // 1: transaction barriers instead of cluster barriers inside the iteration
    int *assert_flag;        // first failed in-kernel bounds check (0 = none)
    double rtol;
    int maxit;
    int32_t *iters, *status;
    double *relres;
};

struct Shared {
    double red[2][4][16];  // [phase parity][value][CTA rank]: partial sums pushed by the cluster peers
    int next;              // next system (valid in CTA rank 0)
    int nhalo;             // v2: halo references collected by this CTA
    int ovf;               // v2: slice / halo did not fit
    int npush;             // v2: halo values this CTA pushes to its peers (filled by the peers at setup)
    int nrpush;            // v2: segment sums this CTA pushes to its peers
    int pad[3];
    unsigned long long mbar[4];   // v2: transaction barriers: reductions (2 slots), halo of p, segment sums
    long long tlast;       // phase timing (debug): last timestamp, per-phase cycle sums
    long long tacc[16];
    long long pad16;       // sizeof(Shared) is a multiple of 16: the arrays behind it stay 16-byte aligned
};
static_assert(sizeof(Shared) % 16 == 0, "shared-memory layout alignment");
// phase timing: thread 0 of every CTA accumulates clock64() deltas per phase when a debug
// buffer is given (tools/pcg_probe.py); costs one predictable branch otherwise
// cheap in-kernel bounds assertions (compute-sanitizer is not available on the target pool):
// the first failing check id is recorded and the access is skipped
#define KCHECK(id, cond) ((cond) ? true : (atomicCAS(a.assert_flag, 0, (id)), false))
#define PHASE(i)                                                   \
    do {                                                           \
        if (PROF && a.dbg && threadIdx.x == 0) {                           \
            long long c_ = clock64();                              \
            sh->tacc[i] += c_ - sh->tlast;                         \
            sh->tlast = c_;                                        \
        }                                                          \
    } while (0)

// ---- point-to-point synchronisation: transaction barriers (mbarrier) in shared memory ------
// A value pushed to a peer with st.async carries the address of an mbarrier in the peer's shared
// memory and completes 8 bytes of that barrier's transaction count when it lands.  The receiver
// posts the number of bytes it expects (one arrive.expect_tx per phase) and waits on its OWN
// barrier: it continues as soon as ITS inputs are there -- no cluster-wide rendezvous, no
// barrier.cluster round trip (~1000 cycles with 16 CTAs here), and the stores need not drain
// before anything.  Bytes that land before expect_tx is posted are accounted for (the count is
// signed); a phase cannot complete before the local arrive.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol error must not hang the GPU -- record it and abort the kernel.  The bound is WALL TIME
// (20 s on the global timer, looked at every 2^16 polls), not a poll count: a peer CTA that is merely slow -- time
// slicing, the twin-engine co-run, a profiler replay -- must not trip it, because a trap poisons the whole context.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity, int *assert_flag, int id) {
    unsigned long long t0 = 0;
    for (unsigned i = 1;; i++) {
        if (mbar_try_wait(bar, parity)) return;
        if ((i & 0xFFFFu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) break;
        }
    }
    atomicCAS(assert_flag, 0, id);
    __trap();
}
__device__ __forceinline__ uint32_t map_cluster_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}

// cluster-wide sum over transaction barriers: same arithmetic (and bits) as cluster_sum_n
template <int N>
__device__ __forceinline__ void cluster_sum_mb(Shared *sh, int slot, uint32_t parity, double (&v)[N],
                                               double *scratch /* 32*N */, int C, int crank, int *assert_flag) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++) v[k] = cnt::warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; k++) scratch[k * 32 + w] = v[k];
    }
    __syncthreads();
    if (w == 0) {
        if (lane == 0) mbar_expect_tx(&sh->mbar[slot], (uint32_t)(C * N * 8));
        const uint32_t rbar = map_cluster_u32(smem_u32(&sh->mbar[slot]), lane < C ? lane : 0);
#pragma unroll
        for (int k = 0; k < N; k++) {
            double t = lane < (CT >> 5) ? scratch[k * 32 + lane] : 0.0;
            t = cnt::warp_sum(t);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (lane < C) st_async_f64(map_cluster_u32(smem_u32(&sh->red[slot][k][crank]), lane), t, rbar);
        }
    }
    mbar_wait(&sh->mbar[slot], parity, assert_flag, 30 + slot);
#pragma unroll
    for (int k = 0; k < N; k++) {
        double t = lane < C ? sh->red[slot][k][lane] : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        v[k] = __shfl_sync(0xffffffffu, t, 0);
    }
}

// cluster-wide sum of N values, bit-identical in every thread of every CTA.
// Warp sums -> warp 0 adds the CTA's 32 warp partials -> lanes 0..C-1 PUSH the CTA's partial
// into slot [rank] of every peer's red[][] through DSMEM (remote stores retire asynchronously,
// the cluster barrier's release/acquire makes them visible) -> after the barrier every warp adds
// the C partials from its OWN shared memory in a fixed tree over ranks.  No remote load and no
// CTA barrier after the cluster barrier: 2 % faster per iteration than pulling the partials.
// `slot` alternates between consecutive reductions: a peer may already be pushing the next
// reduction's partials while this CTA still reads the current ones.  scratch[] needs no guard
// barrier: its readers (warp 0) are done before this reduction's cluster barrier, its next
// writers start after it.
template <int N>
__device__ __forceinline__ void cluster_sum_n(cg::cluster_group &cluster, Shared *sh, int slot, double (&v)[N],
                                                 double *scratch /* 32*N */, int C) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++) v[k] = cnt::warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; k++) scratch[k * 32 + w] = v[k];
    }
    __syncthreads();
    if (w == 0) {
        const int crank = (int)cluster.block_rank();
        Shared *peer = lane < C ? cluster.map_shared_rank(sh, lane) : sh;
#pragma unroll
        for (int k = 0; k < N; k++) {
            double t = lane < (CT >> 5) ? scratch[k * 32 + lane] : 0.0;
            t = cnt::warp_sum(t);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (lane < C) peer->red[slot][k][crank] = t;
        }
    }
    cluster.sync();
#pragma unroll
    for (int k = 0; k < N; k++) {
        double t = lane < C ? sh->red[slot][k][lane] : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        v[k] = __shfl_sync(0xffffffffu, t, 0);
    }
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


// =====================================================================================
// v2: the CTA's slice of the matrix lives in shared memory for the whole solve.
// Per entry one 32-bit word: (slot << 8) | class, where slot indexes p_ext[] -- the CTA's own
// p values followed by the halo values it needs from other CTAs -- and class indexes a
// 256-entry table of -g (conductances take only a handful of distinct values per gate
// configuration: junction kind x gate level).  An iteration then touches global memory only
// to publish p (8 B/row), refresh the halo (one coalesced pass of L2 loads) and update x.
// =====================================================================================
constexpr int MAXR2 = 5120 / CT;       // rows per thread
constexpr int CAP2 = CT * MAXR2;       // rows per CTA (5120)
constexpr int NRS = 4096 / CT;         // row slots of the variant for chunks of up to 4096 rows
constexpr int ECAP = 20224;            // packed entries per CTA
constexpr int HCAP = 2048;             // halo references per CTA

__device__ __forceinline__ int block_excl_scan_i(int v, int *sm /*34 ints*/, int *total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) sm[w] = incl;
    __syncthreads();
    if (w == 0) {
        int sv = lane < (CT >> 5) ? sm[lane] : 0;
        int si = sv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += u;
        }
        sm[lane] = si - sv;
        if (lane == 31) sm[32] = si;
    }
    __syncthreads();
    int r = sm[w] + incl - v;
    *total = sm[32];
    __syncthreads();
    return r;
}

// z = r/d + einv[a] * (sum of r over the row's aggregate).  The aggregate tables are built for
// this cluster size: a segment never straddles two CTAs' row chunks, so every CTA sums ITS
// segments straight from its shared-memory copy of r (fixed order inside a segment); after one
// cluster barrier each row adds the segments of its aggregate in order -> deterministic.
__device__ __forceinline__ void local_segment_sums(const CArgs &a, int sys_index, int crank, int C, int64_t row_base,
                                                   const double *r_s) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q0 = __ldg(a.agg.cta_seg_off + sys_index * C + crank);
    const int q1 = __ldg(a.agg.cta_seg_off + sys_index * C + crank + 1);
    const int4 *list = reinterpret_cast<const int4 *>(a.agg.cta_seg);
    for (int q = q0 + w; q < q1; q += (CT >> 5)) {
        int4 sg = __ldg(list + q);                  // (segment id, first member, count, -)
        double acc = 0.0;
        for (int k = lane; k < sg.z; k += 32) acc += r_s[(int64_t)__ldg(a.agg.mem + sg.y + k) - row_base];
        acc = cnt::warp_sum(acc);
        if (lane == 0) a.segS[sg.x] = acc;
    }
}
__device__ __forceinline__ double cluster_coarse_term(const CArgs &a, int64_t v) {
    int2 rq = __ldg(reinterpret_cast<const int2 *>(a.agg.row_q) + v);
    if (rq.y == 0) return 0.0;
    double S = __ldcg(a.segS + rq.x);                       // most aggregates have one segment
    for (int q = 1; q < rq.y; q++) S += __ldcg(a.segS + rq.x + q);
    return S * __ldg(a.agg.row_einv + v);
}

constexpr int SORTN = 8192;             // power of two >= CAP2 (bitonic sort of the CTA's rows)
constexpr int LCAP = 512;               // two-level: local segments / local aggregates per CTA
constexpr int MCAP = CAP2;              // two-level: aggregated rows per CTA (all of them may be aggregated)

// two-level phases from shared memory
constexpr int RCAP = 512;        // two-level: segment references of the CTA's multi-segment aggregates
// ---- segment sums as a block-wide segmented reduction -------------------------------------
// The CTA's segments are stored back to back in segm[0..M) (member rows, local indices).  Thread
// t owns the K consecutive members t*K .. t*K+K-1 whatever segment they belong to, so every
// thread does the same work for any distribution of segment sizes (a CTA holds a few hundred
// segments of 2..256 rows).  meta (loop invariant, one register): index of the segment of the
// thread's first member | K bits "member q is the last one of its segment" << 16.
// Inside a thread: running sum, closed at every end flag.  Across threads: an open run (members
// after the thread's last end flag, or all of them if it has none) is carried to the thread that
// closes it by a segmented warp scan (shuffles, reset at threads that have an end flag) and one
// more over the 32 warp aggregates.  Fixed summation tree -> deterministic.
// returns the sum of the run that is open when the thread's first member starts
__device__ __forceinline__ double block_open_run_carry(double trail, bool has_end, double *scratch /* 64 */,
                                                       int nactive /* threads that hold elements */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool active = w * 32 < nactive;                            // warp-uniform
    unsigned R = 0;
    double prev = 0.0;
    if (active) {
        R = __ballot_sync(0xffffffffu, has_end);
        const unsigned below = R & (0xFFFFFFFFu >> (31 - lane));     // resets in lanes 0..lane
        const int dist = below ? lane - (31 - __clz(below)) : lane;  // lanes that may be added from the left
        double incl = trail;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (o <= dist) incl += up;
        }
        prev = __shfl_up_sync(0xffffffffu, incl, 1);                 // open run at the end of lane-1
        if (lane == 31) {
            scratch[w] = incl;
            scratch[32 + w] = R ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    if (!active) return 0.0;
    double carry = lane > 0 ? prev : 0.0;
    if ((R & ((1u << lane) - 1u)) == 0u) {
        // no end flag in lanes 0..lane-1: the run was opened in an earlier warp; walk back over the
        // warps it spans (a segment covers a few warps at most)
        for (int j = w - 1; j >= 0; j--) {
            carry += scratch[j];
            if (scratch[32 + j] != 0.0) break;
        }
    }
    return carry;
}

template <int K, bool PAD>
__device__ __forceinline__ void smem_segscan_sums(double *seg_loc, double *scratch, const double *r_s,
                                                  const uint16_t *segm, int meta, int M) {
    const int base = (int)threadIdx.x * K;
    double v[K];
    if (K == 8 && PAD) {
        // PAD: members M .. K*CT-1 of the list point at a slot of r_s that holds 0 -> no bounds checks,
        // and the thread's 8 indices are one 16-byte load
        const uint4 w = *reinterpret_cast<const uint4 *>(segm + base);
        v[0] = r_s[w.x & 0xFFFFu]; v[1] = r_s[w.x >> 16]; v[2] = r_s[w.y & 0xFFFFu]; v[3] = r_s[w.y >> 16];
        v[4] = r_s[w.z & 0xFFFFu]; v[5] = r_s[w.z >> 16]; v[6] = r_s[w.w & 0xFFFFu]; v[7] = r_s[w.w >> 16];
    } else {
#pragma unroll
        for (int q = 0; q < K; q++) v[q] = base + q < M ? r_s[segm[base + q]] : 0.0;
    }
    // branch-free walk: running sum, closed at every end flag; the first closed run still needs the
    // carry from the threads before, later ones are complete and go straight to seg_loc
    const unsigned ef = (unsigned)meta >> 16;
    const int s0 = meta & 0xFFFF;
    double run = 0.0, first = 0.0;
#pragma unroll
    for (int q = 0; q < K; q++) {
        run += v[q];
        const bool fl = (ef >> q) & 1u;
        const unsigned before = ef & ((1u << q) - 1u);
        if (fl && before == 0u) first = run;
        if (fl && before != 0u) seg_loc[s0 + __popc(before)] = run;
        if (fl) run = 0.0;
    }
    const double carry = block_open_run_carry(run, ef != 0u, scratch, (M + K - 1) / K);
    if (ef != 0u) seg_loc[s0] = carry + first;
}

// coarse value of every local aggregate = 1/E * (sum of the aggregate's segment sums, in global
// segment order).  Segment sums live in the shared memory of the CTA that owns the segment.
// Single-segment aggregates (the bulk) read their own CTA's array.  The segment sums of the
// multi-segment aggregates were pushed into ref_val[] by their owners before the barrier (one
// slot per reference, stored back to back per aggregate); they are added per aggregate with the
// same segmented block reduction as the segment sums.  ref_meta (one register): local
// aggregate index | "last reference of its aggregate" << 15.
__device__ __forceinline__ void smem_coarse_values(const double *seg_loc, const uint16_t *la_a,
                                                   const uint16_t *la_qn, const float *la_einv, const double *ref_val,
                                                   double *scratch, double *c_loc, int nla, int nrefs, int ref_meta) {
    const double v = (int)threadIdx.x < nrefs ? ref_val[threadIdx.x] : 0.0;
    for (int i = threadIdx.x; i < nla; i += CT)
        if (la_qn[i] == 1) c_loc[i] = seg_loc[la_a[i]] * (double)la_einv[i];
    const bool is_end = (int)threadIdx.x < nrefs && ((ref_meta >> 15) & 1);
    const double carry = block_open_run_carry(is_end ? 0.0 : v, is_end, scratch, nrefs);
    if (is_end) {
        int ai = ref_meta & 0x7FFF;
        c_loc[ai] = (carry + v) * (double)la_einv[ai];
    }
}

// owner-side pushes (lists built at setup): plain remote stores, made visible by the cluster
// barrier that follows; the caller has made the source values visible inside the CTA
__device__ __forceinline__ void push_halo(cg::cluster_group &cluster, double *p_ext, const uint32_t *push_list,
                                          int npush) {
    for (int k = threadIdx.x; k < npush; k += CT) {
        uint32_t e = push_list[k];
        cluster.map_shared_rank(p_ext, e >> 26)[(e >> 13) & 8191u] = p_ext[e & 8191u];
    }
}
__device__ __forceinline__ void push_halo_mb(double *p_ext, const uint32_t *push_list, int npush,
                                             unsigned long long *bar) {
    const uint32_t pbase = smem_u32(p_ext), lbar = smem_u32(bar);
    for (int k = threadIdx.x; k < npush; k += CT) {
        uint32_t e = push_list[k];
        st_async_f64(map_cluster_u32(pbase + 8u * ((e >> 13) & 8191u), e >> 26), p_ext[e & 8191u],
                     map_cluster_u32(lbar, e >> 26));
    }
}
__device__ __forceinline__ void push_segment_sums_mb(const double *seg_loc, double *ref_val, const uint32_t *rpush_list,
                                                     int nrpush, unsigned long long *bar) {
    const uint32_t rbase = smem_u32(ref_val), lbar = smem_u32(bar);
    for (int k = threadIdx.x; k < nrpush; k += CT) {
        uint32_t e = rpush_list[k];
        st_async_f64(map_cluster_u32(rbase + 8u * ((e >> 9) & 511u), e >> 18), seg_loc[e & 511u],
                     map_cluster_u32(lbar, e >> 18));
    }
}
__device__ __forceinline__ void push_segment_sums(cg::cluster_group &cluster, const double *seg_loc, double *ref_val,
                                                  const uint32_t *rpush_list, int nrpush) {
    for (int k = threadIdx.x; k < nrpush; k += CT) {
        uint32_t e = rpush_list[k];
        cluster.map_shared_rank(ref_val, e >> 18)[(e >> 9) & 511u] = seg_loc[e & 511u];
    }
}

// NR = row slots per thread (registers): 4 when every CTA's chunk has at most 4*CT rows, else 5;
// PROF = per-phase cycle accumulators compiled in (launched only when a phase buffer is set)
template <bool TWO, int NR, bool PROF>
__global__ void __launch_bounds__(CT, 1) k_pcg_cluster2(CArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    Shared *sh = reinterpret_cast<Shared *>(smem_raw);
    double *scratch = reinterpret_cast<double *>(smem_raw + sizeof(Shared));   // 4*32 doubles
    double *gval = scratch + 128;                                               // 256
    double *p_ext = gval + 256;                                                 // CAP2 + HCAP
    double *r_s = p_ext + (CAP2 + HCAP);                                        // CAP2
    uint32_t *ent = reinterpret_cast<uint32_t *>(r_s + CAP2);                   // ECAP (sort keys during setup)
    uint32_t *push_list = ent + ECAP;                                           // HCAP: consumer << 26 | its slot << 13 | own row
    int32_t *halo_tmp = reinterpret_cast<int32_t *>(r_s);                       // setup only: owner << 16 | its row
    int *iscr = reinterpret_cast<int *>(push_list + HCAP);                       // 36
    int *gbase = iscr + 36;                                                     // CAP2 / 32
    uint16_t *slot_s = reinterpret_cast<uint16_t *>(gbase + CAP2 / 32);         // CAP2: row -> local aggregate
    // two-level tables of this CTA (shared memory only in the iteration)
    uint16_t *segm = slot_s + CAP2;                                             // MCAP member rows (positions), 16-byte aligned
    double *seg_loc = reinterpret_cast<double *>(segm + MCAP);                  // LCAP sums of this CTA's segments
    double *c_loc = seg_loc + LCAP;                                             // LCAP coarse values
    double *ref_val = c_loc + LCAP;                                             // RCAP segment sums pushed by their owners
    uint32_t *rpush_list = reinterpret_cast<uint32_t *>(ref_val + RCAP);        // RCAP: consumer << 18 | its reference << 9 | own segment
    uint32_t *refs = reinterpret_cast<uint32_t *>(r_s) + HCAP;                  // setup only: owner CTA << 16 | index
    uint16_t *pos_of = reinterpret_cast<uint16_t *>(r_s) + 2 * (HCAP + RCAP);   // setup only: local row -> position
    float *la_einv = reinterpret_cast<float *>(rpush_list + RCAP);                    // LCAP 1/E (FP32 is plenty for M)
    uint16_t *la_qn = reinterpret_cast<uint16_t *>(la_einv + LCAP);             // LCAP segments of aggregate
    uint16_t *la_a = la_qn + LCAP;                                              // LCAP own segment / first reference
    uint16_t *segoff = la_a + LCAP;                                             // LCAP + 2
    const int t = threadIdx.x, lane = t & 31;
    if (t == 0) {
        for (int i = 0; i < 16; i++) sh->tacc[i] = 0;
        sh->tlast = clock64();
        for (int i = 0; i < 4; i++) mbar_init(&sh->mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // ordered by the first cluster barrier
    }
    uint32_t mpar = 0;                  // bit i: parity of the next phase of mbar[i]
    const bool mb = a.mb != 0;
    // cluster-wide sum / wait for pushed data: transaction barriers, or cluster barriers (reference path)
#define CSUM(N, slot, arr)                                                                                  \
    do {                                                                                                    \
        if (mb) {                                                                                           \
            cluster_sum_mb<N>(sh, slot, (mpar >> (slot)) & 1u, arr, scratch, C, crank, a.assert_flag);      \
            mpar ^= 1u << (slot);                                                                           \
        } else {                                                                                            \
            cluster_sum_n<N>(cluster, sh, slot, arr, scratch, C);                                           \
        }                                                                                                   \
    } while (0)
#define WAIT_PUSHED(i, count)                                                                               \
    do {                                                                                                    \
        if (mb) {                                                                                           \
            if (t == 0) mbar_expect_tx(&sh->mbar[i], (uint32_t)(count) * 8u);                               \
            mbar_wait(&sh->mbar[i], (mpar >> (i)) & 1u, a.assert_flag, 30 + (i));                           \
            mpar ^= 1u << (i);                                                                              \
        } else {                                                                                            \
            cluster.sync();                                                                                 \
        }                                                                                                   \
    } while (0)

    while (true) {
        if (crank == 0 && t == 0) sh->next = atomicAdd(a.counter, 1);
        cluster.sync();
        const int sidx = cluster.map_shared_rank(sh, 0)->next;
        cluster.sync();
        if (sidx >= a.nsys) break;
        const CSys sy = a.sys[sidx];
        const int chunk = (sy.nrows + C - 1) / C;
        const int own0 = crank * chunk;
        const int nown = max(0, min(chunk, sy.nrows - own0));
        const int64_t vbase = (int64_t)sy.slab * a.R + sy.row0;
        const uint8_t *cls = a.cls + (int64_t)sy.slab * a.NNZ;
        const double *diag = a.diag + vbase + own0;
        const double *rhs = a.rhs + vbase + own0;
        double *x = a.x + vbase;                   // network-local index (gathers of x0 use columns)
        const int32_t *rowptr = a.rowptr + sy.row0 + own0;

        // ================= setup (once per system) =================
        PHASE(0);                                               // queue / idle
        if (t == 0) { sh->nhalo = 0; sh->ovf = 0; sh->npush = 0; sh->nrpush = 0; sh->pad[0] = 0; sh->pad[1] = 0; }
        if (t < 256) gval[t] = a.gneg[(int64_t)sy.slab * 256 + t];
        // (1) sort the CTA's rows by length, longest first, ties by row index: every warp then
        //     works on 32 rows of (almost) equal length -> no divergence in the SpMV loop
        for (int i = t; i < SORTN; i += CT) {
            unsigned key = 0;                                   // padding sorts last
            if (i < nown) {
                unsigned len = (unsigned)(__ldg(rowptr + i + 1) - __ldg(rowptr + i));
                key = (len << 13) | (unsigned)(8191 - i);
            }
            ent[i] = 0xFFFFFFFFu - key;
        }
        __syncthreads();
        for (int k = 2; k <= SORTN; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = t; i < SORTN; i += CT) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        unsigned va = ent[i], vb = ent[ixj];
                        bool up = (i & k) == 0;
                        if ((va > vb) == up) { ent[i] = vb; ent[ixj] = va; }
                    }
                }
                __syncthreads();
            }
        }
        // (2) per slot: row, length; group bases of the warp-interleaved entry layout
        // "snake" assignment of the sorted rows to (thread, slot): odd slots run backwards so that
        // every warp gets a similar total row length
        int meta_r[NR];                                      // li | len << 16   (-1: inactive slot)
        int base = 0;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < NR; k++) {
            int p = k * CT + ((k & 1) ? (CT - 1 - t) : t);
            int len = 0;
            meta_r[k] = -1;
            if (p < nown) {
                unsigned key = 0xFFFFFFFFu - ent[p];
                len = (int)(key >> 13);
                int li = 8191 - (int)(key & 8191u);
                if (len > 255) { bad = true; len = 255; }
                meta_r[k] = li | (len << 16);
                pos_of[li] = (uint16_t)p;
            }
            int total;
            // group = 32 consecutive positions; its longest row is the one with the smallest position
            const bool leader = (p & 31) == 0;
            int off = base + block_excl_scan_i(leader ? 32 * len : 0, iscr, &total);
            if (leader) gbase[p >> 5] = off;
            base += total;
        }
        if (bad) sh->ovf = 1;
        const bool fits = base <= ECAP;
        __syncthreads();                                        // sort keys consumed, gbase visible
        // (3) packed entries: (slot << 8) | class at gbase + j*32 + lane
        if (fits) {
#pragma unroll
            for (int k = 0; k < NR; k++) {
                if (meta_r[k] >= 0) {
                    int li = meta_r[k] & 0xFFFF, len = meta_r[k] >> 16;
                    int k0 = __ldg(rowptr + li);
                    const int p = k * CT + ((k & 1) ? (CT - 1 - t) : t);
                    unsigned *dst = ent + gbase[p >> 5] + (p & 31);
                    for (int j = 0; j < len; j++) {
                        int c = __ldg(a.col + k0 + j);
                        unsigned cl = __ldg(cls + k0 + j);
                        unsigned lc = (unsigned)(c - own0);
                        unsigned slot;
                        if (lc < (unsigned)nown) {
                            slot = pos_of[lc];                  // vectors are stored by position
                        } else {
                            int h = atomicAdd(&sh->nhalo, 1);
                            if (h < HCAP) {
                                halo_tmp[h] = ((c / chunk) << 16) | (c % chunk);   // owner CTA, its local row
                                slot = (unsigned)(nown + h);
                            } else {
                                slot = 0;
                                sh->ovf = 1;
                            }
                        }
                        if (KCHECK(4, (dst - ent) + j * 32 < ECAP)) dst[j * 32] = (slot << 14) | (cl << 3);   // byte offsets: p_ext << 11 | gval
                    }
                }
            }
        }
        // (3b) bank-aware entry order: the SpMV gathers p_ext[slot] of the j-th entries of 32 rows at
        // once; 16 random 8-byte words per half-warp hit the 16 bank pairs ~3 deep.  Row sums do not
        // care about the order of a row's entries, so every row (lanes one after the other, first
        // 8 columns) greedily puts into column j the entry whose bank is least used in that column
        // (deterministic; costs ~8 iterations' worth of setup per system, gains 1.3 % per iteration).
        if (fits && a.bankopt) {
            __syncthreads();                                    // all entries packed, halo slots assigned
            uint8_t *tab = reinterpret_cast<uint8_t *>(r_s) + 24576 + (t >> 5) * 256;   // [half][col < 8][bank pair]
#pragma unroll
            for (int k = 0; k < NR; k++) {
                reinterpret_cast<unsigned long long *>(tab)[lane] = 0ull;
                __syncwarp();
                const int p = k * CT + ((k & 1) ? (CT - 1 - t) : t);
                const int n8 = meta_r[k] >= 0 ? min(meta_r[k] >> 16, 8) : 0;
                unsigned *dst = ent + gbase[p >> 5] + (p & 31);
                uint8_t *tb = tab + (lane >> 4) * 128;
                for (int L = 0; L < 32; L++) {
                    if (lane == L) {
                        // halo slots are handed out by an atomic counter: only the entries that point at
                        // the CTA's own rows (deterministic positions) take part; they move to the front
                        // (stable), the halo entries keep their order behind them
                        int nl = 0;
                        for (int i = 0; i < n8; i++) {
                            unsigned u = dst[i * 32];
                            if ((int)(u >> 14) < nown) {
                                for (int m = i; m > nl; m--) dst[m * 32] = dst[(m - 1) * 32];
                                dst[nl * 32] = u;
                                nl++;
                            }
                        }
                        for (int j = 0; j < nl; j++) {
                            int best = j, bestc = 256;
                            for (int i = j; i < nl; i++) {
                                int c = tb[j * 16 + ((dst[i * 32] >> 14) & 15u)];
                                if (c < bestc) { bestc = c; best = i; }
                            }
                            unsigned eb = dst[best * 32];
                            dst[best * 32] = dst[j * 32];
                            dst[j * 32] = eb;
                            tb[j * 16 + ((eb >> 14) & 15u)]++;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        int nseg_loc = 0, nla = 0, nrefs = 0, seg_meta = 0, seg_M = 0, ref_meta = 0;
        if (TWO) {
            // (4) this CTA's segments (member lists as local row indices) and aggregates
            const int cidx = sidx * C + crank;
            const int sq0 = __ldg(a.agg.cta_seg_off + cidx), la0 = __ldg(a.agg.cta_agg_off + cidx);
            nseg_loc = __ldg(a.agg.cta_seg_off + cidx + 1) - sq0;
            nla = __ldg(a.agg.cta_agg_off + cidx + 1) - la0;
            const int4 *list = reinterpret_cast<const int4 *>(a.agg.cta_seg) + sq0;
            const bool cap_ok = nseg_loc <= LCAP && nla <= LCAP;
            int4 sg = make_int4(0, 0, 0, 0);
            if (cap_ok && t < nseg_loc) sg = __ldg(list + t);
            int total;
            int off = block_excl_scan_i(sg.z, iscr, &total);
            bool tw_ok = cap_ok && total <= MCAP;
            // references of the multi-segment aggregates: exclusive scan of their segment counts
            int myq0 = 0, myqn = 0;
            if (cap_ok && t < nla) {
                int aid = __ldg(a.agg.cta_agg + la0 + t);
                myq0 = __ldg(a.agg.agg_seg0 + aid);
                myqn = __ldg(a.agg.agg_seg0 + aid + 1) - myq0;
            }
            int rtotal;
            int roff = block_excl_scan_i(myqn > 1 ? myqn : 0, iscr, &rtotal);
            tw_ok = tw_ok && rtotal <= RCAP;
            nrefs = tw_ok ? rtotal : 0;
            if (!tw_ok) {
                // tables do not fit: the system gets status 3 once the cluster has agreed on it; until
                // then the phases below must see consistent (empty) tables
                if (t == 0) sh->ovf = 1;
                nseg_loc = 0;
                nla = 0;
                for (int i = t; i < nown; i += CT) slot_s[i] = (uint16_t)0xFFFF;
            } else {
                if (t < nseg_loc) {
                    KCHECK(7, sg.x >= 0 && sg.x < a.agg.nseg && sg.y >= 0 && sg.z > 0 && sg.z <= 256);
                    segoff[t] = (uint16_t)off;
                }
                if (t == 0) segoff[nseg_loc] = (uint16_t)total;
                if (t < nla) {
                    int aid = __ldg(a.agg.cta_agg + la0 + t);
                    la_qn[t] = (uint16_t)myqn;
                    la_einv[t] = (float)__ldg(a.agg.agg_einv + aid);
                    if (myqn == 1) {
                        // its only segment is owned by this CTA: index in the local list
                        int home = __ldg(a.agg.seg_home + myq0);
                        KCHECK(5, (home >> 16) == crank && (home & 0xFFFF) < nseg_loc);
                        la_a[t] = (uint16_t)(home & 0xFFFF);
                    } else {
                        la_a[t] = (uint16_t)roff;
                        uint16_t *rmeta = reinterpret_cast<uint16_t *>(c_loc);      // setup-time overlay
                        for (int q = 0; q < myqn; q++) {
                            int home = __ldg(a.agg.seg_home + myq0 + q);
                            KCHECK(6, (home >> 16) < C && (home & 0xFFFF) < LCAP);
                            refs[roff + q] = (uint32_t)home;
                            rmeta[roff + q] = (uint16_t)(t | (q == myqn - 1 ? 0x8000 : 0));
                        }
                    }
                }
                for (int i = t; i < nown; i += CT) {
                    int sl = __ldg(a.agg.row_slot + vbase + own0 + i);
                    KCHECK(3, sl < nla);
                    slot_s[pos_of[i]] = (sl < 0 || sl >= nla) ? (uint16_t)0xFFFF : (uint16_t)sl;
                }
            }
            __syncthreads();
            for (int sgi = (t >> 5); sgi < nseg_loc; sgi += (CT >> 5)) {
                int4 g4 = __ldg(list + sgi);
                int o = segoff[sgi];
                for (int q = lane; q < g4.z; q += 32) {
                    int64_t loc = (int64_t)__ldg(a.agg.mem + g4.y + q) - (vbase + own0);
                    if (KCHECK(1, o + q < MCAP)) segm[o + q] = KCHECK(2, loc >= 0 && loc < nown) ? pos_of[loc] : (uint16_t)0;
                }
            }
            if (t < nrefs) ref_meta = reinterpret_cast<const uint16_t *>(c_loc)[t];
            // segmented-reduction meta of this thread (see smem_segscan_sums)
            {
                const int M = tw_ok ? total : 0;
                const int mb = t * NR;
                seg_meta = 0;
                if (mb < M) {
                    int lo = 0, hi = nseg_loc;                  // segment s with segoff[s] <= mb < segoff[s+1]
                    while (hi - lo > 1) {
                        int mid = (lo + hi) >> 1;
                        if ((int)segoff[mid] <= mb) lo = mid; else hi = mid;
                    }
                    int sc = lo, ef = 0;
#pragma unroll
                    for (int q = 0; q < NR; q++) {
                        if (mb + q < M && mb + q + 1 == (int)segoff[sc + 1]) { ef |= 1 << q; sc++; }
                    }
                    seg_meta = lo | (ef << 16);
                }
                seg_M = M;
                if (NR == NRS) {
                    // padding of the member list (see smem_segscan_sums): positions of this variant stay
                    // below NRS*CT, so the last slot of r_s is free to hold the zero the padding points at
                    for (int i = M + t; i < NRS * CT; i += CT) segm[i] = (uint16_t)(CAP2 - 1);
                    if (t == 0) r_s[CAP2 - 1] = 0.0;
                }
            }
        }
        __syncthreads();
        // ---- turn the consumers' lists around: every owner gets the list of values it has to PUSH
        // (halo copies of p; segment sums of aggregates that continue in other CTAs), so that an
        // iteration has no remote load -- remote stores retire behind the cluster barrier that
        // follows them anyway.  Entry order depends on the atomics' order; every entry writes its
        // own destination, so results do not.
        cluster.sync();                                         // lists complete, counters reset everywhere
        {
            const int nh = min(sh->nhalo, HCAP);
            for (int h = t; h < nh; h += CT) {
                int hc = halo_tmp[h];
                Shared *osh = cluster.map_shared_rank(sh, hc >> 16);
                int k = atomicAdd(&osh->npush, 1);
                if (k < HCAP) {
                    const uint32_t opos = cluster.map_shared_rank(pos_of, hc >> 16)[hc & 0xFFFF];   // owner's position
                    cluster.map_shared_rank(push_list, hc >> 16)[k] =
                        ((uint32_t)crank << 26) | ((uint32_t)(nown + h) << 13) | opos;
                    atomicAdd(&sh->pad[0], 1);                  // halo values that will actually arrive
                } else {
                    osh->ovf = 1;
                }
            }
            if (TWO) {
                for (int j = t; j < nrefs; j += CT) {
                    uint32_t ref = refs[j];
                    Shared *osh = cluster.map_shared_rank(sh, ref >> 16);
                    int k = atomicAdd(&osh->nrpush, 1);
                    if (k < RCAP) {
                        cluster.map_shared_rank(rpush_list, ref >> 16)[k] =
                            ((uint32_t)crank << 18) | ((uint32_t)j << 9) | (ref & 0x1FFu);
                        atomicAdd(&sh->pad[1], 1);              // segment sums that will actually arrive
                    } else {
                        osh->ovf = 1;
                    }
                }
            }
        }
        cluster.sync();                                         // push lists complete
        const int npush = min(sh->npush, HCAP), nrpush = min(sh->nrpush, RCAP);
        const int nhalo_in = sh->pad[0], nref_in = sh->pad[1];
        const double my_ovf = (!fits || sh->ovf) ? 1.0 : 0.0;

        // ================= r = b - A x0, z = M^-1 r, p = z =================
        double d_r[NR], di_r[NR], qz_r[NR];
#define POS(k) ((k) * CT + (((k) & 1) ? (CT - 1 - t) : t))      /* position of the thread's k-th row */
        double acc[4] = {0.0, 0.0, 0.0, my_ovf};    // rz, bb, rr, overflow flag
#pragma unroll
        for (int k = 0; k < NR; k++) {
            d_r[k] = 1.0; di_r[k] = 1.0; qz_r[k] = 0.0;
            if (meta_r[k] >= 0) {
                int li = meta_r[k] & 0xFFFF, len = meta_r[k] >> 16;
                double di = diag[li], xi = x[own0 + li], bi = rhs[li];
                double ax = di * xi;
                int k0 = __ldg(rowptr + li);
                for (int j = 0; j < len; j++) ax += gval[__ldg(cls + k0 + j)] * x[__ldg(a.col + k0 + j)];
                double ri = bi - ax, dinv = 1.0 / di;
                d_r[k] = di; di_r[k] = dinv;
                r_s[POS(k)] = ri;
                acc[1] += bi * bi;
                acc[2] += ri * ri;
                if (!TWO) {
                    double zi = ri * dinv;
                    p_ext[POS(k)] = zi;
                    acc[0] += ri * zi;
                }
            }
        }
        if (TWO) {
            __syncthreads();                                    // r_s complete in this CTA
            smem_segscan_sums<NR, NR == NRS>(seg_loc, scratch, r_s, segm, seg_meta, seg_M);
            __syncthreads();
            if (mb) push_segment_sums_mb(seg_loc, ref_val, rpush_list, nrpush, &sh->mbar[3]);
            else push_segment_sums(cluster, seg_loc, ref_val, rpush_list, nrpush);
            WAIT_PUSHED(3, nref_in);                            // segment sums of the other CTAs are here
            smem_coarse_values(seg_loc, la_a, la_qn, la_einv, ref_val, scratch, c_loc, nla, nrefs, ref_meta);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < NR; k++) {
                if (meta_r[k] >= 0) {
                    double ri = r_s[POS(k)];
                    unsigned sl = slot_s[POS(k)];
                    double zi = ri * di_r[k] + (sl != 0xFFFFu ? c_loc[sl] : 0.0);
                    p_ext[POS(k)] = zi;
                    acc[0] += ri * zi;
                }
            }
        }
        __syncthreads();                                        // own p complete in this CTA
        if (mb) push_halo_mb(p_ext, push_list, npush, &sh->mbar[2]);
        else push_halo(cluster, p_ext, push_list, npush);
        CSUM(4, 0, acc);                                        // (cluster barrier: also publishes the halo copies of p)
        if (mb) WAIT_PUSHED(2, nhalo_in);
        double rz = acc[0];
        const double bb = acc[1];
        double rr = acc[2];
        const double tol2 = a.rtol * a.rtol * bb;
        int it = 0, status = 0;
        bool done = (sy.nrows == 0) || (rr <= tol2);
        if (acc[3] > 0.0) { done = true; status = 3; }          // does not fit on chip: caller re-solves
        if (!done && a.maxit <= 0) { done = true; status = 1; }
        // x lives in position order while the system is being solved (every thread then touches
        // only its own, coalesced, conflict-free locations); everybody has read x0 by now
        const bool entered = !done;
        if (entered) {
            double xin[NR];
#pragma unroll
            for (int k = 0; k < NR; k++) xin[k] = meta_r[k] >= 0 ? __ldcg(x + own0 + (meta_r[k] & 0xFFFF)) : 0.0;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < NR; k++)
                if (meta_r[k] >= 0) x[own0 + POS(k)] = xin[k];
        }
        PHASE(1);                                               // setup + initial residual

        while (!done) {
            PHASE(2);                                           // (halo copies of p arrive as pushes)
            // ---- phase 1: q = A p from shared memory, pq = p.q
            double s1[1] = {0.0};
#pragma unroll
            for (int k = 0; k < NR; k++) {
                if (meta_r[k] >= 0) {
                    const int len = meta_r[k] >> 16;
                    const int p = POS(k);
                    double pi = p_ext[p];
                    double q = d_r[k] * pi;
                    const unsigned *e = ent + gbase[p >> 5] + (p & 31);
                    const char *gb = reinterpret_cast<const char *>(gval), *pb = reinterpret_cast<const char *>(p_ext);
#pragma unroll 2
                    for (int j = 0; j < len; j++) {
                        unsigned u = e[j * 32];             // (byte offset into p_ext) << 11 | byte offset into gval
                        q += *reinterpret_cast<const double *>(gb + (u & 2047u)) * *reinterpret_cast<const double *>(pb + (u >> 11));
                    }
                    qz_r[k] = q;
                    s1[0] += pi * q;
                }
            }
            PHASE(3);                                           // SpMV (thread 0's warp)
            double xv[NR];                                   // x loads fly during the reduction
#pragma unroll
            for (int k = 0; k < NR; k++) xv[k] = meta_r[k] >= 0 ? __ldcg(x + own0 + POS(k)) : 0.0;
            CSUM(1, 1, s1);
            PHASE(4);                                           // reduction 1 incl. waiting for the cluster
            const double pq = s1[0];
            if (!(pq > 0.0)) { status = 2; break; }
            const double alpha = rz / pq;
            // ---- phase 2: x += alpha p (x is stored by position while the system is being solved), r -= alpha q
            double s2[2] = {0.0, 0.0};
#pragma unroll
            for (int k = 0; k < NR; k++) {
                if (meta_r[k] >= 0) {
                    const int p = POS(k);
                    x[own0 + p] = xv[k] + alpha * p_ext[p];
                    double ri = r_s[p] - alpha * qz_r[k];
                    r_s[p] = ri;
                    s2[1] += ri * ri;
                    if (!TWO) {
                        double zi = ri * di_r[k];
                        qz_r[k] = zi;
                        s2[0] += ri * zi;
                    }
                }
            }
            PHASE(5);                                           // x, r update
            if (TWO) {
                __syncthreads();                                // r_s complete in this CTA
                PHASE(12);                                      // (debug) wait for the CTA's other warps
                smem_segscan_sums<NR, NR == NRS>(seg_loc, scratch, r_s, segm, seg_meta, seg_M);
                __syncthreads();
                PHASE(15);                                      // (debug) segmented reduction alone
                if (mb) push_segment_sums_mb(seg_loc, ref_val, rpush_list, nrpush, &sh->mbar[3]);
                else push_segment_sums(cluster, seg_loc, ref_val, rpush_list, nrpush);
                PHASE(6);                                       // pushes of the segment sums
                WAIT_PUSHED(3, nref_in);                        // segment sums of the other CTAs are here
                PHASE(7);                                       // barrier
                smem_coarse_values(seg_loc, la_a, la_qn, la_einv, ref_val, scratch, c_loc, nla, nrefs, ref_meta);
                PHASE(13);                                      // (debug) coarse values, this warp
                __syncthreads();
                PHASE(14);                                      // (debug) wait for the CTA's other warps
#pragma unroll
                for (int k = 0; k < NR; k++) {
                    if (meta_r[k] >= 0) {
                        double ri = r_s[POS(k)];
                        unsigned sl = slot_s[POS(k)];
                        double zi = ri * di_r[k] + (sl != 0xFFFFu ? c_loc[sl] : 0.0);
                        qz_r[k] = zi;
                        s2[0] += ri * zi;
                    }
                }
                PHASE(8);                                       // coarse term + z
            }
            CSUM(2, 0, s2);
            PHASE(9);                                           // reduction 2
            const double beta = s2[0] / rz;
            rz = s2[0];
            rr = s2[1];
            it++;
            if (rr <= tol2) { status = 0; break; }
            if (it >= a.maxit) { status = 1; break; }
            if (!(rr == rr)) { status = 2; break; }
            // ---- phase 3: p = z + beta p, then publish it (coalesced) for the other CTAs' halos
#pragma unroll
            for (int k = 0; k < NR; k++) {
                if (meta_r[k] >= 0) p_ext[POS(k)] = qz_r[k] + beta * p_ext[POS(k)];
            }
            __syncthreads();                                    // own p complete in this CTA
            if (mb) push_halo_mb(p_ext, push_list, npush, &sh->mbar[2]);
            else push_halo(cluster, p_ext, push_list, npush);
            PHASE(10);                                          // p update + halo push
            WAIT_PUSHED(2, nhalo_in);                           // halo copies of the neighbours' p are here
            PHASE(11);                                          // barrier
        }
        if (entered) {                                          // back to row order (a permutation inside the chunk)
            double xf[NR];
#pragma unroll
            for (int k = 0; k < NR; k++) xf[k] = meta_r[k] >= 0 ? __ldcg(x + own0 + POS(k)) : 0.0;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < NR; k++)
                if (meta_r[k] >= 0) x[own0 + (meta_r[k] & 0xFFFF)] = xf[k];
        }
        if (crank == 0 && t == 0) {
            a.iters[sy.out] = it;
            a.status[sy.out] = status;
            a.relres[sy.out] = bb > 0 ? sqrt(rr / bb) : 0.0;
        }
        cluster.sync();
    }
    if (a.dbg && t == 0)
        for (int i = 0; i < 16; i++) a.dbg[(int64_t)blockIdx.x * 16 + i] = sh->tacc[i];
}

constexpr size_t SMEM2_BYTES = sizeof(Shared) + (128 + 256 + (size_t)CAP2 + HCAP + CAP2) * sizeof(double) +
                               (size_t)ECAP * 4 + (size_t)HCAP * 4 + (36 + CAP2 / 32) * 4 + (size_t)CAP2 * 2 +
                               (size_t)LCAP * (8 + 8 + 4 + 2 + 2) + (size_t)RCAP * 4 + ((size_t)LCAP + 2) * 2 +
                               (size_t)MCAP * 2 + (size_t)RCAP * 8 + 16;
static_assert(SMEM2_BYTES <= 232448, "cluster kernel exceeds the 227 KB shared-memory limit");

constexpr size_t SMEM_BYTES = sizeof(Shared) + 96 * sizeof(double) + 3 * (size_t)CAP * sizeof(double);

static int pick_cluster(int max_rows, int cap) {
    int c = 1;
    while (c < MAXC && (int64_t)c * cap < max_rows) c <<= 1;
    return c;
}
}  // namespace

extern "C" int64_t cnt_pcg_cluster_workspace_bytes(int NSYS, int NSLAB, int64_t R) {
    int64_t n = (int64_t)NSLAB * R;
    return cnt::ws_bytes_for(NSYS, sizeof(CSys)) + 2 * cnt::ws_bytes_for(n + 1, 8) + 2 * cnt::ws_bytes_for(1, 4) +
           cnt::ws_bytes_for(n / 2 + n / 256 + 2, 8);
}

// debug: device buffer [grid][16] receiving per-phase cycle sums of the next launches (NULL = off)
static long long *g_phase_dbg = nullptr;
extern "C" void cnt_pcg_cluster_set_phase_buffer(long long *buf) { g_phase_dbg = buf; }

// largest system (rows) the cluster kernel can take
extern "C" int64_t cnt_pcg_cluster_max_rows(void) { return (int64_t)MAXC * CAP2; }

typedef void (*kernel_t)(CArgs);
static int launch_config(kernel_t kern, size_t smem, int C, cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr,
                         cudaStream_t st, int *nclusters) {
    CNT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) CNT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    *cfg = cudaLaunchConfig_t{};
    cfg->blockDim = dim3(CT, 1, 1);
    cfg->dynamicSmemBytes = smem;
    cfg->stream = st;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
    cfg->gridDim = dim3((unsigned)C, 1, 1);
    cudaError_t e = cudaOccupancyMaxActiveClusters(nclusters, kern, cfg);
    if (e != cudaSuccess || *nclusters <= 0) {
        cnt::set_error("cluster size %d not launchable: %s", C, cudaGetErrorString(e));
        cudaGetLastError();
        return CNT_ECUDA;
    }
    return CNT_OK;
}

// smallest power-of-two cluster whose CTAs hold max_rows rows and (v2) the packed entries of a
// system with max_nnz off-diagonal entries (15 % allowance for group padding and uneven chunks)
static int cluster_size_for(int64_t max_rows, int cap, int64_t max_nnz = 0) {
    int C = pick_cluster((int)max_rows, cap);
    while (C < MAXC && (double)C * ECAP < 1.15 * (double)max_nnz) C <<= 1;
    if (const char *e = getenv("CNT_CLUSTER_SIZE")) {
        int c = atoi(e);
        if (c >= 1 && c <= MAXC && (int64_t)c * cap >= max_rows) C = c;     // any size that holds the rows
    }
    return C;
}
// cluster size used for a batch whose largest system has max_rows rows (compressed kernel)
extern "C" int cnt_pcg_cluster_size(int64_t max_rows, int64_t max_nnz) {
    return cluster_size_for(max_rows, CAP2, max_nnz);
}

// number of clusters of C CTAs that can be resident at once on the current device (0 on error)
extern "C" int cnt_pcg_cluster_active(int C) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int n = 0;
    if (C < 1 || C > MAXC) return 0;
    if (launch_config(k_pcg_cluster2<true, MAXR2, false>, SMEM2_BYTES, C, &cfg, attr, nullptr, &n) != CNT_OK) return 0;
    return n;
}

extern "C" int cnt_pcg_cluster_solve(int B, int NSYS, int NSLAB, const int32_t *h_sys_net, const int32_t *h_sys_slab,
                                     const int32_t *h_roff, int64_t R, int64_t NNZ, const int32_t *rowptr,
                                     const int32_t *col, const double *val, const uint8_t *cls, const double *gneg,
                                     const double *diag, const double *rhs,
                                     double *x, double rtol, int maxit, int32_t *iters, double *relres,
                                     int32_t *status, const cnt_aggregates *agg, int cluster_C, void *ws,
                                     int64_t ws_bytes, void *stream) {
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
    const bool v2 = cls && gneg && !getenv("CNT_CLUSTER_V1");
    const bool two = v2 && agg && agg->nseg > 0;
    const int cap = v2 ? CAP2 : CAP;
    if (max_rows > MAXC * cap) {
        cnt::set_error("system with %d rows exceeds the cluster kernel's on-chip capacity (%d)", max_rows, MAXC * cap);
        return CNT_ECAPACITY;
    }
    cnt::Workspace w(ws, ws_bytes);
    CSys *d_sys = w.take<CSys>(NSYS);
    double *pglob = w.take<double>((int64_t)NSLAB * R + 1);
    double *rglob = w.take<double>((int64_t)NSLAB * R + 1);
    int *counter = w.take<int>(1);
    int *assert_flag = w.take<int>(1);
    double *segS = w.take<double>(two ? agg->nseg : 1);
    if (!counter || !segS) {
        cnt::set_error("pcg_cluster workspace too small");
        return CNT_ECAPACITY;
    }
    CNT_CUDA(cudaMemcpyAsync(d_sys, sys.data(), sizeof(CSys) * NSYS, cudaMemcpyHostToDevice, st));
    CNT_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    CNT_CUDA(cudaMemsetAsync(assert_flag, 0, sizeof(int), st));

    int C = cluster_size_for(max_rows, cap);
    if (cluster_C > 0) {
        CNT_CHECK_ARG(cluster_C <= MAXC && (int64_t)cluster_C * cap >= max_rows);
        C = cluster_C;
    }
    if (two && agg->chunk_C != C) {
        cnt::set_error("aggregate tables were built for cluster size %d, the batch needs %d", agg->chunk_C, C);
        return CNT_EINVAL;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int nclusters = 0;
    // chunks of at most 4*CT rows run the variant with 4 row slots per thread (no register spills)
    const bool small = ((int64_t)max_rows + C - 1) / C <= NRS * CT && !getenv("CNT_CLUSTER_NR5");
    kernel_t kern = !v2 ? k_pcg_cluster
                        : g_phase_dbg
                              ? (two ? (small ? k_pcg_cluster2<true, NRS, true> : k_pcg_cluster2<true, MAXR2, true>)
                                     : (small ? k_pcg_cluster2<false, NRS, true> : k_pcg_cluster2<false, MAXR2, true>))
                              : (two ? (small ? k_pcg_cluster2<true, NRS, false> : k_pcg_cluster2<true, MAXR2, false>)
                                     : (small ? k_pcg_cluster2<false, NRS, false> : k_pcg_cluster2<false, MAXR2, false>));
    int rc = launch_config(kern, v2 ? SMEM2_BYTES : SMEM_BYTES, C, &cfg, attr, st, &nclusters);
    if (rc) return rc;
    if (nclusters > NSYS) nclusters = NSYS;
    cfg.gridDim = dim3((unsigned)(nclusters * C), 1, 1);

    CArgs a;
    a.sys = d_sys; a.nsys = NSYS; a.counter = counter; a.R = R; a.NNZ = NNZ;
    a.rowptr = rowptr; a.col = col; a.val = val; a.diag = diag; a.rhs = rhs; a.cls = cls; a.gneg = gneg;
    a.x = x; a.pglob = pglob; a.rtol = rtol; a.maxit = maxit;
    a.rglob = rglob; a.segS = segS;
    a.dbg = g_phase_dbg;
    a.assert_flag = assert_flag;
    a.bankopt = getenv("CNT_CLUSTER_NOBANKOPT") ? 0 : 1;     // measured: 1.3 % per iteration
    a.mb = getenv("CNT_CLUSTER_NOMB") ? 0 : 1;     // reference path: cluster barriers
    if (two) a.agg = *agg; else a.agg = cnt_aggregates{};
    a.iters = iters; a.status = status; a.relres = relres;
    CNT_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    CNT_LAUNCHED();
    if (getenv("CNT_CLUSTER_ASSERT")) {       // debugging aid: synchronise and report in-kernel bounds checks
        int h = 0;
        cudaError_t e = cudaMemcpyAsync(&h, assert_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess || h != 0) {
            cnt::set_error("cluster kernel: %s, in-kernel check %d failed", cudaGetErrorString(e), h);
            return CNT_ECUDA;
        }
    }
    return CNT_OK;
}
