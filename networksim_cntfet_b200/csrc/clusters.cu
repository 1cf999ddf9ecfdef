// Connected components of the junction graph by lock-free union-find with minimum-index
// roots.  Replaces the networkx component search of netsim.py:175-179,197-206 and the
// cluster statistics of measure_perc.py:76-80 / 148-152.
//
// parent[] holds GLOBAL stick ids; a root is only ever hooked under a SMALLER id, so the
// final root of a component is its minimum stick id == the canonical label.
#include "common.cuh"

namespace {

// Parents only ever decrease towards the root.  `s0` is the first
// stick of x's network: the smallest id of the network, hence ALWAYS a root -- it is recognised without
// loading parent[s0].  In a percolating network s0 (the source electrode) is the root of the giant component,
// i.e. the end of almost every chase: that one cache line was read ~130k times per network through L2
// (volatile), serialised at the slice, and made the union pass latency-bound (ncu: 3 % issue slots busy).
// Loads of the chases go through L1 (ld.global.ca): a line may be stale, but a stale parent is still an ancestor
// (parents only move towards the root, and indices strictly decrease along a chain), so a chase over stale lines
// ends at a true ancestor; only the hook needs the truth, and it is an atomicCAS whose return value -- the real
// parent -- is where a failed attempt continues.  The upper levels of the trees, shared by every chase of an SM,
// are L1 hits instead of L2 round trips.
__device__ __forceinline__ int ld_ca(const int *p) {
    int v;
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// root of x after the union pass (k_uf_flatten).  The loads go through L1 here too: the sticks just below the roots are
// on the path of thousands of chases, and as volatile loads they were served one at a time by their L2 slice.  The
// only writes of the pass are shortcuts to the root (still ancestors), so a stale line costs hops, not correctness.
__device__ __forceinline__ int uf_find(int *parent, int x, int s0) {
    if (x == s0) return x;
    int p = ld_ca(parent + x);
    while (p != x && p != s0) {
        x = p;
        p = ld_ca(parent + x);
    }
    return p;
}

__global__ void k_uf_init(int64_t S, int32_t *__restrict__ parent, int32_t *__restrict__ size,
                          uint8_t *__restrict__ hasj) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) {
        parent[s] = (int)s;
        size[s] = 0;
        hasj[s] = 0;
    }
}

// Union pass: Rem's algorithm with splicing (the scheme the ConnectIt study found fastest among the concurrent
// union-finds), UF_K junctions per thread interleaved.  For a junction (u, v) the two climbs advance in lock step --
// always the side whose parent has the larger index -- and stop as soon as both sides see the same parent: a junction
// inside an already connected neighbourhood ends after a hop or two instead of two full chases to the root.  On
// the way every visited stick is spliced onto the other side's (smaller) parent with an atomicMin, which compresses
// the paths.  A root is hooked with an atomicCAS under the other side's parent (smaller index: the root of a tree
// stays its minimum stick id == the canonical label).  Loads go through L1: a stale parent is still an ancestor, a
// stale "root" fails its CAS and continues from the parent the CAS returns.  The network's first stick is the smallest
// id of the network, hence always a root: it is recognised without loading its (hot) line.
#ifndef CNT_UF_K
#define CNT_UF_K 4
#endif
constexpr int UF_K = CNT_UF_K;
__global__ void __launch_bounds__(256) k_uf_union(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ joff,
                                                  int64_t J, const int32_t *__restrict__ stick1,
                                                  const int32_t *__restrict__ stick2, int32_t *parent,
                                                  uint8_t *__restrict__ hasj) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int rx[UF_K], ry[UF_K], px[UF_K], py[UF_K], base[UF_K];
    bool live[UF_K];
    auto load = [&](int x, int s0) { return x == s0 ? s0 : ld_ca(parent + x); };
#pragma unroll
    for (int k = 0; k < UF_K; k++) {
        const int64_t e = first + k * stride;
        live[k] = e < J;
        rx[k] = ry[k] = px[k] = py[k] = base[k] = 0;
        if (live[k]) {
            const int b = cnt::find_segment(joff, B, (int)e);
            const int s0 = soff[b];
            base[k] = s0;
            rx[k] = s0 + stick1[e];
            ry[k] = s0 + stick2[e];
            hasj[rx[k]] = 1;
            hasj[ry[k]] = 1;
        }
    }
#pragma unroll
    for (int k = 0; k < UF_K; k++) {
        if (live[k]) {
            px[k] = load(rx[k], base[k]);
            py[k] = load(ry[k], base[k]);
        }
    }
    for (;;) {
        bool any = false;
#pragma unroll
        for (int k = 0; k < UF_K; k++) {
            if (!live[k]) continue;
            if (px[k] == py[k]) {                              // both sides under one parent: same component
                live[k] = false;
                continue;
            }
            any = true;
            if (px[k] < py[k]) {                               // x: the side whose parent has the larger index
                int t = rx[k]; rx[k] = ry[k]; ry[k] = t;
                t = px[k]; px[k] = py[k]; py[k] = t;
            }
            if (rx[k] == px[k]) {                              // x is a root (as far as this thread knows): hook it
                const int old = atomicCAS(parent + rx[k], rx[k], py[k]);
                if (old == rx[k]) live[k] = false;
                else px[k] = old;                              // it was not (any more): its true parent
            } else {                                           // splice x onto y's parent, climb
                const int z = px[k];
                atomicMin(parent + rx[k], py[k]);
                rx[k] = z;
                px[k] = load(z, base[k]);
            }
        }
        if (!any) break;
    }
}

// One atomic per WARP on the size of the giant component (the root of almost every stick of a percolating network)
// serialised the pass on a few hundred addresses (ncu: 3.1 ms for 22M sticks at 106 GB/s of DRAM traffic): the sticks
// whose root is the first stick of the CTA's first network -- nearly all of them -- are counted in shared memory and
// reach the size table with ONE atomic per CTA.
constexpr int FLAT_T = 1024;
__global__ void __launch_bounds__(FLAT_T) k_uf_flatten(int B, const int32_t *__restrict__ soff, int64_t S, int32_t *parent,
                                                      int32_t *__restrict__ label, int32_t *__restrict__ size) {
    __shared__ int s_hot, s_cnt;
    const int64_t s = (int64_t)blockIdx.x * FLAT_T + threadIdx.x;
    if (threadIdx.x == 0) {
        s_hot = soff[cnt::find_segment(soff, B, (int)((int64_t)blockIdx.x * FLAT_T))];
        s_cnt = 0;
    }
    __syncthreads();
    const int hot = s_hot;
    int r = -1;
    if (s < S) {
        int b = cnt::find_segment(soff, B, (int)s);
        r = uf_find(parent, (int)s, soff[b]);
        parent[s] = r;                                  // shortcut for the chases that pass through s
        label[s] = r - soff[b];
    }
    // component sizes: one atomic per group of equal roots inside the warp
    const unsigned same = __match_any_sync(0xffffffffu, r);
    if (r >= 0 && (threadIdx.x & 31) == __ffs(same) - 1) {
        if (r == hot) atomicAdd(&s_cnt, __popc(same));
        else atomicAdd(size + r, __popc(same));
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(size + hot, s_cnt);
}

// per-network statistics, pass 1 over the sticks: components, multi-stick components, largest, first stick with a junction.
// All sticks of a network are neighbours in the table, so every CTA in flight works on the same two or three networks:
// with one set of global atomics per WARP the pass was serialised on those few 32-byte lines (ncu: 285 cycles of
// long-scoreboard stall per issued instruction, 2.1 ms for 22M sticks).  A CTA that lies inside one network -- all
// but B of them -- reduces in shared memory and issues ONE set.
constexpr int STATS_T = 1024;
__global__ void __launch_bounds__(STATS_T) k_stats_sticks(int B, const int32_t *__restrict__ soff, int64_t S,
                                                         const int32_t *__restrict__ size, const uint8_t *__restrict__ hasj,
                                                         int32_t *__restrict__ acc /* [B][8] */) {
    __shared__ int s_acc[4];
    __shared__ int s_b[2];
    const int64_t s0 = (int64_t)blockIdx.x * STATS_T, s = s0 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        s_acc[0] = s_acc[1] = s_acc[2] = 0;
        s_acc[3] = 0x7fffffff;
        const int64_t last = s0 + STATS_T - 1 < S - 1 ? s0 + STATS_T - 1 : S - 1;
        s_b[0] = cnt::find_segment(soff, B, (int)s0);
        s_b[1] = cnt::find_segment(soff, B, (int)last);
    }
    __syncthreads();
    const bool uniform = s_b[0] == s_b[1];                       // the CTA lies inside one network
    int b = -1, sz = 0, first = 0x7fffffff;
    if (s < S) {
        b = uniform ? s_b[0] : cnt::find_segment(soff, B, (int)s);
        sz = size[s];           // non-zero only at roots
        if (hasj[s]) first = (int)s - soff[b];
    }
    if (uniform) {
        const unsigned roots = __ballot_sync(0xffffffffu, sz > 0), multi = __ballot_sync(0xffffffffu, sz > 1);
        int mx = sz, fi = first;
        for (int o = 16; o > 0; o >>= 1) {
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            fi = min(fi, __shfl_xor_sync(0xffffffffu, fi, o));
        }
        if (lane == 0) {
            if (roots) atomicAdd(&s_acc[0], __popc(roots));
            if (multi) atomicAdd(&s_acc[1], __popc(multi));
            if (mx > 0) atomicMax(&s_acc[2], mx);
            if (fi != 0x7fffffff) atomicMin(&s_acc[3], fi);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int32_t *a = acc + (int64_t)s_b[0] * 8;
            if (s_acc[0]) atomicAdd(a + 0, s_acc[0]);
            if (s_acc[1]) atomicAdd(a + 1, s_acc[1]);
            if (s_acc[2] > 0) atomicMax(a + 2, s_acc[2]);
            if (s_acc[3] != 0x7fffffff) atomicMin(a + 3, s_acc[3]);
        }
        return;
    }
    if (b < 0) return;                                           // a CTA across a network boundary (or the tail): per lane
    int32_t *a = acc + (int64_t)b * 8;
    if (sz > 0) {
        atomicAdd(a + 0, 1);
        if (sz > 1) atomicAdd(a + 1, 1);
        atomicMax(a + 2, sz);
    }
    if (first != 0x7fffffff) atomicMin(a + 3, first);
}
// pass 2: the reference's quirky cluster count needs K = multi-stick components (pass 1); junctions of the
// percolating component
__global__ void k_stats_second(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ joff, int64_t S,
                               int64_t J, const int32_t *__restrict__ stick1, const int32_t *__restrict__ label,
                               const uint8_t *__restrict__ hasj, const int32_t *__restrict__ orig,
                               int32_t *__restrict__ acc) {
    __shared__ int s_hot, s_cnt;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {           // the key most threads of the CTA will add to
        const int64_t t0 = (int64_t)blockIdx.x * blockDim.x;
        s_hot = t0 < S ? cnt::find_segment(soff, B, (int)t0) * 8 + 4
                       : (t0 - S < J ? cnt::find_segment(joff, B, (int)(t0 - S)) * 8 + 5 : -2);
        s_cnt = 0;
    }
    __syncthreads();
    int b = -1, slot = 4;
    if (t < S) {
        // reference quirk: sticks with a junction carry the component ordinal 0..K-1 (K = components
        // with >1 stick), isolated sticks keep their pre-sort index
        if (orig && !hasj[t]) {
            const int bb = cnt::find_segment(soff, B, (int)t);
            if (orig[t] >= acc[(int64_t)bb * 8 + 1]) b = bb;
        }
    } else if (t - S < J) {
        const int64_t e = t - S;
        const int bb = cnt::find_segment(joff, B, (int)e);
        const int s0 = soff[bb], n = soff[bb + 1] - s0;
        const bool perc = n >= 2 && label[s0] == 0 && label[s0 + 1] == 0;
        if (perc && label[s0 + stick1[e]] == 0) b = bb;
        slot = 5;
    }
    // the junctions of a CTA belong to one or two networks: their counts meet in shared memory first (one global atomic per
    // CTA and network instead of one per warp on the same few addresses)
    const int key = b >= 0 ? b * 8 + slot : -1;
    const unsigned same = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && lane == __ffs(same) - 1) {
        if (key == s_hot) atomicAdd(&s_cnt, __popc(same));
        else atomicAdd(acc + key, __popc(same));
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(acc + s_hot, s_cnt);
}
__global__ void k_stats_finish(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ label,
                               const int32_t *__restrict__ size, const int32_t *__restrict__ orig,
                               const int32_t *__restrict__ acc, int32_t *__restrict__ stats) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int s0 = soff[b], n = soff[b + 1] - s0;
    const int32_t *a = acc + (int64_t)b * 8;
    const bool perc = n >= 2 && label[s0] == 0 && label[s0 + 1] == 0;
    int32_t *o = stats + (int64_t)b * CNT_NSTAT;
    o[CNT_STAT_PERCOLATING] = perc ? 1 : 0;
    o[CNT_STAT_NCLUST_TRUE] = a[0];
    o[CNT_STAT_MAXCLUST_TRUE] = a[2];
    o[CNT_STAT_NCLUST_REF] = orig ? a[1] + a[4] : -1;
    o[CNT_STAT_MAXCLUST_REF] = a[3] == 0x7fffffff ? 0 : size[s0 + label[s0 + a[3]]];
    o[CNT_STAT_NCOMP] = perc ? size[s0] : 0;
    o[CNT_STAT_ECOMP] = a[5];
    o[7] = 0;
}
__global__ void k_stats_init(int B, int32_t *__restrict__ acc) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B * 8) acc[t] = (t & 7) == 3 ? 0x7fffffff : 0;
}
}  // namespace

extern "C" int64_t cnt_cluster_workspace_bytes(int64_t S) {
    // parent, size, hasj + the per-network accumulators of the statistics (every network has its two electrodes: B <= S / 2)
    return cnt::ws_bytes_for(S, 4) * 2 + cnt::ws_bytes_for(S, 1) + cnt::ws_bytes_for(4 * S + 16, 4);
}

extern "C" int cnt_label_clusters(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                                  const int32_t *stick1, const int32_t *stick2, const int32_t *orig, int32_t *label,
                                  int32_t *stats, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && joff && label && stats && ws && S >= 0 && J >= 0);
    CNT_CHECK_ARG(J == 0 || (stick1 && stick2));
    cnt::Workspace w(ws, ws_bytes);
    int32_t *parent = w.take<int32_t>(S);
    int32_t *size = w.take<int32_t>(S);
    uint8_t *hasj = w.take<uint8_t>(S);
    int32_t *acc = w.take<int32_t>(8 * (int64_t)B);
    if (!hasj || !acc) {
        cnt::set_error("cluster workspace too small");
        return CNT_ECAPACITY;
    }
    if (S > 0) {
        k_uf_init<<<cnt::grid_for(S, 256), 256, 0, st>>>(S, parent, size, hasj);
        CNT_LAUNCHED();
    }
    if (J > 0) {
        k_uf_union<<<cnt::grid_for((J + UF_K - 1) / UF_K, 256), 256, 0, st>>>(B, soff, joff, J, stick1, stick2, parent, hasj);
        CNT_LAUNCHED();
    }
    if (S > 0) {
        k_uf_flatten<<<cnt::grid_for(S, FLAT_T), FLAT_T, 0, st>>>(B, soff, S, parent, label, size);
        CNT_LAUNCHED();
    }
    k_stats_init<<<cnt::grid_for(8 * (int64_t)B, 256), 256, 0, st>>>(B, acc);
    CNT_LAUNCHED();
    if (S > 0) {
        k_stats_sticks<<<cnt::grid_for(S, STATS_T), STATS_T, 0, st>>>(B, soff, S, size, hasj, acc);
        CNT_LAUNCHED();
        k_stats_second<<<cnt::grid_for(S + J, 256), 256, 0, st>>>(B, soff, joff, S, J, stick1, label, hasj, orig, acc);
        CNT_LAUNCHED();
    }
    k_stats_finish<<<cnt::grid_for(B, 128), 128, 0, st>>>(B, soff, label, size, orig, acc, stats);
    CNT_LAUNCHED();
    return CNT_OK;
}
