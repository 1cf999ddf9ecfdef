// Connected components of the junction graph by lock-free union-find with minimum-index
// roots.  Replaces the networkx component search of netsim.py:175-179,197-206 and the
// cluster statistics of measure_perc.py:76-80 / 148-152.
//
// parent[] holds GLOBAL stick ids; a root is only ever hooked under a SMALLER id, so the
// final root of a component is its minimum stick id == the canonical label.
#include "common.cuh"

namespace {

__device__ __forceinline__ int uf_find(int *parent, int x) {
    // path halving; plain loads/stores are safe: parents only ever decrease towards the root
    int p = ((volatile int *)parent)[x];
    while (p != x) {
        int gp = ((volatile int *)parent)[p];
        if (gp != p) ((volatile int *)parent)[x] = gp;
        x = p;
        p = gp;
    }
    return x;
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

__global__ void k_uf_union(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ joff, int64_t J,
                           const int32_t *__restrict__ stick1, const int32_t *__restrict__ stick2,
                           int32_t *parent, uint8_t *__restrict__ hasj) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= J) return;
    int b = cnt::find_segment(joff, B, (int)e);
    int s0 = soff[b];
    int a = s0 + stick1[e], c = s0 + stick2[e];
    hasj[a] = 1;
    hasj[c] = 1;
    while (true) {
        a = uf_find(parent, a);
        c = uf_find(parent, c);
        if (a == c) break;
        int hi = a > c ? a : c, lo = a > c ? c : a;
        int old = atomicCAS(parent + hi, hi, lo);
        if (old == hi) break;
        a = hi;  // someone else re-parented hi: retry from the current roots
        c = lo;
    }
}

__global__ void k_uf_flatten(int B, const int32_t *__restrict__ soff, int64_t S, int32_t *parent,
                             int32_t *__restrict__ label, int32_t *__restrict__ size) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    int r = uf_find(parent, (int)s);
    label[s] = r - soff[b];
    atomicAdd(size + r, 1);
}

// one block per network
__global__ void __launch_bounds__(256) k_cluster_stats(const int32_t *__restrict__ soff,
                                                       const int32_t *__restrict__ joff,
                                                       const int32_t *__restrict__ stick1,
                                                       const int32_t *__restrict__ label,
                                                       const int32_t *__restrict__ size,
                                                       const uint8_t *__restrict__ hasj,
                                                       const int32_t *__restrict__ orig, int32_t *__restrict__ stats) {
    __shared__ int s_nroots, s_nmulti, s_max, s_first, s_iso, s_ecomp;
    int b = blockIdx.x;
    int s0 = soff[b], n = soff[b + 1] - s0;
    if (threadIdx.x == 0) {
        s_nroots = 0; s_nmulti = 0; s_max = 0; s_first = 0x7fffffff; s_iso = 0; s_ecomp = 0;
    }
    __syncthreads();
    int nroots = 0, nmulti = 0, mx = 0, first = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int sz = size[s0 + i];   // non-zero only at roots
        if (sz > 0) {
            nroots++;
            if (sz > 1) nmulti++;
            mx = max(mx, sz);
        }
        if (hasj[s0 + i]) first = min(first, i);
    }
    atomicAdd(&s_nroots, nroots);
    atomicAdd(&s_nmulti, nmulti);
    atomicMax(&s_max, mx);
    atomicMin(&s_first, first);
    __syncthreads();
    // reference quirk: sticks with a junction carry the component ordinal
    // 0..K-1 (K = components with >1 stick), isolated sticks keep their pre-sort index
    int K = s_nmulti, iso = 0;
    if (orig) {
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (!hasj[s0 + i] && orig[s0 + i] >= K) iso++;
        atomicAdd(&s_iso, iso);
    }
    bool perc = n >= 2 && label[s0] == 0 && label[s0 + 1] == 0;
    int ec = 0;
    if (perc) {
        int j0 = joff[b], j1 = joff[b + 1];
        for (int e = j0 + threadIdx.x; e < j1; e += blockDim.x)
            if (label[s0 + stick1[e]] == 0) ec++;
        atomicAdd(&s_ecomp, ec);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t *o = stats + (int64_t)b * CNT_NSTAT;
        o[CNT_STAT_PERCOLATING] = perc ? 1 : 0;
        o[CNT_STAT_NCLUST_TRUE] = s_nroots;
        o[CNT_STAT_MAXCLUST_TRUE] = s_max;
        o[CNT_STAT_NCLUST_REF] = orig ? K + s_iso : -1;
        o[CNT_STAT_MAXCLUST_REF] = s_first == 0x7fffffff ? 0 : size[s0 + label[s0 + s_first]];
        o[CNT_STAT_NCOMP] = perc ? size[s0] : 0;
        o[CNT_STAT_ECOMP] = s_ecomp;
        o[7] = 0;
    }
}
}  // namespace

extern "C" int64_t cnt_cluster_workspace_bytes(int64_t S) {
    return cnt::ws_bytes_for(S, 4) * 2 + cnt::ws_bytes_for(S, 1);
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
    if (!hasj) {
        cnt::set_error("cluster workspace too small");
        return CNT_ECAPACITY;
    }
    if (S > 0) {
        k_uf_init<<<cnt::grid_for(S, 256), 256, 0, st>>>(S, parent, size, hasj);
        CNT_LAUNCHED();
    }
    if (J > 0) {
        k_uf_union<<<cnt::grid_for(J, 256), 256, 0, st>>>(B, soff, joff, J, stick1, stick2, parent, hasj);
        CNT_LAUNCHED();
    }
    if (S > 0) {
        k_uf_flatten<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, parent, label, size);
        CNT_LAUNCHED();
    }
    k_cluster_stats<<<B, 256, 0, st>>>(soff, joff, stick1, label, size, hasj, orig, stats);
    CNT_LAUNCHED();
    return CNT_OK;
}
