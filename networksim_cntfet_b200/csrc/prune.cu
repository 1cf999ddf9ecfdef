// Exact reduction of the MNA system: dangling sticks carry no current.
//
// A stick of the percolating component that is left with a single junction (after its own
// dangling neighbours are gone) cannot carry current: it sits at the potential of the stick it
// hangs on and its junction can be dropped from the Laplacian without changing any other
// voltage.  Peeling such sticks repeatedly leaves the 2-core of the component plus the two
// electrodes (never peeled).  At 100x100 um / density 10 this removes ~20 % of the unknowns (and
// ~5 % of the PCG iterations).  The reference solves the full component (cnet.py:104-131); the
// results are identical: pruned sticks get the voltage of their attachment point at write-back
// (cnt_currents), their junction currents are exactly 0 (the reference reports rounding noise
// ~1e-20 there).
//
// Edge-parallel peeling rounds: an edge whose one end has degree 1 removes that end, records the
// other end as its parent and decrements the parent's degree.  The surviving set is the unique
// 2-core, and every pruned stick has exactly one live neighbour when it is removed (dangling
// parts are trees), so parents -- hence results -- do not depend on thread scheduling.
#include "common.cuh"

namespace {
__global__ void k_prune_init(int B, const int32_t *__restrict__ soff, int64_t S, const int32_t *__restrict__ label,
                             const int32_t *__restrict__ stats, uint8_t *__restrict__ alive, int32_t *__restrict__ deg,
                             int32_t *__restrict__ parent) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    alive[s] = (stats[(int64_t)b * CNT_NSTAT + CNT_STAT_PERCOLATING] && label[s] == 0) ? 1 : 0;
    deg[s] = 0;
    parent[s] = -1;
}

__global__ void k_prune_degree(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ joff, int64_t J,
                               const int32_t *__restrict__ stick1, const int32_t *__restrict__ stick2,
                               const uint8_t *__restrict__ alive, int32_t *__restrict__ deg) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= J) return;
    int b = cnt::find_segment(joff, B, (int)e);
    int a = soff[b] + stick1[e], c = soff[b] + stick2[e];
    if (alive[a] && alive[c]) {
        atomicAdd(deg + a, 1);
        atomicAdd(deg + c, 1);
    }
}

// one peeling round; *changed is set if anything was removed
__global__ void k_prune_round(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ joff, int64_t J,
                              const int32_t *__restrict__ stick1, const int32_t *__restrict__ stick2, uint8_t *alive,
                              int32_t *deg, int32_t *__restrict__ parent, int32_t *__restrict__ changed) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= J) return;
    int b = cnt::find_segment(joff, B, (int)e);
    int s0 = soff[b];
    int la = stick1[e], lc = stick2[e];
    int a = s0 + la, c = s0 + lc;
    volatile uint8_t *va = alive;
    if (!va[a] || !va[c]) return;
    volatile int32_t *vd = deg;
    // electrodes (local index 0 and 1) are never peeled
    if (la >= 2 && vd[a] == 1) {
        va[a] = 0;
        parent[a] = c;
        atomicSub(deg + c, 1);
        *changed = 1;
    } else if (lc >= 2 && vd[c] == 1) {
        va[c] = 0;
        parent[c] = a;
        atomicSub(deg + a, 1);
        *changed = 1;
    }
}

// attachment point of every pruned stick: follow the parent chain to the first live stick
__global__ void k_prune_attach(int64_t S, const uint8_t *__restrict__ alive, const int32_t *__restrict__ parent,
                               int32_t *__restrict__ attach) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int p = parent[s];
    if (p < 0) {
        attach[s] = -1;
        return;
    }
    while (!alive[p]) p = parent[p];
    attach[s] = p;
}

// ---------------------------------------------------------------------------------------
// Series elimination (second exact reduction).  A stick of the 2-core with exactly two junctions,
// both to ordinary sticks (not electrodes), is a series connection: eliminating it leaves one
// junction between its neighbours a, c with conductance g1 g2 / (g1 + g2), and afterwards
// V = (g1 V_a + g2 V_c) / (g1 + g2).  Only an independent set is eliminated (a candidate goes if
// none of its neighbours is a candidate with a smaller index), so the two neighbours of an
// eliminated stick always stay unknowns and no chains have to be walked: one virtual junction per
// eliminated stick.  Topology only -- the same sticks go in every gate configuration; the
// conductances are combined per configuration in cnt_assemble_values.
// ---------------------------------------------------------------------------------------
__global__ void k_series_adj(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ joff, int64_t J,
                             const int32_t *__restrict__ stick1, const int32_t *__restrict__ stick2,
                             const uint8_t *__restrict__ alive, int32_t *__restrict__ cnt, int32_t *__restrict__ nbr,
                             int32_t *__restrict__ ned) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= J) return;
    int b = cnt::find_segment(joff, B, (int)e);
    int a = soff[b] + stick1[e], c = soff[b] + stick2[e];
    if (!alive[a] || !alive[c]) return;
    int sa = atomicAdd(cnt + a, 1), sc = atomicAdd(cnt + c, 1);
    if (sa < 2) { nbr[2 * (int64_t)a + sa] = c; ned[2 * (int64_t)a + sa] = (int32_t)e; }
    if (sc < 2) { nbr[2 * (int64_t)c + sc] = a; ned[2 * (int64_t)c + sc] = (int32_t)e; }
}
__device__ __forceinline__ bool series_candidate(int B, const int32_t *__restrict__ soff, const uint8_t *__restrict__ alive,
                                                 const int32_t *__restrict__ cnt, const int32_t *__restrict__ nbr,
                                                 int64_t s) {
    if (!alive[s] || cnt[s] != 2) return false;
    int b = cnt::find_segment(soff, B, (int)s);
    int s0 = soff[b];
    return (int)s - s0 >= 2 && nbr[2 * s] - s0 >= 2 && nbr[2 * s + 1] - s0 >= 2;     // no electrode involved
}
__global__ void k_series_cand(int B, const int32_t *__restrict__ soff, int64_t S, const uint8_t *__restrict__ alive,
                              const int32_t *__restrict__ cnt, const int32_t *__restrict__ nbr,
                              uint8_t *__restrict__ cand) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) cand[s] = series_candidate(B, soff, alive, cnt, nbr, s) ? 1 : 0;
}
__global__ void k_series_pick(int64_t S, const uint8_t *__restrict__ cand, const int32_t *__restrict__ nbr,
                              int32_t *__restrict__ flag) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int f = 0;
    if (cand[s]) {
        int n0 = nbr[2 * s], n1 = nbr[2 * s + 1];
        f = !(cand[n0] && n0 < s) && !(cand[n1] && n1 < s);
    }
    flag[s] = f;
}
__global__ void k_series_emit(int B, const int32_t *__restrict__ soff, int64_t S, const int32_t *__restrict__ scan,
                              const int32_t *__restrict__ nbr, const int32_t *__restrict__ ned,
                              int32_t *__restrict__ elim_vid, int32_t *__restrict__ voff, int32_t *__restrict__ va,
                              int32_t *__restrict__ vc, int32_t *__restrict__ ve1, int32_t *__restrict__ ve2) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s <= B) voff[s] = scan[soff[s]];
    if (s >= S) return;
    int v = scan[s];
    if (scan[s + 1] == v) {
        elim_vid[s] = -1;
        return;
    }
    int s0 = soff[cnt::find_segment(soff, B, (int)s)];
    int n0 = nbr[2 * s], n1 = nbr[2 * s + 1], e0 = ned[2 * s], e1 = ned[2 * s + 1];
    if (n0 > n1) {     // the two slots were filled in atomic order: normalise
        int t = n0; n0 = n1; n1 = t;
        t = e0; e0 = e1; e1 = t;
    }
    elim_vid[s] = v;
    va[v] = n0 - s0;
    vc[v] = n1 - s0;
    ve1[v] = e0;
    ve2[v] = e1;
}
}  // namespace

extern "C" int64_t cnt_series_workspace_bytes(int64_t S) {
    return cnt::ws_bytes_for(S, 4) + 2 * cnt::ws_bytes_for(2 * S, 4) + cnt::ws_bytes_for(S, 1) +
           2 * cnt::ws_bytes_for(S + 1, 4) + cnt_scan_workspace_bytes(S + 1);
}

// elim_vid[S]: index of the virtual junction that replaces an eliminated stick, -1 otherwise;
// voff[B+1]: virtual junctions per network (offsets); va, vc [<= S]: the two neighbours (network-local
// stick indices, va < vc); ve1, ve2: the eliminated stick's junctions to va / vc (rows of the junction
// table).  *h_nv: number of virtual junctions.  Synchronises the stream (one 4-byte read).
extern "C" int cnt_series_select(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                                 const int32_t *stick1, const int32_t *stick2, const uint8_t *alive,
                                 int32_t *elim_vid, int32_t *voff, int32_t *va, int32_t *vc, int32_t *ve1,
                                 int32_t *ve2, int64_t *h_nv, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && joff && alive && elim_vid && voff && va && vc && ve1 && ve2 && h_nv && ws && S >= B &&
                  J >= 0);
    *h_nv = 0;
    cnt::Workspace w(ws, ws_bytes);
    int32_t *cnt_ = w.take<int32_t>(S);
    int32_t *nbr = w.take<int32_t>(2 * S), *ned = w.take<int32_t>(2 * S);
    uint8_t *cand = w.take<uint8_t>(S);
    int32_t *flag = w.take<int32_t>(S + 1), *scan = w.take<int32_t>(S + 1);
    int64_t sb = cnt_scan_workspace_bytes(S + 1);
    void *sws = w.take<char>(sb);
    if (!sws) {
        cnt::set_error("series workspace too small");
        return CNT_ECAPACITY;
    }
    CNT_CUDA(cudaMemsetAsync(cnt_, 0, 4 * S, st));
    CNT_CUDA(cudaMemsetAsync(nbr, 0, 8 * S, st));
    if (J > 0) {
        k_series_adj<<<cnt::grid_for(J, 256), 256, 0, st>>>(B, soff, joff, J, stick1, stick2, alive, cnt_, nbr, ned);
        CNT_LAUNCHED();
    }
    const unsigned gs = cnt::grid_for(S + 1, 256);
    k_series_cand<<<gs, 256, 0, st>>>(B, soff, S, alive, cnt_, nbr, cand);
    CNT_LAUNCHED();
    k_series_pick<<<gs, 256, 0, st>>>(S, cand, nbr, flag);
    CNT_LAUNCHED();
    int rc = cnt_exclusive_scan_i32(flag, scan, S, 1, sws, sb, stream);
    if (rc) return rc;
    k_series_emit<<<gs, 256, 0, st>>>(B, soff, S, scan, nbr, ned, elim_vid, voff, va, vc, ve1, ve2);
    CNT_LAUNCHED();
    int32_t h = 0;
    CNT_CUDA(cudaMemcpyAsync(&h, scan + S, 4, cudaMemcpyDeviceToHost, st));
    CNT_CUDA(cudaStreamSynchronize(st));
    *h_nv = h;
    return CNT_OK;
}

extern "C" int64_t cnt_prune_workspace_bytes(int64_t S) { return 2 * cnt::ws_bytes_for(S, 4) + cnt::ws_bytes_for(1, 4); }

// alive[S] (u8): 1 for the sticks of the percolating component that stay in the system (2-core +
// electrodes), 0 otherwise; attach[S]: global stick id a pruned stick hangs on (-1: not pruned).
// Synchronises the stream (a few small D2H reads of the "changed" flag).
extern "C" int cnt_prune_leaves(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                                const int32_t *stick1, const int32_t *stick2, const int32_t *label, const int32_t *stats,
                                uint8_t *alive, int32_t *attach, int32_t *h_rounds, void *ws, int64_t ws_bytes,
                                void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && joff && label && stats && alive && attach && ws && S >= 0 && J >= 0);
    if (h_rounds) *h_rounds = 0;
    if (S == 0) return CNT_OK;
    cnt::Workspace w(ws, ws_bytes);
    int32_t *deg = w.take<int32_t>(S);
    int32_t *parent = w.take<int32_t>(S);
    int32_t *changed = w.take<int32_t>(1);
    if (!changed) {
        cnt::set_error("prune workspace too small");
        return CNT_ECAPACITY;
    }
    k_prune_init<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, label, stats, alive, deg, parent);
    CNT_LAUNCHED();
    int rounds = 0;
    if (J > 0) {
        k_prune_degree<<<cnt::grid_for(J, 256), 256, 0, st>>>(B, soff, joff, J, stick1, stick2, alive, deg);
        CNT_LAUNCHED();
        const int ROUNDS_PER_POLL = 8;
        while (true) {
            CNT_CUDA(cudaMemsetAsync(changed, 0, 4, st));
            for (int k = 0; k < ROUNDS_PER_POLL; k++) {
                k_prune_round<<<cnt::grid_for(J, 256), 256, 0, st>>>(B, soff, joff, J, stick1, stick2, alive, deg, parent,
                                                                    changed);
                CNT_LAUNCHED();
            }
            rounds += ROUNDS_PER_POLL;
            int32_t h = 0;
            CNT_CUDA(cudaMemcpyAsync(&h, changed, 4, cudaMemcpyDeviceToHost, st));
            CNT_CUDA(cudaStreamSynchronize(st));
            if (!h) break;
            if (rounds > 1000000) {
                cnt::set_error("leaf pruning did not terminate");
                return CNT_ECUDA;
            }
        }
    }
    k_prune_attach<<<cnt::grid_for(S, 256), 256, 0, st>>>(S, alive, parent, attach);
    CNT_LAUNCHED();
    if (h_rounds) *h_rounds = rounds;
    return CNT_OK;
}
