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
}  // namespace

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
