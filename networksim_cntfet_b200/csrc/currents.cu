// Write-back of node voltages and junction currents.
// Replaces update_voltages / update_currents (cnet.py:132-146).
#include <math.h>
#include "common.cuh"

namespace {
constexpr double VDS = 0.1;   // netsim.py:182

// voltage of one stick of the percolating component: electrode, unknown, or a stick removed by the
// series elimination (weighted mean of its two neighbours, which are unknowns by construction)
__device__ __forceinline__ double stick_voltage(int t, int s0, const int32_t *__restrict__ urow,
                                                const int32_t *__restrict__ elim_vid, const int32_t *__restrict__ va,
                                                const int32_t *__restrict__ vc, const int32_t *__restrict__ ve1,
                                                const int32_t *__restrict__ ve2, const double *__restrict__ g,
                                                const double *__restrict__ x) {
    int i = t - s0;
    if (i == 0) return VDS;
    if (i == 1) return 0.0;
    if (elim_vid) {
        int v = elim_vid[t];
        if (v >= 0) {
            double g1 = g[ve1[v]], g2 = g[ve2[v]];
            return (g1 * x[urow[s0 + va[v]]] + g2 * x[urow[s0 + vc[v]]]) / (g1 + g2);
        }
    }
    return x[urow[t]];
}

__global__ void k_voltages(int B, const int32_t *__restrict__ soff, int64_t S, const int32_t *__restrict__ label,
                           const int32_t *__restrict__ stats, const int32_t *__restrict__ urow,
                           const int32_t *__restrict__ attach, const int32_t *__restrict__ elim_vid,
                           const int32_t *__restrict__ va, const int32_t *__restrict__ vc,
                           const int32_t *__restrict__ ve1, const int32_t *__restrict__ ve2,
                           const double *__restrict__ g, const double *__restrict__ x, double *__restrict__ V) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    double v = nan("");
    if (stats[(int64_t)b * CNT_NSTAT + CNT_STAT_PERCOLATING] && label[s] == 0) {
        int t = (int)s;
        if (attach && attach[s] >= 0) t = attach[s];   // pruned dangling stick: potential of the stick it hangs on
        v = stick_voltage(t, soff[b], urow, elim_vid, va, vc, ve1, ve2, g, x);
    }
    V[s] = v;
}

// |I| per junction and the three source-current estimators.  A network's junctions are cut into chunks of
// CHUNK_J (counted from the network's own first junction, so the partition -- and with it every bit of the
// sums -- does not depend on what else is in the batch); one CTA per chunk writes its partial sums, one warp
// per network then adds the partials in chunk order.  (One CTA per network, as before, leaves 92 of the
// 148 SMs idle for a batch of 56 networks and takes 75 ms for ONE network of 15M junctions.)
constexpr int CHUNK_J = 8192;
constexpr int CUR_T = 256;

// coff[b] = first chunk of network b (B <= 65535: one CTA)
__global__ void __launch_bounds__(1024) k_cur_chunks(int B, const int32_t *__restrict__ joff, int32_t *__restrict__ coff) {
    __shared__ int s_n[1024];
    const int t = threadIdx.x;
    const int per = (B + 1023) / 1024;
    const int lo = min(t * per, B), hi = min(lo + per, B);
    auto chunks = [&](int b) { return (joff[b + 1] - joff[b] + CHUNK_J - 1) / CHUNK_J; };
    int sn = 0;
    for (int b = lo; b < hi; b++) sn += chunks(b);
    s_n[t] = sn;
    __syncthreads();
    if (t == 0) {
        int acc = 0;
        for (int k = 0; k < 1024; k++) {
            const int c = s_n[k];
            s_n[k] = acc;
            acc += c;
        }
        coff[B] = acc;
    }
    __syncthreads();
    sn = s_n[t];
    for (int b = lo; b < hi; b++) {
        coff[b] = sn;
        sn += chunks(b);
    }
}

__global__ void __launch_bounds__(CUR_T) k_currents(int B, const int32_t *__restrict__ soff, const int32_t *__restrict__ joff,
                                                    const int32_t *__restrict__ coff, const int32_t *__restrict__ stick1,
                                                    const int32_t *__restrict__ stick2,
                                                    const int32_t *__restrict__ stats, const double *__restrict__ g,
                                                    const double *__restrict__ V, double *__restrict__ I,
                                                    double *__restrict__ part /* [chunks][3] */) {
    __shared__ double sm[32];
    const int c = blockIdx.x;
    if (c >= coff[B]) return;
    const int b = cnt::find_segment(coff, B, c);
    const int s0 = soff[b];
    const int j0 = joff[b] + (c - coff[b]) * CHUNK_J, j1 = min(j0 + CHUNK_J, joff[b + 1]);
    const bool perc = stats[(int64_t)b * CNT_NSTAT + CNT_STAT_PERCOLATING] != 0;
    double is = 0, id = 0, pw = 0;
    for (int e = j0 + threadIdx.x; e < j1; e += CUR_T) {
        int a = stick1[e], cc = stick2[e];
        double va = V[s0 + a], vc = V[s0 + cc];
        double cur = nan("");
        if (perc && va == va) {   // inside the percolating component
            double ge = g[e];
            double dv = va - vc;
            cur = fabs(ge * dv);                       // cnet.py:146
            pw += ge * dv * dv;
            if (a == 0) is += ge * (VDS - vc);
            else if (a == 1) id += ge * vc;
        }
        I[e] = cur;
    }
    is = cnt::block_sum(is, sm);
    id = cnt::block_sum(id, sm);
    pw = cnt::block_sum(pw, sm);
    if (threadIdx.x == 0) {
        part[(int64_t)c * 3 + 0] = is;
        part[(int64_t)c * 3 + 1] = id;
        part[(int64_t)c * 3 + 2] = pw;
    }
}

// one warp per network: partials in chunk order (lane-strided, then the fixed shuffle tree)
__global__ void __launch_bounds__(256) k_cur_finish(int B, const int32_t *__restrict__ coff, const double *__restrict__ part,
                                                    double *__restrict__ isrc) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    double is = 0, id = 0, pw = 0;
    for (int c = coff[b] + lane; c < coff[b + 1]; c += 32) {
        is += part[(int64_t)c * 3 + 0];
        id += part[(int64_t)c * 3 + 1];
        pw += part[(int64_t)c * 3 + 2];
    }
    is = cnt::warp_sum(is);
    id = cnt::warp_sum(id);
    pw = cnt::warp_sum(pw);
    if (lane == 0) {
        isrc[b * 3 + 0] = is;
        isrc[b * 3 + 1] = id;
        isrc[b * 3 + 2] = pw / VDS;
    }
}
}  // namespace

static int64_t max_chunks(int B, int64_t J) { return J / CHUNK_J + B; }

extern "C" int64_t cnt_currents_workspace_bytes(int B, int64_t J) {
    if (B <= 0 || J < 0) return 0;
    return cnt::ws_bytes_for(B + 1, 4) + cnt::ws_bytes_for(3 * max_chunks(B, J), 8);
}

extern "C" int cnt_currents(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                            const int32_t *stick1, const int32_t *stick2, const int32_t *label, const int32_t *stats,
                            const int32_t *urow, const int32_t *attach, const int32_t *elim_vid, const int32_t *va,
                            const int32_t *vc, const int32_t *ve1, const int32_t *ve2, const double *g, const double *x,
                            double *V, double *I, double *isrc, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && B <= 65535 && soff && joff && label && stats && urow && V && isrc && S >= 0 && J >= 0 && ws);
    CNT_CHECK_ARG(J == 0 || (stick1 && stick2 && g && I));
    cnt::Workspace w(ws, ws_bytes);
    int32_t *coff = w.take<int32_t>(B + 1);
    double *part = w.take<double>(3 * max_chunks(B, J));
    if (!coff || !part) {
        cnt::set_error("currents workspace too small");
        return CNT_ECAPACITY;
    }
    if (S > 0) {
        k_voltages<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, label, stats, urow, attach, elim_vid, va, vc, ve1, ve2, g, x, V);
        CNT_LAUNCHED();
    }
    k_cur_chunks<<<1, 1024, 0, st>>>(B, joff, coff);
    CNT_LAUNCHED();
    k_currents<<<(unsigned)max_chunks(B, J), CUR_T, 0, st>>>(B, soff, joff, coff, stick1, stick2, stats, g, V, I, part);
    CNT_LAUNCHED();
    k_cur_finish<<<cnt::grid_for((int64_t)B * 32, 256), 256, 0, st>>>(B, coff, part, isrc);
    CNT_LAUNCHED();
    return CNT_OK;
}
