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

// one block per network: |I| per junction and the three source-current estimators,
// reduced in a fixed order (deterministic)
__global__ void __launch_bounds__(1024) k_currents(const int32_t *__restrict__ soff, const int32_t *__restrict__ joff,
                                                   const int32_t *__restrict__ stick1,
                                                   const int32_t *__restrict__ stick2,
                                                   const int32_t *__restrict__ stats, const double *__restrict__ g,
                                                   const double *__restrict__ V, double *__restrict__ I,
                                                   double *__restrict__ isrc) {
    __shared__ double sm[32];
    int b = blockIdx.x;
    int s0 = soff[b];
    int j0 = joff[b], j1 = joff[b + 1];
    bool perc = stats[(int64_t)b * CNT_NSTAT + CNT_STAT_PERCOLATING] != 0;
    double is = 0, id = 0, pw = 0;
    for (int e = j0 + threadIdx.x; e < j1; e += blockDim.x) {
        int a = stick1[e], c = stick2[e];
        double va = V[s0 + a], vc = V[s0 + c];
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
        isrc[b * 3 + 0] = is;
        isrc[b * 3 + 1] = id;
        isrc[b * 3 + 2] = pw / VDS;
    }
}
}  // namespace

extern "C" int cnt_currents(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                            const int32_t *stick1, const int32_t *stick2, const int32_t *label, const int32_t *stats,
                            const int32_t *urow, const int32_t *attach, const int32_t *elim_vid, const int32_t *va,
                            const int32_t *vc, const int32_t *ve1, const int32_t *ve2, const double *g, const double *x,
                            double *V, double *I, double *isrc, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && joff && label && stats && urow && V && isrc && S >= 0 && J >= 0);
    CNT_CHECK_ARG(J == 0 || (stick1 && stick2 && g && I));
    if (S > 0) {
        k_voltages<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, label, stats, urow, attach, elim_vid, va, vc, ve1, ve2, g, x, V);
        CNT_LAUNCHED();
    }
    k_currents<<<B, 1024, 0, st>>>(soff, joff, stick1, stick2, stats, g, V, I, isrc);
    CNT_LAUNCHED();
    return CNT_OK;
}
