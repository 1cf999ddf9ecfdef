// Write-back of node voltages and junction currents.
// Replaces update_voltages / update_currents (cnet.py:132-146).
#include <math.h>
#include "common.cuh"

namespace {
constexpr double VDS = 0.1;   // netsim.py:182

__global__ void k_voltages(int B, const int32_t *__restrict__ soff, int64_t S, const int32_t *__restrict__ label,
                           const int32_t *__restrict__ stats, const int32_t *__restrict__ urow,
                           const int32_t *__restrict__ attach, const double *__restrict__ x,
                           double *__restrict__ V) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    int i = (int)s - soff[b];
    double v = nan("");
    if (stats[(int64_t)b * CNT_NSTAT + CNT_STAT_PERCOLATING] && label[s] == 0) {
        int t = (int)s;
        if (attach && attach[s] >= 0) {   // pruned dangling stick: potential of the stick it hangs on
            t = attach[s];
            i = t - soff[b];
        }
        if (i == 0) v = VDS;
        else if (i == 1) v = 0.0;
        else v = x[urow[t]];
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
                            const int32_t *urow, const int32_t *attach, const double *g, const double *x, double *V,
                            double *I, double *isrc, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && joff && label && stats && urow && V && isrc && S >= 0 && J >= 0);
    CNT_CHECK_ARG(J == 0 || (stick1 && stick2 && g && I));
    if (S > 0) {
        k_voltages<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, label, stats, urow, attach, x, V);
        CNT_LAUNCHED();
    }
    k_currents<<<B, 1024, 0, st>>>(soff, joff, stick1, stick2, stats, g, V, I, isrc);
    CNT_LAUNCHED();
    return CNT_OK;
}
