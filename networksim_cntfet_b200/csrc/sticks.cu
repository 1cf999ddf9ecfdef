// Stick placement from a counter-based RNG + the length sort of the reference.
// Replaces make_stick / get_ends / make_sticks (netsim.py:97-125) and the
// `sort_values('length', ascending=False)` of make_intersects_kdtree (netsim.py:132-134).
//
// RNG: Philox4x32-10, key = the network's 64-bit seed, counter = (stick ordinal, draw block).
// Any stick of any network can be regenerated independently: ensembles are resumable by seed
// and identical no matter how they are batched or sharded over GPUs (the reference hands one
// seed to every network too, measure_perc.py:226-233).  The stream is statistically, not
// bitwise, the reference's (numpy MT19937); parity tests always feed reference sticks.
//
// Sort: segmented stable LSD radix sort (radix.cu).
// Key: source electrode 0, drain electrode 1, body sticks ~bits(length) (longer first);
// stability keeps generation order among equal lengths.  Random lengths: 4 passes over the UPPER 32 bits of the key
// (8 bytes moved per stick and pass), then k_tie_fix orders the few runs of sticks that share those 32 bits (sign,
// exponent, 20 bits of mantissa: ~1e3 pairs per 100k sticks) by the lower 32 bits -- the same permutation as 8 passes
// over the 64-bit key (12 bytes per stick and pass), which constant lengths (one long run) still take.
#include <math.h>
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0,
                                             uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 53-bit uniform in [0,1)
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void k_generate(const unsigned long long *__restrict__ seeds, int B, const int32_t *__restrict__ soff, int64_t S,
                           const double *__restrict__ scaling, double pm, int len_mode, double len_const,
                           double *__restrict__ xc, double *__restrict__ yc, double *__restrict__ ang,
                           double *__restrict__ len, uint8_t *__restrict__ kind, unsigned long long *__restrict__ key,
                           uint32_t *__restrict__ key32, uint32_t *__restrict__ val) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    int i = (int)s - soff[b], n = soff[b + 1] - soff[b];
    val[s] = (uint32_t)s;
    if (i == 0 || i == n - 1) {   // netsim.py:121-124: source first, drain last
        xc[s] = i == 0 ? 0.01 : 0.99;
        yc[s] = 0.5;
        ang[s] = M_PI / 2 - 1e-6;
        len[s] = 100.0;
        kind[s] = 0;
        const unsigned long long ek = (i == 0 && n > 1) || n == 1 ? 0ull : 1ull;
        if (key32) key32[s] = (uint32_t)ek;          // below every body stick: the upper word of ~bits(l > 0) is >= 0x800fffff
        else key[s] = ek;
        return;
    }
    unsigned long long seed = seeds[b];
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t r0[4], r1[4], r2[4];
    philox4x32_10((uint32_t)i, 0u, 0u, 0u, k0, k1, r0);
    philox4x32_10((uint32_t)i, 1u, 0u, 0u, k0, k1, r1);
    philox4x32_10((uint32_t)i, 2u, 0u, 0u, k0, k1, r2);
    double u_pm = u01(r0[0], r0[1]), u_x = u01(r0[2], r0[3]);
    double u_y = u01(r1[0], r1[1]), u_a = u01(r1[2], r1[3]);
    double u_n1 = u01(r2[0], r2[1]), u_n2 = u01(r2[2], r2[3]);
    kind[s] = u_pm <= pm ? 1 : 2;                      // netsim.py:108-109
    xc[s] = u_x;
    yc[s] = u_y;
    ang[s] = u_a * 2 * M_PI;                           // netsim.py:113
    double l;
    if (len_mode == 0) {
        double z = sqrt(-2.0 * log(1.0 - u_n1)) * cos(2.0 * M_PI * u_n2);   // Box-Muller
        l = fabs(0.66 + 0.44 * z) / scaling[b];        // netsim.py:113 abs(normal(0.66,0.44))/scaling
    } else {
        l = len_const / scaling[b];                    // netsim.py:111
    }
    len[s] = l;
    const unsigned long long k = ~(unsigned long long)__double_as_longlong(l);
    if (key32) key32[s] = (uint32_t)(k >> 32);
    else key[s] = k;
}

// after the sort by the upper key word: runs of equal words ordered by the lower word (stable insertion sort by the
// thread at the head of the run; runs are pairs or triples).  Electrodes (keys 0 and 1) are runs of one.
__global__ void k_tie_fix(int B, const int32_t *__restrict__ soff, int64_t S, const uint32_t *__restrict__ key32,
                          uint32_t *__restrict__ perm, const double *__restrict__ ulen) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S || s + 1 >= S) return;
    const uint32_t k = key32[s];
    if (key32[s + 1] != k || k <= 1u) return;
    const int b = cnt::find_segment(soff, B, (int)s);
    const int64_t lo = soff[b], hi = soff[b + 1];
    if (s > lo && key32[s - 1] == k) return;                       // not the head of its run
    int64_t e = s + 1;
    while (e < hi && key32[e] == k) e++;
    auto low = [&](uint32_t p) { return ~(uint32_t)(unsigned long long)__double_as_longlong(ulen[p]); };
    for (int64_t a = s + 1; a < e; a++) {
        const uint32_t pa = perm[a], ka = low(pa);
        int64_t c = a - 1;
        while (c >= s && low(perm[c]) > ka) {
            perm[c + 1] = perm[c];
            c--;
        }
        perm[c + 1] = pa;
    }
}

// final gather into sorted order + end points (netsim.py:97-103)
__global__ void k_gather(int B, const int32_t *__restrict__ soff, int64_t S, const uint32_t *__restrict__ perm,
                         const double *__restrict__ uxc, const double *__restrict__ uyc,
                         const double *__restrict__ uang, const double *__restrict__ ulen,
                         const uint8_t *__restrict__ ukind, double *__restrict__ xc, double *__restrict__ yc,
                         double *__restrict__ ang, double *__restrict__ len, uint8_t *__restrict__ kind,
                         double *__restrict__ xa, double *__restrict__ ya, double *__restrict__ xb,
                         double *__restrict__ yb, int32_t *__restrict__ orig) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int b = cnt::find_segment(soff, B, (int)s);
    uint32_t src = perm[s];
    double x = uxc[src], y = uyc[src], a = uang[src], l = ulen[src];
    xc[s] = x; yc[s] = y; ang[s] = a; len[s] = l;
    kind[s] = ukind[src];
    orig[s] = (int)src - soff[b];
    double c = cos(a), sn = sin(a), h = l / 2;
    xa[s] = x - h * c;
    xb[s] = x + h * c;
    ya[s] = y + h * sn;
    yb[s] = y - h * sn;
}

}  // namespace

extern "C" int64_t cnt_gen_sticks_workspace_bytes(int64_t S, int B) {
    using cnt::ws_bytes_for;
    return 4 * ws_bytes_for(S, 8) + ws_bytes_for(S, 1) + 2 * ws_bytes_for(S, 8) + 2 * ws_bytes_for(S, 4) +
           cnt::segsort_workspace_bytes(S, B);
}

extern "C" int cnt_gen_sticks(const uint64_t *seeds, int B, const int32_t *soff, const int32_t *h_soff,
                              const double *scaling, double pm, int len_mode, double len_const, double *xc,
                              double *yc, double *angle, double *length, uint8_t *kind, double *xa, double *ya,
                              double *xb, double *yb, int32_t *orig, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && seeds && soff && h_soff && scaling && xc && yc && angle && length && kind && xa && ya &&
                  xb && yb && orig && ws);
    CNT_CHECK_ARG(h_soff[0] == 0 && (len_mode == 0 || len_mode == 1));
    for (int b = 0; b < B; b++) CNT_CHECK_ARG(h_soff[b + 1] - h_soff[b] >= 2);   // two electrodes at least
    int64_t S = h_soff[B];
    cnt::Workspace w(ws, ws_bytes);
    double *uxc = w.take<double>(S), *uyc = w.take<double>(S), *uang = w.take<double>(S), *ulen = w.take<double>(S);
    uint8_t *ukind = w.take<uint8_t>(S);
    unsigned long long *keyA = w.take<unsigned long long>(S), *keyB = w.take<unsigned long long>(S);
    uint32_t *valA = w.take<uint32_t>(S), *valB = w.take<uint32_t>(S);
    int64_t sb = cnt::segsort_workspace_bytes(S, B);
    void *sws = w.take<char>(sb);
    if (!sws) {
        cnt::set_error("gen_sticks workspace too small: have %lld need %lld", (long long)ws_bytes,
                       (long long)cnt_gen_sticks_workspace_bytes(S, B));
        return CNT_ECAPACITY;
    }
    k_generate<<<cnt::grid_for(S, 256), 256, 0, st>>>((const unsigned long long *)seeds, B, soff, S, scaling, pm,
                                                      len_mode, len_const, uxc, uyc, uang, ulen, ukind, keyA,
                                                      len_mode == 0 ? (uint32_t *)keyA : nullptr, valA);
    CNT_LAUNCHED();
    uint32_t *perm;
    if (len_mode == 0) {
        uint32_t *ks;
        int rc = cnt::segsort_pairs32(B, h_soff, (uint32_t *)keyA, valA, (uint32_t *)keyB, valB, 4, sws, sb, st, &ks, &perm);
        if (rc) return rc;
        k_tie_fix<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, ks, perm, ulen);
        CNT_LAUNCHED();
    } else {
        unsigned long long *ks;
        int rc = cnt::segsort_pairs(B, h_soff, keyA, valA, keyB, valB, 8, sws, sb, st, &ks, &perm);
        if (rc) return rc;
    }
    k_gather<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, perm, uxc, uyc, uang, ulen, ukind, xc, yc, angle, length,
                                                    kind, xa, ya, xb, yb, orig);
    CNT_LAUNCHED();
    return CNT_OK;
}
