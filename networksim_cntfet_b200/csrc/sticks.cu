// Stick placement from a counter-based RNG + the length sort of the reference.
// Replaces make_stick / get_ends / make_sticks (netsim.py:97-125) and the
// `sort_values('length', ascending=False)` of make_intersects_kdtree (netsim.py:132-134).
//
// RNG: Philox4x32-10, key = seed, counter = (stick ordinal, draw block, network id).  Any
// stick of any network can be regenerated independently: ensembles are resumable by index
// and identical no matter how they are sharded over GPUs.  The stream is statistically, not
// bitwise, the reference's (numpy MT19937); parity tests always feed reference sticks.
//
// Sort: segmented stable LSD radix sort (8-bit digits, 8 passes over a 64-bit key).  Tiles
// never straddle networks and the per-tile digit histograms are laid out
// [network][digit][tile], so ONE exclusive scan of that array yields segmented offsets.
// Key: source electrode 0, drain electrode 1, body sticks ~bits(length) (longer first);
// stability keeps generation order among equal lengths.
#include <math.h>
#include <vector>
#include "common.cuh"

namespace {

constexpr int RS_T = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_T * RS_ITEMS;

struct SortTile {
    int32_t start;     // first element (global)
    int32_t count;
    int32_t hist0;     // index of (digit 0, this tile) in the histogram array
    int32_t stride;    // tiles in this network (distance between consecutive digits)
};

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

__global__ void k_generate(uint64_t seed, int64_t net_id0, int B, const int32_t *__restrict__ soff, int64_t S,
                           const double *__restrict__ scaling, double pm, int len_mode, double len_const,
                           double *__restrict__ xc, double *__restrict__ yc, double *__restrict__ ang,
                           double *__restrict__ len, uint8_t *__restrict__ kind, unsigned long long *__restrict__ key,
                           uint32_t *__restrict__ val) {
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
        key[s] = (i == 0 && n > 1) || n == 1 ? 0ull : 1ull;
        return;
    }
    unsigned long long nid = (unsigned long long)(net_id0 + b);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t r0[4], r1[4], r2[4];
    philox4x32_10((uint32_t)i, 0u, (uint32_t)nid, (uint32_t)(nid >> 32), k0, k1, r0);
    philox4x32_10((uint32_t)i, 1u, (uint32_t)nid, (uint32_t)(nid >> 32), k0, k1, r1);
    philox4x32_10((uint32_t)i, 2u, (uint32_t)nid, (uint32_t)(nid >> 32), k0, k1, r2);
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
    key[s] = ~(unsigned long long)__double_as_longlong(l);
}

// ---------------------------------------------------------------- radix sort passes
__global__ void __launch_bounds__(RS_T) k_rs_hist(const SortTile *__restrict__ tiles,
                                                   const unsigned long long *__restrict__ key, int shift,
                                                   int32_t *__restrict__ hist) {
    __shared__ int h[256];
    SortTile t = tiles[blockIdx.x];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < t.count; k += RS_T) atomicAdd(&h[(int)((key[t.start + k] >> shift) & 255)], 1);
    __syncthreads();
    hist[t.hist0 + threadIdx.x * t.stride] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_T) k_rs_scatter(const SortTile *__restrict__ tiles,
                                                      const unsigned long long *__restrict__ key_in,
                                                      const uint32_t *__restrict__ val_in,
                                                      unsigned long long *__restrict__ key_out,
                                                      uint32_t *__restrict__ val_out, int shift,
                                                      const int32_t *__restrict__ hist_scan) {
    __shared__ int base[256];
    __shared__ int wcnt[RS_T / 32][256];
    SortTile t = tiles[blockIdx.x];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    base[threadIdx.x] = hist_scan[t.hist0 + threadIdx.x * t.stride];
#pragma unroll
    for (int k = 0; k < RS_T / 32; k++) wcnt[k][threadIdx.x] = 0;
    __syncthreads();
    for (int r0 = 0; r0 < t.count; r0 += RS_T) {
        int k = r0 + threadIdx.x;
        bool valid = k < t.count;
        unsigned long long kk = 0;
        uint32_t vv = 0;
        int d = 256 + lane;                         // invalid lanes: unique pseudo-digit
        if (valid) {
            kk = key_in[t.start + k];
            vv = val_in[t.start + k];
            d = (int)((kk >> shift) & 255);
        }
        unsigned m = __match_any_sync(0xffffffffu, d);
        int rank = __popc(m & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[w][d] = __popc(m);
        __syncthreads();
        if (valid) {
            int o = base[d] + rank;
            for (int k2 = 0; k2 < w; k2++) o += wcnt[k2][d];
            key_out[o] = kk;
            val_out[o] = vv;
        }
        __syncthreads();
        int add = 0;
#pragma unroll
        for (int k2 = 0; k2 < RS_T / 32; k2++) {
            add += wcnt[k2][threadIdx.x];
            wcnt[k2][threadIdx.x] = 0;
        }
        base[threadIdx.x] += add;
        __syncthreads();
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

struct SortPlan {
    std::vector<SortTile> tiles;
    int64_t nhist;
};
static void make_sort_plan(int B, const int32_t *h_soff, SortPlan &p) {
    p.tiles.clear();
    int64_t tile_base = 0;
    for (int b = 0; b < B; b++) {
        int n = h_soff[b + 1] - h_soff[b];
        int nt = (n + RS_TILE - 1) / RS_TILE;
        for (int t = 0; t < nt; t++) {
            SortTile st;
            st.start = h_soff[b] + t * RS_TILE;
            st.count = n - t * RS_TILE < RS_TILE ? n - t * RS_TILE : RS_TILE;
            st.hist0 = (int32_t)(tile_base * 256 + t);
            st.stride = nt;
            p.tiles.push_back(st);
        }
        tile_base += nt;
    }
    p.nhist = tile_base * 256;
}
static int64_t max_tiles(int64_t S, int B) { return S / RS_TILE + B + 1; }
}  // namespace

extern "C" int64_t cnt_gen_sticks_workspace_bytes(int64_t S, int B) {
    using cnt::ws_bytes_for;
    int64_t nt = max_tiles(S, B);
    return 4 * ws_bytes_for(S, 8) + ws_bytes_for(S, 1) + 2 * ws_bytes_for(S, 8) + 2 * ws_bytes_for(S, 4) +
           ws_bytes_for(nt, sizeof(SortTile)) + ws_bytes_for(nt * 256 + 1, 4) + cnt_scan_workspace_bytes(nt * 256 + 1);
}

extern "C" int cnt_gen_sticks(uint64_t seed, int64_t net_id0, int B, const int32_t *soff, const int32_t *h_soff,
                              const double *scaling, double pm, int len_mode, double len_const, double *xc,
                              double *yc, double *angle, double *length, uint8_t *kind, double *xa, double *ya,
                              double *xb, double *yb, int32_t *orig, void *ws, int64_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && soff && h_soff && scaling && xc && yc && angle && length && kind && xa && ya && xb && yb &&
                  orig && ws);
    CNT_CHECK_ARG(h_soff[0] == 0 && (len_mode == 0 || len_mode == 1));
    for (int b = 0; b < B; b++) CNT_CHECK_ARG(h_soff[b + 1] - h_soff[b] >= 2);   // two electrodes at least
    int64_t S = h_soff[B];
    SortPlan plan;
    make_sort_plan(B, h_soff, plan);
    int64_t nt = (int64_t)plan.tiles.size(), nt_alloc = max_tiles(S, B);
    cnt::Workspace w(ws, ws_bytes);
    double *uxc = w.take<double>(S), *uyc = w.take<double>(S), *uang = w.take<double>(S), *ulen = w.take<double>(S);
    uint8_t *ukind = w.take<uint8_t>(S);
    unsigned long long *keyA = w.take<unsigned long long>(S), *keyB = w.take<unsigned long long>(S);
    uint32_t *valA = w.take<uint32_t>(S), *valB = w.take<uint32_t>(S);
    SortTile *d_tiles = w.take<SortTile>(nt_alloc);
    int32_t *hist = w.take<int32_t>(nt_alloc * 256 + 1);
    int64_t sb = cnt_scan_workspace_bytes(nt_alloc * 256 + 1);
    void *sws = w.take<char>(sb);
    if (!sws) {
        cnt::set_error("gen_sticks workspace too small: have %lld need %lld", (long long)ws_bytes,
                       (long long)cnt_gen_sticks_workspace_bytes(S, B));
        return CNT_ECAPACITY;
    }
    CNT_CUDA(cudaMemcpyAsync(d_tiles, plan.tiles.data(), sizeof(SortTile) * nt, cudaMemcpyHostToDevice, st));
    k_generate<<<cnt::grid_for(S, 256), 256, 0, st>>>(seed, net_id0, B, soff, S, scaling, pm, len_mode, len_const, uxc,
                                                      uyc, uang, ulen, ukind, keyA, valA);
    CNT_LAUNCHED();
    for (int pass = 0; pass < 8; pass++) {
        int shift = 8 * pass;
        k_rs_hist<<<(unsigned)nt, RS_T, 0, st>>>(d_tiles, keyA, shift, hist);
        CNT_LAUNCHED();
        int rc = cnt_exclusive_scan_i32(hist, hist, plan.nhist, 0, sws, sb, stream);
        if (rc) return rc;
        k_rs_scatter<<<(unsigned)nt, RS_T, 0, st>>>(d_tiles, keyA, valA, keyB, valB, shift, hist);
        CNT_LAUNCHED();
        unsigned long long *tk = keyA; keyA = keyB; keyB = tk;
        uint32_t *tv = valA; valA = valB; valB = tv;
    }
    k_gather<<<cnt::grid_for(S, 256), 256, 0, st>>>(B, soff, S, valA, uxc, uyc, uang, ulen, ukind, xc, yc, angle, length,
                                                    kind, xa, ya, xb, yb, orig);
    CNT_LAUNCHED();
    return CNT_OK;
}
