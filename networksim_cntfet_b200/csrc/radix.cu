// Segmented stable LSD radix sort of (u64 or u32 key, u32 value) pairs, 8-bit digits.
// Used for the reference's length sort (netsim.py:133) and for the spatial (Morton) ordering
// of the unknowns of the MNA system.
//
// Tiles never straddle segments and the per-tile digit histograms are laid out
// [segment][digit][tile], so ONE exclusive scan of that array yields segmented, stable
// scatter offsets: a pass is histogram -> scan -> scatter.
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
    int32_t stride;    // tiles in this segment (distance between consecutive digits)
};

template <class K>
__global__ void __launch_bounds__(RS_T) k_rs_hist(const SortTile *__restrict__ tiles,
                                                   const K *__restrict__ key, int shift,
                                                   int32_t *__restrict__ hist) {
    __shared__ int h[256];
    SortTile t = tiles[blockIdx.x];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < t.count; k += RS_T) atomicAdd(&h[(int)((key[t.start + k] >> shift) & 255)], 1);
    __syncthreads();
    hist[t.hist0 + threadIdx.x * t.stride] = h[threadIdx.x];
}

template <class K>
__global__ void __launch_bounds__(RS_T) k_rs_scatter(const SortTile *__restrict__ tiles,
                                                      const K *__restrict__ key_in,
                                                      const uint32_t *__restrict__ val_in,
                                                      K *__restrict__ key_out,
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
        K kk = 0;
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

static int64_t max_tiles(int64_t n, int nseg) { return n / RS_TILE + nseg + 1; }

template <class K>
int segsort_pairs_t(int nseg, const int32_t *h_off, K *keyA, uint32_t *valA, K *keyB, uint32_t *valB, int npasses, void *ws,
                    int64_t ws_bytes, cudaStream_t st, K **key_out, uint32_t **val_out);
}  // namespace

namespace cnt {
int64_t segsort_workspace_bytes(int64_t n, int nseg) {
    int64_t nt = max_tiles(n, nseg);
    return ws_bytes_for(nt, sizeof(SortTile)) + ws_bytes_for(nt * 256 + 1, 4) + cnt_scan_workspace_bytes(nt * 256 + 1);
}

int segsort_pairs(int nseg, const int32_t *h_off, unsigned long long *keyA, uint32_t *valA, unsigned long long *keyB,
                  uint32_t *valB, int npasses, void *ws, int64_t ws_bytes, cudaStream_t st,
                  unsigned long long **key_out, uint32_t **val_out) {
    return segsort_pairs_t<unsigned long long>(nseg, h_off, keyA, valA, keyB, valB, npasses, ws, ws_bytes, st, key_out, val_out);
}
// the same with 32-bit keys (8 instead of 12 bytes moved per element and pass)
int segsort_pairs32(int nseg, const int32_t *h_off, uint32_t *keyA, uint32_t *valA, uint32_t *keyB, uint32_t *valB, int npasses,
                    void *ws, int64_t ws_bytes, cudaStream_t st, uint32_t **key_out, uint32_t **val_out) {
    return segsort_pairs_t<uint32_t>(nseg, h_off, keyA, valA, keyB, valB, npasses, ws, ws_bytes, st, key_out, val_out);
}
}  // namespace cnt

namespace {
template <class K>
int segsort_pairs_t(int nseg, const int32_t *h_off, K *keyA, uint32_t *valA, K *keyB, uint32_t *valB, int npasses, void *ws,
                    int64_t ws_bytes, cudaStream_t st, K **key_out, uint32_t **val_out) {
    std::vector<SortTile> tiles;
    int64_t tile_base = 0;
    for (int b = 0; b < nseg; b++) {
        int n = h_off[b + 1] - h_off[b];
        int nt = (n + RS_TILE - 1) / RS_TILE;
        for (int t = 0; t < nt; t++) {
            SortTile s;
            s.start = h_off[b] + t * RS_TILE;
            s.count = n - t * RS_TILE < RS_TILE ? n - t * RS_TILE : RS_TILE;
            s.hist0 = (int32_t)(tile_base * 256 + t);
            s.stride = nt;
            tiles.push_back(s);
        }
        tile_base += nt;
    }
    int64_t nhist = tile_base * 256, nt = (int64_t)tiles.size();
    int64_t n = h_off[nseg] - h_off[0], nt_alloc = max_tiles(n, nseg);
    *key_out = keyA;
    *val_out = valA;
    if (nt == 0) return CNT_OK;
    cnt::Workspace w(ws, ws_bytes);
    SortTile *d_tiles = w.take<SortTile>(nt_alloc);
    int32_t *hist = w.take<int32_t>(nt_alloc * 256 + 1);
    int64_t sb = cnt_scan_workspace_bytes(nt_alloc * 256 + 1);
    void *sws = w.take<char>(sb);
    if (!sws) {
        cnt::set_error("segsort workspace too small");
        return CNT_ECAPACITY;
    }
    CNT_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(SortTile) * nt, cudaMemcpyHostToDevice, st));
    for (int pass = 0; pass < npasses; pass++) {
        int shift = 8 * pass;
        k_rs_hist<K><<<(unsigned)nt, RS_T, 0, st>>>(d_tiles, keyA, shift, hist);
        CNT_LAUNCHED();
        int rc = cnt_exclusive_scan_i32(hist, hist, nhist, 0, sws, sb, (void *)st);
        if (rc) return rc;
        k_rs_scatter<K><<<(unsigned)nt, RS_T, 0, st>>>(d_tiles, keyA, valA, keyB, valB, shift, hist);
        CNT_LAUNCHED();
        K *tk = keyA; keyA = keyB; keyB = tk;
        uint32_t *tv = valA; valA = valB; valB = tv;
    }
    *key_out = keyA;
    *val_out = valA;
    return CNT_OK;
}
}  // namespace
