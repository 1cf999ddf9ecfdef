// Symbolic phase of the star-mesh direct solver (replaces solve_mna, cnet.py:147-149; numeric phase
// in starmesh.cu).  Topology only: one plan serves every gate configuration of a batch.
//
// State: an edge pool whose SLOTS never move (round-0 edges, then fill edges in a deterministic
// order), an insert-only open-addressing hash (u, w) -> slot, per-unknown degree, a compact list of
// the live edges and a compact ascending list of the unknowns still in play.  A round is eight
// launches of fixed-size grids that read their sizes from device memory (no host round trip):
//
//   k_mark      live edges: the endpoint with the larger (degree, hash, index) is blocked
//   k_select    unknowns in play: unblocked -> pivot (an independent set of local minima);
//               one single-launch scan (chunked_scan) numbers the pivots, sizes their lists and pair ranges
//               and compacts the survivors
//   k_incident  live edges: an edge of a pivot goes to the pivot's incident list, the rest survive
//   k_pivots    per pivot: neighbours ascending; degrees of the neighbours, unknowns left per network
//   k_items_a   per neighbour pair (u, w) of a pivot: find-or-insert (u, w) in the hash and link the
//               pair into the target's list; per incident entry: link into the neighbour's list
//   k_items_b   the last arrival of every list counts it and names its owner (smallest item)
//   k_items_scan  the same scan over the owners: group numbers, offsets, new slot numbers
//   k_items_c   per owner: members ordered by item index (= by pivot), written as the contribution
//               list of the target edge / the update list of the neighbour; new edges get their slot
//
// Everything the numeric phase sums is ordered by pivot index, and pivot order inside a network does
// not depend on what else is in the batch, so the bits of the solution do not either.  The plan
// layout itself is deterministic (the only nondeterministic orders -- arrival in the lists, position in
// the live-edge list -- are internal).  A network freezes once it has <= dense_max unknowns left; the
// frozen remainders are handed to the dense tail (k_final_*).
#include <stdlib.h>
#include "starmesh_plan.cuh"

namespace {
using cnt::sm::Plan;

constexpr int TB = 256;
constexpr int IPT = 4;                 // items per thread of the chained scans (8 items with the values scanned in place
                                       // in ONE shared array -- half as many look-backs -- measured 2 % slower)
constexpr int TILE = TB * IPT;
constexpr int SAT = 1 << 29;           // saturation of the pair counter (overflow detection)
constexpr int GRID = 148 * 8;          // full grid of the device-sized kernels (8 CTAs of 256 threads per SM)
constexpr int GRID_SCAN = 148 * 4;     // grid of the chunked scans: one chunk per CTA, all CTAs resident (48 / 57 registers)
constexpr int SORT_SMEM = 512;         // longest neighbour list a warp sorts out of shared memory
constexpr unsigned long long DEG1 = 1ull << 46;      // one unit of degree in a priority word

struct __align__(16) HEnt {
    unsigned long long key;            // (u << 32) | w, u < w global rows; all ones = empty
    int slot;                          // edge slot, -1 = inserted this round, not numbered yet
    int head;                          // last arrived pair item of the current round, -1 = none
};
constexpr unsigned long long HEMPTY = ~0ull;

struct __align__(16) V4 {
    int a, b, c, d;
};
__device__ __forceinline__ V4 v4(int a, int b, int c, int d) {
    V4 r;
    r.a = a; r.b = b; r.c = c; r.d = d;
    return r;
}
__device__ __forceinline__ V4 operator+(V4 x, V4 y) {
    V4 r;
    r.a = x.a + y.a;
    r.b = x.b + y.b;
    int c = x.c + y.c;
    r.c = c > SAT ? SAT : c;
    r.d = x.d + y.d;
    return r;
}
__device__ __forceinline__ V4 shfl_up4(V4 v, int o) {
    return v4(__shfl_up_sync(~0u, v.a, o), __shfl_up_sync(~0u, v.b, o), __shfl_up_sync(~0u, v.c, o),
              __shfl_up_sync(~0u, v.d, o));
}
__device__ __forceinline__ V4 shfl_down4(V4 v, int o) {
    return v4(__shfl_down_sync(~0u, v.a, o), __shfl_down_sync(~0u, v.b, o), __shfl_down_sync(~0u, v.c, o),
              __shfl_down_sync(~0u, v.d, o));
}
__device__ __forceinline__ V4 shfl4(V4 v, int l) {
    return v4(__shfl_sync(~0u, v.a, l), __shfl_sync(~0u, v.b, l), __shfl_sync(~0u, v.c, l), __shfl_sync(~0u, v.d, l));
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ V4 ld_v4(const V4 *p) {
    int4 t = __ldcg(reinterpret_cast<const int4 *>(p));
    return v4(t.x, t.y, t.z, t.w);
}

// everything the kernels need, by value
struct Sym {
    int B, R, NNZ, dense_max;
    int capS, capC, capI;
    unsigned hcap;
    const int32_t *roff, *rowptr, *col;
    Plan p;
    // workspace
    int32_t *eu, *ew;                  // [capS] endpoints of every slot (global rows, eu < ew)
    unsigned long long *pri;           // [R] (degree << 46) | (hash of the local index << 26) | local index
    int32_t *mid, *blocked, *rnet, *fill, *nhead, *loc;     // [R]; mid: -1 in play, -2 frozen, >= 0 pivot index
    int32_t *lm;                       // [capS] pivot index of every incident entry
    int32_t *live[2];                  // [capS] live edge slots (ping-pong)
    int2 *luw[2];                      // [capS] ... and their endpoints
    long long *hoff;                   // [B + 1] hash region of every network (keeps the probes of a network together)
    int32_t *nodes[2];                 // [R] unknowns in play, ascending (ping-pong)
    int32_t *nalive;                   // [B] unknowns left per network
    int32_t *dcount;                   // [B] remaining edges per network (dense tail)
    int32_t *nlive;                    // [MAX_ROUNDS + 2] size of the live list at the start of round r
    int32_t *nsurv;                    // [MAX_ROUNDS + 2] ... of which survivors of the round before (the fill edges follow)
    HEnt *htab;
    int32_t *pent, *pnext, *gsz, *ghead, *pm, *pa, *pb, *gidx, *goff, *gnew;     // [capI] per item of a round
    int32_t *pptr;                     // [R + 1] first pair of every pivot of the round
    uint2 *phash;                      // [R] hash region (first entry, size) of every pivot of the round: k_items_a finds it
                                       // with ONE load next to nptr / pptr instead of nbr -> rnet -> hoff (two levels of its chain)
    // chained scans
    int32_t *tickets;                  // one per scan invocation
    int32_t *sflag;
    V4 *sagg, *sinc;
};

__device__ __forceinline__ unsigned long long priority(int deg, int lid) {
    unsigned h = ((unsigned)lid * 2654435761u) >> 12;                 // 20-bit tie-break hash of the network-local index
    unsigned long long d = (unsigned long long)(deg < (1 << 18) - 1 ? deg : (1 << 18) - 1);
    return (d << 46) | ((unsigned long long)h << 26) | (unsigned long long)(unsigned)lid;
}
// entry of key in the hash region of its network, inserted when absent (slot stays -1); -1 when the region is full
__device__ __forceinline__ int hash_find_or_insert(const Sym &s, long long h0, unsigned hsize, unsigned long long key);
__device__ __forceinline__ int hash_find_or_insert(const Sym &s, int net, unsigned long long key) {
    const long long h0 = s.hoff[net];
    return hash_find_or_insert(s, h0, (unsigned)(s.hoff[net + 1] - h0), key);
}
__device__ __forceinline__ int hash_find_or_insert(const Sym &s, long long h0, unsigned hsize, unsigned long long key) {
    unsigned long long z = key * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 29)) * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 32;
    unsigned h = (unsigned)(((z & 0xFFFFFFFFull) * (unsigned long long)hsize) >> 32);
    for (unsigned probes = 0; probes < hsize; probes++) {
        HEnt *ent = s.htab + h0 + h;
        unsigned long long old = ent->key;
        if (old == key) return (int)(h0 + h);
        if (old == HEMPTY) {
            old = atomicCAS(&ent->key, HEMPTY, key);
            if (old == HEMPTY || old == key) return (int)(h0 + h);
        }
        h = h + 1 == hsize ? 0u : h + 1;
    }
    atomicOr(&s.p.info->error, CNT_SM_ERR_HASH);
    return -1;
}
// position of a new element in a list with an atomic counter, one atomic per warp.  Must be called
// by all 32 lanes.
__device__ __forceinline__ int warp_append(int *counter, bool pred) {
    unsigned m = __ballot_sync(~0u, pred);
    if (!m) return -1;
    int lane = threadIdx.x & 31, leader = __ffs(m) - 1, base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(~0u, base, leader);
    return pred ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// the same with ONE atomic per CTA (same-address atomics are served one per L2 clock: a counter hit once per
// warp becomes the bottleneck of a pass over millions of items).  Must be called by all threads of the CTA.
__device__ __forceinline__ int block_append(int *counter, bool pred, int *s_cnt /* TB / 32 + 1 ints */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(~0u, pred);
    if (lane == 0) s_cnt[wid] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < TB / 32; w++) {
            const int c = s_cnt[w];
            s_cnt[w] = tot;
            tot += c;
        }
        s_cnt[TB / 32] = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    const int pos = s_cnt[TB / 32] + s_cnt[wid] + __popc(m & ((1u << lane) - 1u));
    __syncthreads();
    return pred ? pos : -1;
}

// ---------------------------------------------------------------------------------------
// Device-wide exclusive scan of one V4 per item in ONE launch (chained scan with decoupled
// look-back): tiles are handed out by an atomic ticket, so the predecessors of a tile are always
// being processed by running CTAs (no deadlock whatever the grid size); a tile publishes its
// aggregate, looks back over its predecessors a warp at a time, then publishes its inclusive
// prefix.  Flags carry the invocation's epoch, so the status buffers are never cleared.
//   f(i) -> V4 value of item i;  g(i, exclusive prefix, own value);  fin(grand total) once.
// f and g are called with consecutive items on consecutive threads (coalesced); the values go through
// shared memory to the blocked arrangement the scan needs.
// ---------------------------------------------------------------------------------------
template <class F, class G, class H>
__device__ __forceinline__ void chained_scan(int n, int epoch, int *ticket, int *sflag, V4 *sagg, V4 *sinc, int *err,
                                             F f, G g, H fin) {
    __shared__ int s_tile;
    __shared__ V4 s_prefix, s_total;
    __shared__ V4 s_warp[TB / 32];
    __shared__ V4 sv[TILE], se[TILE];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ntiles = (n + TILE - 1) / TILE;
    const V4 zero = v4(0, 0, 0, 0);
    if (ntiles == 0) {
        if (blockIdx.x == 0 && tid == 0) fin(zero);
        return;
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(ticket, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const int i0 = tile * TILE;
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int i = i0 + k * TB + tid;
            sv[k * TB + tid] = i < n ? f(i) : zero;
        }
        __syncthreads();
        V4 sum = zero;
#pragma unroll
        for (int k = 0; k < IPT; k++) sum = sum + sv[tid * IPT + k];
        V4 inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            V4 t = shfl_up4(inc, o);
            if (lane >= o) inc = t + inc;
        }
        V4 ex = shfl_up4(inc, 1);
        if (lane == 0) ex = zero;
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            V4 w = lane < TB / 32 ? s_warp[lane] : zero;
            V4 wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                V4 t = shfl_up4(wi, o);
                if (lane >= o) wi = t + wi;
            }
            V4 wex = shfl_up4(wi, 1);
            if (lane == 0) wex = zero;
            if (lane < TB / 32) s_warp[lane] = wex;
            const V4 total = shfl4(wi, TB / 32 - 1);
            V4 prefix = zero;
            if (tile == 0) {
                if (lane == 0) {
                    sinc[0] = total;
                    st_release(&sflag[0], epoch * 4 + 2);
                }
            } else {
                if (lane == 0) {
                    sagg[tile] = total;
                    st_release(&sflag[tile], epoch * 4 + 1);
                }
                int look = tile - 1;
                for (;;) {
                    const int idx = look - lane;
                    int state = 2;
                    V4 val = zero;
                    if (idx >= 0) {
                        int fl, spins = 0;
                        do {
                            fl = ld_acquire(&sflag[idx]);
                            if ((fl >> 2) != epoch && (++spins & 1023) == 0 &&
                                (spins >= (1 << 20) || (*(volatile int *)err & CNT_SM_ERR_SCAN)))
                                break;
                        } while ((fl >> 2) != epoch);
                        if ((fl >> 2) != epoch) {
                            atomicOr(err, CNT_SM_ERR_SCAN);     // watchdog: never hang the GPU
                        } else {
                            state = fl & 3;
                            val = state == 2 ? ld_v4(&sinc[idx]) : ld_v4(&sagg[idx]);
                        }
                    }
                    const unsigned m = __ballot_sync(~0u, state == 2);
                    const int first = m ? __ffs(m) - 1 : 32;
                    if (lane > first) val = zero;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) val = val + shfl_down4(val, o);
                    prefix = prefix + val;            // meaningful in lane 0
                    if (m) break;
                    look -= 32;
                }
                if (lane == 0) {
                    sinc[tile] = prefix + total;
                    st_release(&sflag[tile], epoch * 4 + 2);
                }
            }
            if (lane == 0) {
                s_prefix = prefix;
                s_total = prefix + total;
            }
        }
        __syncthreads();
        V4 base = s_prefix + s_warp[wid] + ex;
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            se[tid * IPT + k] = base;
            base = base + sv[tid * IPT + k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int i = i0 + k * TB + tid;
            if (i < n) g(i, se[k * TB + tid], sv[k * TB + tid]);
        }
        if (tile == ntiles - 1 && tid == 0) fin(s_total);
    }
}

// ---------------------------------------------------------------------------------------
// The scans of the rounds without the chain: reduce, then scan.  In the chained scan the inclusive prefix of a tile
// hangs on that of a predecessor ~100 tiles back, hop by hop (ncu, round 0 of 224 networks: 63 % of the warp samples
// at the barrier behind warp 0's look-back, two membars and 3.3 look-back steps per tile of 1024 items: 17 us per tile
// and CTA, 60 G items/s however many CTAs are in flight).  Here every CTA takes ONE contiguous chunk of tiles (chunk
// numbers by ticket, so the lower chunks are always held by running CTAs), (1) sums its chunk and publishes the
// total, (2) waits ONCE for the totals of the lower chunks -- nobody waits for anybody before publishing, so there is
// no chain --, (3) walks its chunk again tile by tile with a running prefix and calls g.
//   An item is described by a 32-bit code: code1(i) produces it in pass (1) (loads only), park(i, code) may keep it
//   in an array, code3(i) returns it again in pass (3), dec(code) -> V4.  Both passes give item i to the same thread.
// Pass (3) scans a tile in the striped arrangement the loads come in: one shuffle scan per row of 256 items, the
// IPT * 8 = 32 warp totals scanned by warp 0 -- no transposition through shared memory (the blocked V4 accesses of
// the chained scan are 4-way bank conflicts), two barriers per tile.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ V4 block_sum4(V4 v, V4 *s_warp /* TB / 32 */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + shfl_down4(v, o);
    if (lane == 0) s_warp[wid] = v;
    __syncthreads();
    V4 t = s_warp[0];
#pragma unroll
    for (int w = 1; w < TB / 32; w++) t = t + s_warp[w];
    __syncthreads();
    return t;
}
template <class C1, class P, class C3, class D, class G, class H>
__device__ __forceinline__ void chunked_scan(int n, int epoch, int *ticket, int *sflag, V4 *sagg, int *err, C1 code1, P park,
                                             C3 code3, D dec, G g, H fin) {
    constexpr int NW = TB / 32;
    static_assert(IPT * NW == 32, "the warp totals of a tile are scanned by one warp");
    __shared__ int s_chunk;
    __shared__ V4 s_warp[NW];
    __shared__ V4 s_part[2][32], s_off[2][32], s_tot[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ntiles = (n + TILE - 1) / TILE;
    const V4 zero = v4(0, 0, 0, 0);
    if (ntiles == 0) {
        if (blockIdx.x == 0 && tid == 0) fin(zero);
        return;
    }
    const int nchunks = (int)gridDim.x < ntiles ? (int)gridDim.x : ntiles;
    if (tid == 0) s_chunk = atomicAdd(ticket, 1);
    __syncthreads();
    const int c = s_chunk;
    if (c >= nchunks) return;
    const int q = ntiles / nchunks, rem = ntiles % nchunks;
    const int t0 = c * q + (c < rem ? c : rem), t1 = t0 + q + (c < rem ? 1 : 0);
    // (1) total of the chunk
    V4 sum = zero;
    for (int tile = t0; tile < t1; tile++) {
        int cd[IPT];
#pragma unroll
        for (int k = 0; k < IPT; k++) {                  // the loads of the IPT items first (independent chains in flight) ...
            const int i = tile * TILE + k * TB + tid;
            cd[k] = i < n ? code1(i) : 0;
        }
#pragma unroll
        for (int k = 0; k < IPT; k++) {                  // ... then whatever code1 wants to keep of them
            const int i = tile * TILE + k * TB + tid;
            if (i < n) {
                park(i, cd[k]);
                sum = sum + dec(cd[k]);
            }
        }
    }
    const V4 total = block_sum4(sum, s_warp);
    if (tid == 0) {
        sagg[c] = total;
        st_release(&sflag[c], epoch * 4 + 1);
    }
    // (2) totals of the lower chunks
    V4 pre = zero;
    for (int j = tid; j < c; j += TB) {
        int fl, spins = 0;
        do {
            fl = ld_acquire(&sflag[j]);
            if ((fl >> 2) != epoch && (++spins & 1023) == 0 && (spins >= (1 << 20) || (*(volatile int *)err & CNT_SM_ERR_SCAN)))
                break;
        } while ((fl >> 2) != epoch);
        if ((fl >> 2) != epoch) atomicOr(err, CNT_SM_ERR_SCAN);     // watchdog: never hang the GPU
        else pre = pre + ld_v4(&sagg[j]);
    }
    V4 carry = block_sum4(pre, s_warp);
    // (3) the chunk again, tile by tile
    int buf = 0;
    for (int tile = t0; tile < t1; tile++, buf ^= 1) {
        const int i0 = tile * TILE;
        V4 val[IPT], ex[IPT];
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int i = i0 + k * TB + tid;
            val[k] = i < n ? dec(code3(i)) : zero;
        }
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            V4 inc = val[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                V4 t = shfl_up4(inc, o);
                if (lane >= o) inc = t + inc;
            }
            ex[k] = shfl_up4(inc, 1);
            if (lane == 0) ex[k] = zero;
            if (lane == 31) s_part[buf][k * NW + wid] = inc;
        }
        __syncthreads();
        if (wid == 0) {
            V4 inc = s_part[buf][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                V4 t = shfl_up4(inc, o);
                if (lane >= o) inc = t + inc;
            }
            V4 e = shfl_up4(inc, 1);
            if (lane == 0) e = zero;
            s_off[buf][lane] = e;
            if (lane == 31) s_tot[buf] = inc;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int i = i0 + k * TB + tid;
            if (i < n) g(i, carry + s_off[buf][k * NW + wid] + ex[k], val[k]);
        }
        carry = carry + s_tot[buf];
    }
    if (c == nchunks - 1 && tid == 0) fin(carry);
}

__device__ __forceinline__ int net_of_row(const Sym &s, int r) { return cnt::find_segment(s.roff, s.B, r); }

// ---------------------------------------------------------------------------------------
// initialisation: hash regions; round-0 edge slots (unique stick pairs of the CSR pattern,
// row-major), their value lists (e0_ptr / e0_src), the hash, priorities, the first live / node lists
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_hoff(Sym s) {
    // region of network b: share of the table proportional to its pattern entries (+64)
    __shared__ long long s_n[1024];
    const int t = threadIdx.x;
    const int per = (s.B + 1023) / 1024;
    const int lo = min(t * per, s.B), hi = min(lo + per, s.B);
    const double scale = (double)(s.hcap - 64.0 * s.B) / (double)(s.NNZ > 0 ? s.NNZ : 1);
    auto size_of = [&](int b) {
        const long long nz = s.rowptr[s.roff[b + 1]] - s.rowptr[s.roff[b]];
        return (long long)(scale * (double)nz) + 64;
    };
    long long sn = 0;
    for (int b = lo; b < hi; b++) sn += size_of(b);
    s_n[t] = sn;
    __syncthreads();
    if (t == 0) {
        long long acc = 0;
        for (int k = 0; k < 1024; k++) {
            const long long c = s_n[k];
            s_n[k] = acc;
            acc += c;
        }
        s.hoff[s.B] = acc;
    }
    __syncthreads();
    sn = s_n[t];
    for (int b = lo; b < hi; b++) {
        s.hoff[b] = sn;
        sn += size_of(b);
    }
}

__global__ void __launch_bounds__(TB) k_init(Sym s, int epoch) {
    const int gt = blockIdx.x * TB + threadIdx.x, gn = gridDim.x * TB;
    for (int b = gt; b < s.B; b += gn) s.nalive[b] = s.roff[b + 1] - s.roff[b];
    auto f = [&](int i) {
        const int base = s.roff[net_of_row(s, i)];
        int prev = -1, nu = 0, nk = 0;
        for (int k = s.rowptr[i]; k < s.rowptr[i + 1]; k++) {
            const int c = s.col[k] + base;
            if (c <= i) continue;
            nk++;
            if (c != prev) nu++;
            prev = c;
        }
        return v4(nu, nk, 0, 0);
    };
    auto g = [&](int i, V4 ex, V4 own) {
        const int net = net_of_row(s, i);
        const int base = s.roff[net];
        s.rnet[i] = net;
        s.nodes[0][i] = i;
        int prev = -1, prevu = -1, nd = 0, slot = ex.a - 1, kk = ex.b;
        for (int k = s.rowptr[i]; k < s.rowptr[i + 1]; k++) {
            const int c = s.col[k] + base;
            if (c == i) continue;
            if (c != prev) nd++;
            prev = c;
            if (c < i) continue;
            if (c != prevu) {
                slot++;
                if (slot < s.capS) {
                    s.eu[slot] = i;
                    s.ew[slot] = c;
                    s.live[0][slot] = slot;
                    s.luw[0][slot] = make_int2(i, c);
                    s.p.e0_ptr[slot] = kk;              // the hash entry of the slot: k_init_hash (one thread per slot)
                }
                prevu = c;
            }
            s.p.e0_src[kk++] = k;
        }
        s.pri[i] = priority(nd, i - base);
        if (nd >= (1 << 17)) atomicOr(&s.p.info->error, CNT_SM_ERR_SLOTS);      // degree field of the priority word
    };
    auto fin = [&](V4 t) {
        cnt_sm_info *info = s.p.info;
        info->R = s.R;
        info->nE0 = t.a;
        if (t.a <= s.capS) s.p.e0_ptr[t.a] = t.b;
        else {
            atomicOr(&info->error, CNT_SM_ERR_SLOTS);
            info->done = 1;
        }
        s.nlive[0] = t.a <= s.capS ? t.a : 0;
        s.p.nptr[0] = 0;
        cnt_sm_round *r0 = &info->rounds[0];
        r0->ebase = r0->lbase = r0->cbase = r0->gebase = r0->gnbase = 0;
        r0->sbase = t.a;
        r0->nnodes = s.R;
        r0->nlive = t.a;
    };
    chained_scan(s.R, epoch, s.tickets, s.sflag, s.sagg, s.sinc, &s.p.info->error, f, g, fin);
}

// round-0 edges into the hash, one thread per slot: convergent and fully parallel (inside the per-row loops of
// k_init the probe loops ran with 6 of 32 lanes active)
__global__ void __launch_bounds__(TB) k_init_hash(Sym s) {
    const cnt_sm_info *info = s.p.info;
    if (info->done | info->error) return;
    const int n = info->nE0;
    for (int e = blockIdx.x * TB + threadIdx.x; e < n; e += gridDim.x * TB) {
        const int u = s.eu[e], w = s.ew[e];
        const int h = hash_find_or_insert(s, s.rnet[u], ((unsigned long long)(unsigned)u << 32) | (unsigned)w);
        if (h >= 0) s.htab[h].slot = e;
    }
}

// ---------------------------------------------------------------------------------------
// round kernels
// ---------------------------------------------------------------------------------------
// a frozen network's edges may still be in the list (k_incident drops them): blocking a frozen
// unknown is harmless, k_select no longer looks at it
__global__ void __launch_bounds__(TB) k_mark(Sym s, int r) {
    if (s.p.info->done | s.p.info->error) return;
    const int n = s.nlive[r];
    const int2 *L = s.luw[r & 1];
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB) {
        const int2 e = L[i];
        s.blocked[s.pri[e.x] > s.pri[e.y] ? e.x : e.y] = r + 1;
    }
}

__global__ void __launch_bounds__(TB) k_select(Sym s, int r, int epoch) {
    cnt_sm_info *info = s.p.info;
    if (info->done | info->error) return;
    cnt_sm_round *rd = &info->rounds[r];
    const int n = rd->nnodes, ebase = rd->ebase, lbase = rd->lbase;
    const int32_t *N = s.nodes[r & 1];
    int32_t *Nout = s.nodes[(r + 1) & 1];
    // code of an unknown: 0 frozen (leaves the list), 1 survivor, 2 + degree pivot; parked in s.loc (free until the finals)
    auto code1 = [&](int i) {
        const int v = N[i];
        if (s.nalive[s.rnet[v]] <= s.dense_max) return 0;
        if (s.blocked[v] == r + 1) return 1;
        return 2 + (int)(s.pri[v] >> 46);
    };
    auto park = [&](int i, int code) {
        if (code - 2 >= (1 << 17)) atomicOr(&s.p.info->error, CNT_SM_ERR_SLOTS);    // degree field about to overflow
        s.loc[i] = code;
    };
    auto code3 = [&](int i) { return s.loc[i]; };
    auto dec = [&](int code) {
        if (code < 2) return v4(0, 0, 0, code);
        const int d = code - 2;
        const long long pairs = (long long)d * (d - 1) / 2;
        return v4(1, d, pairs > SAT ? SAT : (int)pairs, 0);
    };
    auto g = [&](int i, V4 ex, V4 own) {
        const int v = N[i];
        if (own.a) {
            const int m = ebase + ex.a;
            s.p.piv[m] = v;
            s.mid[v] = m;
            if ((long long)lbase + ex.b <= s.capS) s.p.nptr[m] = lbase + ex.b;
            s.pptr[ex.a] = ex.c;
        } else if (own.d) {
            Nout[ex.d] = v;
        } else {
            s.mid[v] = -2;                                                      // frozen from now on
        }
    };
    auto fin = [&](V4 t) {
        int err = 0;
        if ((long long)lbase + t.b > s.capS) err |= CNT_SM_ERR_SLOTS;
        if ((long long)t.c + t.b > s.capI) err |= CNT_SM_ERR_ITEMS;
        if ((long long)rd->cbase + t.c > s.capC) err |= CNT_SM_ERR_CONTRIB;
        if (t.a > 0 && r + 1 >= CNT_SM_MAX_ROUNDS) err |= CNT_SM_ERR_ROUNDS;
        if (err) {
            atomicOr(&info->error, err);
            t = v4(0, 0, 0, 0);
        }
        rd->nelim = t.a;
        rd->nlist = t.b;
        rd->npairs = t.c;
        rd->nlive = s.nlive[r];
        if (t.a == 0) {          // nothing left to eliminate (or an error): the rounds are over
            info->nrounds = r;
            __threadfence();
            info->done = 1;
            return;
        }
        s.p.nptr[ebase + t.a] = lbase + t.b;
        s.pptr[t.a] = t.c;
        info->rounds[r + 1].nnodes = t.d;
    };
    chunked_scan(n, epoch, s.tickets + epoch, s.sflag, s.sagg, &info->error, code1, park, code3, dec, g, fin);
}

__global__ void __launch_bounds__(TB) k_incident(Sym s, int r) {
    __shared__ int s_cnt[TB / 32 + 1];
    const cnt_sm_info *info = s.p.info;
    if (info->done | info->error) return;
    const int n = s.nlive[r], ebase = info->rounds[r].ebase;
    const int32_t *L = s.live[r & 1];
    const int2 *Luw = s.luw[r & 1];
    int32_t *Lout = s.live[(r + 1) & 1];
    int2 *Luwout = s.luw[(r + 1) & 1];
    for (int i0 = blockIdx.x * TB; i0 < n; i0 += gridDim.x * TB) {        // CTA-uniform trip count
        const int i = i0 + threadIdx.x;
        bool keep = false;
        int e = 0;
        int2 uw = make_int2(0, 0);
        if (i < n) {
            e = L[i];
            uw = Luw[i];
            const int mu = s.mid[uw.x], mw = s.mid[uw.y];
            if (mu >= ebase) {
                const int k = s.p.nptr[mu] + atomicAdd(&s.fill[mu], 1);
                s.p.nbr[k] = uw.y;
                s.p.eslot[k] = e;
            } else if (mw >= ebase) {
                const int k = s.p.nptr[mw] + atomicAdd(&s.fill[mw], 1);
                s.p.nbr[k] = uw.x;
                s.p.eslot[k] = e;
            } else {
                keep = mu == -1;           // -2: frozen network, its edges leave the list
            }
        }
        const int pos = block_append(&s.nsurv[r + 1], keep, s_cnt);
        if (keep) {
            Lout[pos] = e;
            Luwout[pos] = uw;
        }
    }
}

// neighbours of one pivot ascending, by the whole warp: rank sort (out of shared memory when the list fits)
// through the free tm / tslot ranges of the round (k_items_c fills them later); then lm, the neighbours'
// degrees and the pivot of every pair of the round (pm)
__device__ __forceinline__ void warp_pivot(const Sym &s, int m, int k0, int d, int t0, int npairs, int *sm /* SORT_SMEM */) {
    const int lane = threadIdx.x & 31;
    if (d > 1) {
        const bool in_smem = d <= SORT_SMEM;
        if (in_smem) {
            for (int e = lane; e < d; e += 32) sm[e] = s.p.nbr[k0 + e];
            __syncwarp();
        }
        for (int e = lane; e < d; e += 32) {
            int rank = 0;
            if (in_smem) {
                const int v = sm[e];
#pragma unroll 4
                for (int j = 0; j < d; j++) rank += sm[j] < v;
            } else {
                const int v = s.p.nbr[k0 + e];
                for (int j = 0; j < d; j++) rank += s.p.nbr[k0 + j] < v;
            }
            s.p.tm[k0 + rank] = in_smem ? sm[e] : s.p.nbr[k0 + e];
            s.p.tslot[k0 + rank] = s.p.eslot[k0 + e];
        }
        __syncwarp();
        for (int e = lane; e < d; e += 32) {
            s.p.nbr[k0 + e] = s.p.tm[k0 + e];
            s.p.eslot[k0 + e] = s.p.tslot[k0 + e];
        }
        __syncwarp();
    }
    for (int e = lane; e < d; e += 32) {
        s.lm[k0 + e] = m;
        atomicAdd(&s.pri[s.p.nbr[k0 + e]], 0ull - DEG1);
    }
    for (int t = lane; t < npairs; t += 32) s.pm[t0 + t] = m;
}

__global__ void __launch_bounds__(TB) k_pivots(Sym s, int r) {
    __shared__ int s_sort[TB / 32][SORT_SMEM];
    cnt_sm_info *info = s.p.info;
    if (info->done | info->error) return;
    const int nelim = info->rounds[r].nelim, ebase = info->rounds[r].ebase;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int SMALL = 8;
    const int nwarps = gridDim.x * (TB / 32);
    if (nelim <= nwarps) {
        // few pivots (the late rounds, long lists): one warp each
        const int ml = blockIdx.x * (TB / 32) + wid;
        if (ml >= nelim) return;
        const int m = ebase + ml, k0 = s.p.nptr[m], d = s.p.nptr[m + 1] - k0;
        warp_pivot(s, m, k0, d, s.pptr[ml], s.pptr[ml + 1] - s.pptr[ml], s_sort[wid]);
        if (lane == 0) {
            const int net = s.rnet[s.p.piv[m]];
            const long long h0 = s.hoff[net];
            s.phash[ml] = make_uint2((unsigned)h0, (unsigned)(s.hoff[net + 1] - h0));
            atomicSub(&s.nalive[net], 1);
            if (d > *(volatile int *)&info->max_deg) atomicMax(&info->max_deg, d);
        }
        return;
    }
    const int nceil = (nelim + 31) & ~31;
    int dmax_all = 0;
    for (int ml = blockIdx.x * TB + threadIdx.x; ml < nceil; ml += gridDim.x * TB) {
        int d = 0, net = -1, k0 = 0, t0 = 0, np = 0;
        const int m = ebase + ml;
        if (ml < nelim) {
            k0 = s.p.nptr[m];
            const int k1 = s.p.nptr[m + 1];
            d = k1 - k0;
            t0 = s.pptr[ml];
            np = s.pptr[ml + 1] - t0;
            if (d <= SMALL) {
                // neighbours ascending: the short list is loaded at once (independent loads), ordered in registers by a
                // sorting network and written back (the insertion sort in global memory was 38 % of the kernel's stalls)
                int nb[SMALL], es[SMALL];
#pragma unroll
                for (int a = 0; a < SMALL; a++) {
                    nb[a] = a < d ? s.p.nbr[k0 + a] : 0x7fffffff;
                    es[a] = a < d ? s.p.eslot[k0 + a] : 0;
                }
#define CNT_CE(i, j)                                   \
    if (nb[i] > nb[j]) {                               \
        int t_ = nb[i]; nb[i] = nb[j]; nb[j] = t_;     \
        t_ = es[i]; es[i] = es[j]; es[j] = t_;         \
    }
                if (d > 4) {
                    CNT_CE(0, 1) CNT_CE(2, 3) CNT_CE(4, 5) CNT_CE(6, 7) CNT_CE(0, 2) CNT_CE(1, 3) CNT_CE(4, 6) CNT_CE(5, 7)
                    CNT_CE(1, 2) CNT_CE(5, 6) CNT_CE(0, 4) CNT_CE(3, 7) CNT_CE(1, 5) CNT_CE(2, 6) CNT_CE(1, 4) CNT_CE(3, 6)
                    CNT_CE(2, 4) CNT_CE(3, 5) CNT_CE(3, 4)
                } else if (d > 1) {
                    CNT_CE(0, 1) CNT_CE(2, 3) CNT_CE(0, 2) CNT_CE(1, 3) CNT_CE(1, 2)
                }
#undef CNT_CE
#pragma unroll
                for (int a = 0; a < SMALL; a++) {
                    if (a < d) {
                        s.p.nbr[k0 + a] = nb[a];
                        s.p.eslot[k0 + a] = es[a];
                        s.lm[k0 + a] = m;
                        atomicAdd(&s.pri[nb[a]], 0ull - DEG1);
                    }
                }
                for (int t = 0; t < np; t++) s.pm[t0 + t] = m;
            }
            net = s.rnet[s.p.piv[m]];
            const long long h0 = s.hoff[net];
            s.phash[ml] = make_uint2((unsigned)h0, (unsigned)(s.hoff[net + 1] - h0));
        }
        unsigned big = __ballot_sync(~0u, d > SMALL);
        while (big) {                                         // longer lists: one at a time by the whole warp
            const int src = __ffs(big) - 1;
            big &= big - 1;
            warp_pivot(s, __shfl_sync(~0u, m, src), __shfl_sync(~0u, k0, src), __shfl_sync(~0u, d, src),
                       __shfl_sync(~0u, t0, src), __shfl_sync(~0u, np, src), s_sort[wid]);
        }
        // unknowns left per network: one atomic per run of equal networks inside the warp
        const unsigned same = __match_any_sync(~0u, net);
        if (net >= 0 && lane == __ffs(same) - 1) atomicSub(&s.nalive[net], __popc(same));
        dmax_all = max(dmax_all, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax_all = max(dmax_all, __shfl_xor_sync(~0u, dmax_all, o));
    if (lane == 0 && dmax_all > *(volatile int *)&info->max_deg) atomicMax(&info->max_deg, dmax_all);
}

// item t of a round: t < npairs = neighbour pair of a pivot, otherwise incident entry lbase + (t - npairs)
__global__ void __launch_bounds__(TB) k_items_a(Sym s, int r) {
    const cnt_sm_info *info = s.p.info;
    if (info->done | info->error) return;
    const cnt_sm_round *rd = &info->rounds[r];
    const int np = rd->npairs, nl = rd->nlist, ebase = rd->ebase, lbase = rd->lbase;
    for (int t = blockIdx.x * TB + threadIdx.x; t < np + nl; t += gridDim.x * TB) {
        s.gsz[t] = 0;
        if (t < np) {
            const int m = s.pm[t], k0 = s.p.nptr[m], d = s.p.nptr[m + 1] - k0;
            const uint2 hreg = s.phash[m - ebase];
            // pair q of the pivot -> (i, j), i < j, row-major: row i starts at i (2d - i - 1) / 2
            const int q = t - s.pptr[m - ebase];
            int i = (int)(((2.0 * d - 1.0) - sqrt((2.0 * d - 1.0) * (2.0 * d - 1.0) - 8.0 * (double)q)) * 0.5);
            i = max(0, min(i, d - 2));
            while (i > 0 && (long long)i * (2 * d - i - 1) / 2 > q) i--;
            while ((long long)(i + 1) * (2 * d - i - 2) / 2 <= q) i++;
            const int j = i + 1 + (q - (int)((long long)i * (2 * d - i - 1) / 2));
            const int u = s.p.nbr[k0 + i], w = s.p.nbr[k0 + j];
            const int e = hash_find_or_insert(s, (long long)hreg.x, hreg.y, ((unsigned long long)(unsigned)u << 32) | (unsigned)w);
            s.pent[t] = e;
            s.pa[t] = k0 + i;                 // the two dead edges of the pair by their position in the incident list:
            s.pb[t] = k0 + j;                 // the numeric phase keeps a copy of their values in list order (gl)
            s.pnext[t] = e >= 0 ? atomicExch(&s.htab[e].head, t) : -1;
        } else {
            const int o = s.p.nbr[lbase + (t - np)];
            s.pent[t] = o;
            s.pnext[t] = atomicExch(&s.nhead[o], t);
        }
    }
}

__global__ void __launch_bounds__(TB) k_items_b(Sym s, int r) {
    const cnt_sm_info *info = s.p.info;
    if (info->done | info->error) return;
    const cnt_sm_round *rd = &info->rounds[r];
    const int np = rd->npairs, nl = rd->nlist;
    for (int t = blockIdx.x * TB + threadIdx.x; t < np + nl; t += gridDim.x * TB) {
        const int e = s.pent[t];
        const int head = t < np ? s.htab[e].head : s.nhead[e];
        if (head != t) continue;                          // the last arrival speaks for the whole list
        int cnt = 0, mn = t;
        for (int x = t; x >= 0; x = s.pnext[x]) {
            cnt++;
            mn = min(mn, x);
        }
        const int isnew = t < np && s.htab[e].slot < 0;
        s.gsz[mn] = cnt | (isnew << 30);
        s.ghead[mn] = t;
    }
}

__global__ void __launch_bounds__(TB) k_items_scan(Sym s, int r, int epoch) {
    cnt_sm_info *info = s.p.info;
    if (info->done | info->error) return;
    cnt_sm_round *rd = &info->rounds[r];
    const int np = rd->npairs, nl = rd->nlist;
    auto code = [&](int t) { return s.gsz[t]; };
    auto park = [&](int, int) {};
    auto dec = [&](int z) { return z ? v4(1, z & ((1 << 30) - 1), z >> 30, 0) : v4(0, 0, 0, 0); };
    auto g = [&](int t, V4 ex, V4 own) {
        if (own.a) {
            s.gidx[t] = ex.a;
            s.goff[t] = ex.b;
            s.gnew[t] = ex.c;
        }
        if (t == np) rd->nge = ex.a;                      // owners among the pairs (pairs come first)
    };
    auto fin = [&](V4 t) {
        if (nl == 0) rd->nge = 0;
        rd->ngroups = t.a;                                // k_items_c splits it into nge + ngn
        rd->nnew = t.c;
    };
    chunked_scan(np + nl, epoch, s.tickets + epoch, s.sflag, s.sagg, &info->error, code, park, code, dec, g, fin);
}

__global__ void __launch_bounds__(TB) k_items_c(Sym s, int r) {
    cnt_sm_info *info = s.p.info;
    if (info->done | info->error) return;
    cnt_sm_round *rd = &info->rounds[r];
    const int np = rd->npairs, nl = rd->nlist, nge = rd->nge, ngn = rd->ngroups - rd->nge, nnew = rd->nnew;
    const int lbase = rd->lbase, cbase = rd->cbase, gebase = rd->gebase, gnbase = rd->gnbase, sbase = rd->sbase;
    const bool over = (long long)sbase + nnew > s.capS;
    if (blockIdx.x == 0 && threadIdx.x == 0) {            // books of the next round
        rd->ngn = ngn;
        if (over) {
            atomicOr(&info->error, CNT_SM_ERR_SLOTS);
        } else {
            cnt_sm_round *nx = &info->rounds[r + 1];
            s.nlive[r + 1] = s.nsurv[r + 1] + nnew;       // survivors first, then the fill edges in slot order
            nx->ebase = rd->ebase + rd->nelim;
            nx->lbase = lbase + nl;
            nx->cbase = cbase + np;
            nx->gebase = gebase + nge;
            nx->gnbase = gnbase + ngn;
            nx->sbase = sbase + nnew;
        }
    }
    if (over) return;
    const int n = np + nl, nsurv = s.nsurv[r + 1];
    for (int t = blockIdx.x * TB + threadIdx.x; t < n; t += gridDim.x * TB) {
        const int z = s.gsz[t];
        if (z) {
            const int cnt = z & ((1 << 30) - 1);
            if (t < np) {
                const int e = s.pent[t], off = cbase + s.goff[t], gi = gebase + s.gidx[t];
                if (cnt == 1) {
                    s.p.cm[off] = s.pm[t];
                    s.p.ca[off] = s.pa[t];
                    s.p.cb[off] = s.pb[t];
                } else {
                    int k = 0;
                    for (int x = s.ghead[t]; x >= 0; x = s.pnext[x]) {     // members ordered by item = by pivot
                        int c = k - 1;
                        while (c >= 0 && s.p.cm[off + c] > x) {
                            s.p.cm[off + c + 1] = s.p.cm[off + c];
                            c--;
                        }
                        s.p.cm[off + c + 1] = x;
                        k++;
                    }
                    for (k = 0; k < cnt; k++) {
                        const int x = s.p.cm[off + k];
                        s.p.cm[off + k] = s.pm[x];
                        s.p.ca[off + k] = s.pa[x];
                        s.p.cb[off + k] = s.pb[x];
                    }
                }
                HEnt *ent = s.htab + e;
                int slot = ent->slot;
                if (slot < 0) {                           // a fill edge: numbered in group order
                    const int rank = s.gnew[t];
                    slot = sbase + rank;
                    const unsigned long long key = ent->key;
                    const int2 uw = make_int2((int)(key >> 32), (int)(key & 0xFFFFFFFFull));
                    ent->slot = slot;
                    s.eu[slot] = uw.x;
                    s.ew[slot] = uw.y;
                    atomicAdd(&s.pri[uw.x], DEG1);
                    atomicAdd(&s.pri[uw.y], DEG1);
                    s.live[(r + 1) & 1][nsurv + rank] = slot;
                    s.luw[(r + 1) & 1][nsurv + rank] = uw;
                }
                ent->head = -1;
                s.p.tgt[gi] = slot;
                s.p.cptr[gi] = off;
            } else {
                const int o = s.pent[t], off = lbase + (s.goff[t] - np), gi = gnbase + (s.gidx[t] - nge);
                if (cnt == 1) {
                    const int kk = lbase + (t - np);
                    s.p.tm[off] = s.lm[kk];
                    s.p.tslot[off] = kk;
                } else {
                    int k = 0;
                    for (int x = s.ghead[t]; x >= 0; x = s.pnext[x]) {
                        int c = k - 1;
                        while (c >= 0 && s.p.tm[off + c] > x) {
                            s.p.tm[off + c + 1] = s.p.tm[off + c];
                            c--;
                        }
                        s.p.tm[off + c + 1] = x;
                        k++;
                    }
                    for (k = 0; k < cnt; k++) {
                        const int kk = lbase + (s.p.tm[off + k] - np);
                        s.p.tm[off + k] = s.lm[kk];
                        s.p.tslot[off + k] = kk;
                    }
                }
                s.nhead[o] = -1;
                s.p.tn[gi] = o;
                s.p.tptr[gi] = off;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// after the last round: totals, sentinels, and the dense tail (rows and edges left per network)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_final_a(Sym s) {
    cnt_sm_info *info = s.p.info;
    if (!info->done || info->error) return;
    __shared__ long long s_n[1024], s_q[1024];
    __shared__ int s_max[1024];
    const int t = threadIdx.x;
    const int per = (s.B + 1023) / 1024;
    const int lo = min(t * per, s.B), hi = min(lo + per, s.B);
    long long sn = 0, sq = 0;
    int mx = 0;
    for (int b = lo; b < hi; b++) {
        const int n = s.nalive[b];
        sn += n;
        sq += (long long)n * n;
        mx = max(mx, n);
    }
    s_n[t] = sn;
    s_q[t] = sq;
    s_max[t] = mx;
    __syncthreads();
    if (t == 0) {
        long long an = 0, aq = 0;
        int m = 0;
        for (int k = 0; k < 1024; k++) {
            const long long cn = s_n[k], cq = s_q[k];
            s_n[k] = an;
            s_q[k] = aq;
            an += cn;
            aq += cq;
            m = max(m, s_max[k]);
        }
        const cnt_sm_round *rd = &info->rounds[info->nrounds];
        info->nelim = rd->ebase;
        info->nlist = rd->lbase;
        info->ncontrib = rd->cbase;
        info->nge = rd->gebase;
        info->ngn = rd->gnbase;
        info->nslots = rd->sbase;
        s.p.cptr[rd->gebase] = rd->cbase;
        s.p.tptr[rd->gnbase] = rd->lbase;
        info->dense_nnet = s.B;
        info->dense_nmax = m;
        info->dense_nnodes = (int)an;
        info->dense_gpool = aq;
        s.p.dptr[s.B] = (int)an;
        s.p.doff[s.B] = aq;
    }
    __syncthreads();
    sn = s_n[t];
    sq = s_q[t];
    for (int b = lo; b < hi; b++) {
        const int n = s.nalive[b];
        s.p.dptr[b] = (int)sn;
        s.p.doff[b] = sq;
        sn += n;
        sq += (long long)n * n;
    }
}

__global__ void __launch_bounds__(TB) k_final_nodes(Sym s, int epoch) {
    cnt_sm_info *info = s.p.info;
    if (!info->done || info->error) return;
    auto f = [&](int v) { return v4(s.mid[v] < 0 ? 1 : 0, 0, 0, 0); };
    auto g = [&](int v, V4 ex, V4 own) {
        if (own.a) {
            s.p.dnodes[ex.a] = v;
            s.loc[v] = ex.a - s.p.dptr[s.rnet[v]];
        }
    };
    auto fin = [&](V4) {};
    chained_scan(s.R, epoch, s.tickets + epoch, s.sflag, s.sagg, s.sinc, &info->error, f, g, fin);
}

// remaining edges grouped by network: count, scan (k_final_b), fill
template <bool FILL>
__global__ void __launch_bounds__(TB) k_final_edges(Sym s) {
    cnt_sm_info *info = s.p.info;
    if (!info->done || info->error) return;
    const int n = info->nslots, nceil = (n + 31) & ~31;
    for (int e = blockIdx.x * TB + threadIdx.x; e < nceil; e += gridDim.x * TB) {
        int net = -1, u = 0, w = 0;
        if (e < n) {
            u = s.eu[e];
            w = s.ew[e];
            if (s.mid[u] < 0 && s.mid[w] < 0) net = s.rnet[u];
        }
        const unsigned same = __match_any_sync(~0u, net);
        if (net < 0) continue;
        const int lane = threadIdx.x & 31, leader = __ffs(same) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&s.dcount[net], __popc(same));
        base = __shfl_sync(same, base, leader);
        if (FILL) {
            const int pos = s.p.dedge[net] + base + __popc(same & ((1u << lane) - 1u));
            const long long nn = s.p.dptr[net + 1] - s.p.dptr[net];
            s.p.dslot[pos] = e;
            s.p.depos[pos] = s.p.doff[net] + (long long)s.loc[u] * nn + s.loc[w];
        }
    }
}
__global__ void __launch_bounds__(1024) k_final_b(Sym s) {
    cnt_sm_info *info = s.p.info;
    if (!info->done || info->error) return;
    __shared__ int s_n[1024];
    const int t = threadIdx.x;
    const int per = (s.B + 1023) / 1024;
    const int lo = min(t * per, s.B), hi = min(lo + per, s.B);
    int sn = 0;
    for (int b = lo; b < hi; b++) sn += s.dcount[b];
    s_n[t] = sn;
    __syncthreads();
    if (t == 0) {
        int acc = 0;
        for (int k = 0; k < 1024; k++) {
            const int c = s_n[k];
            s_n[k] = acc;
            acc += c;
        }
        info->dense_nE = acc;
        s.p.dedge[s.B] = acc;
    }
    __syncthreads();
    sn = s_n[t];
    for (int b = lo; b < hi; b++) {
        s.p.dedge[b] = sn;
        sn += s.dcount[b];
        s.dcount[b] = 0;
    }
}

struct WsLayout {
    int64_t eu, ew, lm, live0, live1, luw0, luw1, pri, hoff, rnet, loc, nodes0, nodes1, nalive, pptr, phash, items, sagg, sinc, htab;
    int64_t ones0, ones1;      // [mid, nhead]: memset 0xFF
    int64_t zero0, zero1;      // [blocked, fill, nlive, tickets, sflag]: memset 0
    int64_t mid, nhead, blocked, fill, nlive, nsurv, tickets, sflag, dcount;
    int64_t total;
    unsigned hcap;
    int64_t ntiles;
};

WsLayout ws_layout(int64_t R, int B, int64_t capS, int64_t capI) {
    WsLayout L;
    int64_t pos = 0;
    auto take = [&](int64_t count, int64_t elem) {
        int64_t at = pos;
        pos += cnt::align_up((count > 0 ? count : 1) * elem, 256);
        return at;
    };
    int64_t nmax = R > capS ? R : capS;
    if (capI > nmax) nmax = capI;
    L.ntiles = nmax / TILE + 2;
    L.eu = take(capS, 4);
    L.ew = take(capS, 4);
    L.lm = take(capS, 4);
    L.live0 = take(capS, 4);
    L.live1 = take(capS, 4);
    L.luw0 = take(capS, 8);
    L.luw1 = take(capS, 8);
    L.pri = take(R, 8);
    L.hoff = take((int64_t)B + 1, 8);
    L.rnet = take(R, 4);
    L.loc = take(R, 4);
    L.nodes0 = take(R, 4);
    L.nodes1 = take(R, 4);
    L.nalive = take(B, 4);
    L.pptr = take(R + 1, 4);
    L.phash = take(R, 8);
    L.items = take(10 * capI, 4);
    L.sagg = take(L.ntiles, 16);
    L.sinc = take(L.ntiles, 16);
    L.ones0 = pos;
    L.mid = take(R, 4);
    L.nhead = take(R, 4);
    L.ones1 = pos;
    L.zero0 = pos;
    L.blocked = take(R, 4);
    L.fill = take(R, 4);
    L.nlive = take(CNT_SM_MAX_ROUNDS + 2, 4);
    L.nsurv = take(CNT_SM_MAX_ROUNDS + 2, 4);
    L.tickets = take(2 * CNT_SM_MAX_ROUNDS + 8, 4);
    L.sflag = take(L.ntiles, 4);
    L.dcount = take(B, 4);
    L.zero1 = pos;
    int64_t hc = 2 * capS + 64 * ((int64_t)B + 1);
    L.hcap = (unsigned)hc;
    L.htab = take(hc, (int64_t)sizeof(HEnt));
    L.total = pos;
    return L;
}
}  // namespace

extern "C" int64_t cnt_sm_plan_layout(int64_t R, int B, int64_t NNZ, int64_t cap_slots, int64_t cap_contrib,
                                      int64_t *offsets) {
    int64_t off[CNT_SM_NARR + 1];
    int64_t total = cnt::sm::plan_layout(R, B, NNZ, cap_slots, cap_contrib, off);
    if (offsets)
        for (int k = 0; k <= CNT_SM_NARR; k++) offsets[k] = off[k];
    return total;
}

extern "C" int64_t cnt_sm_symbolic_workspace_bytes(int64_t R, int B, int64_t NNZ, int64_t cap_slots, int64_t cap_contrib,
                                                   int64_t cap_items) {
    (void)NNZ;
    (void)cap_contrib;
    return ws_layout(R, B, cap_slots, cap_items).total;
}

extern "C" int cnt_sm_symbolic(int B, const int32_t *roff, int64_t R, int64_t NNZ, const int32_t *rowptr,
                               const int32_t *col, int dense_max, int rounds_hint, int64_t cap_slots,
                               int64_t cap_contrib, int64_t cap_items, const cnt_sm_info *hint, void *plan,
                               int64_t plan_bytes, void *ws, int64_t ws_bytes, cnt_sm_info *h_info, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CNT_CHECK_ARG(B > 0 && roff && R > 0 && R < (1 << 26) && NNZ >= 0 && rowptr && (col || NNZ == 0) && plan && ws && h_info);
    CNT_CHECK_ARG(dense_max >= 0 && cap_slots >= NNZ / 2 + 1 && cap_slots < (1 << 29) && cap_contrib > 0 &&
                  cap_contrib < (1 << 29) && cap_items > 0 && cap_items < (1 << 29) && NNZ < (1ll << 31));
    int64_t off[CNT_SM_NARR + 1];
    if (cnt::sm::plan_layout(R, B, NNZ, cap_slots, cap_contrib, off) > plan_bytes) {
        cnt::set_error("star-mesh plan buffer too small");
        return CNT_ECAPACITY;
    }
    const WsLayout L = ws_layout(R, B, cap_slots, cap_items);
    if (L.total > ws_bytes) {
        cnt::set_error("star-mesh symbolic workspace too small");
        return CNT_ECAPACITY;
    }
    char *w = (char *)ws;
    Sym s;
    s.B = B; s.R = (int)R; s.NNZ = (int)NNZ; s.dense_max = dense_max;
    s.capS = (int)cap_slots; s.capC = (int)cap_contrib; s.capI = (int)cap_items;
    s.hcap = L.hcap;
    s.roff = roff; s.rowptr = rowptr; s.col = col;
    s.p = cnt::sm::plan_view(plan, off);
    s.eu = (int32_t *)(w + L.eu); s.ew = (int32_t *)(w + L.ew); s.lm = (int32_t *)(w + L.lm);
    s.live[0] = (int32_t *)(w + L.live0); s.live[1] = (int32_t *)(w + L.live1);
    s.pri = (unsigned long long *)(w + L.pri); s.hoff = (long long *)(w + L.hoff);
    s.luw[0] = (int2 *)(w + L.luw0); s.luw[1] = (int2 *)(w + L.luw1);
    s.rnet = (int32_t *)(w + L.rnet); s.loc = (int32_t *)(w + L.loc);
    s.nodes[0] = (int32_t *)(w + L.nodes0); s.nodes[1] = (int32_t *)(w + L.nodes1);
    s.nalive = (int32_t *)(w + L.nalive); s.pptr = (int32_t *)(w + L.pptr);
    s.phash = (uint2 *)(w + L.phash);
    int32_t *items = (int32_t *)(w + L.items);
    s.pent = items; s.pnext = items + cap_items; s.gsz = items + 2 * cap_items; s.ghead = items + 3 * cap_items;
    s.pm = items + 4 * cap_items; s.pa = items + 5 * cap_items; s.pb = items + 6 * cap_items;
    s.gidx = items + 7 * cap_items; s.goff = items + 8 * cap_items; s.gnew = items + 9 * cap_items;
    s.sagg = (V4 *)(w + L.sagg); s.sinc = (V4 *)(w + L.sinc);
    s.mid = (int32_t *)(w + L.mid); s.nhead = (int32_t *)(w + L.nhead);
    s.blocked = (int32_t *)(w + L.blocked); s.fill = (int32_t *)(w + L.fill); s.nlive = (int32_t *)(w + L.nlive); s.nsurv = (int32_t *)(w + L.nsurv);
    s.tickets = (int32_t *)(w + L.tickets); s.sflag = (int32_t *)(w + L.sflag); s.dcount = (int32_t *)(w + L.dcount);
    s.htab = (HEnt *)(w + L.htab);

    CNT_CUDA(cudaMemsetAsync(w + L.ones0, 0xFF, L.ones1 - L.ones0, st));
    CNT_CUDA(cudaMemsetAsync(w + L.zero0, 0, L.zero1 - L.zero0, st));
    CNT_CUDA(cudaMemsetAsync(s.htab, 0xFF, (size_t)L.hcap * sizeof(HEnt), st));
    CNT_CUDA(cudaMemsetAsync(s.p.info, 0, sizeof(cnt_sm_info), st));

    // epochs: 1 = init, 2 + 2r = k_select of round r, 3 + 2r = k_items_scan of round r, then the final scan;
    // ticket counter [epoch] belongs to that invocation
    // CNT_SM_TIMING=1: CUDA-event time of every round kernel, printed to stderr (debugging aid)
    static const bool timing = getenv("CNT_SM_TIMING") != nullptr;
    constexpr int MAXEV = 9 * 128;
    static cudaEvent_t ev[MAXEV];
    static bool ev_made = false;
    int nev = 0;
    if (timing && !ev_made) {
        for (int k = 0; k < MAXEV; k++) cudaEventCreate(&ev[k]);
        ev_made = true;
    }
    static cudaEvent_t evi[4];
    if (timing && !evi[0])
        for (int k = 0; k < 4; k++) cudaEventCreate(&evi[k]);
    if (timing) cudaEventRecord(evi[0], st);
    k_hoff<<<1, 1024, 0, st>>>(s);
    CNT_LAUNCHED();
    k_init<<<GRID, TB, 0, st>>>(s, 1);
    CNT_LAUNCHED();
    if (timing) cudaEventRecord(evi[1], st);
    k_init_hash<<<GRID, TB, 0, st>>>(s);
    CNT_LAUNCHED();
    if (timing) cudaEventRecord(evi[2], st);
    int r = 0;
    int chunk = rounds_hint > 0 ? rounds_hint : 40;
    for (;;) {
        const int rend = r + chunk < CNT_SM_MAX_ROUNDS - 1 ? r + chunk : CNT_SM_MAX_ROUNDS - 1;
        for (; r < rend; r++) {
            // grids: sized from the rounds of the previous batch of the same kind when the caller passes its header
            // (any size is correct -- the kernels stride over device-side counts --, small grids just start faster)
            int gl = GRID, gn = GRID, gp = GRID, gi = GRID;
            if (hint && hint->done && !hint->error && hint->nrounds > 0) {
                const cnt_sm_round *hr = &hint->rounds[r < hint->nrounds ? r : hint->nrounds - 1];
                const double scale = 1.5 * (double)R / (hint->R > 0 ? hint->R : 1);
                auto fit = [&](double n, int per) {
                    double g = scale * n / per + 4;
                    return (int)(g < GRID ? g : GRID);
                };
                gl = fit(hr->nlive, 2 * TB);
                gn = fit(hr->nnodes, TILE);
                gp = fit(hr->nelim, TB);
                if (1.5 * scale * hr->nelim <= GRID * (TB / 32)) gp = fit(hr->nelim, TB / 32);      // a warp per pivot
                gi = fit((double)hr->npairs + hr->nlist, 2 * TB);
            }
            // the two scans: one chunk per CTA, so no more CTAs than fit on the device at once
            if (gn > GRID_SCAN) gn = GRID_SCAN;
            const int gs = gi < gn ? gn : (gi < GRID_SCAN ? gi : GRID_SCAN);
            if (timing && nev + 9 <= MAXEV) cudaEventRecord(ev[nev++], st);
#define CNT_SM_TICK() do { if (timing && nev < MAXEV) cudaEventRecord(ev[nev++], st); } while (0)
            k_mark<<<gl, TB, 0, st>>>(s, r);
            CNT_SM_TICK();
            k_select<<<gn, TB, 0, st>>>(s, r, 2 + 2 * r);
            CNT_SM_TICK();
            k_incident<<<gl, TB, 0, st>>>(s, r);
            CNT_SM_TICK();
            k_pivots<<<gp, TB, 0, st>>>(s, r);
            CNT_SM_TICK();
            k_items_a<<<gi, TB, 0, st>>>(s, r);
            CNT_SM_TICK();
            k_items_b<<<gi, TB, 0, st>>>(s, r);
            CNT_SM_TICK();
            k_items_scan<<<gs, TB, 0, st>>>(s, r, 3 + 2 * r);
            CNT_SM_TICK();
            k_items_c<<<gi, TB, 0, st>>>(s, r);
            CNT_SM_TICK();
            cnt::count_launch(8);
        }
        CNT_CUDA(cudaGetLastError());
        // the finals run only once the rounds are over (they check the flag themselves)
        k_final_a<<<1, 1024, 0, st>>>(s);
        k_final_nodes<<<GRID, TB, 0, st>>>(s, 2 * CNT_SM_MAX_ROUNDS + 4);
        k_final_edges<false><<<GRID, TB, 0, st>>>(s);
        k_final_b<<<1, 1024, 0, st>>>(s);
        k_final_edges<true><<<GRID, TB, 0, st>>>(s);
        cnt::count_launch(5);
        CNT_CUDA(cudaGetLastError());
        CNT_CUDA(cudaMemcpyAsync(h_info, s.p.info, sizeof(cnt_sm_info), cudaMemcpyDeviceToHost, st));
        CNT_CUDA(cudaStreamSynchronize(st));
        if (h_info->done || h_info->error || r >= CNT_SM_MAX_ROUNDS - 1) break;
        // not finished: the final scan did not run (no ticket taken, no flag written), go on
        chunk = 16;
    }
    if (timing) {
        float t01 = 0, t12 = 0;
        cudaEventElapsedTime(&t01, evi[0], evi[1]);
        cudaEventElapsedTime(&t12, evi[1], evi[2]);
        fprintf(stderr, "sm init: k_hoff + k_init %.0f us, k_init_hash %.0f us (B %d R %d nE0 %d rounds %d)\n", t01 * 1e3, t12 * 1e3, B,
                (int)R, h_info->nE0, h_info->nrounds);
        static const char *names[8] = {"mark", "select", "incident", "pivots", "items_a", "items_b", "scan", "items_c"};
        double tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k + 8 < nev; k += 9) {
            fprintf(stderr, "sm round %3d:", k / 9);
            double sum = 0;
            for (int j = 0; j < 8; j++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, ev[k + j], ev[k + j + 1]);
                fprintf(stderr, " %s=%.0f", names[j], ms * 1e3);
                tot[j] += ms;
                sum += ms;
            }
            const cnt_sm_round *rd = &h_info->rounds[k / 9];
            fprintf(stderr, "  sum=%.0f us  (nodes %d live %d piv %d pairs %d list %d)\n", sum * 1e3, rd->nnodes, rd->nlive,
                    rd->nelim, rd->npairs, rd->nlist);
        }
        fprintf(stderr, "sm totals (ms):");
        for (int j = 0; j < 8; j++) fprintf(stderr, " %s=%.3f", names[j], tot[j]);
        fprintf(stderr, "\n");
    }
    if (h_info->error || !h_info->done) {
        cnt::set_error("star-mesh symbolic phase: capacity exceeded (error bits %d, rounds %d)", h_info->error,
                       h_info->nrounds);
        if (!h_info->error) h_info->error = CNT_SM_ERR_ROUNDS;
        return CNT_ECAPACITY;
    }
    return CNT_OK;
}
