/* cntnet_b200 -- C-ABI of the B200-native stick-network hot path.
 *
 * The reference (leobrowning92/networksim-cntfet) is pure Python and has no FFI layer; its
 * seam is the method surface of netsim.py / cnet.py / measure_perc.py.  Each
 * entry point below replaces the *body* of one or more of those methods and is what a
 * ctypes binding inside the reference would call (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer into caller-owned memory (PyTorch CUDA tensors in
 *    the shipped host layer) unless its name starts with h_;
 *  - a batch holds B independent networks; per-stick arrays are concatenated and indexed
 *    through soff[B+1] (int32 prefix offsets), per-junction arrays through joff[B+1];
 *    stick indices stored in junction tables are LOCAL to their network, exactly the
 *    reference's post-sort DataFrame index (netsim.py:133-134);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *    synchronises unless documented; the library allocates no device memory: scratch comes
 *    from caller-provided workspaces sized by the *_workspace_bytes queries;
 *  - every function returns an int status (CNT_OK == 0); nothing throws;
 *    cnt_last_error() gives a thread-local message for the last non-zero status;
 *  - "not percolating" is a normal result (netsim.py:208-218, measure_perc.py:193-197),
 *    reported per network, never an error.
 *  - stick kinds: 0 = 'v' (electrode), 1 = 'm', 2 = 's'; junction kind2 = 3*kind[stick1] +
 *    kind[stick2] (netsim.py:148 `kinds[i]+kinds[j]`).
 */
#ifndef CNTNET_B200_H
#define CNTNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNT_OK 0
#define CNT_EINVAL 1   /* bad argument                                   */
#define CNT_ECUDA 2    /* a CUDA runtime call or launch failed           */
#define CNT_ECAPACITY 3 /* caller-provided buffer / workspace too small   */

#define CNT_ABI_VERSION 2

int cnt_abi_version(void);
const char *cnt_last_error(void);
/* number of kernels this library has launched in this process (bench.py `gpu_launches`). */
int64_t cnt_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * utilities
 * ------------------------------------------------------------------------------------- */
/* exclusive prefix sum of n int32 values; out may alias in; out[n] receives the total when
 * write_total != 0 (out must then hold n+1 values).  ws: cnt_scan_workspace_bytes(n). */
int64_t cnt_scan_workspace_bytes(int64_t n);
int cnt_exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, int write_total,
                           void *ws, int64_t ws_bytes, void *stream);

/* FP64 FMA peak of the device (TFLOP/s, best of `reps` launches of a register-resident DFMA
 * kernel): roofline denominator of the junction kernel.  scratch: >= 606208 doubles.  Synchronises. */
int cnt_fp64_peak(double *scratch, int reps, double *h_tflops, void *stream);

/* ---------------------------------------------------------------------------------------
 * (1) stick placement -- replaces make_stick / get_ends / make_sticks (netsim.py:97-125)
 *     and the length sort of make_intersects_kdtree (netsim.py:132-134).
 * Counter-based Philox4x32-10 keyed by the network's own 64-bit seed seeds[b] (the reference
 * also hands one seed per network, measure_perc.py:226-233), counter = stick ordinal; law:
 * kind 'm' iff u<=pm, xc,yc~U(0,1), angle~U(0,2pi), length=|N(0.66,0.44)|/scaling[b]
 * (len_mode 0) or len_const/scaling[b] (len_mode 1).  Each network gets soff[b+1]-soff[b]
 * sticks INCLUDING the two electrodes [0.01|0.99, 0.5, pi/2-1e-6, 100, 'v'].
 * Output order is canonical: source electrode (x=0.01) at local index 0, drain (x=0.99) at
 * 1, then the body sticks by length descending, ties by generation ordinal; orig[] is the
 * reference's pre-sort DataFrame index (netsim.py:132: source 0, body 1..n, drain n+1).
 * ------------------------------------------------------------------------------------- */
int64_t cnt_gen_sticks_workspace_bytes(int64_t n_sticks_total, int B);
int cnt_gen_sticks(const uint64_t *seeds, int B, const int32_t *soff, const int32_t *h_soff,
                   const double *scaling, double pm, int len_mode, double len_const,
                   double *xc, double *yc, double *angle, double *length, uint8_t *kind,
                   double *xa, double *ya, double *xb, double *yb, int32_t *orig,
                   void *ws, int64_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * (2) junction finder -- replaces check_intersect (netsim.py:75-94) and the pair loop of
 *     make_intersects_kdtree (netsim.py:135-150).
 * A pair (i<j, local post-sort indices) is TESTED iff the squared centre distance is
 * <= length[i]^2 (the KD-tree ball of netsim.py:142) and is a JUNCTION iff the FP64
 * slope/intercept predicate holds and 0<=x<=1, 0<=y<=1.  IEEE round-to-nearest, no FMA
 * contraction: junction sets and coordinates are bit-identical to the reference.
 * Two calls: _count fills jcount[S] (junctions whose stick1 is that stick) and
 * ntests[B]; the caller scans jcount into jstart[S+1], allocates J = jstart[S] rows and
 * calls _fill (SAME workspace, untouched in between: the binned records are reused), which
 * writes rows sorted by (network, stick1, stick2) and joff[B+1].  jcount holds S+1 ints.
 * h_soff is the host copy of soff (grid geometry is derived from it on the host).
 * ------------------------------------------------------------------------------------- */
int64_t cnt_junction_workspace_bytes(int B, const int32_t *h_soff);
int cnt_junction_count(int B, const int32_t *soff, const int32_t *h_soff,
                       const double *xa, const double *ya, const double *xb, const double *yb,
                       const double *xc, const double *yc, const double *length,
                       int32_t *jcount, unsigned long long *ntests,
                       void *ws, int64_t ws_bytes, void *stream);
int cnt_junction_fill(int B, const int32_t *soff, const int32_t *h_soff,
                      const double *xa, const double *ya, const double *xb, const double *yb,
                      const double *xc, const double *yc, const double *length, const uint8_t *kind,
                      const int32_t *jstart, int64_t J,
                      int32_t *stick1, int32_t *stick2, double *jx, double *jy, uint8_t *kind2,
                      int32_t *joff, void *ws, int64_t ws_bytes, void *stream);

/* The same search sharded by spatial slab over `shard_world` ranks (one process per GPU; BASELINE config 5, a single
 * huge network -- the reference cannot spread one make_intersects_kdtree over workers).  The unit box is cut into
 * vertical slabs of whole tile columns (8 cells); rank r evaluates the home sticks whose centre lies in slab r
 * against the sticks around them (the halo across the slab boundary is read from the stick table, which every
 * rank holds), rank 0 also the long sticks (electrodes).  _count_slab leaves 0 in jcount for the rows of other
 * ranks: the caller ADDS jcount and ntests over the ranks (all-reduce), scans, and calls _fill_slab on zero-filled
 * tables; the tables of the ranks have disjoint supports, so adding them (all-reduce; the doubles as int64 bit
 * patterns) gives bit for bit the table of the unsharded search.  shard_world = 1 is cnt_junction_count / _fill. */
int cnt_junction_count_slab(int B, const int32_t *soff, const int32_t *h_soff,
                            const double *xa, const double *ya, const double *xb, const double *yb,
                            const double *xc, const double *yc, const double *length,
                            int32_t *jcount, unsigned long long *ntests,
                            void *ws, int64_t ws_bytes, int shard_rank, int shard_world, void *stream);
int cnt_junction_fill_slab(int B, const int32_t *soff, const int32_t *h_soff,
                           const double *xa, const double *ya, const double *xb, const double *yb,
                           const double *xc, const double *yc, const double *length, const uint8_t *kind,
                           const int32_t *jstart, int64_t J, int32_t *stick1, int32_t *stick2, double *jx,
                           double *jy, uint8_t *kind2, int32_t *joff, void *ws, int64_t ws_bytes,
                           int shard_rank, int shard_world, void *stream);
/* device pointer (inside ws) of the sortedness flag written by the last _count: non-zero if
 * some network's table was not sorted by length descending (netsim.py:133 invariant). */
int cnt_junction_errflag_ptr(int B, const int32_t *h_soff, void *ws, int64_t ws_bytes, int32_t **out);

/* ---------------------------------------------------------------------------------------
 * (3) clusters -- replaces make_graph's component search (netsim.py:175-179), label_clusters
 *     (netsim.py:197-206) and the nclust / maxclust lines of measure_perc.py:76-80,148-152.
 * label[s] = smallest local stick index of the connected component (isolated sticks label
 * themselves).  stats[b*CNT_NSTAT + k]:
 * ------------------------------------------------------------------------------------- */
#define CNT_STAT_PERCOLATING 0  /* sticks 0 and 1 in one component (netsim.py:177-178)       */
#define CNT_STAT_NCLUST_TRUE 1  /* number of distinct labels                                 */
#define CNT_STAT_MAXCLUST_TRUE 2 /* largest component                                        */
#define CNT_STAT_NCLUST_REF 3   /* len(sticks.cluster.drop_duplicates()) (measure_perc.py:76) */
#define CNT_STAT_MAXCLUST_REF 4 /* len(max(nx.connected_components(g))) (measure_perc.py:78)  */
#define CNT_STAT_NCOMP 5        /* sticks in the percolating component (0 if none)           */
#define CNT_STAT_ECOMP 6        /* junctions in the percolating component                    */
#define CNT_NSTAT 8
int64_t cnt_cluster_workspace_bytes(int64_t n_sticks_total);
int cnt_label_clusters(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                       const int32_t *stick1, const int32_t *stick2, const int32_t *orig /* may be NULL */,
                       int32_t *label, int32_t *stats, void *ws, int64_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * (3b) optional exact reduction before assembly: dangling sticks (left with one junction once
 *     their own dangling neighbours are removed) carry no current, so they are peeled off the
 *     percolating component until its 2-core + the two electrodes remain.  The reference keeps
 *     them in the matrix (cnet.py:104-131); the solution on the remaining sticks is the same and
 *     a peeled stick takes the voltage of the stick it hangs on (cnt_currents, attach[]).
 *     alive[S] u8: 1 = stays an MNA node (electrodes always); attach[S]: global stick index of
 *     the attachment point of a peeled stick, -1 otherwise.  h_rounds (host, optional): peeling
 *     rounds run.  Synchronises the stream.
 * ------------------------------------------------------------------------------------- */
int64_t cnt_prune_workspace_bytes(int64_t n_sticks_total);
int cnt_prune_leaves(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                     const int32_t *stick1, const int32_t *stick2, const int32_t *label, const int32_t *stats,
                     uint8_t *alive, int32_t *attach, int32_t *h_rounds, void *ws, int64_t ws_bytes, void *stream);

/* (3c) optional second exact reduction: series elimination.  A stick of the 2-core with exactly
 *     two junctions, both to ordinary sticks, is replaced by ONE virtual junction between its two
 *     neighbours (conductance g1 g2 / (g1 + g2), evaluated per gate configuration by
 *     cnt_assemble_values); afterwards V = (g1 V_a + g2 V_c) / (g1 + g2) (cnt_currents).  Only an
 *     independent set is eliminated (no chains), chosen from the topology alone.
 *     elim_vid[S]: virtual junction of an eliminated stick, -1 otherwise; voff[B+1]: virtual
 *     junctions per network; va, vc [capacity S]: the neighbours (network-local, va < vc);
 *     ve1, ve2: the stick's junctions to va / vc (rows of the junction table).  Entries of virtual
 *     junction v carry the junction id J + v in nz_eid.  Synchronises the stream. */
int64_t cnt_series_workspace_bytes(int64_t n_sticks_total);
int cnt_series_select(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                      const int32_t *stick1, const int32_t *stick2, const uint8_t *alive,
                      int32_t *elim_vid, int32_t *voff, int32_t *va, int32_t *vc, int32_t *ve1, int32_t *ve2,
                      int64_t *h_nv, void *ws, int64_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * (4) MNA assembly -- replaces populate_graph (netsim.py:192-195), update_conductivity,
 *     make_G, make_A, make_z (cnet.py:99-131) in the reduced SPD form
 *         L_UU V_U = 0.1 * g_{U,source}
 *     (nodes ascending by stick index, source = stick 0 at 0.1 V, ground = stick 1).
 * _pattern builds, for every percolating network of the batch, the CSR pattern of L_UU
 * without its diagonal: rowptr[R+1], col[NNZ] (column = row index LOCAL to the network),
 * nz_eid[NNZ] (global junction row that owns the entry), per-row src_eid / drn_eid (the
 * junction to stick 0 / stick 1, -1 if none), urow[S] (global row of a stick, -1 if not an
 * unknown), roff[B+1] (row offsets per network; empty range if not percolating).
 * Two-phase like the junction finder: _pattern_count fills rowcnt[R]; caller scans.
 * _values evaluates one gate configuration: g[e] = gtable[kind2[e]*nlevels + level[e]]
 * (table computed on the host with the reference's own formula so conductances are
 * bit-identical), val = -g[nz_eid], diag = row sum incl. electrode contacts, rhs = 0.1*g_src;
 * cls[k] = kind2*nlevels + level of the entry's junction (so val[k] == -gtable[cls[k]]; needs
 * 9*nlevels <= 256), the compressed form the cluster PCG kernel keeps in shared memory.
 * gtable == NULL: g[] is an input (conductances evaluated by the caller), cls must be NULL.
 * ------------------------------------------------------------------------------------- */
int64_t cnt_assemble_workspace_bytes(int64_t n_sticks_total);
int cnt_assemble_rows(int B, const int32_t *soff, int64_t S, const int32_t *label, const int32_t *stats,
                      const uint8_t *alive /* from cnt_prune_leaves, or NULL = whole component */,
                      const int32_t *elim_vid /* from cnt_series_select, or NULL */,
                      int32_t *urow, int32_t *roff, void *ws, int64_t ws_bytes, void *stream);
/* optional, between _rows and _pattern_count: renumber every network's unknown rows in Z-order
 * (Morton code) of the stick centres so that neighbouring sticks get neighbouring rows (SpMV
 * gather locality; lets the cluster PCG kernel keep most gathers in shared memory).  Only
 * urow changes; roff stays.  The reference's ascending node order (cnet.py:137-139) is a
 * reporting convention that the write-back reproduces per stick. */
int64_t cnt_assemble_reorder_workspace_bytes(int64_t R, int B);
int cnt_assemble_reorder(int B, const int32_t *h_roff, int64_t S, int64_t R, const double *xc, const double *yc,
                         int32_t *urow, void *ws, int64_t ws_bytes, void *stream);
int cnt_assemble_pattern_count(int B, const int32_t *soff, const int32_t *joff, int64_t J,
                               const int32_t *stick1, const int32_t *stick2,
                               const int32_t *urow, int64_t R, int32_t *rowcnt,
                               int64_t NV, const int32_t *voff, const int32_t *va, const int32_t *vc /* virtual junctions or 0/NULL */,
                               void *stream);
int cnt_assemble_pattern_fill(int B, const int32_t *soff, const int32_t *joff, int64_t J,
                              const int32_t *stick1, const int32_t *stick2,
                              const int32_t *urow, const int32_t *roff, int64_t R, const int32_t *rowptr,
                              int32_t *col, int32_t *nz_eid, int32_t *src_eid, int32_t *drn_eid,
                              int64_t NV, const int32_t *voff, const int32_t *va, const int32_t *vc,
                              void *ws, int64_t ws_bytes, void *stream);
/* row_stick[r] = local stick index of unknown row r (inverse of urow). */
int cnt_assemble_row_stick(int B, const int32_t *soff, int64_t S, const int32_t *urow,
                           int32_t *row_stick, void *stream);
int cnt_gate_levels(int64_t J, const double *jx, const double *jy, uint8_t *level,
                    int global_set, uint8_t global_level,
                    int rect_set, double cx, double cy, double w, double h, uint8_t rect_level, void *stream);
int cnt_assemble_values(int64_t J, const uint8_t *kind2, const uint8_t *level, const double *gtable, int nlevels,
                        int64_t R, const int32_t *rowptr, const int32_t *nz_eid,
                        const int32_t *src_eid, const int32_t *drn_eid,
                        double *g, double *val, double *diag, double *rhs, uint8_t *cls /* may be NULL */,
                        const int32_t *ve1, const int32_t *ve2 /* from cnt_series_select, or NULL; then the classes
                        9*nlevels ... are the pairs (c1 <= c2) of base classes in row-major triangular order */,
                        void *stream);

/* ---------------------------------------------------------------------------------------
 * (5) solve -- replaces solve_mna (cnet.py:147-149, SuperLU) by a batched Jacobi-PCG.
 * Value arrays come in NSLAB slabs (one per gate configuration): val + q*NNZ,
 * diag/rhs/x + q*R, all sharing the batch pattern.  System s = (network h_sys_net[s], slab
 * h_sys_slab[s]); a (network, slab) pair may appear at most once.  Stops when
 * ||r||_2 <= rtol*||b||_2 (recurrence residual; see DESIGN.md "PCG accuracy" for why there is
 * no residual replacement) or after maxit.  x holds the initial guess on entry.
 * Per system: iters (int32), relres (double), status (0 converged, 1 maxit, 2 breakdown).
 * Synchronises the stream every `check_every` iterations to poll for completion.
 * ------------------------------------------------------------------------------------- */
/* Two-level preconditioner (optional, both solvers): M^-1 = D^-1 + Z diag(Z^T A Z)^-1 Z^T with
 * Z the indicator vectors of the aggregates of strongly coupled unknowns (connected components
 * of the entries with |val| > theta * max|val| of the system).  Jacobi on the fine level plus
 * Jacobi on the aggregate level: per iteration one segmented sum and one broadcast.  It cuts
 * the iteration count 5-60x on gated networks (DESIGN.md "Two-level preconditioner") and is
 * deterministic.  cnt_pcg_aggregates builds the tables (it synchronises the stream once);
 * system indices stored in seg_sys refer to the h_sys_* arrays given to it, so the solver must
 * be called with the same system list.  Capacities: agg_of_row, mem: NSLAB*R; seg_*:
 * NSLAB*R/2 + NSLAB*R/256 + 1; agg_seg0: NSLAB*R/2 + 2; agg_einv: NSLAB*R/2 + 1. */
typedef struct cnt_aggregates {
    const int32_t *agg_of_row;   /* [NSLAB*R] aggregate of slab row (slab*R + row), -1 = none      */
    const int32_t *mem;          /* [NSLAB*R] slab rows sorted by aggregate                        */
    const int32_t *seg_start;    /* [nseg] first member (index into mem) of a segment (<= 256 rows) */
    const int32_t *seg_count;    /* [nseg]                                                          */
    const int32_t *seg_agg;      /* [nseg] aggregate of the segment                                 */
    const int32_t *seg_sys;      /* [nseg] system of the segment                                    */
    const int32_t *agg_seg0;     /* [nagg+1] first segment of an aggregate                          */
    const double *agg_einv;      /* [nagg] 1 / (Z^T A Z)_aa                                         */
    const int32_t *sys_seg;      /* [2*NSYS] segments of system s: [sys_seg[2s], sys_seg[2s+1])     */
    /* cluster solver only (chunk_C > 0): segments never straddle the row chunks of a C-CTA cluster */
    const int32_t *row_q;        /* [2*NSLAB*R] (first segment, segment count) of the row's aggregate */
    const double *row_einv;      /* [NSLAB*R] 1/E of the row's aggregate (0 = none)                  */
    const int32_t *cta_seg_off;  /* [NSYS*chunk_C+1] segments owned by CTA c of system s             */
    const int32_t *cta_seg;      /* [4*nseg] (segment id, first member, count, 0) in CTA order        */
    const int32_t *cta_agg_off;  /* [NSYS*chunk_C+1] aggregates touched by CTA c of system s          */
    const int32_t *cta_agg;      /* aggregate ids in CTA order (<= nseg entries)                      */
    const int32_t *row_slot;     /* [NSLAB*R] slot of the row's aggregate in its CTA's list, -1 none  */
    const int32_t *seg_home;     /* [nseg] (owning CTA << 16) | index in that CTA's segment list      */
    int64_t nagg, nseg;
    int32_t chunk_C, pad;
} cnt_aggregates;
int64_t cnt_pcg_aggregate_workspace_bytes(int NSYS, int NSLAB, int64_t R);
int cnt_pcg_aggregates(int B, int NSYS, int NSLAB, const int32_t *h_sys_net, const int32_t *h_sys_slab,
                       const int32_t *h_roff, const int32_t *roff, int64_t R, int64_t NNZ,
                       const int32_t *rowptr, const int32_t *col, const double *val, const double *diag,
                       double theta, int32_t *agg_of_row, int32_t *mem, int32_t *seg_start,
                       int32_t *seg_count, int32_t *seg_agg, int32_t *seg_sys, int32_t *agg_seg0,
                       double *agg_einv, int32_t *sys_seg /* [2*NSYS] */,
                       int chunk_C /* cluster size the tables are built for (cnt_pcg_cluster_size), 0 = flat only */,
                       int32_t *row_q, double *row_einv, int32_t *cta_seg_off, int32_t *cta_seg,
                       int32_t *cta_agg_off, int32_t *cta_agg /* capacity of seg_* */, int32_t *row_slot,
                       int32_t *seg_home /* capacity of seg_* */,
                       int64_t seg_capacity /* entries of the seg_* / cta_agg / seg_home arrays (cta_seg: 4x + 4 per CTA):
                                               NSLAB*R/2 + NSLAB*R/256 + 2 is enough for chunk_C == 0, NSLAB*R + 2 always;
                                               CNT_ECAPACITY when the tables need more */,
                       int64_t *h_nagg, int64_t *h_nseg,
                       void *ws, int64_t ws_bytes, void *stream);

int64_t cnt_pcg_workspace_bytes(int NSYS, int NSLAB, int64_t R);
int cnt_pcg_solve(int B, int NSYS, int NSLAB, const int32_t *h_sys_net, const int32_t *h_sys_slab,
                  const int32_t *h_roff, int64_t R, int64_t NNZ, const int32_t *rowptr, const int32_t *col,
                  const double *val, const double *diag, const double *rhs, double *x,
                  double rtol, int maxit, int check_every,
                  int32_t *iters, double *relres, int32_t *status,
                  const cnt_aggregates *agg /* NULL = plain Jacobi */,
                  void *ws, int64_t ws_bytes, void *stream);

/* Persistent thread-block-cluster variant of the same solve (pcg_cluster.cu): ONE launch for
 * the whole batch; each cluster of CTAs pulls systems from a work queue and keeps p, r, diag
 * in shared memory and x, q/z in registers for the whole solve (dot products through
 * distributed shared memory, phase boundaries = hardware cluster barriers).  Same arguments,
 * stopping rule and per-system outputs as cnt_pcg_solve; nothing synchronises the stream.
 * Returns CNT_ECAPACITY if a system has more than cnt_pcg_cluster_max_rows() rows (use
 * cnt_pcg_solve for those).  With cls/gneg (from cnt_assemble_values) every CTA packs its
 * slice of the matrix into shared memory (16-bit slot + 8-bit class per entry) so that an
 * iteration touches global memory only for the halo exchange of p; a system whose slice or
 * halo does not fit gets status 3 and must be re-solved with cnt_pcg_solve. */
int64_t cnt_pcg_cluster_max_rows(void);
/* debug: device buffer [grid][16] of int64 that receives per-phase cycle sums (thread 0 of every CTA) of the
 * following cluster launches; NULL switches the instrumentation off (default) */
void cnt_pcg_cluster_set_phase_buffer(long long *buf);
/* cluster size (CTAs) for a batch whose largest system has max_rows rows and max_nnz off-diagonal
 * entries (0 = unknown: rows only); pass it as cluster_C below and as chunk_C to cnt_pcg_aggregates */
int cnt_pcg_cluster_size(int64_t max_rows, int64_t max_nnz);
/* clusters of C CTAs (C = 1,2,4,8,16) that can be resident at once on the current device */
int cnt_pcg_cluster_active(int C);
int64_t cnt_pcg_cluster_workspace_bytes(int NSYS, int NSLAB, int64_t R);
int cnt_pcg_cluster_solve(int B, int NSYS, int NSLAB, const int32_t *h_sys_net, const int32_t *h_sys_slab,
                          const int32_t *h_roff, int64_t R, int64_t NNZ, const int32_t *rowptr, const int32_t *col,
                          const double *val, const uint8_t *cls /* [NSLAB*NNZ] or NULL */,
                          const double *gneg /* [NSLAB*256] = -gtable, or NULL */,
                          const double *diag, const double *rhs, double *x,
                          double rtol, int maxit,
                          int32_t *iters, double *relres, int32_t *status,
                          const cnt_aggregates *agg /* NULL = plain Jacobi; needs cls/gneg */,
                          int cluster_C /* CTAs per cluster, 0 = smallest that holds the rows */,
                          void *ws, int64_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * (5b) ONE very large network sharded over several GPUs (BASELINE config 5; no counterpart in the
 *      reference -- extends solve_mna, cnet.py:147-149).  Every rank owns a contiguous range of
 *      the Z-ordered unknown rows (a compact spatial patch) of the same system, keeps its slice of
 *      the matrix with LOCAL column indices (own rows 0..n_own-1, then the n_halo remote rows it
 *      references) and advances the NS gate configurations together.  The library provides the
 *      per-rank phases; the caller performs the exchange named after each phase between them
 *      (networksim_cntfet_b200/slab.py does it with torch.distributed over NCCL):
 *        cnt_slab_init                    -> all-reduce(red2)   -> cnt_slab_direction(first=1) -> all-gather(sendbuf -> recvbuf)
 *        per iteration: cnt_slab_spmv     -> all-reduce(red_pq) -> cnt_slab_update             -> all-reduce(red2)
 *                       -> cnt_slab_direction(first=0)          -> all-gather(sendbuf -> recvbuf)
 *      Nothing synchronises the stream.  Aggregates (two-level preconditioner, cnt_pcg_aggregates on
 *      the replicated system) are renumbered per rank: local aggregate = one with members among the
 *      own rows; shared = one with members on more than one rank (same numbering on all ranks).
 *      All [NS, ...] arrays are dense with the stated row length.
 * ------------------------------------------------------------------------------------- */
enum { CNT_SLAB_RZ = 0, CNT_SLAB_BB = 1, CNT_SLAB_RR = 2, CNT_SLAB_BETA = 3, CNT_SLAB_ITERS = 4,
       CNT_SLAB_STATUS = 5, CNT_SLAB_NSCAL = 8 };
#define CNT_SLAB_BIGAGG 256     /* local aggregates with more members are summed by one block each */
typedef struct cnt_slab_ctx {
    int32_t NS, nblk;                 /* gate configurations; ceil(max(n_own,1) / cnt_slab_block_rows()) */
    int64_t n_own, n_halo;
    int64_t val_stride;               /* elements between the value arrays of two configurations */
    const int32_t *rowptr;            /* [n_own+1] local */
    const int32_t *col;               /* local column: < n_own own row, else n_own + halo index */
    const double *val;                /* configuration q at val + q*val_stride */
    const double *diag, *rhs;         /* [NS, n_own] */
    double *x, *r, *q, *dinv;         /* [NS, n_own] */
    double *p;                        /* [NS, n_own + n_halo] */
    /* aggregates of the own rows */
    int64_t maxL, maxSh, maxBig;      /* row lengths (>= 1) of the aggregate tables below */
    const int32_t *nloc, *nsh, *nbig; /* [NS] local / shared / big local aggregates of configuration q */
    int32_t max_nbig, pad0;           /* max over q of nbig[q] (host copy) */
    const int32_t *aggptr;            /* [NS, maxL+1] members of local aggregate a: aggmem[aggptr[a] .. aggptr[a+1]) */
    const int32_t *aggmem;            /* [NS, n_own] own rows sorted by local aggregate */
    const int32_t *row_agg;           /* [NS, n_own] local aggregate of the row, -1 = none */
    const int32_t *agg_shared;        /* [NS, maxL] shared slot of the local aggregate, -1 = lives on this rank only */
    const double *einv;               /* [NS, maxL] 1 / (Z^T A Z)_aa */
    const int32_t *shared_local;      /* [NS, maxSh] local aggregate of shared slot s, -1 = no member here */
    const double *einv_sh;            /* [NS, maxSh] */
    const int32_t *bigagg;            /* [NS, maxBig] local aggregates with > CNT_SLAB_BIGAGG members */
    double *S;                        /* [NS, maxL] sums of r over the aggregates */
    /* reductions */
    double *part;                     /* [NS, 3, nblk] block partials */
    double *red_pq;                   /* [NS]              all-reduce (sum) after cnt_slab_spmv */
    double *red2;                     /* [NS, 2 + maxSh]   all-reduce (sum) after cnt_slab_init / cnt_slab_update */
    double *scal;                     /* [NS, CNT_SLAB_NSCAL] */
    int32_t *done;                    /* [NS] */
    /* halo exchange of p */
    int64_t n_exp, emax;              /* exported own rows; row length of the send buffer (max n_exp over ranks, >= 1) */
    const int32_t *exp_rows;          /* [n_exp] own rows other ranks reference */
    double *sendbuf;                  /* [NS, emax]        all-gather after cnt_slab_direction */
    const double *recvbuf;            /* [world, NS, emax] */
    const int64_t *halo_src;          /* [n_halo] index into recvbuf of halo row h for configuration 0 */
    double rtol;
    int32_t maxit, pad1;
} cnt_slab_ctx;
int cnt_slab_block_rows(void);
int cnt_slab_init(const cnt_slab_ctx *c, void *stream);
int cnt_slab_direction(const cnt_slab_ctx *c, int first, void *stream);
int cnt_slab_spmv(const cnt_slab_ctx *c, void *stream);
int cnt_slab_update(const cnt_slab_ctx *c, void *stream);

/* The SpMV of the PCG solvers on its own (the operator application inside the SuperLU-replacing solve,
 * cnet.py:147-149): y[q] = diag[q] .* x[q] + A[q] x[q] for NS value sets over ONE CSR pattern of off-diagonal
 * entries (rowptr [n_rows+1], col [nnz] = row index in 0..n_rows-1 -- or beyond, when x is longer than
 * n_rows (halo copies behind the own rows); val / diag / x / y: configuration q at base + q * stride;
 * diag may be NULL).  Streaming kernel (csrc/spmv_stream.cuh): a CTA reads the contiguous entry range of
 * its 256 rows with coalesced 16-byte loads (needs col and val + q * val_stride 16-byte aligned; otherwise
 * a scalar path), products through shared memory, one thread per row sums in CSR order (deterministic).
 * Algorithmic bytes per configuration: 12 nnz + 20 n_rows (+ 8 n_rows with diag). */
int cnt_spmv_csr(int NS, int64_t n_rows, int64_t nnz, const int32_t *rowptr, const int32_t *col, const double *val,
                 int64_t val_stride, const double *diag, int64_t diag_stride, const double *x, int64_t x_stride,
                 double *y, int64_t y_stride, void *stream);

/* ---------------------------------------------------------------------------------------
 * (5c) direct solve by parallel star-mesh elimination -- replaces solve_mna (cnet.py:147-149) by
 *      sparse Cholesky written on the conductances (every quantity a sum of positive terms):
 *        eliminate v (neighbours u_i, conductances g_i, electrode coupling c_v):
 *          d_v = c_v + sum g_i;  g(u_i,u_j) += g_i g_j / d_v;  c_ui += g_i c_v / d_v;  b_ui += g_i b_v / d_v
 *        back-substitution:  x_v = (b_v + sum g_i x_ui) / d_v
 *      cnt_sm_symbolic (csrc/starmesh_sym.cu; topology only, shared by all gate configurations)
 *      eliminates per round the independent set of unknowns whose (degree, hash, index) is the
 *      smallest among their neighbours, for all networks of the batch at once, and writes the PLAN:
 *      index lists over edge SLOTS that never move (round-0 edges first, fill edges appended in a
 *      deterministic order).  A network stops as soon as it has <= dense_max unknowns left; what is
 *      left is finished by one dense elimination per (network, configuration).  cnt_sm_solve is the
 *      numeric phase for NS gate configurations at once: every sum runs in a fixed order
 *      (deterministic bits, independent of what else is in the batch).
 *
 *      Plan = one caller-owned device buffer of cnt_sm_plan_layout(...) bytes holding the int32
 *      arrays below (offsets from cnt_sm_plan_layout; all indices global to the batch):
 *        PIV[nelim] eliminated row by pivot index m (rounds back to back, rows ascending inside a round)
 *        NPTR[nelim+1], NBR[nlist], ESLOT[nlist]: neighbours of pivot m (ascending) and the edge slots to them
 *        TGT[nge], CPTR[nge+1], CM/CA/CB[ncontrib]: edge slot receiving the contributions
 *            gl[CA] gl[CB] / d[CM], contributions of a target ordered by pivot.  CA, CB (and TSLOT) name a
 *            dead edge by its POSITION in NBR/ESLOT: gl[k] = g[ESLOT[k]] is final once its pivot is
 *            eliminated, and the numeric phase keeps it in list order (sequential reads instead of gathers)
 *        TN[ngn], TPTR[ngn+1], TM/TSLOT[nlist]: row receiving gl[TSLOT] {c,b}[PIV[TM]] / d[TM], ordered by pivot
 *        E0PTR[nE0+1], E0SRC: round-0 edge e = sum of the CSR entries E0SRC[E0PTR[e] .. E0PTR[e+1])
 *        DPTR[B+1], DNODES: rows left to the dense tail per network (ascending); DSLOT/DEPOS[dense_nE]:
 *            remaining edge slots and their position in the pool of dense matrices, grouped by network
 *            (DEDGE[B+1]); DOFF[B+1] (int64): offset of every network's n x n matrix in the pool
 *        INFO: device copy of cnt_sm_info
 * ------------------------------------------------------------------------------------- */
#define CNT_SM_MAX_ROUNDS 1024
enum {
    CNT_SM_PIV = 0, CNT_SM_NPTR, CNT_SM_NBR, CNT_SM_ESLOT, CNT_SM_TGT, CNT_SM_CPTR, CNT_SM_CM, CNT_SM_CA, CNT_SM_CB,
    CNT_SM_TN, CNT_SM_TPTR, CNT_SM_TM, CNT_SM_TSLOT, CNT_SM_E0PTR, CNT_SM_E0SRC, CNT_SM_DPTR, CNT_SM_DNODES,
    CNT_SM_DSLOT, CNT_SM_DOFF, CNT_SM_DEPOS, CNT_SM_INFO, CNT_SM_DEDGE, CNT_SM_NARR
};
/* error bits of cnt_sm_info.error */
#define CNT_SM_ERR_SLOTS 1     /* more edge slots than cap_slots */
#define CNT_SM_ERR_CONTRIB 2   /* more star-mesh contributions than cap_contrib (the fill budget) */
#define CNT_SM_ERR_ITEMS 4     /* a single round has more (pairs + incident entries) than cap_items */
#define CNT_SM_ERR_HASH 8      /* edge hash table full */
#define CNT_SM_ERR_SCAN 16     /* device-wide scan watchdog */
#define CNT_SM_ERR_ROUNDS 32   /* more than CNT_SM_MAX_ROUNDS - 1 rounds */
typedef struct cnt_sm_round {
    int32_t nelim, nlist, npairs;       /* pivots, incident entries (sum of degrees), neighbour pairs of this round */
    int32_t nge, ngn, nnew;             /* edge-target groups, node-target groups, new (fill) edge slots */
    int32_t ebase, lbase, cbase;        /* running totals before this round: pivots, incident entries, contributions */
    int32_t gebase, gnbase, sbase;      /* ... edge groups, node groups, edge slots */
    int32_t nlive, nnodes;              /* live edges / unknowns still in play at the start of the round */
    int32_t ngroups, pad;               /* nge + ngn */
} cnt_sm_round;
typedef struct cnt_sm_info {
    int32_t done, error, nrounds, max_deg;
    int32_t R, nE0, nslots, nelim;      /* unknowns, round-0 edges, edge slots, eliminated by the sparse rounds */
    int32_t nlist, ncontrib, nge, ngn;  /* totals over the rounds (nlist = nnz of the factor) */
    int32_t dense_nnet, dense_nmax, dense_nnodes, dense_nE;
    int64_t dense_gpool;                /* sum over the networks of (remaining unknowns)^2 */
    int64_t reserved;
    cnt_sm_round rounds[CNT_SM_MAX_ROUNDS];
} cnt_sm_info;
/* byte offsets of the CNT_SM_NARR plan arrays into offsets[0..CNT_SM_NARR); returns the buffer size */
int64_t cnt_sm_plan_layout(int64_t R, int B, int64_t NNZ, int64_t cap_slots, int64_t cap_contrib, int64_t *offsets);
int64_t cnt_sm_symbolic_workspace_bytes(int64_t R, int B, int64_t NNZ, int64_t cap_slots, int64_t cap_contrib,
                                        int64_t cap_items);
/* roff[B+1] (device): first row of every network; rowptr[R+1] / col[NNZ] (device): symmetric off-diagonal
 * pattern, columns network-local and ascending inside a row (cnt_assemble_pattern_fill).  rounds_hint: number of
 * rounds launched before the header is read back for the first time (one stream synchronisation per read-back;
 * rounds beyond the last needed one exit immediately).  hint (may be NULL): header of an earlier plan of a
 * similar batch; its per-round sizes size the grids of the round kernels (performance only: every kernel
 * strides over device-side counts).  Returns CNT_ECAPACITY with h_info->error set when a
 * capacity was exceeded (the caller retries with larger capacities or takes an iterative solver). */
int cnt_sm_symbolic(int B, const int32_t *roff, int64_t R, int64_t NNZ, const int32_t *rowptr, const int32_t *col,
                    int dense_max, int rounds_hint, int64_t cap_slots, int64_t cap_contrib, int64_t cap_items,
                    const cnt_sm_info *hint, void *plan, int64_t plan_bytes, void *ws, int64_t ws_bytes,
                    cnt_sm_info *h_info, void *stream);
int64_t cnt_sm_solve_workspace_bytes(int NS, const cnt_sm_info *h_info);
/* numeric phase.  val [NS, val_stride]: CSR off-diagonal values (-g) of cnt_assemble_values; electrode coupling of
 * row r = g[src_eid[r]] + g[drn_eid[r]] (either may be < 0 = none; src_eid / drn_eid / g may be NULL) + coupling[q, r]
 * (optional [NS, R]); leak != 0 adds the rounding error of the assembled FP64 diagonal (cnt_assemble_values'
 * summation order), i.e. the direct solve then answers the same floating-point matrix as the iterative solvers;
 * rhs [NS, rhs_stride]; x [NS, R] out.  Nothing synchronises the stream. */
int cnt_sm_solve(int NS, int B, int64_t R, int64_t NNZ, const cnt_sm_info *h_info, const void *plan,
                 int64_t cap_slots, int64_t cap_contrib, const int32_t *rowptr, const double *val, int64_t val_stride,
                 const int32_t *src_eid, const int32_t *drn_eid, const double *g, int64_t g_stride,
                 const double *coupling, const double *rhs, int64_t rhs_stride, int leak, double *x, void *ws,
                 int64_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * (6) write-back -- replaces update_voltages / update_currents (cnet.py:132-146).
 * V[S]: stick voltage (0.1 source, 0 drain, NaN outside the percolating component);
 * I[J]: |g (V1 - V2)| per junction (NaN outside the component);
 * isrc[b*3+0] = sum_j g_0j (0.1 - V_j)  (== x[-1] of the MNA system, cnet.py:135)
 * isrc[b*3+1] = sum_j g_1j V_j          (current into the ground electrode)
 * isrc[b*3+2] = sum_e g_e dV_e^2 / 0.1  (dissipated power / Vds)
 * The sums run over chunks of 8192 junctions counted from the network's own first junction, then over the
 * chunks in order: fixed order, independent of the batch.  ws: cnt_currents_workspace_bytes(B, J) bytes.
 * ------------------------------------------------------------------------------------- */
int64_t cnt_currents_workspace_bytes(int B, int64_t J);
int cnt_currents(int B, const int32_t *soff, const int32_t *joff, int64_t S, int64_t J,
                 const int32_t *stick1, const int32_t *stick2, const int32_t *label, const int32_t *stats,
                 const int32_t *urow, const int32_t *attach /* from cnt_prune_leaves, or NULL */,
                 const int32_t *elim_vid, const int32_t *va, const int32_t *vc, const int32_t *ve1,
                 const int32_t *ve2 /* from cnt_series_select, or all NULL */,
                 const double *g, const double *x, double *V, double *I, double *isrc, void *ws, int64_t ws_bytes,
                 void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CNTNET_B200_H */
