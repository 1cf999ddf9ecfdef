// FP64 segment-intersection predicate, bit-compatible with netsim.py:75-94.
//
// The reference evaluates, for stick 1 = lower (longer) index and stick 2:
//     m  = (ya - yb) / (xa - xb)            per stick      (netsim.py:81-82)
//     b  = ya - m * xa                      per stick      (netsim.py:84-85)
//     m1 == m2 -> no junction                              (netsim.py:86-87)
//     xi = (b2 - b1) / (m1 - m2)                           (netsim.py:89)
//     yi = (b2*m1 - b1*m2) / (m1 - m2)                     (netsim.py:90)
//     junction iff min(x1) < xi < max(x1) and min(x2) < xi < max(x2)   (netsim.py:91)
// then make_intersects_kdtree keeps it iff 0<=xi<=1 and 0<=yi<=1 (netsim.py:147).
// m and b depend on one stick only, so they are computed once per stick (same operations,
// same rounding) and the pair test needs two divisions instead of four.  Every operation is
// an explicitly rounded IEEE double op -- no FMA contraction -- so xi, yi and the decision
// are bit-identical to CPython/numpy float arithmetic.  Vertical sticks give inf/nan and
// fall out as "no junction" through the comparisons, exactly as in the reference.
// The reference's early-out (netsim.py:77-78) cannot change the result (if max(x1)<min(x2)
// no xi can satisfy both strict intervals) and is not reproduced.
#pragma once

#if defined(__CUDA_ARCH__)
#define CNT_HD __host__ __device__ __forceinline__
#define CNT_MUL(a, b) __dmul_rn((a), (b))
#define CNT_SUB(a, b) __dsub_rn((a), (b))
#define CNT_ADD(a, b) __dadd_rn((a), (b))
#define CNT_DIV(a, b) __ddiv_rn((a), (b))
#else
#if defined(__CUDACC__)
#define CNT_HD __host__ __device__ __forceinline__
#else
#define CNT_HD static inline
#endif
// host build: compile with -ffp-contract=off
#define CNT_MUL(a, b) ((a) * (b))
#define CNT_SUB(a, b) ((a) - (b))
#define CNT_ADD(a, b) ((a) + (b))
#define CNT_DIV(a, b) ((a) / (b))
#endif

// per-stick line record: slope, intercept, x-interval
struct CntLine {
    double m, b, xmin, xmax;
};

CNT_HD CntLine cnt_make_line(double xa, double ya, double xb, double yb) {
    CntLine l;
    l.m = CNT_DIV(CNT_SUB(ya, yb), CNT_SUB(xa, xb));
    l.b = CNT_SUB(ya, CNT_MUL(l.m, xa));
    l.xmin = xa < xb ? xa : xb;
    l.xmax = xa < xb ? xb : xa;
    return l;
}

// KD-tree ball of netsim.py:142: centre distance <= length of the lower-index stick
CNT_HD bool cnt_in_ball(double xc1, double yc1, double len1, double xc2, double yc2) {
    double dx = CNT_SUB(xc1, xc2), dy = CNT_SUB(yc1, yc2);
    return CNT_ADD(CNT_MUL(dx, dx), CNT_MUL(dy, dy)) <= CNT_MUL(len1, len1);
}

// pair predicate + unit-box filter; l1 is the lower index
CNT_HD bool cnt_junction(const CntLine &l1, const CntLine &l2, double *xi_out, double *yi_out) {
    // exact necessary condition of the strict double interval test: cheap reject, no division
    double lo = l1.xmin > l2.xmin ? l1.xmin : l2.xmin;
    double hi = l1.xmax < l2.xmax ? l1.xmax : l2.xmax;
    if (!(lo < hi)) return false;
    if (l1.m == l2.m) return false;
    double d = CNT_SUB(l1.m, l2.m);
    double xi = CNT_DIV(CNT_SUB(l2.b, l1.b), d);
    if (!(lo < xi && xi < hi)) return false;
    double yi = CNT_DIV(CNT_SUB(CNT_MUL(l2.b, l1.m), CNT_MUL(l1.b, l2.m)), d);
    if (!(0.0 <= xi && xi <= 1.0 && 0.0 <= yi && yi <= 1.0)) return false;
    *xi_out = xi;
    *yi_out = yi;
    return true;
}
