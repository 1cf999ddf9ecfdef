/* Host build of the device predicate (networksim_cntfet_b200/csrc/predicate.cuh) for the CPU test
 * suite: g++ -O2 -ffp-contract=off -shared -fPIC.  Same source as the CUDA kernels, so a bit-exact
 * match with the reference's golden junction tables on the CPU pins the arithmetic sequence. */
#include <stdint.h>
#include "../../networksim_cntfet_b200/csrc/predicate.cuh"

/* all pairs i<j inside the KD-ball of stick i; writes up to cap hits; returns the number of hits
 * (which may exceed cap) and the number of tested pairs through *ntests */
extern "C" long predicate_all_pairs(long n, const double *xa, const double *ya, const double *xb, const double *yb,
                         const double *xc, const double *yc, const double *len, long cap, int32_t *s1, int32_t *s2,
                         double *jx, double *jy, long *ntests) {
    long hits = 0, tests = 0;
    for (long i = 0; i < n; i++) {
        CntLine l1 = cnt_make_line(xa[i], ya[i], xb[i], yb[i]);
        for (long j = i + 1; j < n; j++) {
            if (!cnt_in_ball(xc[i], yc[i], len[i], xc[j], yc[j])) continue;
            tests++;
            CntLine l2 = cnt_make_line(xa[j], ya[j], xb[j], yb[j]);
            double xi, yi;
            if (cnt_junction(l1, l2, &xi, &yi)) {
                if (hits < cap) { s1[hits] = (int32_t)i; s2[hits] = (int32_t)j; jx[hits] = xi; jy[hits] = yi; }
                hits++;
            }
        }
    }
    *ntests = tests;
    return hits;
}
