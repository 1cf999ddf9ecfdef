// Layout of the star-mesh plan buffer shared by the symbolic phase (starmesh_sym.cu) and the numeric
// phase (starmesh.cu).  The plan is ONE caller-owned device buffer; the positions of its arrays
// depend only on (R, B, NNZ, cap_slots, cap_contrib) and are exported through cnt_sm_plan_layout.
#pragma once
#include "common.cuh"

namespace cnt {
namespace sm {

struct Plan {
    int32_t *piv, *nptr, *nbr, *eslot;             // pivots (by pivot index) and their incident lists
    int32_t *tgt, *cptr, *cm, *ca, *cb;            // edge-target groups and their star-mesh contributions
    int32_t *tn, *tptr, *tm, *tslot;               // node-target groups (c, b updates of the neighbours)
    int32_t *e0_ptr, *e0_src;                      // round-0 edge values: sums of CSR entries
    int32_t *dptr, *dnodes, *dslot, *dedge;        // dense tail: rows of every network, remaining edge slots (grouped by network)
    long long *doff, *depos;                       // ... matrix offsets, position of every remaining edge
    cnt_sm_info *info;                             // device copy of the header
};

// byte offsets of the CNT_SM_NARR arrays (+ total size in off[CNT_SM_NARR])
inline int64_t plan_layout(int64_t R, int B, int64_t NNZ, int64_t capS, int64_t capC, int64_t *off) {
    const int64_t n[CNT_SM_NARR] = {
        R, R + 1, capS, capS,                      // piv nptr nbr eslot
        capC, capC + 1, capC, capC, capC,          // tgt cptr cm ca cb
        capS, capS + 1, capS, capS,                // tn tptr tm tslot
        NNZ + 2, NNZ + 1,                          // e0_ptr e0_src
        (int64_t)B + 1, R, capS,                   // dptr dnodes dslot
        2 * ((int64_t)B + 1), 2 * capS,            // doff depos (64-bit entries)
        (int64_t)((sizeof(cnt_sm_info) + 3) / 4),  // info
        (int64_t)B + 1,                            // dedge
    };
    int64_t pos = 0;
    for (int k = 0; k < CNT_SM_NARR; k++) {
        off[k] = pos;
        pos += align_up(n[k] * 4, 256);
    }
    off[CNT_SM_NARR] = pos;
    return pos;
}

inline Plan plan_view(void *buf, const int64_t *off) {
    char *p = (char *)buf;
    Plan v;
    v.piv = (int32_t *)(p + off[CNT_SM_PIV]);
    v.nptr = (int32_t *)(p + off[CNT_SM_NPTR]);
    v.nbr = (int32_t *)(p + off[CNT_SM_NBR]);
    v.eslot = (int32_t *)(p + off[CNT_SM_ESLOT]);
    v.tgt = (int32_t *)(p + off[CNT_SM_TGT]);
    v.cptr = (int32_t *)(p + off[CNT_SM_CPTR]);
    v.cm = (int32_t *)(p + off[CNT_SM_CM]);
    v.ca = (int32_t *)(p + off[CNT_SM_CA]);
    v.cb = (int32_t *)(p + off[CNT_SM_CB]);
    v.tn = (int32_t *)(p + off[CNT_SM_TN]);
    v.tptr = (int32_t *)(p + off[CNT_SM_TPTR]);
    v.tm = (int32_t *)(p + off[CNT_SM_TM]);
    v.tslot = (int32_t *)(p + off[CNT_SM_TSLOT]);
    v.e0_ptr = (int32_t *)(p + off[CNT_SM_E0PTR]);
    v.e0_src = (int32_t *)(p + off[CNT_SM_E0SRC]);
    v.dptr = (int32_t *)(p + off[CNT_SM_DPTR]);
    v.dnodes = (int32_t *)(p + off[CNT_SM_DNODES]);
    v.dslot = (int32_t *)(p + off[CNT_SM_DSLOT]);
    v.doff = (long long *)(p + off[CNT_SM_DOFF]);
    v.depos = (long long *)(p + off[CNT_SM_DEPOS]);
    v.info = (cnt_sm_info *)(p + off[CNT_SM_INFO]);
    v.dedge = (int32_t *)(p + off[CNT_SM_DEDGE]);
    return v;
}

}  // namespace sm
}  // namespace cnt
