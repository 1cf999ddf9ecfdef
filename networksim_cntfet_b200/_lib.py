"""ctypes binding of libcntnet_b200.so (the C-ABI declared in include/cntnet_b200.h).

There is NO fallback: if the shared library is missing or a call returns a non-zero status
this module raises.  Build it with `python -m networksim_cntfet_b200.build`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CNT_B200_LIB: another build of the same library (A/B measurements of kernel variants); never a fallback
LIB_PATH = os.environ.get("CNT_B200_LIB") or os.path.join(HERE, "libcntnet_b200.so")

CNT_NSTAT = 8
STAT_PERCOLATING, STAT_NCLUST_TRUE, STAT_MAXCLUST_TRUE, STAT_NCLUST_REF, STAT_MAXCLUST_REF, STAT_NCOMP, STAT_ECOMP = range(7)

# cnt_slab_ctx (sharded solve of one large network): scalar slots and the size above which a local
# aggregate is summed by a whole block
CNT_SLAB_RZ, CNT_SLAB_BB, CNT_SLAB_RR, CNT_SLAB_BETA, CNT_SLAB_ITERS, CNT_SLAB_STATUS = range(6)
CNT_SLAB_NSCAL = 8
CNT_SLAB_BIGAGG = 256

vp, i32, i64, u64, f64, u8 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_uint8
I32P = C.POINTER(C.c_int32)

# name -> (restype, argtypes); mirrors include/cntnet_b200.h one to one
PROTOTYPES = {
    "cnt_abi_version": (i32, []),
    "cnt_last_error": (C.c_char_p, []),
    "cnt_launch_count": (i64, []),
    "cnt_fp64_peak": (i32, [vp, i32, C.POINTER(f64), vp]),
    "cnt_scan_workspace_bytes": (i64, [i64]),
    "cnt_exclusive_scan_i32": (i32, [vp, vp, i64, i32, vp, i64, vp]),
    "cnt_gen_sticks_workspace_bytes": (i64, [i64, i32]),
    "cnt_gen_sticks": (i32, [vp, i32, vp, I32P, vp, f64, i32, f64] + [vp] * 10 + [vp, i64, vp]),
    "cnt_junction_workspace_bytes": (i64, [i32, I32P]),
    "cnt_junction_count": (i32, [i32, vp, I32P] + [vp] * 7 + [vp, vp, vp, i64, vp]),
    "cnt_junction_fill": (i32, [i32, vp, I32P] + [vp] * 8 + [vp, i64] + [vp] * 5 + [vp, vp, i64, vp]),
    "cnt_junction_count_slab": (i32, [i32, vp, I32P] + [vp] * 7 + [vp, vp, vp, i64, i32, i32, vp]),
    "cnt_junction_fill_slab": (i32, [i32, vp, I32P] + [vp] * 8 + [vp, i64] + [vp] * 5 + [vp, vp, i64, i32, i32, vp]),
    "cnt_junction_errflag_ptr": (i32, [i32, I32P, vp, i64, C.POINTER(vp)]),
    "cnt_cluster_workspace_bytes": (i64, [i64]),
    "cnt_label_clusters": (i32, [i32, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, i64, vp]),
    "cnt_assemble_workspace_bytes": (i64, [i64]),
    "cnt_prune_workspace_bytes": (i64, [i64]),
    "cnt_prune_leaves": (i32, [i32, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "cnt_series_workspace_bytes": (i64, [i64]),
    "cnt_series_select": (i32, [i32, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "cnt_assemble_rows": (i32, [i32, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    "cnt_assemble_reorder_workspace_bytes": (i64, [i64, i32]),
    "cnt_assemble_reorder": (i32, [i32, I32P, i64, i64, vp, vp, vp, vp, i64, vp]),
    "cnt_assemble_pattern_count": (i32, [i32, vp, vp, i64, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp]),
    "cnt_assemble_pattern_fill": (i32, [i32, vp, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, vp]),
    "cnt_assemble_row_stick": (i32, [i32, vp, i64, vp, vp, vp]),
    "cnt_gate_levels": (i32, [i64, vp, vp, vp, i32, u8, i32, f64, f64, f64, f64, u8, vp]),
    "cnt_assemble_values": (i32, [i64, vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "cnt_pcg_aggregate_workspace_bytes": (i64, [i32, i32, i64]),
    "cnt_pcg_aggregates": (i32, [i32, i32, i32, I32P, I32P, I32P, vp, i64, i64, vp, vp, vp, vp, f64] + [vp] * 9 +
                           [i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, C.POINTER(i64), C.POINTER(i64), vp, i64, vp]),
    "cnt_pcg_workspace_bytes": (i64, [i32, i32, i64]),
    "cnt_pcg_solve": (i32, [i32, i32, i32, I32P, I32P, I32P, i64, i64, vp, vp, vp, vp, vp, vp, f64, i32, i32, vp, vp, vp, vp, vp, i64, vp]),
    "cnt_pcg_cluster_max_rows": (i64, []),
    "cnt_pcg_cluster_set_phase_buffer": (None, [vp]),
    "cnt_pcg_cluster_active": (i32, [i32]),
    "cnt_pcg_cluster_size": (i32, [i64, i64]),
    "cnt_pcg_cluster_workspace_bytes": (i64, [i32, i32, i64]),
    "cnt_pcg_cluster_solve": (i32, [i32, i32, i32, I32P, I32P, I32P, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, f64, i32, vp, vp, vp, vp, i32, vp, i64, vp]),
    "cnt_sm_plan_layout": (i64, [i64, i32, i64, i64, i64, C.POINTER(i64)]),
    "cnt_sm_symbolic_workspace_bytes": (i64, [i64, i32, i64, i64, i64, i64]),
    "cnt_sm_symbolic": (i32, [i32, vp, i64, i64, vp, vp, i32, i32, i64, i64, i64, vp, vp, i64, vp, i64, vp, vp]),
    "cnt_sm_solve_workspace_bytes": (i64, [i32, vp]),
    "cnt_sm_solve": (i32, [i32, i32, i64, i64, vp, vp, i64, i64, vp, vp, i64, vp, vp, vp, i64, vp, vp, i64, i32, vp, vp,
                           i64, vp]),
    "cnt_slab_block_rows": (i32, []),
    "cnt_slab_init": (i32, [vp, vp]),
    "cnt_slab_direction": (i32, [vp, i32, vp]),
    "cnt_slab_spmv": (i32, [vp, vp]),
    "cnt_slab_update": (i32, [vp, vp]),
    "cnt_spmv_csr": (i32, [i32, i64, i64, vp, vp, vp, i64, vp, i64, vp, i64, vp, i64, vp]),
    "cnt_currents_workspace_bytes": (i64, [i32, i64]),
    "cnt_currents": (i32, [i32, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
}
# entry points that may be absent from an older build of the library (checked by tests
# against the header, never silently replaced)
OPTIONAL = set()


class Aggregates(C.Structure):
    """mirror of `struct cnt_aggregates` (include/cntnet_b200.h)"""
    _fields_ = [("agg_of_row", vp), ("mem", vp), ("seg_start", vp), ("seg_count", vp), ("seg_agg", vp),
                ("seg_sys", vp), ("agg_seg0", vp), ("agg_einv", vp), ("sys_seg", vp), ("row_q", vp), ("row_einv", vp),
                ("cta_seg_off", vp), ("cta_seg", vp), ("cta_agg_off", vp), ("cta_agg", vp), ("row_slot", vp),
                ("seg_home", vp),
                ("nagg", i64), ("nseg", i64), ("chunk_C", C.c_int32),
                ("pad", C.c_int32)]


class CntError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library and attach prototypes; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CntError("%s not found: run `python -m networksim_cntfet_b200.build` (there is no CPU fallback)"
                       % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if name in OPTIONAL:
                continue
            raise CntError("libcntnet_b200.so does not export %s (stale build?)" % name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().cnt_last_error().decode(errors="replace")
        raise CntError("cntnet_b200 status %d: %s" % (rc, msg))


def i32_array(values):
    """host int32 array for the h_* arguments"""
    arr = (C.c_int32 * len(values))(*[int(v) for v in values])
    return arr
