"""CPU: the C-ABI library builds, loads, and exports every symbol include/cntnet_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "cntnet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cnt_[a-z0-9_]+)\s*\(", txt)))


def test_build_and_symbols():
    from networksim_cntfet_b200 import build
    lib_path = build.build()
    assert os.path.isfile(lib_path)
    lib = ctypes.CDLL(lib_path)
    syms = header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.cnt_abi_version.restype = ctypes.c_int
    assert lib.cnt_abi_version() == 2


def test_binding_covers_header():
    from networksim_cntfet_b200 import _lib
    assert sorted(set(_lib.PROTOTYPES) - _lib.OPTIONAL) == header_symbols()
    lib = _lib.load()
    lib.cnt_launch_count.restype = ctypes.c_int64
    assert lib.cnt_launch_count() >= 0


def test_workspace_queries_are_host_only():
    from networksim_cntfet_b200 import _lib
    lib = _lib.load()
    assert lib.cnt_scan_workspace_bytes(10 ** 6) >= 4 * 245
    assert lib.cnt_cluster_workspace_bytes(1000) >= 9000
    soff = _lib.i32_array([0, 302, 6704])
    assert lib.cnt_junction_workspace_bytes(2, soff) > 6704 * 64


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    from networksim_cntfet_b200 import _lib
    from networksim_cntfet_b200.engine import Engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.CntError):
        Engine()


def header_param_counts():
    txt = open(os.path.join(ROOT, "include", "cntnet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(cnt_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_prototype_arity_matches_header():
    """every ctypes prototype has exactly as many arguments as the C declaration"""
    from networksim_cntfet_b200 import _lib
    counts = header_param_counts()
    bad = {n: (len(a), counts.get(n)) for n, (_, a) in _lib.PROTOTYPES.items() if counts.get(n) != len(a)}
    assert not bad, bad


def test_aggregates_struct_matches_header():
    from networksim_cntfet_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "cntnet_b200.h")).read()
    body = re.search(r"typedef struct cnt_aggregates \{(.*?)\} cnt_aggregates;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip().replace("const ", "")
        if decl:
            names += [n.strip().lstrip("*").strip() for n in decl.split(" ", 1)[1].split(",")]
    assert names == [f[0] for f in _lib.Aggregates._fields_], (names, [f[0] for f in _lib.Aggregates._fields_])


def test_slab_ctx_struct_matches_header():
    """field names and order of slab.SlabCtx mirror struct cnt_slab_ctx"""
    from networksim_cntfet_b200 import slab
    txt = open(os.path.join(ROOT, "include", "cntnet_b200.h")).read()
    body = re.search(r"typedef struct cnt_slab_ctx \{(.*?)\} cnt_slab_ctx;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip().replace("const ", "")
        if decl:
            names += [n.strip().lstrip("*").strip() for n in decl.split(" ", 1)[1].split(",")]
    assert names == [f[0] for f in slab.SlabCtx._fields_], (names, [f[0] for f in slab.SlabCtx._fields_])
    import ctypes
    # 8-byte members only after the two leading int32 pairs: no hidden padding
    assert ctypes.sizeof(slab.SlabCtx) == sum(ctypes.sizeof(f[1]) for f in slab.SlabCtx._fields_)


def _struct_fields(name):
    txt = open(os.path.join(ROOT, "include", "cntnet_b200.h")).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip().replace("const ", "")
        if decl:
            names += [re.sub(r"\[.*\]", "", n).strip().lstrip("*").strip() for n in decl.split(" ", 1)[1].split(",")]
    return names


def test_starmesh_structs_match_header():
    """starmesh.SmRound / SmInfo mirror cnt_sm_round / cnt_sm_info; the plan layout query is host-only"""
    from networksim_cntfet_b200 import _lib, starmesh
    assert _struct_fields("cnt_sm_round") == [f[0] for f in starmesh.SmRound._fields_]
    assert _struct_fields("cnt_sm_info") == [f[0] for f in starmesh.SmInfo._fields_]
    assert ctypes.sizeof(starmesh.SmRound) == 64
    assert ctypes.sizeof(starmesh.SmInfo) == 16 * 4 + 16 + 64 * starmesh.MAX_ROUNDS
    txt = open(os.path.join(ROOT, "include", "cntnet_b200.h")).read()
    assert int(re.search(r"#define CNT_SM_MAX_ROUNDS (\d+)", txt).group(1)) == starmesh.MAX_ROUNDS
    enum = re.search(r"enum \{\s*CNT_SM_PIV = 0,(.*?)CNT_SM_NARR", txt, flags=re.S).group(1)
    names = ["piv"] + [n.strip()[7:].lower().replace("e0ptr", "e0_ptr").replace("e0src", "e0_src") for n in enum.split(",") if n.strip()]
    assert tuple(names) == starmesh.ARRAYS
    lib = _lib.load()
    off = (ctypes.c_int64 * (len(starmesh.ARRAYS) + 1))()
    total = lib.cnt_sm_plan_layout(1000, 3, 5000, 6000, 20000, off)
    assert total == off[len(starmesh.ARRAYS)] and list(off) == sorted(off) and off[0] == 0
    assert off[starmesh.ARRAYS.index("cm") + 1] - off[starmesh.ARRAYS.index("cm")] >= 4 * 20000
    assert lib.cnt_sm_symbolic_workspace_bytes(1000, 3, 5000, 6000, 20000, 12000) > 16 * 2 * 6000
