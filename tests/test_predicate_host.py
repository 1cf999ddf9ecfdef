"""CPU: the predicate header shared with the CUDA kernels (csrc/predicate.cuh), compiled for the host
with g++ -ffp-contract=off, reproduces the reference's junction tables BIT-exactly (netsim.py:75-94,
:142, :147) -- pins the operation order / rounding of the device code without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden

SRC = os.path.join(ROOT, "tests", "host", "predicate_host.cpp")


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host") / "libpredicate_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, SRC], check=True)
    lib = ctypes.CDLL(out)
    dp, ip, lp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_long)
    lib.predicate_all_pairs.restype = ctypes.c_long
    lib.predicate_all_pairs.argtypes = [ctypes.c_long] + [dp] * 7 + [ctypes.c_long, ip, ip, dp, dp, lp]
    return lib


@pytest.mark.parametrize("name", [n for n in golden_names() if "n6400" not in n])
def test_host_predicate_bit_exact(hostlib, name):
    g = load_golden(name)
    n = len(g["xc"])
    arr = [np.ascontiguousarray(g[k], dtype=np.float64) for k in ("xa", "ya", "xb", "yb", "xc", "yc", "length")]
    cap = len(g["stick1"]) + 16
    s1 = np.zeros(cap, np.int32)
    s2 = np.zeros(cap, np.int32)
    jx = np.zeros(cap)
    jy = np.zeros(cap)
    nt = ctypes.c_long(0)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    hits = hostlib.predicate_all_pairs(n, *[a.ctypes.data_as(dp) for a in arr], cap, s1.ctypes.data_as(ip),
                                       s2.ctypes.data_as(ip), jx.ctypes.data_as(dp), jy.ctypes.data_as(dp),
                                       ctypes.byref(nt))
    assert hits == len(g["stick1"])
    assert np.array_equal(s1[:hits], g["stick1"]) and np.array_equal(s2[:hits], g["stick2"])   # (i, j) order
    assert np.array_equal(jx[:hits], g["jx"]) and np.array_equal(jy[:hits], g["jy"])           # bit-exact
    from oracle import cntnet_oracle as O
    from conftest import golden_sticks
    assert nt.value == O.find_junctions(golden_sticks(g))["ntests"]
