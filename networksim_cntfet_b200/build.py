"""In-tree build of libcntnet_b200.so with nvcc for sm_100a (no JIT cache: the built .so
travels with the repo snapshot to the GPU box).

    python -m networksim_cntfet_b200.build [--force] [--verbose]
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.environ.get("CNT_B200_OBJDIR") or os.path.join(HERE, "_build")     # variants for A/B runs: own object directory
LIB = os.environ.get("CNT_B200_LIB") or os.path.join(HERE, "libcntnet_b200.so")   # ... and library name
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["util.cu", "radix.cu", "sticks.cu", "junctions.cu", "clusters.cu", "prune.cu", "assemble.cu", "aggregates.cu", "pcg.cu", "pcg_cluster.cu", "pcg_slab.cu", "spmv.cu", "starmesh.cu", "starmesh_sym.cu",
           "currents.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE]
# bit-exact FP64 predicate: no FMA contraction anywhere in that translation unit (the kernels
# also use explicit __d*_rn intrinsics; this is belt and braces)
PER_FILE = {"junctions.cu": ["-fmad=false"]}
if os.environ.get("CNT_TILE_FLAGS"):          # A/B of the tile kernel's shape: e.g. "-DCNT_TILE_CT=256 -DCNT_TILE_Q=1024"
    PER_FILE["junctions.cu"] = PER_FILE["junctions.cu"] + os.environ["CNT_TILE_FLAGS"].split()
if os.environ.get("CNT_UF_K"):                # A/B: junctions per thread of the union pass
    PER_FILE["clusters.cu"] = ["-DCNT_UF_K=" + os.environ["CNT_UF_K"]]
if os.environ.get("CNT_CLUSTER_THREADS"):
    PER_FILE["pcg_cluster.cu"] = ["-DCNT_CLUSTER_THREADS=" + os.environ["CNT_CLUSTER_THREADS"]]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths, extra):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def sources():
    return [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns the .so path."""
    nvcc = nvcc_path()
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "cntnet_b200.h"))
    objs = []
    jobs = []
    for src in sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        flags = COMMON + ARCH + PER_FILE.get(src, [])
        if verbose:
            flags = flags + ["-Xptxas", "-v"]
        stamp = obj + ".sha"
        dig = _digest([path] + headers, " ".join(flags))
        objs.append(obj)
        if not force and os.path.isfile(obj) and os.path.isfile(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((src, [nvcc] + flags + ["-c", path, "-o", obj], stamp, dig))

    def compile_one(job):
        src, cmd, stamp, dig = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
        open(stamp, "w").write(dig)

    if jobs:       # translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1, 8)) as pool:
            list(pool.map(compile_one, jobs))
    rebuilt = bool(jobs)
    if rebuilt or force or not os.path.isfile(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
