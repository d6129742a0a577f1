"""Compiles the CUDA sources of this package for sm_100a into lib/libhive_b200.so (in-tree, so the
built library travels with the repo snapshot to the GPU box)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.environ.get("HIVE_B200_LIB") or os.path.join(LIB_DIR, "libhive_b200.so")   # override: A/B builds

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


# per-file flags: the search mirrors the reference's IEEE operation order, so no FMA contraction there
EXTRA_FLAGS = {"hive_mcts.cu": ["--fmad=false"], "hive_env.cu": ["-Xcompiler", "-mpopcnt"]}


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _source_digest():
    """sha256 over everything the library is built from (sources, headers, public header, flags): the staleness test
    is a content hash, not file times -- a checkout or a copy to another box never triggers or hides a rebuild."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(os.path.dirname(HERE), "include", "hive_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    h.update(repr((NVCC_FLAGS, sorted(EXTRA_FLAGS.items()))).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> lib/libhive_b200.so. Idempotent."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if os.environ.get("HIVE_B200_LIB"):
        return LIB_PATH                      # externally built variant: use as is
    digest, stamp = _source_digest(), LIB_PATH + ".sha256"
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libhive_b200.so")
    objs = []
    obj_dir = os.path.join(HERE, "build")
    shutil.rmtree(obj_dir, ignore_errors=True)              # no stale objects of earlier variants
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in sources():                                   # the translation units compile side by side
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        flags = [f for f in NVCC_FLAGS if f != "-shared"] + EXTRA_FLAGS.get(os.path.basename(src), [])
        procs.append((src, subprocess.Popen([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])))
        objs.append(obj)
    for src, pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("nvcc failed on " + src)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs)
    with open(stamp, "w") as f:
        f.write(digest + "\n")
    return LIB_PATH
