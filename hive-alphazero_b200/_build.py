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


def _newest_source_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), "include", "hive_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> lib/libhive_b200.so. Idempotent."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if os.environ.get("HIVE_B200_LIB"):
        return LIB_PATH                      # externally built variant: use as is
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libhive_b200.so")
    objs = []
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        flags = [f for f in NVCC_FLAGS if f != "-shared"] + EXTRA_FLAGS.get(os.path.basename(src), [])
        subprocess.check_call([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
        objs.append(obj)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs)
    return LIB_PATH
