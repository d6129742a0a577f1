"""A/B builds of libhive_b200.so with different compile-time tunables -> hive-alphazero_b200/lib/variants/lib_<name>.so
(git-ignored; they travel to the GPU box with the snapshot).  usage: build_variants.py name:-DX=1,-DY=2 ..."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = importlib.import_module("hive-alphazero_b200._build")
out_dir = os.path.join(B.LIB_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    defs = [d for d in defs.split(",") if d]
    objs = []
    for src in B.sources():
        base = os.path.basename(src)
        obj = os.path.join(B.HERE, "build", "%s_%s.o" % (name, base[:-3]))
        flags = [f for f in B.NVCC_FLAGS if f != "-shared"] + B.EXTRA_FLAGS.get(base, []) + defs
        subprocess.check_call(["nvcc"] + flags + ["-c", "-o", obj, src])
        objs.append(obj)
    lib = os.path.join(out_dir, "lib_%s.so" % name)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs)
    print(lib)
