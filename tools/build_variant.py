"""Development aid: builds lib/libmagi_<name>.so = the objects of the fast build with flow_inst_0.cu (FN) recompiled with
extra -D flags (plus -DMAGI_DEV_KNOBS: the getenv-driven A/B switches and phase clocks, compiled out of the product library), so that kernel variants can be measured side by side in ONE gpurun call (MAGI_LIB_NAME selects the library).
usage: python tools/build_variant.py <name> [--srcs=a.cu,b.cu] [-DFOO=1 ...]     (--srcs: recompile these sources instead, e.g. gemm_f64.cu)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "manifold_constrained_gaussian_process_inference_b200")
sys.path.insert(0, ROOT)
from manifold_constrained_gaussian_process_inference_b200 import build as B

name, defs = sys.argv[1], sys.argv[2:]
srcs = ("flow_inst_0.cu", "banded_inst_0.cu", "magi_abi.cu")       # the FN kernels and the dispatch code (MAGI_DEV_KNOBS lives there)
for a in list(defs):
    if a.startswith("--srcs="):
        srcs = tuple(a[7:].split(","))
        defs.remove(a)
B.build(fast=True)                                           # refreshes build_fast/*.o (and the default fast .so)
objdir = os.path.join(PKG, "build_fast")
vdir = os.path.join(PKG, "build_var", name)
os.makedirs(vdir, exist_ok=True)
objs = []
for src in srcs:
    obj = os.path.join(vdir, src[:-3] + ".o")
    subprocess.check_call([B.NVCC] + B.FLAGS + ["-DMAGI_FAST_BUILD", "-DMAGI_DEV_KNOBS"] + defs + ["-c", os.path.join(B.CSRC, src), "-o", obj])
    objs.append(obj)
skip = {os.path.basename(o) for o in objs}
others = [os.path.join(objdir, f) for f in sorted(os.listdir(objdir)) if f.endswith(".o") and f not in skip]
lib = os.path.join(PKG, "lib", "libmagi_%s.so" % name)
subprocess.check_call([B.NVCC, "-shared", "-o", lib] + others + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
print(lib)
