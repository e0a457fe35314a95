"""Builds libmagi_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, os.environ.get("MAGI_OBJ_DIR", "build"))
LIB = os.path.join(LIBDIR, os.environ.get("MAGI_LIB_NAME", "libmagi_b200.so"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime(path, seen=None):
    """Newest modification time of `path` and of every header it includes (transitively, quoted includes only)."""
    import re
    seen = seen if seen is not None else set()
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return 0.0
    seen.add(path)
    m = os.path.getmtime(path)
    with open(path) as f:
        for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', f.read(), flags=re.M):
            m = max(m, _deps_mtime(os.path.join(os.path.dirname(path), inc), seen))
    return m


def _compile(src, extra):
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) >= _deps_mtime(srcp) and not extra.get("force"):
        return obj, False
    cmd = [NVCC] + FLAGS + extra.get("defs", []) + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, True


def build(force: bool = False, fast: bool = False, verbose: bool = False) -> str:
    global OBJDIR, LIB
    if fast and "MAGI_OBJ_DIR" not in os.environ:
        OBJDIR = os.path.join(HERE, "build_fast")      # objects compiled with -DMAGI_FAST_BUILD must never be linked into a full build
    if fast and "MAGI_LIB_NAME" not in os.environ:
        LIB = os.path.join(LIBDIR, "libmagi_fast.so")  # ... and the development library never replaces the product library
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    extra = {"force": force, "defs": (["-DMAGI_FAST_BUILD"] if fast else []) + os.environ.get("MAGI_EXTRA_DEFS", "").split()}
    srcs = sources()
    if fast:   # development builds: only the FN / Hes1 / LV kernels (the full build adds the other models)
        keep = ("_inst_0.cu", "_inst_1.cu", "_inst_7.cu")
        srcs = [s_ for s_ in srcs if not (s_.startswith(("banded_inst_", "flow_inst_")) and not s_.endswith(keep))]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, extra), srcs))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not os.path.exists(LIB) or force:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, fast="--fast" in sys.argv, verbose=True))
