"""Builds the in-tree native library (CUDA kernels + C ABI + TensorRT-shaped plugin classes) for sm_100a.

    python -m b200_whisper._build      (or __graft_entry__.build())

nvcc cross-compiles without a GPU.  The .so lands in <package>/lib/ (git-ignored, travels with gpurun).
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
PLUG = os.path.join(PKG, "plugins")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "build")
LIBNAME = "libb200_whisper.so"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
          "-I" + PLUG]
# B200_TC_DEBUG=1: compile the clock64 / %globaltimer stamps into the tcgen05 GEMM (tools/tc_timing*.py,
# tools/gemm_timeline.py).  Off by default: the stamps cost about 1 % of the decoder step even when unused.
TC_DEBUG = os.environ.get("B200_TC_DEBUG", "0") == "1"
# B200_DS_DEBUG=1: compile the per-phase %globaltimer / clock64 stamps into the persistent decoder-step kernel
# (tools/step_phases.py).
DS_DEBUG = os.environ.get("B200_DS_DEBUG", "0") == "1"
DEBUG_STAMPS = TC_DEBUG or DS_DEBUG
if TC_DEBUG:
    COMMON = COMMON + ["-DB200_TC_DEBUG=1"]
if DS_DEBUG:
    COMMON = COMMON + ["-DB200_DS_DEBUG=1"]
# experiment switches (bisecting): B200_EXTRA_DEFS="-DX -DY"; -DB200_XA_DEBUG=1 compiles per-CTA %globaltimer stamps into
# the whole-pair cross-attention kernel (tools/xa_timeline.py)
EXTRA_DEFS = os.environ.get("B200_EXTRA_DEFS", "").split()
if EXTRA_DEFS:
    COMMON = COMMON + EXTRA_DEFS
    DEBUG_STAMPS = True


def _sources():
    srcs = []
    for d in (CSRC, PLUG):
        if os.path.isdir(d):
            for root, _, files in os.walk(d):
                for f in sorted(files):
                    if f.endswith((".cu", ".cpp")):
                        srcs.append(os.path.join(root, f))
    return srcs


def _headers_stamp():
    h = hashlib.sha256()
    for d in (CSRC, PLUG, os.path.join(ROOT, "include")):
        if not os.path.isdir(d):
            continue
        for root, _, files in os.walk(d):
            for f in sorted(files):
                if f.endswith((".h", ".cuh", ".hpp")):
                    p = os.path.join(root, f)
                    h.update(p.encode())
                    h.update(str(os.path.getmtime(p)).encode())
    return h.hexdigest()[:16]


def _compile(src, stamp, verbose):
    rel = os.path.relpath(src, PKG).replace(os.sep, "_")
    obj = os.path.join(OBJDIR, f"{rel}.{stamp}{'.tcdbg' if TC_DEBUG else ''}{'.dsdbg' if DS_DEBUG else ''}{('.' + hashlib.sha256(' '.join(EXTRA_DEFS).encode()).hexdigest()[:8]) if EXTRA_DEFS else ''}.o")
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return obj
    cmd = [NVCC] + ARCH + COMMON + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    stamp = _headers_stamp()
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, stamp, verbose), srcs))
    out = os.path.join(LIBDIR, LIBNAME)
    flavour = os.path.join(LIBDIR, ".flavour")
    want = ("tc-debug " if TC_DEBUG else "") + ("ds-debug " if DS_DEBUG else "") + " ".join(EXTRA_DEFS) if DEBUG_STAMPS else "release"
    have = open(flavour).read().strip() if os.path.exists(flavour) else ""
    if (not os.path.exists(out)) or have != want or any(os.path.getmtime(o) > os.path.getmtime(out) for o in objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", out] + objs + ["-Xlinker", "--version-script=" + os.path.join(PKG, "exports.map")]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(flavour, "w") as f:
            f.write(want)
    # drop stale objects
    keep = set(objs)
    for f in os.listdir(OBJDIR):
        p = os.path.join(OBJDIR, f)
        if p not in keep:
            os.remove(p)
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
