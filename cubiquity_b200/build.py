"""Builds cubiquity_b200/lib/libcubiquity_b200.so in-tree with nvcc for sm_100a.

    python -m cubiquity_b200.build [--force]

The kernels are compiled with -fmad=false: bit-exact agreement with the reference's CPU ray cast
needs `(float(p) - o) * inv` and `o + d * t` to stay un-fused (see csrc/traverse.cuh).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcubiquity_b200.so")
SOURCES = ["api.cu", "trace_kernels.cu", "wavefront_kernels.cu", "viewer_kernels.cu", "voxelize_kernels.cu", "bake_kernels.cu", "edit_kernels.cu", "edit.cpp", "host_shim.cpp", "voxelize_host.cpp"]
HEADERS = ["traverse.cuh", "shading.cuh", "meshmath.cuh", "cbq_internal.h", os.path.join("..", "..", "include", "cubiquity_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-fmad=false",                      # arithmetic contract, see module docstring
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function,-ffp-contract=off",   # host code: no fused multiply-add either (host_shim.cpp)
    "--diag-suppress", "549",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cubiquity_b200 has no CPU build")


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build_library(force=False, verbose=False, defines=(), name=None):
    """`defines`/`name` build an experimental variant (lib/<name>.so) for A/B runs; the default build
    takes neither."""
    global LIB
    lib = LIB if name is None else os.path.join(LIBDIR, name + ".so")
    if name is None and not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = find_nvcc()
    objs = []
    procs = []
    objdir = os.path.join(LIBDIR, "obj" if name is None else "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    # link next to the target and rename: a reader (or a repository snapshot) never sees a half-written library
    tmp = lib + ".partial"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs + ["-lpthread"], check=True)
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    names = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--name=")]
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, name=names[0] if names else None))
