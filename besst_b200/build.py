"""Build libbesst_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so
travels with the source tree).  `python -m besst_b200.build`"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libbesst_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# per-file extra flags: the scoring math must round like the reference (no FMA contraction)
SOURCES = [("besst_api.cu", []), ("besst_links.cu", []), ("besst_sort.cu", []),
           ("besst_edges.cu", ["-fmad=false"]), ("besst_metrics.cu", []), ("besst_bamdev.cu", []), ("besst_paths.cu", [])]


BAMIO_SO = os.path.join(HERE, "libbesst_bamio.so")


def build_bamio(force=False):
    """Host-only ingest library (g++, zlib, threads): BAM file -> record columns."""
    src = os.path.join(CSRC, "besst_bamio.cpp")
    hdr = os.path.join(HERE, "..", "include", "besst_bamio.h")
    if force or _stale(BAMIO_SO, [src, hdr]):
        cxx = os.environ.get("CXX", "g++")
        cmd = [cxx, "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wall", "-Wextra", "-o", BAMIO_SO, src, "-lz"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed on besst_bamio.cpp")
    return BAMIO_SO


HOSTCHECK_SO = os.path.join(HERE, "libbesst_bgzf_hostcheck.so")


def build_hostcheck(force=False):
    """TEST TOOLING: the device ingest's inflate / CRC / record scan / window loop compiled for the host (lane-serial),
    so the CPU suite can check the code the GPU runs against zlib.  No product path loads it."""
    deps = [os.path.join(CSRC, f) for f in ("bgzf_hostcheck.cpp", "bgzf_core.cuh", "bam_ingest.hpp")]
    if force or _stale(HOSTCHECK_SO, deps):
        cxx = os.environ.get("CXX", "g++")
        cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Wextra", "-o", HOSTCHECK_SO, deps[0]]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed on bgzf_hostcheck.cpp")
    return HOSTCHECK_SO


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    headers = [os.path.join(CSRC, "besst_internal.cuh"), os.path.join(HERE, "..", "include", "besst_b200.h"),
               os.path.join(CSRC, "bgzf_core.cuh"), os.path.join(CSRC, "bam_ingest.hpp"), os.path.join(CSRC, "paths_core.cuh")]
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    objs, log = [], []
    for name, extra in SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(objdir, name.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s" % name)
            if verbose:
                sys.stderr.write(r.stderr)
        objs.append(obj)
    if force or _stale(SO, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", SO] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    with open(os.path.join(objdir, "ptxas.log"), "a") as fh:
        fh.write("".join(log))
    build_bamio(force)
    build_hostcheck(force)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
