"""The C-ABI library loads without a GPU and exports every entry point that
include/besst_b200.h declares; the ctypes struct mirrors have the C layout; the
product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from besst_b200 import _lib, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "besst_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(besst_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.SO_PATH):
        from besst_b200 import build
        build.build()
    L = C.CDLL(_lib.SO_PATH)
    names = declared_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(L, name), "libbesst_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == names, "besst_b200/_lib.py EXPORTS out of sync with the header"
    L.besst_abi_version.restype = C.c_int
    assert L.besst_abi_version() == abi.ABI_VERSION


def test_bamio_library_exports_every_declared_symbol():
    from besst_b200 import bamio, build
    if not os.path.exists(bamio.BAMIO_SO):
        build.build_bamio()
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "besst_bamio.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(besst_[a-z0-9_]+)\s*\(", text)))
    L = C.CDLL(bamio.BAMIO_SO)
    for name in names:
        assert hasattr(L, name), "libbesst_bamio.so does not export %s" % name
    assert sorted(bamio.BAMIO_EXPORTS) == names
    L.besst_bamio_abi_version.restype = C.c_int
    assert L.besst_bamio_abi_version() == 2


def test_ctypes_structs_match_c_layout():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "besst_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(besst_contig_row), sizeof(besst_records), sizeof(besst_lib_params),
         sizeof(besst_graph_sizes), sizeof(besst_graph_out), sizeof(besst_link_tuple), sizeof(besst_libmetrics_out));
  printf("%zu %zu %zu %zu\n", offsetof(besst_lib_params, read_len), offsetof(besst_lib_params, halo_prev_obs1),
         offsetof(besst_records, on_device), offsetof(besst_graph_out, counters));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "layout.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "layout")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(x) for x in out]
    assert sizes[:7] == [abi.CONTIG_ROW_DTYPE.itemsize, C.sizeof(abi.Records), C.sizeof(abi.LibParams),
                         C.sizeof(abi.GraphSizes), C.sizeof(abi.GraphOut), abi.LINK_TUPLE_DTYPE.itemsize,
                         C.sizeof(abi.LibMetricsOut)]
    assert sizes[7:] == [abi.LibParams.read_len.offset, abi.LibParams.halo_prev_obs1.offset,
                         abi.Records.on_device.offset, abi.GraphOut.counters.offset]


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from besst_b200.engine import CudaEngine
    with pytest.raises(_lib.BesstLibraryError) as e:
        CudaEngine()
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "besst_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|oracle_lib|oracle_engine|ref_harness)\b", text, flags=re.M), f
                assert "besst_oracle_" not in text, f
