"""ctypes loader for libbesst_b200.so -- the thin shim between the Python entry
points and the CUDA engine, in the style of the reference's own ctypes
precedent (BESST/diploid/wrapper_sw.py:12-24).  Fails loudly: there is no CPU
fallback behind this module."""
from __future__ import annotations

import ctypes as C
import os

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BESST_B200_LIB") or os.path.join(HERE, "libbesst_b200.so")   # override: A/B builds of the same ABI

EXPORTS = ["besst_abi_version", "besst_create", "besst_destroy", "besst_last_error", "besst_set_contigs",
           "besst_graph_build", "besst_graph_fetch", "besst_links_extract", "besst_links_tuples_device",
           "besst_links_fishy_device", "besst_links_partials", "besst_links_to_graph", "besst_libmetrics",
           "besst_gapest_batch", "besst_last_timing", "besst_kernel_launches", "besst_set_profiling",
           "besst_kernel_profile", "besst_links_partials_device", "besst_links_fetch", "besst_links_partition",
           "besst_trsk_sd_batch", "besst_set_stream", "besst_graph_view", "besst_links_group", "besst_runs_route",
           "besst_runs_pack", "besst_runs_to_graph", "besst_runs_pack_peer", "besst_gapest_func_batch", "besst_runs_obs_bytes", "besst_contigs_select", "besst_csr_prune_dense", "besst_gapest_lognormal_batch", "besst_exchange_prepare",
           "besst_bam_ingest", "besst_bam_ingest_part", "besst_bam_ingest_n_refs", "besst_bam_ingest_ref_name", "besst_bam_ingest_ref_length",
           "besst_bam_ingest_head", "besst_device_read",
           "besst_paths_between", "besst_paths_count", "besst_paths_hit_threshold", "besst_paths_pops", "besst_paths_arrays", "besst_paths_free", "besst_scaffold_prune_ambiguous"]

_lib = None


class BesstLibraryError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise BesstLibraryError(
            "libbesst_b200.so is not built (%s). Build it with `python -m besst_b200.build` (needs nvcc, sm_100a). "
            "There is no CPU fallback for this path." % SO_PATH)
    L = C.CDLL(SO_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise BesstLibraryError("libbesst_b200.so does not export %s (stale build?)" % name)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.besst_abi_version.restype = C.c_int
    L.besst_create.restype = vp
    L.besst_create.argtypes = [C.c_int]
    L.besst_destroy.argtypes = [vp]
    L.besst_last_error.restype = C.c_char_p
    L.besst_last_error.argtypes = [vp]
    L.besst_set_contigs.argtypes = [vp, vp, i64, i64, i64]
    L.besst_contigs_select.argtypes = [vp, i32]
    L.besst_exchange_prepare.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.besst_gapest_lognormal_batch.argtypes = [vp, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp, i64, vp]
    L.besst_csr_prune_dense.restype = i64
    L.besst_csr_prune_dense.argtypes = [i64, vp, vp, vp, i32, vp]
    L.besst_graph_build.argtypes = [vp, C.POINTER(abi.LibParams), C.POINTER(abi.Records), C.POINTER(abi.GraphSizes)]
    L.besst_graph_fetch.argtypes = [vp, C.POINTER(abi.GraphOut)]
    L.besst_links_extract.argtypes = [vp, C.POINTER(abi.LibParams), C.POINTER(abi.Records), C.POINTER(i64)]
    L.besst_links_tuples_device.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.besst_links_fishy_device.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.besst_links_partials.argtypes = [vp, vp, vp]
    L.besst_links_to_graph.argtypes = [vp, C.POINTER(abi.LibParams), vp, i64, vp, i64, C.POINTER(abi.GraphSizes)]
    L.besst_libmetrics.argtypes = [vp, C.POINTER(abi.LibParams), C.POINTER(abi.Records), vp, i64, i32,
                                   C.POINTER(abi.LibMetricsOut), vp, i64]
    L.besst_gapest_batch.argtypes = [vp, C.POINTER(abi.LibParams), vp, vp, vp, i64, vp, vp]
    L.besst_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.besst_kernel_launches.argtypes = [vp, C.POINTER(i64)]
    L.besst_set_profiling.argtypes = [vp, C.c_int]
    L.besst_kernel_profile.argtypes = [vp, vp, vp, i32]
    L.besst_links_partials_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.besst_links_fetch.argtypes = [vp, vp, vp]
    L.besst_links_partition.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.besst_trsk_sd_batch.argtypes = [vp, C.POINTER(abi.LibParams), vp, vp, vp, i64, vp]
    L.besst_set_stream.argtypes = [vp, vp]
    L.besst_graph_view.argtypes = [vp, C.POINTER(abi.GraphOut)]
    L.besst_links_group.argtypes = [vp, C.POINTER(i64)]
    L.besst_runs_route.argtypes = [vp, i32, vp, vp]
    L.besst_runs_pack.argtypes = [vp, i32, vp, vp]
    L.besst_runs_pack_peer.argtypes = [vp, i32, vp, vp]
    L.besst_runs_obs_bytes.argtypes = [C.POINTER(abi.LibParams)]
    L.besst_gapest_func_batch.argtypes = [vp, C.POINTER(abi.LibParams), vp, vp, vp, i64, vp]
    L.besst_runs_to_graph.argtypes = [vp, C.POINTER(abi.LibParams), vp, i64, vp, i64, i32, i32, vp, vp, vp, vp, i64,
                                      C.POINTER(abi.GraphSizes)]
    L.besst_bam_ingest.argtypes = [vp, C.c_char_p, i64, i32, C.POINTER(abi.Records), C.POINTER(abi.BamIngestStats)]
    L.besst_bam_ingest_part.argtypes = [vp, C.c_char_p, i64, i32, i32, i32, i64, C.POINTER(abi.Records), C.POINTER(abi.BamIngestStats),
                                        C.POINTER(i64), C.POINTER(i64)]
    L.besst_bam_ingest_n_refs.restype = i64
    L.besst_bam_ingest_n_refs.argtypes = [vp]
    L.besst_bam_ingest_ref_name.restype = C.c_char_p
    L.besst_bam_ingest_ref_name.argtypes = [vp, i64]
    L.besst_bam_ingest_ref_length.restype = i64
    L.besst_bam_ingest_ref_length.argtypes = [vp, i64]
    L.besst_bam_ingest_head.restype = i64
    L.besst_bam_ingest_head.argtypes = [vp, vp, vp, i64]
    L.besst_device_read.argtypes = [vp, vp, vp, i64]
    L.besst_paths_between.restype = vp
    L.besst_paths_between.argtypes = [i64, vp, vp, vp, vp, vp, i64, i64, C.c_double, i32, i32, i32]
    L.besst_paths_count.restype = i64
    L.besst_paths_count.argtypes = [vp]
    L.besst_paths_hit_threshold.argtypes = [vp]
    L.besst_paths_pops.restype = i64
    L.besst_paths_pops.argtypes = [vp]
    L.besst_paths_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.besst_paths_free.argtypes = [vp]
    L.besst_scaffold_prune_ambiguous.restype = i64
    L.besst_scaffold_prune_ambiguous.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp]
    if L.besst_abi_version() != abi.ABI_VERSION:
        raise BesstLibraryError("ABI version mismatch: library %d, binding %d" % (L.besst_abi_version(), abi.ABI_VERSION))
    _lib = L
    return L
