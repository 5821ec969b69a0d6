"""CUDA engine: the Python face of the C ABI (one ctx per process / GPU)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from ._lib import BesstLibraryError, load

STAGE_NAMES = ["extract_links", "radix_bucket", "segment_heads", "edge_reduce", "ks_gapest_score"]


class CudaEngine(object):
    name = "cuda"

    def __init__(self, device=-1):
        self._L = load()
        self._ctx = self._L.besst_create(int(device))
        if not self._ctx:
            raise BesstLibraryError("besst_create failed: %s (no CPU fallback)"
                                    % self._L.besst_last_error(None).decode())
        self._table_id = None
        self._slot = 0
        self._slot_contigs = {}

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.besst_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise BesstLibraryError("%s failed (%d, %s): %s" % (
                what, rc, abi.ERRORS.get(rc, "?"), self._L.besst_last_error(self._ctx).decode()))
        return rc

    # -- contig table ---------------------------------------------------------------
    def set_contigs(self, rows, n_scaffolds, n_large_scaffolds):
        rows = np.ascontiguousarray(rows, dtype=abi.CONTIG_ROW_DTYPE)
        self._check(self._L.besst_set_contigs(self._ctx, rows.ctypes.data, rows.shape[0], int(n_scaffolds),
                                              int(n_large_scaffolds)), "besst_set_contigs")
        self._n_contigs = rows.shape[0]
        self._slot_contigs[self._slot] = rows.shape[0]

    def select_table(self, slot):
        """Make contig table `slot` (0..7) current: set_contigs writes it, builds read it.  The libraries of a run
        see different Contigs/Scaffolds states (runBESST:143-231); their tables stay resident side by side."""
        self._check(self._L.besst_contigs_select(self._ctx, int(slot)), "besst_contigs_select")
        self._slot = int(slot)
        self._n_contigs = self._slot_contigs.get(self._slot, 0)
        self._table_id = None

    def set_table(self, table):
        self.set_contigs(table.rows, table.n_scaffolds, table.n_large_scaffolds)

    # -- graph build ------------------------------------------------------------------
    def build(self, params, records):
        """records: abi.Records (host or device pointers).  Leaves the result in HBM."""
        sizes = abi.GraphSizes()
        self._check(self._L.besst_graph_build(self._ctx, C.byref(params), C.byref(records), C.byref(sizes)),
                    "besst_graph_build")
        return sizes

    def fetch(self, sizes):
        out, arrays = abi.alloc_graph_out(sizes)
        self._check(self._L.besst_graph_fetch(self._ctx, C.byref(out)), "besst_graph_fetch")
        return abi.graph_result(out, arrays)

    def fetch_view(self, sizes):
        """Like fetch, but the arrays are zero-copy views of pinned buffers owned by the engine:
        valid until the next fetch_view / close.  The fast path for a caller that consumes the
        result right away (CreateGraph.PE materialising networkx graphs, bench.py)."""
        out = abi.GraphOut()
        self._check(self._L.besst_graph_view(self._ctx, C.byref(out)), "besst_graph_view")
        return abi.graph_result(out, abi.view_graph_out(out, sizes))

    def graph_build(self, table, params, batch, view=False):
        """Host-buffer call used by CreateGraph.PE: table + RecordBatch in, GraphResult out.
        view=True: the arrays are views of the engine's pinned buffers (valid until the next view)."""
        self.set_table(table)
        keep = []
        rec = abi.make_records(batch, keepalive=keep)
        sizes = self.build(params, rec)
        return self.fetch_view(sizes) if view else self.fetch(sizes)

    def make_dist_backend(self, table):
        """The device side of besst_b200.dist.DistributedGraphBuild for this engine (CreateGraph.PE under torchrun)."""
        import torch
        from .dist import CudaBackend
        self.set_table(table)
        dev = torch.device("cuda", torch.cuda.current_device())
        return CudaBackend(self, dev)

    # -- the two halves around the multi-GPU exchange --------------------------------------
    def links_extract(self, params, records):
        n = C.c_int64()
        self._check(self._L.besst_links_extract(self._ctx, C.byref(params), C.byref(records), C.byref(n)),
                    "besst_links_extract")
        return n.value

    def links_device(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self._L.besst_links_tuples_device(self._ctx, C.byref(p), C.byref(n)), "besst_links_tuples_device")
        fp, fn = C.c_void_p(), C.c_int64()
        self._check(self._L.besst_links_fishy_device(self._ctx, C.byref(fp), C.byref(fn)), "besst_links_fishy_device")
        return (p.value or 0, n.value), (fp.value or 0, fn.value)

    def fishy_device(self):
        """(device pointer, count) of the fishy keys of the last extraction (does not materialise the tuple array)."""
        fp, fn = C.c_void_p(), C.c_int64()
        self._check(self._L.besst_links_fishy_device(self._ctx, C.byref(fp), C.byref(fn)), "besst_links_fishy_device")
        return fp.value or 0, fn.value

    def links_partials(self):
        aligned = np.zeros(self._n_contigs, dtype=np.int64)
        counters = np.zeros(abi.N_COUNTERS, dtype=np.int64)
        self._check(self._L.besst_links_partials(self._ctx, aligned.ctypes.data, counters.ctypes.data),
                    "besst_links_partials")
        return aligned, counters

    def links_partials_device(self):
        """(device pointer of aligned_len[C] int64, device pointer of counters[16] int64)"""
        a, c = C.c_void_p(), C.c_void_p()
        self._check(self._L.besst_links_partials_device(self._ctx, C.byref(a), C.byref(c)), "besst_links_partials_device")
        return a.value, c.value

    def links_tuples_host(self):
        (_, n), _ = self.links_device()
        out = np.zeros(n, dtype=abi.LINK_TUPLE_DTYPE)
        self._check(self._L.besst_links_fetch(self._ctx, out.ctypes.data, None), "besst_links_fetch")
        return out

    def links_fishy_host(self):
        _, (_, n) = self.links_device()
        out = np.zeros(n, dtype=np.uint64)
        self._check(self._L.besst_links_fetch(self._ctx, None, out.ctypes.data), "besst_links_fetch")
        return out

    def links_partition(self, world, out_tuples_ptr, out_fishy_ptr, out_ordinals_ptr=None):
        """Stable bucketing of the extracted tuples / fishy keys by destination rank into
        caller-provided device buffers.  -> (tuple_counts[world], fishy_counts[world])"""
        tc = np.zeros(world, dtype=np.int64)
        fc = np.zeros(world, dtype=np.int64)
        self._check(self._L.besst_links_partition(self._ctx, int(world), out_tuples_ptr, out_ordinals_ptr, out_fishy_ptr,
                                                  tc.ctypes.data, fc.ctypes.data), "besst_links_partition")
        return tc, fc

    # -- run-level exchange (preferred multi-GPU path) ----------------------------------------------
    def links_group(self):
        """Group the extracted links into runs.  -> number of runs, or None when the stream has no
        local order (use the tuple-level exchange)."""
        n = C.c_int64()
        rc = self._check(self._L.besst_links_group(self._ctx, C.byref(n)), "besst_links_group")
        return None if rc == 1 else n.value

    def exchange_prepare(self, world, out_fishy_ptr):
        """group + route counts + fishy partition with one host read.  -> (summary int64[8], link_counts, run_counts, fishy_counts)"""
        summary = np.zeros(8, dtype=np.int64)
        lc, rc_, fc = (np.zeros(world, dtype=np.int64) for _ in range(3))
        self._check(self._L.besst_exchange_prepare(self._ctx, int(world), out_fishy_ptr, summary.ctypes.data, lc.ctypes.data,
                                                   rc_.ctypes.data, fc.ctypes.data), "besst_exchange_prepare")
        return summary, lc, rc_, fc

    def runs_route(self, world):
        lc = np.zeros(world, dtype=np.int64)
        rc = np.zeros(world, dtype=np.int64)
        self._check(self._L.besst_runs_route(self._ctx, int(world), lc.ctypes.data, rc.ctypes.data), "besst_runs_route")
        return lc, rc

    def runs_pack(self, world, out_obs_ptr, out_desc_ptr):
        self._check(self._L.besst_runs_pack(self._ctx, int(world), out_obs_ptr, out_desc_ptr), "besst_runs_pack")

    def runs_obs_bytes(self, params):
        return int(self._L.besst_runs_obs_bytes(C.byref(params)))

    def runs_pack_peer(self, world, obs_ptrs, desc_ptrs):
        """obs_ptrs[d] / desc_ptrs[d]: device addresses (ints) of this rank's segment in destination d's buffers."""
        a = (C.c_void_p * world)(*[int(p) for p in obs_ptrs])
        b = (C.c_void_p * world)(*[int(p) for p in desc_ptrs])
        self._check(self._L.besst_runs_pack_peer(self._ctx, int(world), a, b), "besst_runs_pack_peer")

    def runs_to_graph(self, params, obs_ptr, n_links, desc_ptr, n_runs, world, block_bits, src_run_counts, src_link_counts,
                      src_first_base, fishy_ptr, n_fishy):
        sizes = abi.GraphSizes()
        a = np.ascontiguousarray(src_run_counts, dtype=np.int64)
        b = np.ascontiguousarray(src_link_counts, dtype=np.int64)
        c = np.ascontiguousarray(src_first_base, dtype=np.int64)
        self._check(self._L.besst_runs_to_graph(self._ctx, C.byref(params), obs_ptr, int(n_links), desc_ptr, int(n_runs), int(world),
                                                int(block_bits), a.ctypes.data, b.ctypes.data, c.ctypes.data, fishy_ptr, int(n_fishy),
                                                C.byref(sizes)), "besst_runs_to_graph")
        return sizes

    def links_counters(self):
        counters = np.zeros(abi.N_COUNTERS, dtype=np.int64)
        self._check(self._L.besst_links_partials(self._ctx, None, counters.ctypes.data), "besst_links_partials")
        return counters

    def set_stream(self, cuda_stream_ptr):
        """Run the engine on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0/None restores its own."""
        self._check(self._L.besst_set_stream(self._ctx, cuda_stream_ptr or None), "besst_set_stream")

    def links_to_graph(self, params, tuples_ptr, n_tuples, fishy_ptr, n_fishy):
        sizes = abi.GraphSizes()
        self._check(self._L.besst_links_to_graph(self._ctx, C.byref(params), tuples_ptr, int(n_tuples), fishy_ptr,
                                                 int(n_fishy), C.byref(sizes)), "besst_links_to_graph")
        return sizes

    # -- BAM ingest on the device -----------------------------------------------------------
    def ingest_bam(self, path, head_records=1000, check_crc=True, blind_seeds=False, part=(0, 1), start_voffset=-1):
        """BAM file -> DeviceRecordBatch: BGZF inflate and record decode on the GPU, the columns stay in HBM (owned by the
        engine, valid until the next ingest_bam / close).  Pass the batch to libmetrics / graph_build like a RecordBatch.
        part = (r, n): only the r-th of n parts of the file (multi-GPU ingest, besst_bam_ingest_part); the batch then
        carries first_voffset / landing_voffset for the consistency check between neighbouring parts
        (besst_b200.dist.ingest_bam_distributed does the whole protocol)."""
        import os
        from .records import DeviceRecordBatch
        rec, st = abi.Records(), abi.BamIngestStats()
        flags = (0 if check_crc else abi.BAM_NO_CRC) | (abi.BAM_BLIND_SEEDS if blind_seeds else 0)
        first, landing = C.c_int64(-1), C.c_int64(-1)
        self._check(self._L.besst_bam_ingest_part(self._ctx, os.fsencode(path), int(head_records), flags, int(part[0]), int(part[1]),
                                                  int(start_voffset), C.byref(rec), C.byref(st), C.byref(first), C.byref(landing)),
                    "besst_bam_ingest_part")
        n_ref = int(self._L.besst_bam_ingest_n_refs(self._ctx))
        references = [self._L.besst_bam_ingest_ref_name(self._ctx, i).decode("ascii") for i in range(n_ref)]
        lengths = [int(self._L.besst_bam_ingest_ref_length(self._ctx, i)) for i in range(n_ref)]
        nh = min(int(rec.n), int(head_records))
        rlen, alen = np.zeros(nh, np.int32), np.zeros(nh, np.int32)
        got = self._L.besst_bam_ingest_head(self._ctx, rlen.ctypes.data, alen.ctypes.data, nh)
        if got != nh:
            raise BesstLibraryError("besst_bam_ingest_head returned %d, expected %d" % (got, nh))
        batch = DeviceRecordBatch(self, rec, references, lengths, rlen, alen, {k: getattr(st, k) for k, _ in abi.BamIngestStats._fields_})
        batch.first_voffset, batch.landing_voffset = int(first.value), int(landing.value)
        return batch

    def device_read(self, ptr, count, dtype):
        """numpy copy of `count` elements at device address `ptr` (columns the engine owns)."""
        out = np.empty(int(count), dtype=dtype)
        if count:
            self._check(self._L.besst_device_read(self._ctx, ptr, out.ctypes.data, out.nbytes), "besst_device_read")
        return out

    # -- library metrics ------------------------------------------------------------------
    def libmetrics(self, rows, params, batch, ref_lengths, want_isize, cap=1 << 20, records=None):
        """-> (rc, abi.LibMetricsOut, adjusted_distribution).  rc == 1: fewer than
        1001 insert-size samples (libmetrics.py:311-314)."""
        self.set_contigs(rows, 1, 0)
        keep = []
        rec = records if records is not None else abi.make_records(batch, keepalive=keep)
        lens = np.ascontiguousarray(ref_lengths, dtype=np.int64)
        out = abi.LibMetricsOut()
        adj = np.zeros(cap, dtype=np.float64)
        rc = self._check(self._L.besst_libmetrics(self._ctx, C.byref(params), C.byref(rec), lens.ctypes.data,
                                                  lens.shape[0], int(want_isize), C.byref(out), adj.ctypes.data, cap),
                         "besst_libmetrics")
        return rc, out, adj[:min(cap, out.n_bins)]

    # -- batched GapEstimator --------------------------------------------------------------
    def gapest_batch(self, params, mean_obs, len1, len2):
        mean_obs = np.ascontiguousarray(mean_obs, dtype=np.float64)
        len1 = np.ascontiguousarray(len1, dtype=np.float64)
        len2 = np.ascontiguousarray(len2, dtype=np.float64)
        gap = np.zeros(mean_obs.shape[0], dtype=np.int32)
        sd = np.zeros(mean_obs.shape[0], dtype=np.float64)
        self._check(self._L.besst_gapest_batch(self._ctx, C.byref(params), mean_obs.ctypes.data, len1.ctypes.data,
                                               len2.ctypes.data, mean_obs.shape[0], gap.ctypes.data, sd.ctypes.data),
                    "besst_gapest_batch")
        return gap, sd

    def gapest_lognormal_batch(self, mu, sigma, read_len, samples, row_ptr, len1, len2):
        """ML gaps under the lognormal model from the raw observations of every edge (CSR layout)."""
        samples = np.ascontiguousarray(samples, dtype=np.int32)
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        len1 = np.ascontiguousarray(len1, dtype=np.float64)
        len2 = np.ascontiguousarray(len2, dtype=np.float64)
        n = row_ptr.shape[0] - 1
        gap = np.zeros(n, dtype=np.int32)
        self._check(self._L.besst_gapest_lognormal_batch(self._ctx, float(mu), float(sigma), float(read_len), samples.ctypes.data,
                                                         row_ptr.ctypes.data, len1.ctypes.data, len2.ctypes.data, n, gap.ctypes.data),
                    "besst_gapest_lognormal_batch")
        return gap

    def func_of_d_batch(self, params, d, len1, len2):
        d = np.ascontiguousarray(d, dtype=np.float64)
        len1 = np.ascontiguousarray(len1, dtype=np.float64)
        len2 = np.ascontiguousarray(len2, dtype=np.float64)
        out = np.zeros(d.shape[0], dtype=np.float64)
        self._check(self._L.besst_gapest_func_batch(self._ctx, C.byref(params), d.ctypes.data, len1.ctypes.data,
                                                    len2.ctypes.data, d.shape[0], out.ctypes.data), "besst_gapest_func_batch")
        return out

    def trsk_sd_batch(self, params, gap, len1, len2):
        gap = np.ascontiguousarray(gap, dtype=np.float64)
        len1 = np.ascontiguousarray(len1, dtype=np.float64)
        len2 = np.ascontiguousarray(len2, dtype=np.float64)
        sd = np.zeros(gap.shape[0], dtype=np.float64)
        self._check(self._L.besst_trsk_sd_batch(self._ctx, C.byref(params), gap.ctypes.data, len1.ctypes.data,
                                                len2.ctypes.data, gap.shape[0], sd.ctypes.data), "besst_trsk_sd_batch")
        return sd

    # -- timing / accounting -------------------------------------------------------------
    def timing(self):
        total = C.c_float()
        stages = (C.c_float * abi.N_STAGES)()
        self._check(self._L.besst_last_timing(self._ctx, C.byref(total), stages), "besst_last_timing")
        return total.value, dict(zip(STAGE_NAMES, list(stages)[:len(STAGE_NAMES)]))

    def set_profiling(self, on, accumulate=False):
        self._check(self._L.besst_set_profiling(self._ctx, 2 if (on and accumulate) else int(bool(on))), "besst_set_profiling")

    def kernel_profile(self, cap=1 << 16):
        """[(kernel name, ms)] per launch of the last build (needs set_profiling(True))."""
        ids = np.zeros(cap, dtype=np.int32)
        ms = np.zeros(cap, dtype=np.float32)
        n = self._check(self._L.besst_kernel_profile(self._ctx, ids.ctypes.data, ms.ctypes.data, cap), "besst_kernel_profile")
        return [(KERNEL_NAMES[int(ids[i])], float(ms[i])) for i in range(n)]

    def kernel_launches(self):
        n = C.c_int64()
        self._check(self._L.besst_kernel_launches(self._ctx, C.byref(n)), "besst_kernel_launches")
        return n.value


KERNEL_NAMES = ["k_extract_links", "k_radix_hist", "k_radix_scan_hist", "k_radix_sweep", "k_heads", "k_edge_reduce",
                "k_score_keys", "k_fishy_rekey", "k_metrics", "k_gapest", "k_tile_scan", "k_compact_tuples",
                "k_partition", "k_ks_eval", "k_ks_sort", "k_group_blocks", "k_runs", "k_ks_block", "k_bgzf_inflate", "k_bam_scan",
                "k_bam_decode"]

_default = None


def default_engine():
    """Process-wide engine on the current CUDA device.  Raises if the CUDA
    library or a GPU is missing -- the product path never falls back to a CPU
    implementation."""
    global _default
    if _default is None:
        _default = CudaEngine()
    return _default
