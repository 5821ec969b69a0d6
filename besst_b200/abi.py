"""ctypes mirror of include/besst_b200.h (struct layouts, constants) and the
numpy marshalling helpers shared by the CUDA binding (`_lib.py`) and by the
test-side oracle binding (`oracle/oracle_lib.py`)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

ABI_VERSION = 5

CTG_ABSENT, CTG_LARGE, CTG_SMALL = 0, 1, 2
ORIENT_FR, ORIENT_RF = 0, 1
ERF_AS7126, ERF_LIBM = 0, 1

CNT_COUNT, CNT_NON_UNIQUE, CNT_NON_UNIQUE_SCAF, CNT_DUPLICATES, CNT_TOO_LONG = 0, 1, 2, 3, 4
CNT_FISHY, CNT_CALLS, CNT_VALID, CNT_LAST_OBS1, CNT_LAST_OBS2 = 5, 6, 7, 8, 9
CNT_FIRST_OBS1, CNT_FIRST_OBS2 = 10, 11
CNT_POS_TILES = 12
N_COUNTERS = 16
N_STAGES = 8

EDGE_LL, EDGE_SCORED, EDGE_NEGGAP, EDGE_BIG, EDGE_CPLX = 1, 2, 4, 8, 16

ERRORS = {-1: "invalid argument", -2: "CUDA error", -3: "out of memory", -4: "bad call order",
          -5: "no CUDA device"}

CONTIG_ROW_DTYPE = np.dtype([("state", "<i4"), ("scaffold", "<i4"), ("direction", "<i4"), ("position", "<i4"),
                             ("length", "<i4"), ("scaf_length", "<i4"), ("in_largest", "<i4"), ("reserved", "<i4")])
RUN_DESC_DTYPE = np.dtype([("u", "<u4"), ("v", "<u4"), ("count", "<u4"), ("first", "<u4"), ("offset", "<u4"), ("block", "<u4")])
RUN_BLOCK = 2048   # links per grouping block (GB_TILE in besst_edges.cu)
LINK_TUPLE_DTYPE = np.dtype([("u", "<u4"), ("v", "<u4"), ("obs_u", "<i4"), ("obs_v", "<i4")])


class Records(C.Structure):
    _fields_ = [("n", C.c_int64), ("tid", C.c_void_p), ("mtid", C.c_void_p), ("pos", C.c_void_p),
                ("mpos", C.c_void_p), ("tlen", C.c_void_p), ("qlen", C.c_void_p), ("flag", C.c_void_p),
                ("mapq", C.c_void_p), ("on_device", C.c_int32), ("reserved", C.c_int32), ("packed", C.c_void_p)]


def pack_record_columns(flag, mapq, qlen):
    """flag | mapq << 12 | qlen << 20 (include/besst_b200.h BESST_PACK_RECORD): the one column the graph build reads
    instead of the three.  None when a value does not fit (flag >= 4096 or qlen >= 4096)."""
    flag = np.asarray(flag).astype(np.uint32)
    qlen = np.asarray(qlen)
    if flag.size and (int(flag.max()) >= 4096 or int(qlen.max()) >= 4096 or int(qlen.min()) < 0):
        return None
    return flag | (np.asarray(mapq).astype(np.uint32) << np.uint32(12)) | (qlen.astype(np.uint32) << np.uint32(20))


class LibParams(C.Structure):
    _fields_ = [("orientation", C.c_int32), ("min_mapq", C.c_int32), ("detect_duplicate", C.c_int32),
                ("extend_paths", C.c_int32), ("no_score", C.c_int32), ("erf_variant", C.c_int32),
                ("read_len", C.c_double), ("mean_ins_size", C.c_double), ("std_dev_ins_size", C.c_double),
                ("ins_size_threshold", C.c_double), ("halo_prev_obs1", C.c_int32), ("halo_prev_obs2", C.c_int32)]


class BamIngestStats(C.Structure):
    """besst_bam_ingest_stats (include/besst_b200.h)"""
    _fields_ = [("compressed_bytes", C.c_int64), ("uncompressed_bytes", C.c_int64), ("blocks", C.c_int64),
                ("records", C.c_int64), ("windows", C.c_int64), ("rescans", C.c_int64), ("seconds_total", C.c_double),
                ("seconds_read", C.c_double), ("ms_inflate", C.c_float), ("ms_scan", C.c_float), ("ms_decode", C.c_float),
                ("crc_checked", C.c_int32)]


BAM_NO_CRC, BAM_BLIND_SEEDS = 1, 2


class GraphSizes(C.Structure):
    _fields_ = [("n_edges", C.c_int64), ("n_links", C.c_int64), ("n_contigs", C.c_int64), ("n_fishy", C.c_int64),
                ("n_ll_links", C.c_int64)]


_GRAPH_FIELDS = [("edge_u", np.uint32, "E"), ("edge_v", np.uint32, "E"), ("nr_links", np.int32, "E"),
                 ("obs_sum", np.int64, "E"), ("obs_sq", np.int64, "E"), ("first_idx", np.int64, "E"),
                 ("row_ptr", np.int64, "E1"), ("gap", np.int32, "E"), ("score", np.float64, "E"),
                 ("ks", np.float64, "E"), ("sd_obs", np.float64, "E"), ("sd_model", np.float64, "E"),
                 ("fishy", np.int32, "E"), ("flags", np.uint8, "E"), ("obs_u", np.int32, "L"),
                 ("obs_v", np.int32, "L"), ("aligned_len", np.int64, "C")]


class GraphOut(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _, _ in _GRAPH_FIELDS] + [("counters", C.c_int64 * N_COUNTERS)]


class LibMetricsOut(C.Structure):
    _fields_ = [("n_samples", C.c_int64), ("n_trimmed", C.c_int64), ("mean_before", C.c_double),
                ("sd_before", C.c_double), ("mean_converged", C.c_double), ("sd_converged", C.c_double),
                ("skewness", C.c_double), ("mu_adj", C.c_double), ("sigma_adj", C.c_double),
                ("skew_adj", C.c_double), ("median_adj", C.c_int64), ("mode_adj", C.c_int64),
                ("n_bins", C.c_int64), ("cont_mapped", C.c_int64), ("cont_n", C.c_int64),
                ("cont_mean", C.c_double), ("cont_sd", C.c_double), ("records_scanned", C.c_int64),
                ("cont_n_before", C.c_int64), ("cont_mean_before", C.c_double), ("cont_sd_before", C.c_double)]


@dataclass
class GraphResult:
    """Host copy of one graph build (the CSR edge list of include/besst_b200.h)."""
    edge_u: np.ndarray
    edge_v: np.ndarray
    nr_links: np.ndarray
    obs_sum: np.ndarray
    obs_sq: np.ndarray
    first_idx: np.ndarray
    row_ptr: np.ndarray
    gap: np.ndarray
    score: np.ndarray
    ks: np.ndarray
    sd_obs: np.ndarray
    sd_model: np.ndarray
    fishy: np.ndarray
    flags: np.ndarray
    obs_u: np.ndarray
    obs_v: np.ndarray
    aligned_len: np.ndarray
    counters: np.ndarray

    @property
    def n_edges(self):
        return int(self.edge_u.shape[0])

    @property
    def n_links(self):
        return int(self.obs_u.shape[0])


def alloc_graph_out(sizes):
    """Allocate numpy result arrays for `sizes` and a GraphOut pointing at them."""
    dims = {"E": int(sizes.n_edges), "E1": int(sizes.n_edges) + 1, "L": int(sizes.n_links), "C": int(sizes.n_contigs)}
    arrays = {name: np.empty(dims[d], dtype=dt) for name, dt, d in _GRAPH_FIELDS}
    out = GraphOut()
    for name, _, _ in _GRAPH_FIELDS:
        setattr(out, name, arrays[name].ctypes.data)
    return out, arrays


def view_graph_out(out, sizes):
    """numpy views over the library-owned pinned buffers a besst_graph_view call left in `out`
    (no copy; valid until the next view / destroy on that ctx)."""
    dims = {"E": int(sizes.n_edges), "E1": int(sizes.n_edges) + 1, "L": int(sizes.n_links), "C": int(sizes.n_contigs)}
    arrays = {}
    for name, dt, d in _GRAPH_FIELDS:
        n = dims[d]
        ptr = getattr(out, name)
        if n == 0 or not ptr:
            arrays[name] = np.empty(0, dtype=dt)
            continue
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        arrays[name] = np.frombuffer(buf, dtype=dt, count=n)
    return arrays


def graph_result(out, arrays):
    return GraphResult(counters=np.array(list(out.counters), dtype=np.int64), **arrays)


def make_records(batch_or_arrays, on_device=False, keepalive=None):
    """Fill a Records struct from a RecordBatch (host numpy) or from a dict of
    raw device pointers (`{name: int_ptr, 'n': N}`) when on_device."""
    if hasattr(batch_or_arrays, "abi_records"):   # a DeviceRecordBatch: columns the engine already holds in HBM
        return batch_or_arrays.abi_records
    r = Records()
    if on_device:
        r.n = int(batch_or_arrays["n"])
        for name in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq"):
            setattr(r, name, int(batch_or_arrays.get(name) or 0) or None)
        r.packed = int(batch_or_arrays.get("packed") or 0) or None
        r.on_device = 1
        return r
    arrs = batch_or_arrays.device_arrays() if hasattr(batch_or_arrays, "device_arrays") else batch_or_arrays
    r.n = int(arrs["tid"].shape[0])
    for name in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq"):
        a = arrs[name]
        assert a.flags["C_CONTIGUOUS"]
        setattr(r, name, a.ctypes.data)
        if keepalive is not None:
            keepalive.append(a)
    packed = getattr(batch_or_arrays, "packed", None) if not isinstance(batch_or_arrays, dict) else batch_or_arrays.get("packed")
    if packed is not None:
        assert packed.flags["C_CONTIGUOUS"] and packed.dtype == np.uint32 and packed.shape[0] == r.n
        r.packed = packed.ctypes.data
        if keepalive is not None:
            keepalive.append(packed)
    r.on_device = 0
    return r


def make_params(orientation, min_mapq, read_len, mean_ins_size, std_dev_ins_size, ins_size_threshold,
                detect_duplicate=True, extend_paths=True, no_score=False, erf_variant=ERF_AS7126,
                halo=(-1, -1)):
    p = LibParams()
    p.orientation = ORIENT_FR if orientation in ("fr", ORIENT_FR) else ORIENT_RF
    p.min_mapq = int(min_mapq)
    p.detect_duplicate = int(bool(detect_duplicate))
    p.extend_paths = int(bool(extend_paths))
    p.no_score = int(bool(no_score))
    p.erf_variant = int(erf_variant)
    p.read_len = float(read_len if read_len is not None else 0.0)
    p.mean_ins_size = float(mean_ins_size if mean_ins_size is not None else 0.0)
    p.std_dev_ins_size = float(std_dev_ins_size if std_dev_ins_size is not None else 0.0)
    p.ins_size_threshold = float(ins_size_threshold if ins_size_threshold is not None else 0.0)
    p.halo_prev_obs1, p.halo_prev_obs2 = int(halo[0]), int(halo[1])
    return p
