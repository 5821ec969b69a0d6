"""ORACLE / TEST INFRASTRUCTURE -- ctypes binding of oracle/besst_oracle.c.

Importable only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  Builds the .so with `make -C oracle` on first use.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from besst_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libbesst_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "besst_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "besst_b200.h")
    stale = (not os.path.exists(SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.besst_oracle_graph_build.restype = C.c_void_p
        L.besst_oracle_graph_build.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(abi.LibParams),
                                               C.POINTER(abi.Records)]
        L.besst_oracle_graph_sizes.argtypes = [C.c_void_p, C.POINTER(abi.GraphSizes), C.POINTER(C.c_int64),
                                               C.POINTER(C.c_int)]
        L.besst_oracle_graph_fetch.argtypes = [C.c_void_p, C.POINTER(abi.GraphOut), C.c_void_p, C.c_void_p, C.c_void_p]
        L.besst_oracle_graph_free.argtypes = [C.c_void_p]
        L.besst_oracle_libmetrics.argtypes = [C.c_void_p, C.c_int64, C.POINTER(abi.LibParams), C.POINTER(abi.Records),
                                              C.c_void_p, C.c_int64, C.c_int32, C.POINTER(abi.LibMetricsOut),
                                              C.c_void_p, C.c_int64]
        L.besst_oracle_gapest_batch.argtypes = [C.POINTER(abi.LibParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_int64, C.c_void_p, C.c_void_p]
        L.besst_oracle_gap_estimator.restype = C.c_int32
        L.besst_oracle_gap_estimator.argtypes = [C.c_double] * 6 + [C.c_int]
        L.besst_oracle_tr_sk_std_dev.restype = C.c_double
        L.besst_oracle_tr_sk_std_dev.argtypes = [C.c_double] * 6 + [C.c_int]
        L.besst_oracle_max_obs_distr.restype = C.c_double
        L.besst_oracle_max_obs_distr.argtypes = [C.c_double, C.c_double]
        L.besst_oracle_ks_2samp.restype = C.c_double
        L.besst_oracle_ks_2samp.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        _lib = L
    return _lib


def graph_build(rows, n_scaffolds, params, batch):
    """-> (abi.GraphResult, tuples[LINK_TUPLE_DTYPE], fishy {key: count}, consistent)"""
    L = lib()
    rows = np.ascontiguousarray(rows)
    keep = []
    rec = abi.make_records(batch, keepalive=keep)
    h = L.besst_oracle_graph_build(rows.ctypes.data, rows.shape[0], int(n_scaffolds), C.byref(params), C.byref(rec))
    try:
        sizes, nt, cons = abi.GraphSizes(), C.c_int64(), C.c_int()
        L.besst_oracle_graph_sizes(h, C.byref(sizes), C.byref(nt), C.byref(cons))
        out, arrays = abi.alloc_graph_out(sizes)
        tuples = np.zeros(nt.value, dtype=abi.LINK_TUPLE_DTYPE)
        fk = np.zeros(sizes.n_fishy, dtype=np.uint64)
        fc = np.zeros(sizes.n_fishy, dtype=np.int32)
        L.besst_oracle_graph_fetch(h, C.byref(out), tuples.ctypes.data, fk.ctypes.data, fc.ctypes.data)
    finally:
        L.besst_oracle_graph_free(h)
    return abi.graph_result(out, arrays), tuples, dict(zip(fk.tolist(), fc.tolist())), bool(cons.value)


def gapest_lognormal_batch(mu, sigma, read_len, samples, row_ptr, len1, len2):
    L = lib()
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    len1 = np.ascontiguousarray(len1, dtype=np.float64)
    len2 = np.ascontiguousarray(len2, dtype=np.float64)
    n = row_ptr.shape[0] - 1
    gap = np.zeros(n, dtype=np.int32)
    L.besst_oracle_gapest_lognormal_batch.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_int64, C.c_void_p]
    L.besst_oracle_gapest_lognormal_batch(float(mu), float(sigma), float(read_len), samples.ctypes.data, row_ptr.ctypes.data,
                                          len1.ctypes.data, len2.ctypes.data, n, gap.ctypes.data)
    return gap


def libmetrics(rows, params, batch, ref_lengths, want_isize, cap=1 << 20):
    L = lib()
    rows = np.ascontiguousarray(rows)
    keep = []
    rec = abi.make_records(batch, keepalive=keep)
    lens = np.ascontiguousarray(ref_lengths, dtype=np.int64)
    out = abi.LibMetricsOut()
    adj = np.zeros(cap, dtype=np.float64)
    rc = L.besst_oracle_libmetrics(rows.ctypes.data, rows.shape[0], C.byref(params), C.byref(rec), lens.ctypes.data,
                                   lens.shape[0], int(want_isize), C.byref(out), adj.ctypes.data, cap)
    return rc, out, adj[:min(cap, out.n_bins)]


def gapest_batch(params, mean_obs, len1, len2):
    L = lib()
    mean_obs = np.ascontiguousarray(mean_obs, dtype=np.float64)
    len1 = np.ascontiguousarray(len1, dtype=np.float64)
    len2 = np.ascontiguousarray(len2, dtype=np.float64)
    gap = np.zeros(mean_obs.shape[0], dtype=np.int32)
    sd = np.zeros(mean_obs.shape[0], dtype=np.float64)
    L.besst_oracle_gapest_batch(C.byref(params), mean_obs.ctypes.data, len1.ctypes.data, len2.ctypes.data,
                                mean_obs.shape[0], gap.ctypes.data, sd.ctypes.data)
    return gap, sd
