"""ORACLE / TEST INFRASTRUCTURE -- the sequential C oracle run slice by slice, one BAM-order slice per
rank, merged into the result of the whole library (besst_b200/digest.py tables).

The reference is ONE sequential scan (CreateGraph.py:111-211); the only state that crosses a record is
"(obs1, obs2) of the previous CreateEdge call" (:835-838, 869-870).  So the scan over slice s is exact if
it starts from the last call made in slices 0..s-1.  Every rank scans its slice once with a start state
that matches nothing, the first/last calls are gathered, and only a slice whose first call equals the call
before it (a duplicate pair split by the cut) is scanned again with its true start state.

Lets bench.py / the tests check a multi-GPU build at FULL size in parallel on the ranks' host cores
(the single-pass oracle needs the whole library on one host and ~5 s per 100 M pairs).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import oracle_lib
from besst_b200 import abi, digest

NO_MATCH = -(2 ** 31) + 1


def _with_halo(params, halo):
    p = abi.LibParams()
    C.memmove(C.byref(p), C.byref(params), C.sizeof(abi.LibParams))
    p.halo_prev_obs1, p.halo_prev_obs2 = int(halo[0]), int(halo[1])
    return p


def sliced_oracle(dist, rank, world, rows, n_scaffolds, params, batch_slice, group=None):
    """Collective over `group`.  -> on rank 0: dict(table=merged digest table of the whole library,
    counters=int64[16] (sums; last/first call slots global), aligned_len=int64[C], seconds=max oracle time);
    None on the other ranks."""
    import time
    t0 = time.perf_counter()
    res, _, fishy, consistent = oracle_lib.graph_build(rows, n_scaffolds, _with_halo(params, (NO_MATCH, NO_MATCH)), batch_slice)
    c = res.counters
    mine = [int(c[abi.CNT_CALLS] > 0), int(c[abi.CNT_LAST_OBS1]), int(c[abi.CNT_LAST_OBS2]), int(c[abi.CNT_FIRST_OBS1]),
            int(c[abi.CNT_FIRST_OBS2])]
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    halos, cur = [], (int(params.halo_prev_obs1), int(params.halo_prev_obs2))
    for r in range(world):
        halos.append(cur)
        if everyone[r][0]:
            cur = (everyone[r][1], everyone[r][2])
    if everyone[rank][0] and (everyone[rank][3], everyone[rank][4]) == halos[rank]:
        res, _, fishy, consistent = oracle_lib.graph_build(rows, n_scaffolds, _with_halo(params, halos[rank]), batch_slice)
    seconds = time.perf_counter() - t0
    n_links = [None] * world
    dist.all_gather_object(n_links, int(res.n_links), group=group)
    table = digest.edge_table(res, first_base=sum(n_links[:rank]))
    payload = dict(table=table, fishy=fishy, counters=np.asarray(res.counters, dtype=np.int64), aligned=np.asarray(res.aligned_len, dtype=np.int64),
                   consistent=bool(consistent), seconds=seconds)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged = digest.apply_fishy(digest.merge_slices([g["table"] for g in gathered]), [g["fishy"] for g in gathered])
    counters = np.sum([g["counters"] for g in gathered], axis=0)
    with_calls = [r for r in range(world) if everyone[r][0]]
    counters[abi.CNT_LAST_OBS1], counters[abi.CNT_LAST_OBS2] = cur
    counters[abi.CNT_FIRST_OBS1], counters[abi.CNT_FIRST_OBS2] = (everyone[with_calls[0]][3], everyone[with_calls[0]][4]) if with_calls else (0, 0)
    return dict(table=merged, counters=counters, aligned_len=np.sum([g["aligned"] for g in gathered], axis=0),
                consistent=all(g["consistent"] for g in gathered), seconds=max(g["seconds"] for g in gathered))


def gather_owned(dist, rank, world, local_result, group=None):
    """The ranks' DISJOINT shares of a distributed build (GraphResult with global first_idx) -> on rank 0 the
    digest table of all edges, plus the (already reduced) counters / aligned_len of rank 0's result."""
    table = digest.edge_table(local_result)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(table, gathered, dst=0, group=group)
    if rank != 0:
        return None
    return dict(table=digest.concat_owned(gathered), counters=np.asarray(local_result.counters, dtype=np.int64),
                aligned_len=np.asarray(local_result.aligned_len, dtype=np.int64))


def compare(got, want):
    report = digest.compare(got["table"], want["table"])
    report["counters_equal"] = bool(np.array_equal(got["counters"][:12], want["counters"][:12]))
    report["aligned_len_equal"] = bool(np.array_equal(got["aligned_len"], want["aligned_len"]))
    report["integers_bit_exact"] = bool(report["integers_bit_exact"] and report["counters_equal"] and report["aligned_len_equal"])
    return report
