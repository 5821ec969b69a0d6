"""Per-edge digests of a graph-build result, mergeable across BAM-order slices.

Used to check a multi-GPU build at full size against per-slice runs of a sequential
implementation without gathering hundreds of millions of observations on one host: every link
of an edge enters an order-sensitive polynomial hash  H = sum_j w_j * P**j (mod 2**64)  of its
observation pair, so that the hash of a concatenation is  H_a + P**n_a * H_b.  Partial tables
(one per slice, in slice order) merge exactly like the reference's `CreateEdge` accumulates
(CreateGraph.py:842-862): link counts, observation sums and squares add up, the observation
lists concatenate in BAM order, the first slice that has the edge defines its first appearance.

Pure numpy bookkeeping; knows nothing about who produced the results it compares.
"""
from __future__ import annotations

import numpy as np

P = np.uint64(0x9FB21C651E98DF25)      # odd
A = np.uint64(0x9E3779B97F4A7C15)
B = np.uint64(0xC2B2AE3D27D4EB4F)
C0 = np.uint64(0x165667B19E3779F9)


def _powers(n):
    """P**0 .. P**n (mod 2**64)"""
    t = np.full(int(n) + 1, P, dtype=np.uint64)
    t[0] = 1
    with np.errstate(over="ignore"):
        return np.cumprod(t, dtype=np.uint64)


def edge_table(res, first_base=0):
    """GraphResult -> dict of per-edge arrays: key (u << 32 | v), n, obs_sum, obs_sq, first (+ first_base),
    fishy, h (list hash), plus the pass-through per-edge results gap / score / flags."""
    E = int(res.n_edges)
    row_ptr = np.asarray(res.row_ptr, dtype=np.int64)
    nr = np.asarray(res.nr_links, dtype=np.int64)
    key = (np.asarray(res.edge_u).astype(np.uint64) << np.uint64(32)) | np.asarray(res.edge_v).astype(np.uint64)
    h = np.zeros(E, dtype=np.uint64)
    if E and int(row_ptr[-1]) > 0:
        with np.errstate(over="ignore"):
            w = np.asarray(res.obs_u).astype(np.int64).view(np.uint64) * A + np.asarray(res.obs_v).astype(np.int64).view(np.uint64) * B + C0
            j = np.arange(int(row_ptr[-1]), dtype=np.int64) - np.repeat(row_ptr[:-1], nr)
            w *= _powers(int(nr.max()))[j]
            h = np.add.reduceat(w, row_ptr[:-1]).astype(np.uint64)
    return dict(key=key, n=nr.copy(), obs_sum=np.asarray(res.obs_sum, dtype=np.int64).copy(),
                obs_sq=np.asarray(res.obs_sq, dtype=np.int64).copy(),
                first=np.asarray(res.first_idx, dtype=np.int64) + int(first_base), fishy=np.asarray(res.fishy, dtype=np.int64).copy(),
                h=h, gap=np.asarray(res.gap).copy(), score=np.asarray(res.score).copy(), flags=np.asarray(res.flags).copy(),
                parts=np.ones(E, dtype=np.int64))


def merge_slices(tables):
    """Partial tables of consecutive BAM-order slices (slice order!) -> the table of the whole library.
    `parts` counts the slices that contributed to an edge; gap / score / flags are those of the first
    contributor and are only meaningful where parts == 1."""
    tables = [t for t in tables if t is not None]
    key = np.concatenate([t["key"] for t in tables])
    src = np.concatenate([np.full(t["key"].shape[0], i, dtype=np.int64) for i, t in enumerate(tables)])
    order = np.lexsort((src, key))          # by key, slices in order inside a key
    key_s = key[order]
    head = np.ones(key_s.shape[0], dtype=bool)
    head[1:] = key_s[1:] != key_s[:-1]
    starts = np.nonzero(head)[0]
    cat = {f: np.concatenate([t[f] for t in tables])[order] for f in ("n", "obs_sum", "obs_sq", "first", "fishy", "h", "gap", "score", "flags")}
    out = {"key": key_s[head]}
    if key_s.shape[0] == 0:
        for f in cat:
            out[f] = cat[f]
        out["parts"] = np.zeros(0, dtype=np.int64)
        return out
    n_before = np.cumsum(cat["n"]) - cat["n"]
    n_before -= np.repeat(n_before[starts], np.diff(np.append(starts, key_s.shape[0])))   # links of the edge in earlier slices
    with np.errstate(over="ignore"):
        shifted = cat["h"] * _powers(int(n_before.max()))[n_before]
    out["h"] = np.add.reduceat(shifted, starts).astype(np.uint64)
    for f in ("n", "obs_sum", "obs_sq"):
        out[f] = np.add.reduceat(cat[f], starts)
    # fishy pairs are counted per build over ALL fishy records, whichever slice they sit in: sum
    out["fishy"] = np.add.reduceat(cat["fishy"], starts)
    for f in ("first", "gap", "score", "flags"):
        out[f] = cat[f][starts]
    out["parts"] = np.diff(np.append(starts, key_s.shape[0]))
    return out


def apply_fishy(table, fishy_dicts):
    """Per-slice {node-pair key: count} maps (unmapped-read1 records, CreateGraph.py:141-163) -> the per-edge
    totals of the whole library, whichever slice the fishy records sit in."""
    keys = np.concatenate([np.fromiter(d.keys(), dtype=np.uint64, count=len(d)) for d in fishy_dicts] + [np.zeros(0, np.uint64)])
    cnts = np.concatenate([np.fromiter(d.values(), dtype=np.int64, count=len(d)) for d in fishy_dicts] + [np.zeros(0, np.int64)])
    total = np.zeros(table["key"].shape[0], dtype=np.int64)
    if keys.shape[0] and total.shape[0]:
        pos = np.searchsorted(table["key"], keys)
        pos = np.minimum(pos, total.shape[0] - 1)
        hit = table["key"][pos] == keys
        np.add.at(total, pos[hit], cnts[hit])
    table["fishy"] = total
    return table


def concat_owned(tables):
    """Tables of DISJOINT edge sets (one per owner rank) -> one table sorted by key."""
    tables = [t for t in tables if t is not None]
    key = np.concatenate([t["key"] for t in tables])
    order = np.argsort(key, kind="stable")
    out = {f: np.concatenate([t[f] for t in tables])[order] for f in tables[0]}
    return out


INT_FIELDS = ("key", "n", "obs_sum", "obs_sq", "first", "h", "fishy")


def compare(got, want, rtol=1e-6):
    """-> dict: integers_bit_exact (edge set, link counts, sums, first appearance, observation lists in
    order), gap_equal / score_max_rel_diff over the edges of `want` built from a single slice."""
    same_shape = got["key"].shape == want["key"].shape
    report = {"edges": int(want["key"].shape[0]), "links": int(want["n"].sum()), "edges_got": int(got["key"].shape[0])}
    exact = same_shape and all(np.array_equal(got[f], want[f]) for f in INT_FIELDS)
    report["integers_bit_exact"] = bool(exact)
    if not exact:
        report["first_mismatch"] = next((f for f in INT_FIELDS if not same_shape or not np.array_equal(got[f], want[f])), None)
        return report
    single = want["parts"] == 1
    scored = single & ((want["flags"] & 2) != 0)
    report["multi_slice_edges"] = int((~single).sum())
    report["gap_equal"] = bool(np.array_equal(got["gap"][scored], want["gap"][scored]))
    a, b = got["score"][scored], want["score"][scored]
    ok = ~np.isnan(b)
    report["score_nan_pattern_equal"] = bool(np.array_equal(np.isnan(a), np.isnan(b)))
    report["score_max_rel_diff"] = float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300))) if ok.any() else 0.0
    report["scores_within_tolerance"] = bool(report["score_max_rel_diff"] <= rtol)
    return report
