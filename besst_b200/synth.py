"""Synthetic libraries for parity tests and the benchmark.

Written with torch ops only so the same code builds a few thousand pairs on the
CPU for the test-suite and 2e8 pairs directly in HBM for bench.py (no network,
no datasets: `data: synthetic`).  Not part of the hot path.

Model: contig lengths 500+Exp(4500) clipped to [500, 100000], random strand,
true gaps clip(round(N(500,300)), 0, 1500), genome = concatenation; fragments
start uniformly on the genome with N(mu, sigma) length, 100 bp reads that must
both lie inside contigs; PE = fr innies, MP = rf outies, optional PE
contamination of an MP library (fr innies N(350,100)); 1 % exact duplicate
pairs; mapq 60 / 0 / 1..10 with weights .90/.05/.05; 0.5 % pairs with one mate
unmapped (a fifth of those reported BWA-style on their own contig, which feeds
the "fishy" branch, CreateGraph.py:141-163); two records per pair with
consistent flags / tlen / mate fields, sorted by (tid, pos).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

READ = 100

CONFIGS = {
    # name: (n_contigs, n_pairs, orientation, mu, sigma, contamination)
    "tiny": (60, 20000, "fr", 550.0, 50.0, 0.0),
    "small_pe": (400, 200000, "fr", 550.0, 50.0, 0.0),
    "small_mp": (400, 200000, "rf", 3000.0, 500.0, 0.0),
    "small_mp_cont": (400, 200000, "rf", 3000.0, 500.0, 0.25),
    "config2": (10000, 20000000, "fr", 550.0, 50.0, 0.0),
    "config3": (100000, 200000000, "rf", 3000.0, 500.0, 0.0),
    "config4_pe": (100000, 200000000, "fr", 550.0, 50.0, 0.0),
    "config4_mp": (100000, 200000000, "rf", 3000.0, 500.0, 0.25),
}
SEED0 = 20261017


@dataclass
class SynthLibrary:
    cols: dict            # name -> torch tensor (tid, mtid, pos, mpos, tlen, qlen int32; flag int16 bits; mapq uint8)
    lengths: torch.Tensor  # contig lengths (int64, CPU)
    names: list
    n_pairs: int
    orientation: str
    mu: float
    sigma: float

    @property
    def n_records(self):
        return int(self.cols["tid"].shape[0])

    def to_batch(self):
        from .records import RecordBatch
        c = {k: v.cpu().numpy() for k, v in self.cols.items()}
        c["flag"] = c["flag"].astype(np.uint16)
        return RecordBatch(references=self.names, lengths=[int(x) for x in self.lengths.tolist()], **c)


def make_contigs(n_contigs, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    lengths = (500 + torch.empty(n_contigs, dtype=torch.float64).exponential_(1.0 / 4500.0, generator=g)).clamp_(500, 100000).to(torch.int64)
    strand = torch.randint(0, 2, (n_contigs,), generator=g)
    gaps = torch.normal(500.0, 300.0, (n_contigs,), generator=g).round().clamp_(0, 1500).to(torch.int64)
    starts = torch.cumsum(lengths + gaps, 0) - (lengths + gaps)
    names = None
    return lengths, strand, starts, names


def contig_names(lengths, strand, starts):
    s, l, r = starts.tolist(), lengths.tolist(), strand.tolist()
    return ["c%d,pos:%d-%d,rc:%d" % (i, s[i], s[i] + l[i], r[i]) for i in range(len(l))]


def _draw_fragments(n_try, lo, span, contamination, mu, sigma, g, dev, max_len=None):
    """n_try candidate fragments with start uniform in [lo, lo + span): (start, length, is_contamination)"""
    f = lo + (torch.rand(n_try, generator=g, device=dev, dtype=torch.float64) * span).to(torch.int64)
    is_cont = torch.rand(n_try, generator=g, device=dev) < contamination
    L = torch.normal(0.0, 1.0, (n_try,), generator=g, device=dev, dtype=torch.float64)
    L = torch.where(is_cont, 350.0 + 100.0 * L, mu + sigma * L).round().to(torch.int64).clamp_(2 * READ, max_len)
    return f, L, is_cont


def _place_reads(f, L, is_cont, n_pairs, d_len, d_start, n_contigs, genome):
    """Keep the first n_pairs fragments whose two reads lie inside contigs."""
    a0 = f                      # left read  [a0, a0+READ)
    b0 = f + L - READ           # right read [b0, b0+READ)
    ca = torch.searchsorted(d_start, a0, right=True) - 1
    cb = (torch.searchsorted(d_start, b0, right=True) - 1).clamp_(0, n_contigs - 1)
    ok = (a0 + READ <= d_start[ca] + d_len[ca]) & (b0 >= d_start[cb]) & (b0 + READ <= d_start[cb] + d_len[cb]) & (b0 + READ <= genome)
    keep = torch.nonzero(ok).squeeze(1)[:n_pairs]
    return a0[keep], b0[keep], ca[keep], cb[keep], is_cont[keep]


def _pairs_to_records(a0, b0, ca, cb, is_cont, orientation, d_len, d_strand, d_start, g, dev):
    """Placed pairs -> the two BAM records of every pair (unsorted columns), after adding 1 % exact
    duplicates.  Consumes `g` in a fixed order: the libraries of the golden fixtures depend on it."""
    n = int(a0.shape[0])
    # 1 % exact duplicates: copies of earlier pairs
    n_dup = n // 100
    if n_dup:
        src = torch.randint(0, n, (n_dup,), generator=g, device=dev)
        a0 = torch.cat([a0, a0[src]]); b0 = torch.cat([b0, b0[src]])
        ca = torch.cat([ca, ca[src]]); cb = torch.cat([cb, cb[src]])
        is_cont = torch.cat([is_cont, is_cont[src]])
        n += n_dup

    # genome strand of the two reads: innie (left fwd, right rev) for fr and contamination, outie for rf
    outie = torch.full((n,), orientation == "rf", device=dev) & ~is_cont
    rev_a_g, rev_b_g = outie, ~outie

    def to_contig(g0, c, rev_g):
        off = g0 - d_start[c]
        rc = d_strand[c] == 1
        pos = torch.where(rc, d_len[c] - off - READ, off)
        return pos.to(torch.int32), rev_g ^ rc

    pa, ra = to_contig(a0, ca, rev_a_g)
    pb, rb = to_contig(b0, cb, rev_b_g)
    first_is_a = torch.rand(n, generator=g, device=dev) < 0.5

    # one mate unmapped in 0.5 % of the pairs; a fifth of those BWA-style on its own contig
    u = torch.rand(n, generator=g, device=dev)
    unm = u < 0.005
    fishy_style = u < 0.001
    unm_is_a = torch.rand(n, generator=g, device=dev) < 0.5

    def mapq(k):
        r = torch.rand(k, generator=g, device=dev)
        low = torch.randint(1, 11, (k,), generator=g, device=dev)
        return torch.where(r < 0.90, torch.full_like(low, 60), torch.where(r < 0.95, torch.zeros_like(low), low)).to(torch.uint8)

    def record(c_self, p_self, r_self, c_mate, p_mate, r_mate, is_first, self_unm, mate_unm, left):
        same = c_self == c_mate
        flag = torch.full((n,), 0x1, dtype=torch.int32, device=dev)
        flag |= torch.where(same & ~self_unm & ~mate_unm, 0x2, 0)
        flag |= torch.where(self_unm, 0x4, 0) | torch.where(mate_unm, 0x8, 0)
        flag |= torch.where(r_self, 0x10, 0) | torch.where(r_mate, 0x20, 0)
        flag |= torch.where(is_first, 0x40, 0x80)
        # an unmapped read sits at its mate's coordinates unless it is reported BWA-style
        relocate = self_unm & ~fishy_style
        tid = torch.where(relocate, c_mate, c_self).to(torch.int32)
        pos = torch.where(relocate, p_mate, p_self)
        mtid = c_mate.to(torch.int32)
        mpos = p_mate
        mate_reloc = mate_unm & ~fishy_style
        mtid = torch.where(mate_reloc, tid, mtid)
        mpos = torch.where(mate_reloc, pos, mpos)
        span = (torch.maximum(p_self, p_mate) + READ - torch.minimum(p_self, p_mate)).to(torch.int32)
        leftmost = (p_self < p_mate) | ((p_self == p_mate) & left)
        tlen = torch.where(same & ~self_unm & ~mate_unm, torch.where(leftmost, span, -span), torch.zeros_like(span))
        return tid, mtid, pos, mpos, tlen, flag

    a_unm, b_unm = unm & unm_is_a, unm & ~unm_is_a
    recs_a = record(ca, pa, ra, cb, pb, rb, first_is_a, a_unm, b_unm, torch.ones(n, dtype=torch.bool, device=dev))
    recs_b = record(cb, pb, rb, ca, pa, ra, ~first_is_a, b_unm, a_unm, torch.zeros(n, dtype=torch.bool, device=dev))
    tid = torch.cat([recs_a[0], recs_b[0]]); mtid = torch.cat([recs_a[1], recs_b[1]])
    pos = torch.cat([recs_a[2], recs_b[2]]); mpos = torch.cat([recs_a[3], recs_b[3]])
    tlen = torch.cat([recs_a[4], recs_b[4]]); flag = torch.cat([recs_a[5], recs_b[5]])
    mq = mapq(2 * n)
    return {"tid": tid, "mtid": mtid, "pos": pos, "mpos": mpos, "tlen": tlen, "flag": flag, "mapq": mq}, n


def _sorted_columns(cols, dev):
    """BAM order: stable sort by (tid, pos)"""
    key = cols["tid"].to(torch.int64) * (1 << 32) + cols["pos"].to(torch.int64)
    order = torch.sort(key, stable=True).indices
    del key
    m = int(order.shape[0])
    return {"tid": cols["tid"][order], "mtid": cols["mtid"][order], "pos": cols["pos"][order], "mpos": cols["mpos"][order],
            "tlen": cols["tlen"][order], "qlen": torch.full((m,), READ, dtype=torch.int32, device=dev),
            "flag": cols["flag"][order].to(torch.int16), "mapq": cols["mapq"][order]}


def make_library(n_contigs, n_pairs, orientation="fr", mu=550.0, sigma=50.0, contamination=0.0, seed=SEED0,
                 device="cpu", with_names=True, oversample=None):
    lengths, strand, starts, _ = make_contigs(n_contigs, seed)
    names = contig_names(lengths, strand, starts) if with_names else ["c%d" % i for i in range(n_contigs)]
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    d_len, d_strand, d_start = lengths.to(dev), strand.to(dev), starts.to(dev)
    genome = int((starts[-1] + lengths[-1]).item())
    inside_frac = float(lengths.sum().item()) / genome
    if oversample is None:
        oversample = 1.25 / max(inside_frac * inside_frac * 0.9, 0.05)
    n_try = int(n_pairs * oversample) + 1024

    f, L, is_cont = _draw_fragments(n_try, 0, genome - READ, contamination, mu, sigma, g, dev)
    a0, b0, ca, cb, is_cont = _place_reads(f, L, is_cont, n_pairs, d_len, d_start, n_contigs, genome)
    del f, L
    cols, n = _pairs_to_records(a0, b0, ca, cb, is_cont, orientation, d_len, d_strand, d_start, g, dev)
    cols = _sorted_columns(cols, dev)
    return SynthLibrary(cols=cols, lengths=lengths, names=names, n_pairs=n, orientation=orientation, mu=mu, sigma=sigma)


# ---- ONE global library, generated slice by slice (multi-GPU workloads) ------------------------------
ZONE_W = 16384   # width of a boundary zone in genome coordinates; fragments are clamped below it


def slice_bounds(n_contigs, world):
    """Contig ranges of the BAM-order range partition: rank r owns tids [bounds[r], bounds[r+1])"""
    return [(n_contigs * r) // world for r in range(world)] + [n_contigs]


def make_library_slice(n_contigs, n_pairs, orientation, mu, sigma, contamination, seed, rank, world, device="cpu", read_seed=0):
    """Rank `rank`'s slice of ONE global BAM-ordered library of n_contigs contigs and ~n_pairs pairs, range-
    partitioned by contig: the concatenation of the slices of ranks 0..world-1 is a single sorted BAM whose
    pairs span the cuts (a pair starting left of a cut with its right read on the other side leaves one
    record in either slice; duplicates of such pairs are split too).  The genome is divided into interior
    zones (one per rank, own seed) and boundary zones of ZONE_W bases left of every cut (own seed, generated
    IDENTICALLY by both neighbours, each keeping the records on its own contigs), so no rank ever needs
    another rank's records.  The library is defined by (seed, read_seed, world): `seed` fixes the contigs,
    `read_seed` tells the libraries of one assembly apart, a different world size is a different (equally
    distributed) library.
    -> SynthLibrary (lengths = ALL contigs; n_pairs = records of the slice / 2)"""
    lengths, strand, starts, _ = make_contigs(n_contigs, seed)
    dev = torch.device(device)
    d_len, d_strand, d_start = lengths.to(dev), strand.to(dev), starts.to(dev)
    genome = int((starts[-1] + lengths[-1]).item())
    inside_frac = float(lengths.sum().item()) / genome
    oversample = 1.25 / max(inside_frac * inside_frac * 0.9, 0.05)
    max_len = ZONE_W - 1
    if mu + 6 * sigma >= max_len:
        raise ValueError("insert size distribution too wide for the boundary zones")
    bounds = slice_bounds(n_contigs, world)
    cut = [int(starts[b].item()) if b < n_contigs else genome for b in bounds]   # genome coordinate of every cut
    cut[0] = 0

    def zone(lo, hi, zseed):
        if hi - lo <= 0:
            return None
        want = int(round(n_pairs * (hi - lo) / float(genome)))
        if want <= 0:
            return None
        g = torch.Generator(device=dev).manual_seed(zseed)
        n_try = int(want * oversample) + 1024
        f, L, is_cont = _draw_fragments(n_try, lo, hi - lo, contamination, mu, sigma, g, dev, max_len=max_len)
        a0, b0, ca, cb, is_cont = _place_reads(f, L, is_cont, want, d_len, d_start, n_contigs, genome)
        cols, _ = _pairs_to_records(a0, b0, ca, cb, is_cont, orientation, d_len, d_strand, d_start, g, dev)
        return cols

    parts = []
    lo = cut[rank]
    hi = cut[rank + 1] - (ZONE_W if rank + 1 < world else 0)   # the last rank's interior runs to the genome end
    if rank + 1 == world:
        hi -= READ
    rs = seed + 15485863 * int(read_seed)
    parts.append(zone(lo, max(lo, hi), rs + 7919 * (rank + 1)))
    if rank > 0:
        parts.append(zone(cut[rank] - ZONE_W, cut[rank], rs + 104729 * rank))
    if rank + 1 < world:
        parts.append(zone(max(cut[rank], cut[rank + 1] - ZONE_W), cut[rank + 1], rs + 104729 * (rank + 1)))
    parts = [p for p in parts if p is not None]
    keys = ("tid", "mtid", "pos", "mpos", "tlen", "flag", "mapq")
    cols = {k: torch.cat([p[k] for p in parts]) if parts else torch.zeros(0, dtype=torch.int32 if k != "mapq" else torch.uint8, device=dev) for k in keys}
    mine = (cols["tid"] >= bounds[rank]) & (cols["tid"] < bounds[rank + 1])
    cols = {k: v[mine] for k, v in cols.items()}
    cols = _sorted_columns(cols, dev)
    n_rec = int(cols["tid"].shape[0])
    return SynthLibrary(cols=cols, lengths=lengths, names=["c%d" % i for i in range(n_contigs)], n_pairs=n_rec // 2,
                        orientation=orientation, mu=mu, sigma=sigma)


def later_library_rows(lengths, contig_threshold, seed, join=3, drop_every=17):
    """The contig table a LATER library sees (runBESST:143-231: library k+1 starts from the scaffolds
    MakeScaffolds.Algorithm built from library k): runs of 1..join consecutive contigs joined into
    multi-contig scaffolds with random directions and gaps of 1..399, every drop_every-th contig removed
    (repeats), scaffolds shorter than contig_threshold demoted to small (CleanObjects, CreateGraph.py:788-810).
    Vectorised counterpart of tests/helpers.later_library_objects + ContigTable for tables of 1e5..1e6
    contigs.  -> (rows, n_scaffolds, n_large_scaffolds)"""
    from .abi import CONTIG_ROW_DTYPE, CTG_LARGE, CTG_SMALL
    from .contig_table import largest_reference_mask
    lengths = np.asarray(lengths, dtype=np.int64)
    n = lengths.shape[0]
    rng = np.random.default_rng(seed)
    # group boundaries: a random break before every contig, and a forced one after `join` members
    brk = rng.random(n) < 0.5
    brk[0] = True
    gid0 = np.cumsum(brk) - 1
    first = np.nonzero(brk)[0]
    within = np.arange(n) - first[gid0]
    brk |= (within % join) == 0
    gid = np.cumsum(brk) - 1
    first = np.nonzero(brk)[0]
    present = np.ones(n, dtype=bool)
    if drop_every:
        present[np.arange(n) % drop_every == drop_every - 1] = False
    direction = rng.integers(0, 2, n)
    gap = rng.integers(1, 400, n)
    # position inside the scaffold: cumulative (length + gap) over the PRESENT members before it
    step = np.where(present, lengths + gap, 0)
    cum = np.cumsum(step) - step
    position = cum - cum[first[gid]]
    n_groups = int(gid[-1]) + 1
    end = np.where(present, position + lengths, 0)
    s_len = np.zeros(n_groups, dtype=np.int64)
    np.maximum.at(s_len, gid, end)
    has_member = np.zeros(n_groups, dtype=bool)
    has_member[gid[present]] = True
    large_g = has_member & (s_len >= contig_threshold)
    small_g = has_member & ~large_g
    index = np.full(n_groups, -1, dtype=np.int64)
    n_large = int(large_g.sum())
    index[large_g] = np.arange(n_large)
    index[small_g] = n_large + np.arange(int(small_g.sum()))
    rows = np.zeros(n, dtype=CONTIG_ROW_DTYPE)
    rows["state"] = np.where(present, np.where(large_g[gid], CTG_LARGE, CTG_SMALL), 0)
    rows["scaffold"] = np.where(present, index[gid], 0)
    rows["direction"] = np.where(present, direction, 0)
    rows["position"] = np.where(present, position, 0)
    rows["length"] = np.where(present, lengths, 0)
    rows["scaf_length"] = np.where(present, s_len[gid], 0)
    rows["in_largest"] = largest_reference_mask(lengths)
    return rows, n_large + int(small_g.sum()), n_large


def make_config(name, device="cpu", seed_offset=0, scale=1.0, with_names=True):
    n_contigs, n_pairs, orientation, mu, sigma, cont = CONFIGS[name]
    idx = list(CONFIGS).index(name)
    return make_library(max(2, int(n_contigs * scale)), max(1000, int(n_pairs * scale)), orientation, mu, sigma, cont,
                        seed=SEED0 + idx + seed_offset, device=device, with_names=with_names)
