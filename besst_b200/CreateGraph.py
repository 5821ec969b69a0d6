"""Drop-in for `BESST.CreateGraph.PE` (reference CreateGraph.py:45-321).

Same signature, same side effects on the four object dicts and on `param`,
same lines written to `Information`, same returned `(G, G_prime)` networkx
graphs -- but the per-record loop (:111-211), CreateEdge (:812-871), the
PosDir calculators (:1024-1076) and the per-edge statistics of
GiveScoreOnEdges (:498-614) run as sm_100a CUDA kernels behind the C ABI
(include/besst_b200.h).  This module is the host side above that ABI: it
flattens the objects into the contig table, calls the engine once, replays the
reference's post-filters (:237-321) on the CSR edge list it gets back
(besst_b200/csr_post.py: masks over the edge list, the order-dependent pruning in C)
and only then builds the two networkx graphs, for the surviving edges.

There is no CPU fallback: without the CUDA library `PE` raises.
"""
from __future__ import annotations

import os
import sys
from time import time

import numpy as np

from . import abi
from . import e_nr_links
from .contig_table import ContigTable
from .normal import MaxObsDistr
from .objects import classes
from .records import as_batch, as_file


def _new_graph():
    # looked up at call time: the graphs are handed to BESST's consumers (MakeScaffolds.Algorithm, runBESST:199),
    # so they must be of whatever `networkx` the host process runs BESST with (1.x per docs/INSTALL.md:42)
    import networkx
    return networkx.Graph()


def _sample_sd(values, mean):
    # the reference's expression, evaluated left to right (CreateGraph.py:918)
    n = float(len(values))
    return (sum([x ** 2 - 2 * x * mean + mean ** 2 for x in values]) / (n - 1)) ** 0.5


def CalculateStats(sorted_contig_lengths, sorted_contig_lengths_small, param, Information):
    """N50/L50 (CreateGraph.py:874-896)."""
    half = param.tot_assembly_length / 2.0
    cur, nr, N50, L50 = 0, 0, 0, 0
    for group in (sorted_contig_lengths, sorted_contig_lengths_small):
        if N50 != 0:
            break
        for length in group:
            cur += length
            nr += 1
            if cur >= half:
                N50, L50 = length, nr
                break
    print('L50: ', L50, 'N50: ', N50, 'Initial contig assembly length: ', param.tot_assembly_length, file=Information)
    return N50, L50


def InitializeObjects(bam_file, Contigs, Scaffolds, param, Information, G_prime, small_contigs, small_scaffolds, C_dict):
    """One contig + one single-contig scaffold per BAM reference present in the
    FASTA, large if length >= contig_threshold (CreateGraph.py:729-786)."""
    contig_cls, scaffold_cls = classes()
    lengths = [int(x) for x in bam_file.lengths]
    names = bam_file.references
    seq_lengths = [len(s) for s in C_dict.values()]
    param.tot_assembly_length = sum(seq_lengths)
    N50, L50 = CalculateStats(sorted(seq_lengths, reverse=True), [], param, Information)
    param.current_L50, param.current_N50 = L50, N50
    threshold = param.contig_threshold
    start = time()
    for i, name in enumerate(names):
        if (i + 1) % 100000 == 0:
            print('Time adding 100k keys', time() - start, file=Information)
            start = time()
        if name not in C_dict:
            continue
        length = lengths[i]
        large = length >= threshold
        if not large and not length > 0:
            continue
        c = contig_cls(name)
        c.length = length
        c.sequence = C_dict.pop(name)
        c.direction = True
        c.position = 0
        s = scaffold_cls(param.scaffold_indexer, [c], length)
        c.scaffold = s.name
        if large:
            Contigs[name] = c
            Scaffolds[s.name] = s
        else:
            small_contigs[name] = c
            small_scaffolds[s.name] = s
        param.scaffold_indexer += 1


def CleanObjects(Contigs, Scaffolds, param, Information, small_contigs, small_scaffolds):
    """Demote scaffolds shorter than this library's contig_threshold
    (CreateGraph.py:788-810)."""
    large = sorted((s.s_length for s in Scaffolds.values()), reverse=True)
    small = sorted((s.s_length for s in small_scaffolds.values()), reverse=True)
    N50, L50 = CalculateStats(large, small, param, Information)
    param.current_L50, param.current_N50 = L50, N50
    moved = 0
    for name in list(Scaffolds.keys()):
        s = Scaffolds[name]
        if s.s_length < param.contig_threshold:
            for c in s.contigs:
                del Contigs[c.name]
                small_contigs[c.name] = c
            small_scaffolds[name] = s
            del Scaffolds[name]
            moved += 1
    print('Nr of contigs/scaffolds that was singeled out due to length constraints ' + str(moved), file=Information)


def InitializeGraph(dict_with_scaffolds, graph, Information):
    """Two nodes per scaffold joined by an nr_links=None edge
    (CreateGraph.py:710-722)."""
    start = time()
    for cnt, (name, s) in enumerate(dict_with_scaffolds.items(), 1):
        # add_node(..., length=) works under networkx 1.x (where graph.nodes is a method) and 2.x/3.x alike;
        # same node and adjacency insertion order as add_edge followed by the two attribute writes (:713-716)
        graph.add_node((name, 'L'), length=s.s_length)
        graph.add_node((name, 'R'), length=s.s_length)
        graph.add_edge((name, 'L'), (name, 'R'), nr_links=None)
        if cnt % 100000 == 0:
            print('Total nr of keys added: ', cnt, 'Time for adding last 100 000 keys: ', time() - start, file=Information)
            start = time()


def _write_fasta(path, contigs):
    with open(path, 'w') as fh:
        for c in contigs:
            print('>' + c.name, file=fh)
            seq = c.sequence
            for i in range(0, len(seq), 60):
                print(seq[i:i + 60], file=fh)


def _forget(contigs, Contigs, small_contigs):
    for c in contigs:
        if c.name in Contigs:
            del Contigs[c.name]
        else:
            del small_contigs[c.name]


def filter_low_coverage_contigs(Contigs, Scaffolds, graphs, param, small_contigs, small_scaffolds, Information):
    """-z_min filter (CreateGraph.py:407-433; GenerateOutput.py:68-79)."""
    print('Removing low coverage contigs if -z_min specified..', file=Information)
    low = []
    for c in Contigs.values():
        if c.coverage < param.lower_cov_cutoff:
            low.append(c)
            del Scaffolds[c.scaffold]
            graphs.remove_scaffold(c.scaffold, large=True)
    for c in small_contigs.values():
        if c.coverage < param.lower_cov_cutoff:
            low.append(c)
            del small_scaffolds[c.scaffold]
            graphs.remove_scaffold(c.scaffold, large=False)
    _write_fasta(param.output_directory + '/low_coverage_contigs.fa', low)
    _forget(low, Contigs, small_contigs)
    print('Removed a total of: ', len(low), ' low coverage contigs. With coverage lower than ', param.lower_cov_cutoff, file=Information)


def RemoveOutliers(mean_cov, std_dev, cov_list):
    k = MaxObsDistr(len(cov_list), 0.95)
    kept = [x for x in cov_list if x < mean_cov + k * std_dev and x < 2 * mean_cov]
    return len(cov_list) > len(kept), kept


def CalculateMeanCoverage(Contigs, Information, param):
    """Mean/sd of coverage over the <=50000 longest large contigs with iterative
    outlier removal (CreateGraph.py:898-949)."""
    by_length = sorted(((c.length, name) for name, c in Contigs.items()), key=lambda t: t[0], reverse=True)[:50000]
    cov = [Contigs[name].coverage for _, name in by_length if Contigs[name].coverage > 0]
    if len(cov) <= 1:
        sys.exit("Too few contigs to calculate coverage on. Got: {0} contigs. If you have specified  -z_min or --min_mapq, consider lower them. If not, check the BAM file for proper alignments. Exiting here before scaffolding...".format(len(cov)))
    n = float(len(cov))
    mean_cov = sum(cov) / n
    std_dev = _sample_sd(cov, mean_cov)
    print('Mean coverage before filtering out extreme observations = ', mean_cov, file=Information)
    print('Std dev of coverage before filtering out extreme observations= ', std_dev, file=Information)
    print('Number of contigs used in calc of coverage before filtering: ', n, file=Information)
    again = True
    while again:
        again, kept = RemoveOutliers(mean_cov, std_dev, cov)
        n = float(len(kept))
        if n == 0 or sum(kept) == 0:
            break
        mean_cov = sum(kept) / n
        std_dev = _sample_sd(kept, mean_cov)
        cov = kept
    print('Mean coverage after filtering = ', mean_cov, file=Information)
    print('Std coverage after filtering = ', std_dev, file=Information)
    print('Number of contigs used in calc of coverage after filtering: ', n, file=Information)
    print('Length of longest contig in calc of coverage: ', by_length[0][0], file=Information)
    print('Length of shortest contig in calc of coverage: ', by_length[-1][0], file=Information)
    return mean_cov, std_dev


def RepeatDetector(Contigs, Scaffolds, graphs, param, small_contigs, small_scaffolds, Information):
    """Coverage-based repeat removal (CreateGraph.py:959-1018;
    GenerateOutput.py:47-66)."""
    mean_cov, std_dev = param.mean_coverage, param.std_dev_coverage
    k = MaxObsDistr(len(Contigs), 0.95)
    thresh = param.cov_cutoff if param.cov_cutoff else max(mean_cov + k * std_dev, 2 * mean_cov - 3 * std_dev)
    print('Detecting repeats..', file=Information)
    repeats, count_hapl = [], 0
    hapl_limit = None
    if param.detect_haplotype:
        hapl_limit = mean_cov / 2.0 + param.hapl_threshold * std_dev
    for c in Contigs.values():
        if c.coverage > thresh:
            repeats.append(c)
            del Scaffolds[c.scaffold]
            graphs.remove_scaffold(c.scaffold, large=True)
        if hapl_limit is not None and c.coverage < hapl_limit:
            count_hapl += 1
            c.is_haplotype = True
    for c in small_contigs.values():
        if c.coverage > thresh:
            repeats.append(c)
            del small_scaffolds[c.scaffold]
            graphs.remove_scaffold(c.scaffold, large=False)
        if hapl_limit is not None and c.coverage < hapl_limit:
            count_hapl += 1
            c.is_haplotype = True
    with open(param.output_directory + '/repeats_log.tsv', 'w') as fh:
        print("contig_accession\tlength\tcoverage\tcov/mean_cov(exp number of placements)\tlib_mean\tplacable", file=fh)
        for c in sorted(repeats, key=lambda x: x.coverage, reverse=True):
            placable = "Yes" if param.mean_ins_size > c.length else 'No'
            print("{0}\t{1}\t{2}\t{3}\t{4}\t{5}".format(c.name, c.length, round(c.coverage, 1), round(c.coverage / param.mean_coverage, 0), round(param.mean_ins_size, 0), placable), file=fh)
    _write_fasta(param.output_directory + '/repeats.fa', repeats)
    _forget(repeats, Contigs, small_contigs)
    print('Removed a total of: ', len(repeats), ' repeats. With coverage larger than ', thresh, file=Information)
    if param.detect_haplotype:
        print('Marked a total of: ', count_hapl, ' potential haplotypes.', file=Information)


def infer_spurious_link_count_threshold(graphs, param):
    """Expected link count over a gap of mean+sd-2r between two 100 kb contigs
    -> param.expected_links_over_mean_plus_stddev (CreateGraph.py:323-353)."""
    nr_nodes = graphs.number_of_nodes("G_prime") / 2
    ratio = param.contamination_ratio if param.contamination_ratio else 0
    cov = param.mean_coverage * (1 - ratio)
    link_params = e_nr_links.Param(param.mean_ins_size, param.std_dev_ins_size, cov, param.read_len, 0)
    gap = param.mean_ins_size + param.std_dev_ins_size - 2 * param.read_len
    expected = e_nr_links.ExpectedLinks(100000, 100000, gap, link_params)
    for link_number, total in graphs.link_count_profile():
        print('Nodes: {0}.\t Total edges with over {1} links:{2}. \tAverage density: {3}'.format(nr_nodes, link_number, total, total / float(nr_nodes)), file=param.information_file)
    param.expected_links_over_mean_plus_stddev = 5 if expected < 5 else int(expected)
    print('Letting filtering threshold in high complexity regions be {0} for this library.'.format(param.expected_links_over_mean_plus_stddev), file=param.information_file)


def remove_edges_below_threshold(graphs, param):
    """Order-dependent pruning of G_prime (CreateGraph.py:355-404): in
    `graph.edges()` order, drop a weak edge only while both endpoints still have
    more than 4 neighbours; then drop everything under -e."""
    print('Remove edges in high complexity areas.', file=param.information_file)
    removed = graphs.prune_dense_regions(param.expected_links_over_mean_plus_stddev)
    print('Removed total of {0} edges in high density areas.'.format(removed), file=param.information_file)
    low = graphs.drop_low_support("G_prime", param.edgesupport)
    print('Removed an additional of {0} edges with low support from full graph G_prime of all contigs.'.format(low), file=param.information_file)


def _score_file_names(Scaffolds, node, first):
    """Contig names / directions column of score_file_pass_N.tsv (CreateGraph.py:622-650); the reference's
    direction strings are constant by construction ('+' if True else '-')."""
    forward = (node[1] == 'R') if first else (node[1] == 'L')
    contigs = Scaffolds[node[0]].contigs if forward else Scaffolds[node[0]].contigs[::-1]
    return ";".join(c.name for c in contigs), ";".join(('+' if forward else '-') for _ in contigs)


def GiveScoreOnEdges(G, res, table, param, Information, Scaffolds=None):
    """The per-edge gap and score (the arithmetic of CreateGraph.py:498-614 ran on the GPU) are attached when
    the graphs are materialised (csr_post.CsrGraphs.materialise); this writes score_file_pass_N.tsv when
    param.print_scores (:481-483,622-654) and the closing counter line."""
    if getattr(param, 'plots', False):
        print('plots requested: the score histograms of CreateGraph.py:656-662 are not produced by besst_b200', file=Information)
    if getattr(param, 'print_scores', False) and Scaffolds is not None:
        index = {}
        for e in np.nonzero(res.flags & abi.EDGE_SCORED)[0].tolist():
            index[(int(res.edge_u[e]), int(res.edge_v[e]))] = e
        with open(os.path.join(param.output_directory, "score_file_pass_{0}.tsv".format(param.pass_number)), "w") as score_file:
            print("{0}\t{1}\t{2}\t{3}\t{4}\t{5}\t{6}\t{7}".format("scf1/ctg1", "o1", "scf2/ctg2", "o2", "gap", "link_variation_score", "link_dispersity_score", "number_of_links"), file=score_file)
            for n0, n1, d in G.edges(data=True):
                if d['nr_links'] is None:
                    continue
                a, b = table.node_id(n0), table.node_id(n1)
                e = index[(a, b) if a < b else (b, a)]
                if res.flags[e] & abi.EDGE_NEGGAP:
                    continue
                n = int(res.nr_links[e])
                sd, sd0 = float(res.sd_obs[e]), float(res.sd_model[e])
                std_dev_score = 0 if (sd == 0.0 or sd0 == 0.0 or sd != sd) else min(sd / sd0, sd0 / sd)
                span_score = 0 if n < 5 else 1 - float(res.ks[e])
                # `gap` of the reference is GapEstimator's int for two long scaffolds, else the float naive estimate (:536-539)
                gap = int(res.gap[e]) if res.flags[e] & abi.EDGE_BIG else (n * param.mean_ins_size - int(res.obs_sum[e])) / float(n)
                scf1, dir1 = _score_file_names(Scaffolds, n0, True)
                scf2, dir2 = _score_file_names(Scaffolds, n1, False)
                print("{0}\t{1}\t{2}\t{3}\t{4}\t{5}\t{6}\t{7}".format(scf1, dir1, scf2, dir2, gap, std_dev_score, span_score, n), file=score_file)
    print('Number of significantly spurious edges:', 0, file=Information)


def get_conditional_stddevs(steps, empirical_isize_distr, max_isize):
    """sd of the insert-size distribution conditioned on spanning a gap, per gap size (CreateGraph.py:436-468):
    density f(x) * max(0, x - gap + 1) for every gap in `steps`, each value repeated up to the next step."""
    xs = np.fromiter(empirical_isize_distr.keys(), dtype=np.int64, count=len(empirical_isize_distr))
    fx = np.fromiter(empirical_isize_distr.values(), dtype=np.float64, count=len(empirical_isize_distr))
    out, previous_gap = [], 0
    for gap in steps:
        dens = np.zeros(max_isize + 1, dtype=np.float64)
        w = np.maximum(0, xs - gap + 1)
        dens[xs[w > 0]] = (fx * w)[w > 0]
        tot = float(dens.sum())
        idx = np.arange(max_isize + 1, dtype=np.float64)
        mu_c = float((idx * dens).sum()) / tot
        sd_c = float(np.sqrt((((idx - mu_c) ** 2) * dens).sum() / tot))
        out.extend([sd_c] if gap == 0 else [sd_c] * (gap - previous_gap))
        previous_gap = gap
    return out


def lognormal_rescore(res, table, param, engine):
    """The lognormal branch of GiveScoreOnEdges (CreateGraph.py:485-493, 523-531, 549-553) on the CSR result:
    the gap of every scored edge between two long scaffolds comes from the lognormal ML estimator over the
    edge's raw observations (one warp per edge, besst_gapest_lognormal_batch), capped at the end of the
    conditional-sd table, and the model sd is that table's entry; sample sd and KS statistic are the ones the
    build computed.  The reference's own branch only runs under Python 2 (`range` with a float step, :490):
    `max_isize / 50` is read as the integer division it was there."""
    emp = param.empirical_distribution
    max_isize = max(emp.keys())
    steps = list(range(0, int(max_isize * 0.8), max(1, max_isize // 50)))
    cond = np.asarray(get_conditional_stddevs(steps, emp, max_isize), dtype=np.float64)
    log_norm_max_gap = cond.shape[0] - 1
    scored = np.nonzero(res.flags & abi.EDGE_SCORED)[0]
    if scored.shape[0] == 0:
        return
    big = (res.flags[scored] & abi.EDGE_BIG) != 0
    n = res.nr_links[scored].astype(np.float64)
    len1 = table.scaffold_lengths[res.edge_u[scored] >> 1].astype(np.float64)
    len2 = table.scaffold_lengths[res.edge_v[scored] >> 1].astype(np.float64)
    gap = (n * param.mean_ins_size - res.obs_sum[scored]) / n                     # data_observation (:511)
    if big.any():
        eb = scored[big]
        nr = res.nr_links[eb].astype(np.int64)
        rp = np.concatenate([[0], np.cumsum(nr)])
        idx = np.repeat(res.row_ptr[eb] - rp[:-1], nr) + np.arange(int(rp[-1]), dtype=np.int64)
        samples = (res.obs_u[idx].astype(np.int64) + res.obs_v[idx]).astype(np.int32)
        g = engine.gapest_lognormal_batch(param.lognormal_mean, param.lognormal_sigma, param.read_len, samples, rp, len1[big], len2[big])
        gap[big] = np.minimum(g, log_norm_max_gap)                               # :526-528
    neg = (-gap > len1) | (-gap > len2)                                           # :542
    sd0 = np.full(scored.shape[0], 4294967296.0)
    sd0[big] = cond[np.where(gap[big] > 0, gap[big].astype(np.int64), 0)]         # :549-553
    sd, ks = res.sd_obs[scored], res.ks[scored]
    with np.errstate(divide="ignore", invalid="ignore"):
        sds = np.where((sd0 == 0) | (sd == 0) | np.isnan(sd), 0.0, np.minimum(sd / sd0, sd0 / sd))
    span = np.where(n < 5, 0.0, 1 - ks)
    score = np.where((sds > 0.5) & (span > 0.5), sds + span, 0.0)
    score[neg] = 0.0
    res.gap[scored] = np.trunc(gap).astype(np.int32)                              # int(gap) :541
    res.score[scored] = score
    res.sd_model[scored] = np.where(neg, np.nan, sd0)
    res.flags[scored] = (res.flags[scored] & ~np.uint8(abi.EDGE_NEGGAP)) | np.where(neg, abi.EDGE_NEGGAP, 0).astype(np.uint8)


def engine_params(param, halo=(-1, -1)):
    return abi.make_params(param.orientation, param.min_mapq, param.read_len, param.mean_ins_size,
                           param.std_dev_ins_size, param.ins_size_threshold,
                           detect_duplicate=param.detect_duplicate, extend_paths=param.extend_paths,
                           no_score=param.no_score, halo=halo)


def _world():
    """(rank, world size) of the torch.distributed job this process belongs to, (0, 1) outside one."""
    from .records import world
    return world()


def graph_build_distributed(engine, table, params, batch, rank, world):
    """The multi-GPU build behind PE: every process of a torchrun job calls PE with the same
    arguments; rank r extracts the r-th BAM-order slice of the records on its GPU, runs of links are exchanged
    once by edge hash, every rank builds its share of the edges, and the shares are merged on every rank
    (fetch_global), so that all processes continue with the identical CSR -- the order-dependent post-filters
    below and BESST's consumers run replicated, as the single-process program would."""
    from .dist import DistributedGraphBuild
    backend = engine.make_dist_backend(table)
    runner = DistributedGraphBuild(backend, rank, world)
    if getattr(batch, "dist_info", None):   # the rank's own part of the file, as the distributed ingest left it (device or host columns)
        mine = batch
    else:
        n = len(batch)
        bounds = [(n * r // world) - ((n * r // world) % 128) for r in range(world)] + [n]
        mine = batch.slice(bounds[rank], bounds[rank + 1])
    runner.step(params, backend.records(mine) if hasattr(backend, "records") else mine)
    return runner.fetch_global()


def _number_of_edges(graph):
    """graph.number_of_edges() over the raw adjacency dicts (networkx sums a degree VIEW: a Python call per node)"""
    adj = getattr(graph, "_adj", None)
    if adj is None:
        adj = graph.adj
    loops = sum(1 for u, nbrs in adj.items() if u in nbrs)
    return (sum(map(len, adj.values())) + loops) // 2


class _gc_paused(object):
    """PE allocates a few container objects per contig, scaffold, node and edge -- millions for a real assembly,
    none of them garbage: pause the cyclic collector's generation scans (~40 % of the host time otherwise)."""

    def __enter__(self):
        import gc
        self.was = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        import gc
        if self.was:
            gc.enable()
        return False


def PE(Contigs, Scaffolds, Information, C_dict, param, small_contigs, small_scaffolds, bam_file, engine=None):
    with _gc_paused():
        return _PE(Contigs, Scaffolds, Information, C_dict, param, small_contigs, small_scaffolds, bam_file, engine)


def _PE(Contigs, Scaffolds, Information, C_dict, param, small_contigs, small_scaffolds, bam_file, engine=None):
    from .csr_post import CsrGraphs
    bam_file = as_file(bam_file, engine)   # a path: decoded once by the native ingest library (shared with get_metrics)
    print('Parsing BAM file...', file=Information)
    if param.first_lib:
        start = time()
        InitializeObjects(bam_file, Contigs, Scaffolds, param, Information, None, small_contigs, small_scaffolds, C_dict)
        print('Time initializing BESST objects: ', time() - start, file=Information)
    else:
        start = time()
        CleanObjects(Contigs, Scaffolds, param, Information, small_contigs, small_scaffolds)
        print('Time cleaning BESST objects for next library: ', time() - start, file=Information)

    if len(Scaffolds) == 0:
        if not os.path.isfile(param.output_directory + '/repeats.fa'):
            open(param.output_directory + '/repeats.fa', 'w').close()
        return (_new_graph(), _new_graph())

    # InitializeGraph (:85-96) happens at materialisation; its progress lines go out here
    start = time()
    for d in ((small_scaffolds, Scaffolds) if param.no_score else
              (Scaffolds, small_scaffolds, Scaffolds) if param.extend_paths else (Scaffolds,)):
        for cnt in range(100000, len(d) + 1, 100000):
            print('Total nr of keys added: ', cnt, 'Time for adding last 100 000 keys: ', time() - start, file=Information)
    print('Total time elapsed for initializing Graph: ', time() - start, file=Information)

    print('Reading bam file and creating scaffold graph...', file=Information)
    start = time()
    if engine is None:
        from .engine import default_engine
        engine = default_engine()
    batch = as_batch(bam_file)
    table = ContigTable(bam_file.references, bam_file.lengths, Contigs, small_contigs, Scaffolds, small_scaffolds)
    rank, world = _world()
    if world > 1:
        res = graph_build_distributed(engine, table, engine_params(param), batch, rank, world)
    else:
        res = engine.graph_build(table, engine_params(param), batch, view=True)   # consumed below, before the next build
    graphs = CsrGraphs(res, table, param)   # (G, G_prime) as masks over the CSR edge list
    cnt = res.counters
    print('ELAPSED reading file:', time() - start, file=Information)
    print('NR OF FISHY READ LINKS: ', int(cnt[abi.CNT_FISHY]), file=Information)
    print('Number of USEFUL READS (reads mapping to different contigs uniquly): ', int(cnt[abi.CNT_COUNT]), file=Information)
    print('Number of non unique reads (at least one read non-unique in read pair) that maps to different contigs (filtered out from scaffolding): ', int(cnt[abi.CNT_NON_UNIQUE]), file=Information)
    print('Reads with too large insert size from "USEFUL READS" (filtered out): ', int(cnt[abi.CNT_TOO_LONG]), file=Information)
    print('Initial number of edges in G (the graph with large contigs): ', graphs.number_of_edges("G"), file=Information)
    print('Initial number of edges in G_prime (the full graph of all contigs before removal of repats): ', graphs.number_of_edges("G_prime"), file=Information)
    if param.detect_duplicate:
        print('Number of duplicated reads indicated and removed: ', int(cnt[abi.CNT_DUPLICATES]), file=Information)

    # coverage of every contig under this library (CreateGraph.py:237-244)
    tid_of = {name: i for i, name in enumerate(bam_file.references)}
    for d in (Contigs, small_contigs):
        for name, c in d.items():
            tid = tid_of.get(name)
            aligned = int(res.aligned_len[tid]) if tid is not None else 0
            c.coverage = aligned / float(c.length)

    if param.first_lib and param.lower_cov_cutoff:
        filter_low_coverage_contigs(Contigs, Scaffolds, graphs, param, small_contigs, small_scaffolds, Information)

    param.mean_coverage, param.std_dev_coverage = CalculateMeanCoverage(Contigs, Information, param)
    if param.first_lib:
        RepeatDetector(Contigs, Scaffolds, graphs, param, small_contigs, small_scaffolds, Information)
    print('Number of edges in G (after repeat removal): ', graphs.number_of_edges("G"), file=Information)
    print('Number of edges in G_prime (after repeat removal): ', graphs.number_of_edges("G_prime"), file=Information)

    print('Number of BWA buggy edges removed: ', graphs.remove_bug_edges(), file=Information)
    print('Number of edges in G (after filtering for buggy flag stats reporting): ', graphs.number_of_edges("G"), file=Information)
    print('Number of edges in G_prime  (after filtering for buggy flag stats reporting): ', graphs.number_of_edges("G_prime"), file=Information)

    infer_spurious_link_count_threshold(graphs, param)
    if not param.edgesupport:
        param.edgesupport = 5
        print('Letting -e be {0} for this library.'.format(param.edgesupport), file=Information)
    else:
        print('User has set -e to be {0} for this library.'.format(param.edgesupport), file=Information)
    removed = graphs.drop_low_support("G", param.edgesupport)
    print('Removed {0} edges from graph G of border contigs.'.format(removed), file=Information)
    remove_edges_below_threshold(graphs, param)

    if not param.no_score and param.lognormal:
        for f in ("gap", "score", "sd_model", "flags"):   # overwritten in place below
            if not getattr(res, f).flags.writeable:
                setattr(res, f, getattr(res, f).copy())
        lognormal_rescore(res, table, param, engine)
    # networkx objects for the surviving edges only, with CreateEdge's attributes and GiveScoreOnEdges' gap / score
    G, G_prime = graphs.materialise(_new_graph, Scaffolds, small_scaffolds,
                                    lazy_observations=bool(getattr(param, 'lazy_observations', False)))
    if not param.no_score:
        GiveScoreOnEdges(G, res, table, param, Information, Scaffolds)
    print('Number of edges in G_prime  (after removing edges under -e threshold (if not specified, default is -e 3): ', _number_of_edges(G_prime), file=Information)
    print("\n -------------------------------------------------------------\n", file=Information)
    print('Nr of contigs/scaffolds included in this pass: ' + str(len(Scaffolds) + len(small_scaffolds)), file=Information)
    print('Out of which {0} acts as border contigs.'.format(len(Scaffolds)), file=Information)
    return (G, G_prime)
