#!/usr/bin/env python
"""Benchmark of the hot path: read-pairs/s through graph build + GapEst.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path (records -> CSR edges + link statistics + KS/GapEst score) over
every library of the workload, in sequence (SURVEY.md 8d generator).  Workloads follow BASELINE.json's
`configs`:

  N=1  config3  100 k contigs / 200 M MP pairs (rf), 1 library  -- the configuration the north-star
                target (>= 100 M read-pairs/s) is quoted on
  N=2  config3  scaled x2 (weak): ONE global library of 200 k contigs / 400 M pairs
  N=4  config4  100 k contigs / 400 M pairs: a PE library, then a PE-contaminated MP library that sees
                multi-contig scaffolds (the state a previous pass leaves behind)
  N=8  config5  1 M contigs / 2 B pairs: three libraries (PE 550/50, MP 3000/500, MP 8000/1200)

At N>1 every library is ONE global BAM-ordered file range-partitioned over the ranks (pairs and
duplicates span the cuts); whole runs of links are exchanged once by edge hash (stored straight into
the destination GPU's memory over NVLink), every rank builds its share of the edges.

Prints ONE JSON line (rank 0).  `value` = whole-job read-pairs/s with the records resident in HBM;
`e2e` = the same through the C-ABI call with pinned HOST buffers (H2D of the record columns and D2H of
the result inside the timed region); `roofline` = the dominant kernel against the measured HBM peak;
`cpu_baseline` = the C oracle (a port of the reference's Python path) on host cores; `parity` = the CUDA
result of the WHOLE workload against that oracle.  `--impl reference` times that CPU port alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL's banner (NCCL_DEBUG=VERSION/WARN/INFO) would land on stdout next to the one JSON line: send it to stderr
if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"

import numpy as np  # noqa: E402

# name: contigs, libraries (orientation, mu, sigma, contamination, pairs), weak = scale contigs/pairs with the world size
WORKLOADS = {
    "config3": dict(desc="synthetic 100k contigs / 200 M MP pairs (rf), 1 library", contigs=100_000, weak=True,
                    libs=[("rf", 3000.0, 500.0, 0.0, 200_000_000)]),
    "config2": dict(desc="synthetic 10k contigs / 20 M PE pairs, 1 library", contigs=10_000, weak=True,
                    libs=[("fr", 550.0, 50.0, 0.0, 20_000_000)]),
    "config4": dict(desc="two libraries PE + MP with PE-contamination, 100k contigs / 400 M pairs", contigs=100_000, weak=False,
                    libs=[("fr", 550.0, 50.0, 0.0, 200_000_000), ("rf", 3000.0, 500.0, 0.25, 200_000_000)]),
    "config5": dict(desc="1 M contigs / 2 B read pairs, 3 libraries", contigs=1_000_000, weak=False,
                    libs=[("fr", 550.0, 50.0, 0.0, 667_000_000), ("rf", 3000.0, 500.0, 0.0, 667_000_000),
                          ("rf", 8000.0, 1200.0, 0.0, 666_000_000)]),
    "small": dict(desc="synthetic 400 contigs / 200 k MP pairs (debug)", contigs=400, weak=True,
                  libs=[("rf", 3000.0, 500.0, 0.0, 200_000)]),
    "small2": dict(desc="synthetic 3000 contigs, PE then contaminated MP, 2 x 1 M pairs (debug)", contigs=3000, weak=False,
                   libs=[("fr", 550.0, 50.0, 0.0, 1_000_000), ("rf", 3000.0, 500.0, 0.25, 1_000_000)]),
}
DEFAULT_WORKLOAD = {1: "config3", 2: "config3", 4: "config4", 8: "config5"}
RECORD_BYTES = 4 + 4 + 4 + 4 + 4 + 2 + 1   # tid mtid pos mpos qlen flag mapq (tlen is only read by libmetrics)
TUPLE_BYTES = 16
REFERENCE_SAMPLE_PAIRS = 10_000_000
RESULT_FIELDS = ("edge_u", "edge_v", "nr_links", "obs_sum", "obs_sq", "first_idx", "row_ptr", "gap", "score", "ks", "sd_obs",
                 "sd_model", "fishy", "flags", "obs_u", "obs_v", "aligned_len")


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append(line.strip())
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def library_params(abi, orientation, mu, sigma):
    return abi.make_params(orientation=orientation, min_mapq=11, read_len=100.0, mean_ins_size=mu,
                           std_dev_ins_size=sigma, ins_size_threshold=mu + 6 * sigma)


def checksum(res):
    """Order-sensitive digest of a GraphResult's integer content."""
    import hashlib
    h = hashlib.sha256()
    for f in ("edge_u", "edge_v", "nr_links", "obs_sum", "obs_sq", "first_idx", "row_ptr", "fishy", "obs_u", "obs_v",
              "aligned_len"):
        h.update(np.ascontiguousarray(getattr(res, f)).tobytes())
    h.update(res.counters[:10].tobytes())
    return h.hexdigest()[:16]


def result_bytes(res, abi):
    return sum(int(getattr(res, f).nbytes) for f in RESULT_FIELDS) + 8 * abi.N_COUNTERS


class Library(object):
    """One library of the workload on this rank: device-resident record columns + its contig table slot."""

    def __init__(self, index, spec, n_contigs, world, rank, dev, synth, abi, first_library_rows, legacy_single, packed):
        import torch
        orientation, mu, sigma, cont, n_pairs = spec
        self.index, self.orientation, self.mu, self.sigma, self.cont = index, orientation, mu, sigma, cont
        seed = synth.SEED0 + 2 + 100 * index
        if legacy_single:   # N=1 single-library workloads: the very library of the round-1 bench lines
            lib = synth.make_library(n_contigs, n_pairs, orientation, mu, sigma, cont, seed=seed, device=dev, with_names=False)
        else:
            lib = synth.make_library_slice(n_contigs, n_pairs, orientation, mu, sigma, cont, synth.SEED0 + 2, rank, world, device=dev,
                                           read_seed=index)
        self.tlen = lib.cols["tlen"] if index == 0 and world == 1 else None   # only the libmetrics leg reads it
        self.plain = {k: lib.cols[k] for k in ("flag", "mapq", "qlen")} if (self.tlen is not None or not packed) else None
        if packed:   # flag | mapq << 12 | qlen << 20: the one column the ingest library writes for the graph build (20 B/record)
            pk = (lib.cols["flag"].to(torch.int32) & 0xfff) | (lib.cols["mapq"].to(torch.int32) << 12) | (lib.cols["qlen"] << 20)
            self.cols = {"tid": lib.cols["tid"], "mtid": lib.cols["mtid"], "pos": lib.cols["pos"], "mpos": lib.cols["mpos"], "packed": pk}
        else:
            self.cols = {k: v for k, v in lib.cols.items() if k != "tlen"}
        self.packed = packed
        self.n_rec = lib.n_records
        self.lengths = lib.lengths.numpy()
        threshold = mu + 4 * sigma
        if index == 0:
            self.rows, self.n_scaf, self.n_large = first_library_rows(self.lengths, threshold)
            self.state = "first library (one contig per scaffold)"
        else:
            self.rows, self.n_scaf, self.n_large = synth.later_library_rows(self.lengths, threshold, seed=synth.SEED0 + 31 * index)
            self.state = "later library (runs of 1-3 contigs joined into scaffolds, 1/17 removed)"
        self.params = library_params(abi, orientation, mu, sigma)
        ptrs = {k: v.data_ptr() for k, v in self.cols.items()}
        if self.plain is not None:
            ptrs.update({k: v.data_ptr() for k, v in self.plain.items()})
        ptrs["tlen"] = self.tlen.data_ptr() if self.tlen is not None else 0
        ptrs["n"] = self.n_rec
        self.rec_dev = abi.make_records(ptrs, on_device=True)
        del lib
        torch.cuda.synchronize()

    def describe(self):
        return {"orientation": self.orientation, "mean_ins_size": self.mu, "std_dev_ins_size": self.sigma,
                "contamination": self.cont, "records_this_rank": self.n_rec, "contig_table": self.state}


class _Seq(object):
    """contig sequence stand-in for the PE-level leg: only len() is used on this path"""
    __slots__ = ("n",)

    def __init__(self, n):
        self.n = int(n)

    def __len__(self):
        return self.n

    def __getitem__(self, s):
        lo, hi, _ = s.indices(self.n)
        return "N" * max(0, hi - lo)


class _RunParam(object):
    """what runBESST:88-158 sets on its parameter object for one library run without -m/-s"""

    def __init__(self, orientation, outdir, info, lazy):
        self.scaffold_indexer, self.no_score, self.min_mapq, self.max_contig_overlap = 1, False, 11, 200
        self.cov_cutoff, self.lower_cov_cutoff, self.plots, self.development, self.print_scores = None, 0.001, False, False, False
        self.first_lib, self.pass_number, self.bamfile, self.orientation = True, 1, "in_memory.bam", orientation
        self.mean_ins_size = self.std_dev_ins_size = self.ins_size_threshold = self.contig_threshold = None
        self.edgesupport = self.read_len = None
        self.output_directory, self.information_file = outdir, info
        self.detect_haplotype, self.hapl_ratio, self.hapl_threshold = False, 1.3, 3
        self.detect_duplicate, self.extend_paths, self.lognormal = True, True, False
        self.contamination_ratio = self.tot_assembly_length = None
        self.lazy_observations = lazy


def pe_level_leg(eng, L, host_cols, lazy):
    """The reference-facing entry points end to end for one library (runBESST:168,182): get_metrics (library
    parameters estimated, no -m/-s) + CreateGraph.PE from host records to the two networkx graphs."""
    import io
    import tempfile
    from besst_b200 import CreateGraph as CG, libmetrics
    from besst_b200.records import BatchFile, RecordBatch
    arrs = {k: v[:L.n_rec].numpy() for k, v in host_cols.items()}
    arrs["flag"] = arrs["flag"].view(np.uint16)
    if "packed" in arrs:
        arrs["packed"] = arrs["packed"].view(np.uint32)
    names = ["c%d" % i for i in range(L.lengths.shape[0])]
    batch = RecordBatch(references=names, lengths=[int(x) for x in L.lengths.tolist()],
                        rlen=np.full(min(L.n_rec, 1000), 100, np.int32), alen=np.full(min(L.n_rec, 1000), 100, np.int32), **arrs)
    bam = BatchFile(batch)
    info = io.StringIO()
    param = _RunParam(L.orientation, tempfile.mkdtemp(prefix="besst_bench_"), info, lazy)
    C_dict = {n: _Seq(x) for n, x in zip(names, batch.lengths)}
    Contigs, Scaffolds, small_contigs, small_scaffolds = {}, {}, {}, {}
    import contextlib
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        libmetrics.get_metrics(bam, param, info, engine=eng)
        t1 = time.perf_counter()
        G, G_prime = CG.PE(Contigs, Scaffolds, info, C_dict, param, small_contigs, small_scaffolds, bam, engine=eng)
    t2 = time.perf_counter()
    return {"get_metrics_s": round(t1 - t0, 3), "PE_s": round(t2 - t1, 3), "read_pairs_per_s": L.n_rec / 2.0 / (t2 - t0),
            "G_edges": G.number_of_edges(), "G_prime_edges": G_prime.number_of_edges(),
            "estimated": {"mean_ins_size": param.mean_ins_size, "std_dev_ins_size": param.std_dev_ins_size},
            "observations": "lazy views (param.lazy_observations)" if lazy else "python lists (as the reference stores them)"}


def ingest_leg(eng, pairs):
    """FILE -> graph (N=1): a synthetic MP library of `pairs` read pairs written as a sorted BAM (write_bam_columns:
    206-byte records, htslib block layout, compresses ~2.3x), then
      device         besst_bam_ingest: compressed file -> PCIe -> one warp per BGZF block -> record columns in HBM
      host_threads   libbesst_bamio.so (zlib on every host core) -> host columns         [the round-1 ingest]
      file_to_graph  device ingest + besst_libmetrics + besst_graph_build + result views: no record in host memory
    Wall clock of the calls (the file sits in the page cache on every leg: disk speed is not measured)."""
    import tempfile
    from besst_b200 import abi, bamio, synth
    from besst_b200.contig_table import first_library_rows
    from besst_b200.libmetrics import metric_rows
    lib = synth.make_library(max(50, pairs // 2000), pairs, "rf", 3000.0, 500.0, 0.0, seed=20261099)
    batch = lib.to_batch()
    n = len(batch)
    path = os.path.join(tempfile.mkdtemp(prefix="besst_bench_"), "library.bam")
    t0 = time.perf_counter()
    bamio.write_bam_columns(path, batch)
    t_write = time.perf_counter() - t0
    eng.ingest_bam(path)   # warm-up: page cache, window buffers, column allocation
    runs = [dict(eng.ingest_bam(path).stats) for _ in range(3)]
    best = min(runs, key=lambda r: r["seconds_total"])
    t0 = time.perf_counter()
    nat = bamio.read_bam_native(path)
    t_host = time.perf_counter() - t0
    lengths = np.asarray(batch.lengths, dtype=np.int64)
    mrows = metric_rows(lengths)
    eng.select_table(7)

    def file_to_graph():
        t0 = time.perf_counter()
        dev = eng.ingest_bam(path)
        t1 = time.perf_counter()
        rc, m, _ = eng.libmetrics(mrows, abi.make_params("rf", 11, 100.0, 0.0, 0.0, 0.0), dev, lengths, True)
        t2 = time.perf_counter()
        mu, sd = m.mu_adj, m.sigma_adj
        rows, n_scaf, n_large = first_library_rows(lengths, mu + 4 * sd)
        eng.set_contigs(rows, n_scaf, n_large)
        params = abi.make_params("rf", 11, 100.0, mu, sd, mu + 6 * sd)
        res = eng.fetch_view(eng.build(params, dev.abi_records))
        t3 = time.perf_counter()
        return {"ingest_ms": round(1e3 * (t1 - t0), 2), "libmetrics_ms": round(1e3 * (t2 - t1), 2), "graph_build_fetch_ms": round(1e3 * (t3 - t2), 2),
                "total_ms": round(1e3 * (t3 - t0), 2), "read_pairs_per_s": n / 2.0 / (t3 - t0), "edges": res.n_edges, "links": res.n_links,
                "estimated": {"mean_ins_size": mu, "std_dev_ins_size": sd}}
    file_to_graph()
    ftg = min((file_to_graph() for _ in range(3)), key=lambda r: r["total_ms"])
    same = all(np.array_equal(getattr(nat, f), getattr(batch, f)) for f in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq"))
    dev_host = eng.ingest_bam(path).to_host()
    same_dev = all(np.array_equal(getattr(dev_host, f), getattr(batch, f)) for f in ("tid", "mtid", "pos", "mpos", "tlen", "qlen", "flag", "mapq"))
    os.remove(path)
    return {"file": {"records": n, "compressed_MB": round(best["compressed_bytes"] / 1e6, 1), "inflated_MB": round(best["uncompressed_bytes"] / 1e6, 1),
                     "bgzf_blocks": best["blocks"], "write_s": round(t_write, 2)},
            "device": {"wall_ms": round(1e3 * best["seconds_total"], 2), "records_per_s": n / best["seconds_total"],
                       "host_read_ms": round(1e3 * best["seconds_read"], 2), "inflate_ms": round(best["ms_inflate"], 3),
                       "inflate_GBps_out": round(best["uncompressed_bytes"] / 1e6 / max(best["ms_inflate"], 1e-9), 2),
                       "scan_ms": round(best["ms_scan"], 3), "decode_ms": round(best["ms_decode"], 3), "windows": best["windows"],
                       "crc_checked": bool(best["crc_checked"]), "columns_equal_source": bool(same_dev)},
            "host_threads": {"wall_ms": round(1e3 * t_host, 2), "records_per_s": n / t_host, "threads": int(nat.stats["threads"]),
                             "columns_equal_source": bool(same)},
            "file_to_graph": ftg}


_T0 = time.time()


def crumb(msg):
    """progress line on stderr (rank 0): says where a run was when something outside this program stopped it"""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write("[bench %6.1fs] %s\n" % (time.time() - _T0, msg))
        sys.stderr.flush()


def arm_watchdog(seconds):
    """A collective that never completes (a peer died, the fabric hiccuped) blocks inside C where no Python handler runs:
    let the kernel end the process instead (SIGALRM, default action), so the launcher tears the job down and the GPUs are
    released within a bounded time."""
    import signal
    if seconds > 0:
        signal.signal(signal.SIGALRM, signal.SIG_DFL)
        signal.alarm(int(seconds))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from besst_b200 import abi, synth
    from besst_b200.contig_table import first_library_rows
    from besst_b200.engine import CudaEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    crumb("start: world %d" % world)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # host-side object gathers of the parity leg
        crumb("process groups up")
    name = args.workload or DEFAULT_WORKLOAD.get(world, "config3")
    W = WORKLOADS[name]
    mult = world if W["weak"] else 1
    n_contigs = max(2, int(W["contigs"] * mult * args.scale))
    specs = [(o, mu, sd, c, max(1000, int(p * mult * args.scale))) for o, mu, sd, c, p in W["libs"]]
    legacy_single = world == 1 and len(specs) == 1

    t_gen = time.time()
    packed = args.records == "packed"
    record_bytes = 20 if packed else RECORD_BYTES
    libs = [Library(i, s, n_contigs, world, rank, dev, synth, abi, first_library_rows, legacy_single, packed) for i, s in enumerate(specs)]
    t_gen = time.time() - t_gen
    crumb("libraries generated (%.1f s)" % t_gen)
    n_rec = sum(L.n_rec for L in libs)
    pairs_this_rank = n_rec / 2.0

    eng = CudaEngine(local_rank)
    for L in libs:
        eng.select_table(L.index)
        eng.set_contigs(L.rows, L.n_scaf, L.n_large)
    # run the library on torch's current stream: CUDA events recorded there bracket its kernels and
    # the NCCL collectives alike
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    eng.set_stream(torch.cuda.current_stream().cuda_stream)

    runner = None
    if world > 1:
        from besst_b200.dist import CudaBackend, DistributedGraphBuild
        runner = DistributedGraphBuild(CudaBackend(eng, dev), rank, world)

    def build(L, rec):
        eng.select_table(L.index)
        return runner.step(L.params, rec) if runner is not None else eng.build(L.params, rec)

    pos_tiles = {}   # per library: record tiles whose pos / mpos columns K1 fetched (BESST_CNT_POS_TILES)

    def step():
        out = []
        for L in libs:
            out.append(build(L, L.rec_dev))
            if L.index not in pos_tiles:
                torch.cuda.synchronize()   # N>1: the counters are all-reduced in place (sum over ranks)
                pos_tiles[L.index] = int(eng.links_counters()[abi.CNT_POS_TILES]) // world
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.set_profiling(True, accumulate=True)   # per-launch event pairs, read once after the timed region
    for _ in range(args.warmup):
        sizes = step()
    barrier()
    crumb("warm-up done")
    launches_warm = eng.kernel_launches()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    prof = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.kernel_profile()   # drop the warm-up launches
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        for L in libs:
            build(L, L.rec_dev)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    for kname, ms in eng.kernel_profile():
        prof.setdefault(kname, []).append(ms)
    eng.set_profiling(True)
    crumb("timed region done: %.3f ms/step on this rank" % (1e3 * (t1 - t0) / args.steps))
    exchange = None
    if world > 1 and rank == 0:   # what the exchange puts on NVLink, from the count matrix (`nvidia-smi nvlink -gt d` reports nothing on these boxes)
        exchange = {"model_bytes_out_per_library_rank0": int(getattr(runner, "exchange_bytes_out", 0)),
                    "model": "4 B (2 x u16) or 8 B per link + 24 B per run + 8 B per fishy key to other ranks; plus NCCL: count matrix, coverage + counter all-reduce"}
    launches_timed = eng.kernel_launches() - launches_warm   # kernels of this library inside the timed region
    # nvidia-smi samples every 100 ms and the timed region lasts a few tens of ms: keep the same load running
    # (untimed) until there are enough clock samples under load
    t_load = time.perf_counter()
    if world > 1:   # collectives inside: every rank must run the same number of steps
        for _ in range(max(4, min(80, int(0.6 / max((t1 - t0) / args.steps, 1e-3))))):
            step()
    else:
        while len(sampler.rows) < 6 and time.perf_counter() - t_load < 3.0:
            step()
    torch.cuda.synchronize()
    clocks = sampler.finish()
    wall = t1 - t0
    elapsed = ev0.elapsed_time(ev1) * 1e-3    # device time on the launching stream
    n_links = sum(int(s.n_links) for s in sizes)
    n_edges = sum(int(s.n_edges) for s in sizes)
    n_ll = sum(int(s.n_ll_links) for s in sizes)
    n_rec_total = n_rec
    if world > 1:
        t = torch.tensor([elapsed, wall], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, wall = float(t[0].item()), float(t[1].item())
        tot = torch.tensor([pairs_this_rank, launches_timed, n_links, n_edges, n_ll, n_rec], device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_pairs, launches_timed = float(tot[0].item()), int(tot[1].item())
        g_links, g_edges, g_ll, n_rec_total = int(tot[2].item()), int(tot[3].item()), int(tot[4].item()), int(tot[5].item())
    else:
        total_pairs = float(pairs_this_rank)
        g_links, g_edges, g_ll = n_links, n_edges, n_ll
    ms_per_step = 1e3 * elapsed / args.steps
    value = total_pairs * args.steps / elapsed

    # one more (untimed) step with device-synchronised phase marks: where the multi-GPU step spends its time
    dist_phases = None
    if runner is not None:
        runner.timing = True
        acc = {}
        for L in libs:
            build(L, L.rec_dev)
            for k, v in runner.phase_ms.items():
                acc[k] = acc.get(k, 0.0) + v
        runner.timing = False
        dist_phases = {k: round(v, 3) for k, v in acc.items()}
        barrier()

    # ---- roofline of the dominant kernel (per-launch CUDA events, timed region; this rank's share) ---------
    per_step = {k: float(np.sum(v)) / args.steps for k, v in prof.items()}
    n_launch = {k: len(v) / args.steps for k, v in prof.items()}
    n_tiles = sum((L.n_rec + 127) // 128 for L in libs)
    alg_bytes_step = {   # algorithmic bytes per STEP of each kernel family on this rank (DESIGN.md "Kernels")
        # K1 reads tid / mtid / flags of every record, pos / mpos only of the tiles that can hold a CreateEdge candidate
        # (counted by the kernel), writes one tuple per accepted link and one aggregate per tile
        "k_extract_links": (record_bytes - 8) * n_rec + 8 * 128 * sum(pos_tiles.values()) + TUPLE_BYTES * (n_links if world == 1 else 0) + 48 * n_tiles,
        "k_compact_tuples": 2 * TUPLE_BYTES * n_links + 8 * n_tiles,
        "k_radix_sweep": (8 + 8) * n_links,       # packed sort word (key | BAM index): 8 B in, 8 B out per pass
        "k_radix_hist": 8 * n_links,
        "k_edge_reduce": (8 + 8) * n_links + 64 * n_edges,    # k_edge_gather: grouped (o1,o2) in, obs_u/obs_v out
        "k_group_blocks": (16 + 8) * n_links + 8 * n_tiles,   # scratch tuples + tile offsets in, grouped observations out
        "k_score_keys": (8 + 8) * n_ll / 3.0 + 13 * n_edges,   # LL scan (over edges) + key build
        "k_ks_block": (4 + 4) * n_ll + 8 * n_edges,           # obs_u, obs_v of the scored links in, one double per edge out
        "k_ks_sort": (4 + 4) * n_ll,
        "k_ks_eval": (4 + 4) * n_ll + 8 * n_edges,
        "k_gapest": 64 * n_edges,
        "k_heads": 8 * n_links,
    }
    if world > 1:   # this rank's extraction produced the links it SENT, not the ones it owns now: use the global mean
        alg_bytes_step["k_extract_links"] += TUPLE_BYTES * (g_links // world)
    traffic = {}
    try:   # DRAM bytes per launch from the committed ncu capture of this workload (profiles/ncu_traffic.json)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("workload") == name and args.scale == 1.0 and world == 1 and tj.get("records", "plain") == args.records:
            traffic = tj
    except Exception:
        pass
    dominant = max(per_step, key=per_step.get) if per_step else None
    peak, peak_src = measured_peak_gbs()
    roofline = None
    if dominant is not None:
        launches = max(n_launch[dominant], 1)
        avg_ms = per_step[dominant] / launches
        per_launch = alg_bytes_step.get(dominant, 0) / launches
        ach = per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic.get(dominant), "peak_source": peak_src,
                    "avg_launch_ms": round(avg_ms, 4), "launches_per_step": n_launch[dominant],
                    "algorithmic_bytes_per_launch": int(per_launch),
                    "share_of_step": round(per_step[dominant] / max(sum(per_step.values()), 1e-9), 3),
                    "whole_step": {"bytes_alg": int(48 * pairs_this_rank + 8 * n_links + 64 * n_edges + 32 * n_contigs * len(libs)),
                                   "note": "SURVEY 8d: 48 B/pair + 8 B/link + 64 B/edge + 32 B/contig over the device-resident step"}}
        roofline["whole_step"]["frac"] = round(roofline["whole_step"]["bytes_alg"] / (ms_per_step * 1e-3) / 1e9 / peak, 4)
        # SURVEY 8d asks for both denominators: the measured copy bandwidth (`frac`) and the 8.0 TB/s nominal HBM3e figure
        roofline["frac_of_nominal_8TBps"] = round(ach / 8000.0, 4)
        roofline["whole_step"]["frac_of_nominal_8TBps"] = round(roofline["whole_step"]["bytes_alg"] / (ms_per_step * 1e-3) / 1e9 / 8000.0, 4)
    kernels = {k: {"ms_per_step": round(per_step[k], 4), "launches_per_step": n_launch[k],
                   "GBps": round(alg_bytes_step[k] / (per_step[k] * 1e-3) / 1e9, 1) if k in alg_bytes_step and per_step[k] > 0 else None}
               for k in sorted(per_step, key=per_step.get, reverse=True)}

    # ---- library metrics (K7) on the first library, records resident: the metric "when mu, sigma are not given" --
    libmetrics = None
    if world == 1 and libs[0].tlen is not None and not args.no_libmetrics:
        try:
            L0 = libs[0]
            eng.select_table(7)   # besst_libmetrics uploads its own table (in_largest flags): keep the builds' tables intact
            from besst_b200.libmetrics import metric_rows
            mrows = metric_rows(L0.lengths)   # the table get_metrics uploads: only the 1000-longest flags matter (libmetrics.py:231-233)
            eng.libmetrics(mrows, L0.params, None, L0.lengths, True, records=L0.rec_dev)   # warm-up
            torch.cuda.synchronize()
            reps = 3
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                rc, m, _ = eng.libmetrics(mrows, L0.params, None, L0.lengths, True, records=L0.rec_dev)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            kms = [x for kname, x in eng.kernel_profile() if kname == "k_metrics"]
            scanned = int(m.records_scanned)
            libmetrics = {"ms_per_call": round(ms, 4), "records_scanned": scanned, "samples": int(m.n_samples),
                          "kernel_ms": round(float(np.sum(kms)), 4), "kernel_launches": len(kms),
                          "algorithmic_bytes": 15 * scanned,
                          "GBps_kernels": round(15 * scanned / max(float(np.sum(kms)), 1e-6) / 1e6, 1),
                          "note": "capped BAM-order sampling (libmetrics.py:283-343): the scan stops at the reference's 1e6-sample caps; "
                                  "host call includes the O(bins) statistics and the read-back"}
        except Exception as exc:
            libmetrics = {"error": repr(exc)}

    # ---- end to end through the C ABI with pinned host buffers + CPU baseline + parity, library by library -------
    e2e = None
    cpu_baseline = None
    parity = None
    pe_level = None
    try:
        max_rec = max(L.n_rec for L in libs)
        host = {k: torch.empty((max_rec,), dtype=v.dtype, pin_memory=True) for k, v in libs[0].cols.items()}
        e2e_steps = max(1, min(args.steps, 3))
        e2e_time, h2d, d2h = 0.0, 0, 0
        oracle_s, oracle_pairs = 0.0, 0.0
        reports = []
        if (rank == 0 or world > 1) and not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle_lib
            if rank == 0:
                oracle_lib.build()
            barrier()
        for L in libs:
            crumb("library %d: e2e / cpu / parity legs" % L.index)
            for k in host:
                host[k][:L.n_rec].copy_(L.cols[k])
            torch.cuda.synchronize()
            hp = {k: v.data_ptr() for k, v in host.items()}
            hp["tlen"] = 0
            hp["n"] = L.n_rec
            rec_host = abi.make_records(hp, on_device=True)   # HOST pointers: the flag is flipped below
            rec_host.on_device = 0

            def e2e_step():   # host columns in (sliced H2D overlapped with K1 inside the library), this rank's CSR out
                s = build(L, rec_host)
                return runner.fetch_local(view=True) if runner is not None else eng.fetch_view(s)
            if not args.no_e2e:
                e2e_step()   # warm-up: staging buffers, pinned result buffers
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    res = e2e_step()
                barrier()
                e2e_time += time.perf_counter() - t0
                h2d += record_bytes * L.n_rec
                d2h += result_bytes(res, abi)
            # ---- CPU: the C oracle over the WHOLE library (every rank its own BAM-order slice), then parity ----
            if not args.no_cpu:
                from besst_b200.records import RecordBatch
                arrs = {k: host[k][:L.n_rec].numpy() for k in host if k != "packed"}
                if packed:   # the oracle reads the three plain columns
                    pk = host["packed"][:L.n_rec].numpy().view(np.uint32)
                    arrs["flag"] = (pk & np.uint32(0xfff)).astype(np.uint16)
                    arrs["mapq"] = ((pk >> np.uint32(12)) & np.uint32(0xff)).astype(np.uint8)
                    arrs["qlen"] = (pk >> np.uint32(20)).astype(np.int32)
                else:
                    arrs["flag"] = arrs["flag"].view(np.uint16)
                batch = RecordBatch(tlen=np.zeros(L.n_rec, np.int32), **arrs)
                if world == 1:
                    t0 = time.perf_counter()
                    want, _, _, _ = oracle_lib.graph_build(L.rows, L.n_scaf, L.params, batch)
                    t_cpu = time.perf_counter() - t0
                    got = eng.fetch(build(L, L.rec_dev))
                    scored = (want.flags & abi.EDGE_SCORED) != 0
                    reports.append({"library": L.index, "records": L.n_rec, "edges": int(want.n_edges), "links": int(want.n_links),
                                    "oracle_digest": checksum(want), "cuda_digest": checksum(got),
                                    "integers_bit_exact": checksum(want) == checksum(got),
                                    "gap_equal": bool(np.array_equal(got.gap[scored], want.gap[scored])),
                                    "score_max_rel_diff": float(np.nanmax(np.abs(got.score[scored] - want.score[scored]) / np.maximum(np.abs(want.score[scored]), 1e-300))) if scored.any() else 0.0})
                    oracle_s += t_cpu
                    oracle_pairs += L.n_rec / 2.0
                else:
                    import slice_oracle
                    build(L, L.rec_dev)
                    got = slice_oracle.gather_owned(dist, rank, world, runner.fetch_local(), group=host_group)
                    want = slice_oracle.sliced_oracle(dist, rank, world, L.rows, L.n_scaf, L.params, batch, group=host_group)
                    if rank == 0:
                        rep = slice_oracle.compare(got, want)
                        rep["library"] = L.index
                        reports.append(rep)
                        oracle_s += want["seconds"]
                    oracle_pairs += L.n_rec / 2.0
                del batch
            # ---- the reference-facing entry points end to end (N=1, first library): get_metrics + PE ----------------
            if world == 1 and L.index == 0 and L.tlen is not None and args.pe_level:
                try:
                    hc = {k: v for k, v in host.items() if k != "packed"}
                    for k, v in dict(L.plain, tlen=L.tlen).items():
                        hc[k] = torch.empty((L.n_rec,), dtype=v.dtype, pin_memory=True)
                        hc[k].copy_(v)
                    if packed:
                        hc["packed"] = host["packed"]
                    torch.cuda.synchronize()
                    pe_level_leg(eng, L, hc, True)   # warm-up: imports (networkx), staging buffers
                    pe_level = {"lazy": pe_level_leg(eng, L, hc, True)}
                    if L.n_rec <= 60_000_000:   # python lists of every observation: minutes and tens of GB beyond that
                        pe_level["lists"] = pe_level_leg(eng, L, hc, False)
                    del hc
                    for Lk in libs:   # PE uploaded its own contig table into the current slot
                        eng.select_table(Lk.index)
                        eng.set_contigs(Lk.rows, Lk.n_scaf, Lk.n_large)
                except Exception as exc:
                    import traceback
                    traceback.print_exc(file=sys.stderr)
                    pe_level = {"error": repr(exc)}
        if not args.no_e2e:
            if world > 1:
                t = torch.tensor([e2e_time], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                b = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
                dist.all_reduce(b, op=dist.ReduceOp.SUM)
                e2e_time, h2d, d2h = float(t.item()), int(b[0].item()), int(b[1].item())
            e2e = {"value": total_pairs * e2e_steps / e2e_time, "unit": "read-pairs/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": round(1e3 * e2e_time / e2e_steps, 3), "steps": e2e_steps,
                   "h2d_GBps_aggregate": round(h2d * e2e_steps / e2e_time / 1e9, 1)}
        if not args.no_cpu:
            if world > 1:
                t = torch.tensor([oracle_pairs], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                oracle_pairs = float(t.item())
            if rank == 0:
                cpu_baseline = {"value": oracle_pairs / max(oracle_s, 1e-9), "unit": "read-pairs/s", "cores": world, "kind": "port",
                                "sample": "the whole workload (%d pairs), C oracle (oracle/besst_oracle.c), %s, %.1f s"
                                          % (int(oracle_pairs), "one sequential pass on one core" if world == 1 else
                                             "every rank scans its BAM-order slice on its own core, slices merged (oracle/slice_oracle.py)", oracle_s)}
                parity = {"coverage": "whole workload: every library, every record, every link",
                          "integers_bit_exact": all(r["integers_bit_exact"] for r in reports),
                          "gap_equal": all(r.get("gap_equal", False) for r in reports),
                          "score_max_rel_diff": max(r.get("score_max_rel_diff", 0.0) for r in reports), "libraries": reports}
    except Exception as exc:   # pinned allocation can fail on small hosts: say so instead of inventing a number
        import traceback
        traceback.print_exc(file=sys.stderr)
        e2e = e2e or {"value": None, "unit": "read-pairs/s", "error": repr(exc)}

    ingest = None
    if world == 1 and args.ingest_pairs > 0:
        crumb("ingest leg")
        try:
            ingest = ingest_leg(eng, int(args.ingest_pairs))
            for Lk in libs:   # the leg used table slot 7 and the current-slot selection
                eng.select_table(Lk.index)
        except Exception as exc:
            import traceback
            traceback.print_exc(file=sys.stderr)
            ingest = {"error": repr(exc)}
    crumb("legs done")
    if rank == 0:
        line = {
            "metric": "read-pairs/s through graph-build+GapEst", "value": value, "unit": "read-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak" if (W["weak"] or world == 1) else "strong", "vs_baseline": None,
            "dtype": "int32/int64 + f64 (GapEst)", "data": "synthetic",
            "config": {"workload": "%s: %s%s" % (name, W["desc"], (" -- scaled x%d: ONE global library of %d contigs, range-partitioned in BAM order"
                                                                  % (world, n_contigs)) if (W["weak"] and world > 1) else ""),
                       "contigs": n_contigs, "pairs_per_step": int(total_pairs), "records_per_step": n_rec_total,
                       "libraries": [L.describe() for L in libs],
                       "accepted_links": g_links, "edges": g_edges, "scored_links": g_ll,
                       "record_tiles_with_positions": "%d of %d on rank 0" % (sum(pos_tiles.values()), n_tiles),
                       "record_format": ("packed: tid, mtid, pos, mpos int32 + one uint32 flag|mapq<<12|qlen<<20 = 20 B/record (besst_records.packed, "
                                         "written by the ingest library)") if packed else "plain: tid, mtid, pos, mpos, qlen int32 + flag u16 + mapq u8 = 23 B/record",
                       "l2_policy": "inputs (%.1f GB per rank) larger than the 126 MB L2" % (record_bytes * n_rec / 1e9),
                       "parallelism": ("1 process per GPU; every library ONE global BAM range-partitioned over the ranks; runs of links "
                                       "routed by edge hash and stored into the destination GPU's memory by the pack kernel (NVLink peer "
                                       "stores), one all_gather of sizes, one all_reduce of coverage+counters") if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_timed),
            "wall_ms_per_step": round(1e3 * wall / args.steps, 4),
            "timing": "CUDA events on the launching stream (the library runs on torch's current stream), max over ranks",
            "dist_phases_ms": dist_phases, "exchange": exchange,
            "roofline": roofline, "kernels": kernels, "libmetrics": libmetrics, "pe_level": pe_level, "ingest": ingest, "cpu_baseline": cpu_baseline, "parity": parity,
            "generate_s": round(t_gen, 2), "impl": "ours",
        }
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU implementation of the path: its Python cannot travel to
    the GPU box (and needs pysam/mathstats), so this times the C oracle port of it
    on a bounded sample of the same workload, one core (the reference is single-threaded)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    from besst_b200 import abi, synth
    from besst_b200.contig_table import first_library_rows
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    name = args.workload or DEFAULT_WORKLOAD.get(world, "config3")
    W = WORKLOADS[name]
    total = float(sum(s[4] for s in W["libs"]))
    frac = min(1.0, REFERENCE_SAMPLE_PAIRS / total) * args.scale   # 10 M pairs over all libraries: ~0.5 s per step on one core
    n_contigs = max(2, int(W["contigs"] * frac))
    libs = []
    for i, (orientation, mu, sigma, cont, pairs) in enumerate(W["libs"]):
        lib = synth.make_library(n_contigs, max(1000, int(pairs * frac)), orientation, mu, sigma, cont,
                                 seed=synth.SEED0 + 2 + 100 * i, device="cpu", with_names=False)
        lengths = lib.lengths.numpy()
        rows, n_scaf, _ = first_library_rows(lengths, mu + 4 * sigma) if i == 0 else synth.later_library_rows(lengths, mu + 4 * sigma, seed=synth.SEED0 + 31 * i)
        libs.append((lib.to_batch(), rows, n_scaf, library_params(abi, orientation, mu, sigma), lib.n_pairs))
    n_pairs = sum(l[4] for l in libs)

    def step():
        for batch, rows, n_scaf, params, _ in libs:
            oracle_lib.graph_build(rows, n_scaf, params, batch)
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n_pairs * args.steps / dt
    sample = "%d contigs / %d pairs in %d libraries (%.2f%% of the workload, same generator), C oracle port, 1 core" % (
        n_contigs, n_pairs, len(libs), 100.0 * frac)
    print(json.dumps({
        "impl": "reference", "metric": "read-pairs/s through graph-build+GapEst", "value": value,
        "unit": "read-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True,
        "scaling": "weak" if (W["weak"] or world == 1) else "strong", "vs_baseline": None,
        "dtype": "int32/int64 + f64 (GapEst)", "data": "synthetic",
        "config": {"workload": "%s: %s" % (name, W["desc"]), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "read-pairs/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "read-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: config3 at 1 and 2 GPUs, config4 at 4, config5 at 8 (BASELINE.json configs)")
    ap.add_argument("--records", default="packed", choices=["packed", "plain"],
                    help="record layout handed to the graph build: packed = 20 B/record (default), plain = the 23 B/record columns of round 1")
    ap.add_argument("--scale", type=float, default=1.0, help="debug: shrink the workload")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiler runs only)")
    ap.add_argument("--no-libmetrics", action="store_true", help="skip the library-metrics leg")
    ap.add_argument("--no-pe-level", dest="pe_level", action="store_false",
                    help="skip the leg that times get_metrics + CreateGraph.PE (host records -> networkx graphs) at N=1")
    ap.add_argument("--ingest-pairs", type=int, default=2000000,
                    help="N=1: read pairs of the BAM file written for the file -> graph leg (0: skip the leg)")
    ap.add_argument("--watchdog", type=int, default=900, help="seconds after which a stuck run ends itself (0: never)")
    args = ap.parse_args()
    arm_watchdog(args.watchdog)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
