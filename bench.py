#!/usr/bin/env python
"""Benchmark of the hot path: read-pairs/s through graph build + GapEst.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (records -> CSR edges + link statistics +
KS/GapEst score) over one synthetic library (SURVEY.md 8d generator).  At N=1
the workload is BASELINE.json's config 3 (100k contigs / 200 M MP pairs, rf):
the configuration the north-star target (>= 100 M read-pairs/s) is quoted on.
At N>1 every rank owns one such library slice (weak scaling), link tuples are
exchanged once with an NCCL all-to-all keyed by the edge hash.

Prints ONE JSON line (rank 0).  `value` = whole-job read-pairs/s with the
records resident in HBM; `e2e` = the same through the C-ABI call with pinned
HOST buffers (H2D of the record columns and D2H of the result inside the timed
region); `roofline` = the dominant kernel against the measured HBM peak;
`cpu_baseline` = the C oracle (a port of the reference's Python path) on one
host core over a bounded sample.  `--impl reference` times that CPU port alone.
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
# NCCL's version banner (NCCL_DEBUG=VERSION/WARN/INFO) goes to stdout: keep the one JSON line alone
os.environ.pop("NCCL_DEBUG", None)
if os.environ.get("BESST_NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = os.environ["BESST_NCCL_DEBUG"]

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (synth config, description)
    "config3": ("config3", "synthetic 100k contigs / 200 M MP pairs (rf), 1 library"),
    "config2": ("config2", "synthetic 10k contigs / 20 M PE pairs, 1 library"),
    "small": ("small_mp", "synthetic 400 contigs / 200 k MP pairs (debug)"),
}
RECORD_BYTES = 4 + 4 + 4 + 4 + 4 + 2 + 1   # tid mtid pos mpos qlen flag mapq (tlen is only read by libmetrics)
TUPLE_BYTES = 16
CPU_SAMPLE_RECORDS = 40_000_000


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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from besst_b200 import abi, synth
    from besst_b200.contig_table import first_library_rows
    from besst_b200.engine import CudaEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg_name, desc = WORKLOADS[args.workload]
    n_contigs, n_pairs, orientation, mu, sigma, cont = synth.CONFIGS[cfg_name]
    n_contigs = max(2, int(n_contigs * args.scale))
    n_pairs = max(1000, int(n_pairs * args.scale))
    t_gen = time.time()
    lib = synth.make_library(n_contigs, n_pairs, orientation, mu, sigma, cont,
                             seed=synth.SEED0 + 2 + 1000 * rank, device=dev, with_names=False)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    n_rec = lib.n_records
    pairs_per_rank = lib.n_pairs
    contig_threshold = mu + 4 * sigma
    lengths_all = [lib.lengths.numpy()]
    if world > 1:   # global contig table = concatenation of every rank's contig block
        lengths_all = [synth.make_contigs(n_contigs, synth.SEED0 + 2 + 1000 * r)[0].numpy() for r in range(world)]
    rows, n_scaf, n_large = first_library_rows(np.concatenate(lengths_all), contig_threshold)
    cols = dict(lib.cols)
    if world > 1:
        cols["tid"] = cols["tid"] + rank * n_contigs
        cols["mtid"] = cols["mtid"] + rank * n_contigs

    eng = CudaEngine(local_rank)
    eng.set_contigs(rows, n_scaf, n_large)
    # run the library on torch's current stream: CUDA events recorded there bracket its kernels and
    # the NCCL collectives alike
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    params = library_params(abi, orientation, mu, sigma)
    ptrs = {k: v.data_ptr() for k, v in cols.items()}
    ptrs["n"] = n_rec
    rec_dev = abi.make_records(ptrs, on_device=True)

    runner = None
    if world > 1:
        from besst_b200.dist import CudaBackend, DistributedGraphBuild
        runner = DistributedGraphBuild(CudaBackend(eng, dev), rank, world)
        step = lambda: runner.step(params, rec_dev)      # noqa: E731
    else:
        step = lambda: eng.build(params, rec_dev)        # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.set_profiling(True)
    for _ in range(args.warmup):
        sizes = step()
    barrier()
    launches_warm = eng.kernel_launches()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    prof = {}
    dev_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        sizes = step()
        for name, ms in eng.kernel_profile():
            prof.setdefault(name, []).append(ms)
        dev_ms.append(eng.timing()[0] if world == 1 else 0.0)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    launches_timed = eng.kernel_launches() - launches_warm   # kernels of this library inside the timed region
    # nvidia-smi samples every 100 ms and the timed region lasts a few tens of ms: keep the same load running
    # (untimed) until there are enough clock samples under load
    t_load = time.perf_counter()
    if world > 1:   # collectives inside: every rank must run the same number of steps
        for _ in range(80):
            step()
    else:
        while len(sampler.rows) < 6 and time.perf_counter() - t_load < 3.0:
            step()
    torch.cuda.synchronize()
    clocks = sampler.finish()
    wall = t1 - t0
    elapsed = ev0.elapsed_time(ev1) * 1e-3    # device time on the launching stream
    if world > 1:
        t = torch.tensor([elapsed, wall], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, wall = float(t[0].item()), float(t[1].item())
        tot = torch.tensor([pairs_per_rank, launches_timed, int(sizes.n_links), int(sizes.n_edges), int(sizes.n_ll_links)],
                           device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_pairs, launches_timed = float(tot[0].item()), int(tot[1].item())
    else:
        total_pairs = float(pairs_per_rank)
    ms_per_step = 1e3 * elapsed / args.steps
    value = total_pairs * args.steps / elapsed

    # ---- roofline of the dominant kernel (per-launch CUDA events, timed region) ---------
    per_step = {k: float(np.sum(v)) / args.steps for k, v in prof.items()}
    n_launch = {k: len(v) / args.steps for k, v in prof.items()}
    n_links, n_edges, n_ll = int(sizes.n_links), int(sizes.n_edges), int(sizes.n_ll_links)
    alg_bytes = {   # algorithmic bytes per launch of each kernel (DESIGN.md "Kernels")
        "k_extract_links": RECORD_BYTES * n_rec + TUPLE_BYTES * n_links + 48 * ((n_rec + 127) // 128),
        "k_compact_tuples": 2 * TUPLE_BYTES * n_links + 8 * ((n_rec + 127) // 128),
        "k_radix_sweep": (8 + 8) * n_links,       # packed sort word (key | BAM index): 8 B in, 8 B out per pass
        "k_radix_hist": 8 * n_links,
        "k_edge_reduce": (8 + 8) * n_links + 64 * n_edges,    # k_edge_gather: grouped (o1,o2) in, obs_u/obs_v out
        "k_group_blocks": (16 + 8) * n_links + 8 * ((n_rec + 127) // 128),   # scratch tuples + tile offsets in, grouped observations out
        "k_score_keys": (8 + 8) * n_ll / 3.0 + 13 * n_edges,   # 3 launches: LL scan (2, over edges) + key build
        "k_ks_block": (4 + 4) * n_ll + 8 * n_edges,           # obs_u, obs_v of the scored links in, one double per edge out
        "k_ks_sort": (4 + 4) * n_ll,
        "k_ks_eval": (4 + 4) * n_ll + 8 * n_edges,
        "k_gapest": 64 * n_edges,
        "k_heads": 8 * n_links,
    }
    traffic = {}
    try:   # DRAM bytes per launch from the committed ncu capture of this workload (profiles/ncu_traffic.json)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("workload") == args.workload and args.scale == 1.0:
            traffic = tj
    except Exception:
        pass
    dominant = max(per_step, key=per_step.get) if per_step else None
    peak, peak_src = measured_peak_gbs()
    roofline = None
    if dominant is not None:
        avg_ms = per_step[dominant] / max(n_launch[dominant], 1)
        ach = alg_bytes.get(dominant, 0) / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic.get(dominant), "peak_source": peak_src,
                    "avg_launch_ms": round(avg_ms, 4), "launches_per_step": n_launch[dominant],
                    "algorithmic_bytes_per_launch": int(alg_bytes.get(dominant, 0)),
                    "share_of_step": round(per_step[dominant] / max(sum(per_step.values()), 1e-9), 3)}
    kernels = {k: {"ms_per_step": round(per_step[k], 4), "launches_per_step": n_launch[k],
                   "GBps": round(alg_bytes[k] * n_launch[k] / (per_step[k] * 1e-3) / 1e9, 1) if k in alg_bytes and per_step[k] > 0 else None}
               for k in sorted(per_step, key=per_step.get, reverse=True)}

    # ---- end to end through the C ABI with pinned host buffers (N=1 path per rank) -------
    e2e = None
    cpu_baseline = None
    parity = None
    try:
        host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in cols.items() if k != "tlen"}
        for k in host:
            host[k].copy_(cols[k])
        torch.cuda.synchronize()
        hp = {k: v.data_ptr() for k, v in host.items()}
        hp["tlen"] = 0
        hp["n"] = n_rec
        rec_host = abi.make_records(hp, on_device=True)
        rec_host.on_device = 0
        e2e_steps = max(1, min(args.steps, 3))
        if world == 1 and not args.no_e2e:
            res = None
            eng.fetch_view(eng.build(params, rec_host))   # warm-up: staging buffers, pinned result buffers
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                s = eng.build(params, rec_host)
                res = eng.fetch_view(s)
            barrier()
            t1 = time.perf_counter()
            d2h = sum(int(getattr(res, f).nbytes) for f in ("edge_u", "edge_v", "nr_links", "obs_sum", "obs_sq",
                      "first_idx", "row_ptr", "gap", "score", "ks", "sd_obs", "sd_model", "fishy", "flags", "obs_u",
                      "obs_v", "aligned_len")) + 8 * abi.N_COUNTERS
            e2e = {"value": pairs_per_rank * e2e_steps / (t1 - t0), "unit": "read-pairs/s",
                   "h2d_bytes_per_step": RECORD_BYTES * n_rec, "d2h_bytes_per_step": d2h,
                   "ms_per_step": round(1e3 * (t1 - t0) / e2e_steps, 3), "steps": e2e_steps}
        # ---- CPU baseline: the C oracle on one host core over a bounded sample -------------
        if rank == 0 and world == 1 and not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle_lib
            from besst_b200.records import RecordBatch
            m = min(n_rec, CPU_SAMPLE_RECORDS)
            arrs = {k: host[k][:m].numpy() for k in host}
            arrs["flag"] = arrs["flag"].view(np.uint16)
            sample = RecordBatch(tlen=np.zeros(m, np.int32), **arrs)
            t0 = time.perf_counter()
            want, _, _, _ = oracle_lib.graph_build(rows, n_scaf, params, sample)
            t_cpu = time.perf_counter() - t0
            cpu_baseline = {"value": (m / 2) / t_cpu, "unit": "read-pairs/s", "cores": 1, "kind": "port",
                            "sample": "first %d records (%d pairs) of the same library, C oracle "
                                      "(oracle/besst_oracle.c), %.1f s" % (m, m // 2, t_cpu)}
            keep = []
            got = eng.fetch(eng.build(params, abi.make_records(sample, keepalive=keep)))
            parity = {"sample_records": m, "oracle_digest": checksum(want), "cuda_digest": checksum(got),
                      "integers_bit_exact": checksum(want) == checksum(got),
                      "gap_equal": bool(np.array_equal(got.gap, want.gap)),
                      "score_max_rel_diff": float(np.nanmax(np.abs(got.score - want.score) / np.maximum(np.abs(want.score), 1e-300))) if want.n_edges else 0.0}
    except Exception as exc:   # pinned allocation can fail on small hosts: say so instead of inventing a number
        e2e = e2e or {"value": None, "unit": "read-pairs/s", "error": repr(exc)}

    if world > 1:
        try:
            e2e_steps = max(1, min(args.steps, 3))

            def e2e_step():   # host columns in (sliced H2D overlapped with K1 inside the library), this rank's CSR out
                runner.step(params, rec_host)
                return runner.fetch_local(view=True)
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                res = e2e_step()
            barrier()
            dt = time.perf_counter() - t0
            d2h = sum(int(getattr(res, f).nbytes) for f in ("edge_u", "edge_v", "nr_links", "obs_sum", "obs_sq", "first_idx",
                      "row_ptr", "gap", "score", "ks", "sd_obs", "sd_model", "fishy", "flags", "obs_u", "obs_v", "aligned_len"))
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            b = torch.tensor([RECORD_BYTES * n_rec, d2h], device=dev, dtype=torch.float64)
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
            e2e = {"value": total_pairs * e2e_steps / float(t.item()), "unit": "read-pairs/s",
                   "h2d_bytes_per_step": int(b[0].item()), "d2h_bytes_per_step": int(b[1].item()),
                   "ms_per_step": round(1e3 * float(t.item()) / e2e_steps, 3), "steps": e2e_steps}
        except Exception as exc:
            e2e = {"value": None, "unit": "read-pairs/s", "error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "read-pairs/s through graph-build+GapEst", "value": value, "unit": "read-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 + f64 (GapEst)",
            "data": "synthetic",
            "config": {"workload": desc + (" x%d ranks" % world if world > 1 else ""), "contigs_per_rank": n_contigs,
                       "pairs_per_rank": pairs_per_rank, "records_per_rank": n_rec, "orientation": orientation,
                       "mean_ins_size": mu, "std_dev_ins_size": sigma, "accepted_links": n_links, "edges": n_edges, "scored_links": n_ll,
                       "l2_policy": "inputs (%.1f GB) larger than the 126 MB L2" % (RECORD_BYTES * n_rec / 1e9),
                       "parallelism": "1 process per GPU; tuples all-to-all by edge hash" if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_timed),
            "device_ms_per_step": round(float(np.mean(dev_ms)), 4) if world == 1 else None,
            "wall_ms_per_step": round(1e3 * wall / args.steps, 4),
            "timing": "CUDA events on the launching stream (the library runs on torch's current stream), max over ranks",
            "dist_phases_ms": ({k: round(v, 3) for k, v in runner.phase_ms.items()} if (runner is not None and runner.timing) else None),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "parity": parity,
            "generate_s": round(t_gen, 2), "impl": "ours",
        }
        print(json.dumps(line))
    if world > 1:
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
    cfg_name, desc = WORKLOADS[args.workload]
    n_contigs, n_pairs, orientation, mu, sigma, cont = synth.CONFIGS[cfg_name]
    frac = min(1.0, (CPU_SAMPLE_RECORDS / 2) / (2.0 * n_pairs)) * args.scale   # 10 M pairs: ~0.5 s per step on one core
    lib = synth.make_library(max(2, int(n_contigs * frac)), max(1000, int(n_pairs * frac)), orientation, mu, sigma, cont,
                             seed=synth.SEED0 + 2, device="cpu", with_names=False)
    batch = lib.to_batch()
    rows, n_scaf, n_large = first_library_rows(lib.lengths.numpy(), mu + 4 * sigma)
    params = library_params(abi, orientation, mu, sigma)
    for _ in range(min(args.warmup, 1)):
        oracle_lib.graph_build(rows, n_scaf, params, batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res, _, _, _ = oracle_lib.graph_build(rows, n_scaf, params, batch)
    dt = time.perf_counter() - t0
    value = lib.n_pairs * args.steps / dt
    sample = "%d contigs / %d pairs (%.2f%% of the workload, same generator), C oracle port, 1 core" % (
        len(lib.names), lib.n_pairs, 100.0 * frac)
    print(json.dumps({
        "impl": "reference", "metric": "read-pairs/s through graph-build+GapEst", "value": value,
        "unit": "read-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32/int64 + f64 (GapEst)", "data": "synthetic",
        "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "read-pairs/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "read-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="debug: shrink the workload")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiler runs only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
