"""Multi-GPU graph build: one process per GPU, records range-partitioned in BAM
order, ONE exchange step.

The reference is a single sequential scan (CreateGraph.py:111-211); the only
state that crosses a record boundary is "(obs1,obs2) of the previous CreateEdge
call" (:835-838,869-870).  The decomposition therefore is:

  1. halo    every rank finds the last CreateEdge observation of its own slice
             (extract over the slice tail), an all_gather of 3 integers gives
             each rank the last call of all preceding ranks;
  2. extract records -> accepted link tuples of the slice (besst_links_extract,
             with the halo) -- identical to what the sequential scan would do;
  3. exchange tuples are bucketed by hash(edge) mod world (besst_links_partition,
             stable) and sent with ONE all_to_all over NVLink; buckets arrive in
             source-rank order = global BAM order;
  4. build   every rank owns a disjoint edge set: bucket -> CSR -> KS -> GapEst
             (besst_links_to_graph), no further communication;
  5. reduce  per-contig aligned length and the counters: one all_reduce(sum).

`torch.distributed` is plumbing (NCCL on GPUs, gloo in the CPU tests); all data
movement on the device side goes through the C ABI.  The backend object hides
the engine so that the host logic can be exercised on CPU with gloo.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

NO_MATCH = -(2 ** 31) + 1   # halo value no observation pair can equal


class _SymmUnavailable(RuntimeError):
    pass


def _copy_params(params, halo):
    p = abi.LibParams()
    C.memmove(C.byref(p), C.byref(params), C.sizeof(abi.LibParams))
    p.halo_prev_obs1, p.halo_prev_obs2 = int(halo[0]), int(halo[1])
    return p


def records_tail(rec, count):
    """View of the last `count` records of a device-resident Records struct
    (start rounded down to a multiple of 16 records so that every column stays
    16-byte aligned for the 128-bit loads)."""
    n = int(rec.n)
    lo = max(0, n - count)
    lo -= lo % 16
    r = abi.Records()
    r.n = n - lo
    for name, width in (("tid", 4), ("mtid", 4), ("pos", 4), ("mpos", 4), ("tlen", 4), ("qlen", 4), ("flag", 2), ("mapq", 1)):
        base = getattr(rec, name)
        setattr(r, name, (base + lo * width) if base else None)
    r.on_device = rec.on_device
    return r


class CudaBackend(object):
    """The engine + torch device buffers behind DistributedGraphBuild."""

    def __init__(self, engine, device):
        import torch
        self.torch = torch
        self.engine = engine
        self.device = device
        self._bufs = {}
        self._symm = {}
        self._stream = None
        self.bind_stream()

    def bind_stream(self):
        """Run the engine on torch's CURRENT stream of this device: the NCCL collectives, the symmetric-memory
        barriers and torch's own ops in DistributedGraphBuild are enqueued there, and the engine's kernels read
        and write the same buffers (receive buffers, partial sums).  Called at construction and at the start
        of every step (a no-op unless the caller switched streams in between)."""
        s = int(self.torch.cuda.current_stream(self.device).cuda_stream)
        if s != self._stream:
            # stream 0 is the legacy default stream: the ABI spells it cudaStreamLegacy (0x1), NULL means "ctx-owned"
            self.engine.set_stream(s if s else 1)
            self._stream = s

    def _buf(self, name, n, dtype):
        t = self._bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = self.torch.empty(max(int(n * 1.1) + 1024, 1024), dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def records(self, batch):
        """abi.Records over a host RecordBatch (kept alive by the backend until the next call)"""
        self._keep = []
        return abi.make_records(batch, keepalive=self._keep)

    def tail_last_call(self, params, rec):
        """(has, obs1, obs2) of the last CreateEdge call in the slice."""
        n = int(rec.n)
        count = 1 << 16
        p = _copy_params(params, (NO_MATCH, NO_MATCH))
        while True:
            self.engine.links_extract(p, records_tail(rec, count))
            c = self.engine.links_counters()
            if c[abi.CNT_CALLS] > 0:
                return 1, int(c[abi.CNT_LAST_OBS1]), int(c[abi.CNT_LAST_OBS2])
            if count >= n:
                return 0, 0, 0
            count *= 16

    def extract(self, params, rec):
        return self.engine.links_extract(params, rec)

    def partition(self, world):
        torch = self.torch
        (_, n), (_, nf) = self.engine.links_device()
        send_t = self._buf("send_t", 4 * n, torch.int32)
        send_o = self._buf("send_o", n, torch.int32)
        send_f = self._buf("send_f", nf, torch.int64)
        tc, fc = self.engine.links_partition(world, send_t.data_ptr(), send_f.data_ptr(), send_o.data_ptr())
        return send_t[:4 * n].view(-1, 4), send_o[:n], send_f[:nf], tc, fc

    # run-level exchange
    def call_summary(self):
        """(number of CreateEdge calls, (obs1, obs2) of the last one, of the first one) of the last extraction"""
        c = self.engine.links_counters()
        return (int(c[abi.CNT_CALLS]), (int(c[abi.CNT_LAST_OBS1]), int(c[abi.CNT_LAST_OBS2])),
                (int(c[abi.CNT_FIRST_OBS1]), int(c[abi.CNT_FIRST_OBS2])))

    def group(self):
        return self.engine.links_group()

    def route(self, world):
        return self.engine.runs_route(world)

    def partition_fishy(self, world):
        torch = self.torch
        _, nf = self.engine.fishy_device()
        send_f = self._buf("send_f", nf, torch.int64)
        _, fc = self.engine.links_partition(world, None, send_f.data_ptr(), None)
        return send_f[:nf], fc

    def prepare_exchange(self, world, n_local):
        """call_summary + group + route + partition_fishy in one engine call with ONE host read.
        -> (calls, last, first, n_runs or None, link_counts, run_counts, send_f, fishy_counts)"""
        torch = self.torch
        _, nf = self.engine.fishy_device()
        send_f = self._buf("send_f", nf, torch.int64)
        s, lc, rc, fc = self.engine.exchange_prepare(world, send_f.data_ptr() if nf else None)
        return (int(s[2]), (int(s[3]), int(s[4])), (int(s[5]), int(s[6])), int(s[1]) if s[0] else None, lc, rc, send_f[:nf], fc)

    def obs_words(self, params):
        """int32 words per exchanged link: 1 (two 16-bit observations) or 2"""
        return self.engine.runs_obs_bytes(params) // 4

    def pack(self, world, n_links, n_runs, w=2):
        torch = self.torch
        send_obs = self._buf("send_obs", w * n_links, torch.int32)
        send_desc = self._buf("send_desc", 6 * n_runs, torch.int32)
        self.engine.runs_pack(world, send_obs.data_ptr(), send_desc.data_ptr())
        return send_obs[:w * n_links].view(-1, w), send_desc[:6 * n_runs].view(-1, 6)

    def recv_run_buffers(self, n_links, n_runs, n_fishy, w=2):
        torch = self.torch
        return (self._buf("recv_obs", w * n_links, torch.int32)[:w * n_links].view(-1, w),
                self._buf("recv_desc", 6 * n_runs, torch.int32)[:6 * n_runs].view(-1, 6),
                self._buf("recv_f", n_fishy, torch.int64)[:n_fishy])

    def runs_to_graph(self, params, recv_obs, recv_desc, world, block_bits, src_runs, src_links, src_first, recv_f):
        return self.engine.runs_to_graph(params, recv_obs.data_ptr() if recv_obs.numel() else None, recv_obs.shape[0],
                                         recv_desc.data_ptr() if recv_desc.numel() else None, recv_desc.shape[0], world,
                                         block_bits, src_runs, src_links, src_first,
                                         recv_f.data_ptr() if recv_f.numel() else None, recv_f.shape[0])

    # peer-mapped receive buffers (torch symmetric memory over NVLink) for the fused pack + exchange
    def symm_buffer(self, name, n_elems, dtype, group):
        """A symmetric buffer of at least n_elems (every rank passes the same n_elems: the allocation and
        the rendezvous are collective).  -> (local tensor, handle with .buffer_ptrs / .barrier / .get_buffer)"""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as sm
        cur = self._symm.get(name)
        if cur is None or cur[0].numel() < n_elems or cur[0].dtype != dtype:
            cap = int(n_elems * 1.25) + 4096
            t = sm.empty(cap, dtype=dtype, device=self.device)
            h = sm.rendezvous(t, group=group if group is not None else dist.group.WORLD)
            self._symm[name] = (t, h)
        return self._symm[name]

    def pack_peer(self, world, obs_ptrs, desc_ptrs):
        self.engine.runs_pack_peer(world, obs_ptrs, desc_ptrs)

    def recv_buffers(self, n_tuples, n_fishy):
        torch = self.torch
        return (self._buf("recv_t", 4 * n_tuples, torch.int32)[:4 * n_tuples].view(-1, 4),
                self._buf("recv_o", n_tuples, torch.int32)[:n_tuples],
                self._buf("recv_f", n_fishy, torch.int64)[:n_fishy])

    def to_graph(self, params, recv_t, recv_f):
        return self.engine.links_to_graph(params, recv_t.data_ptr() if recv_t.numel() else None, recv_t.shape[0],
                                          recv_f.data_ptr() if recv_f.numel() else None, recv_f.shape[0])

    def partial_tensors(self):
        """(aligned_len[C], counters[16]) as in-place reducible device tensors."""
        a_ptr, c_ptr = self.engine.links_partials_device()
        return (_wrap_device_i64(self.torch, a_ptr, self.engine._n_contigs, self.device),
                _wrap_device_i64(self.torch, c_ptr, abi.N_COUNTERS, self.device))

    def partial_span(self):
        """aligned_len[C], one spare word and counters[16] as ONE contiguous tensor (they share an allocation), with views of
        the two parts: a single all-reduce covers both.  -> (span, aligned, counters)"""
        a_ptr, c_ptr = self.engine.links_partials_device()
        C_ = self.engine._n_contigs
        if not a_ptr or c_ptr != a_ptr + 8 * (C_ + 1):
            return None
        span = _wrap_device_i64(self.torch, a_ptr, C_ + 1 + abi.N_COUNTERS, self.device)
        return span, span[:C_], span[C_ + 1:]

    def counts_tensor(self, values):
        return self.torch.tensor(values, dtype=self.torch.int64, device=self.device)

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

    def fetch(self, sizes, view=False):
        return self.engine.fetch_view(sizes) if view else self.engine.fetch(sizes)


class _CudaArray(object):
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def _wrap_device_i64(torch, ptr, n, device):
    if n == 0 or not ptr:
        return torch.zeros(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_CudaArray(ptr, n), device=device)


class DistributedGraphBuild(object):
    def __init__(self, backend, rank, world, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.b = backend
        self.rank, self.world, self.group = rank, world, group
        self.last = None
        # BESST_DIST_EXCHANGE=tuples forces the tuple-level exchange (A/B, tests); default: runs
        import os
        mode = os.environ.get("BESST_DIST_EXCHANGE", "peer")   # peer | runs | tuples
        self.exchange_runs = mode != "tuples"
        self.exchange_peer = mode == "peer" and hasattr(backend, "symm_buffer") and world > 1
        # BESST_DIST_TIMING=1: device-synchronised wall time per phase of the last step in self.phase_ms (diagnostics)
        self.timing = os.environ.get("BESST_DIST_TIMING", "0") == "1"
        self.phase_ms = {}
        self._t_last = None

    def _mark(self, name):
        if not self.timing:
            return
        import time
        sync = getattr(self.b, "synchronize", None)
        if sync:
            sync()
        t = time.perf_counter()
        if self._t_last is not None and name:
            self.phase_ms[name] = self.phase_ms.get(name, 0.0) + 1e3 * (t - self._t_last)
        self._t_last = t

    # -- step 1: the last CreateEdge call of all preceding ranks --------------------------------
    def halo(self, params, rec):
        has, o1, o2 = self.b.tail_last_call(params, rec)
        mine = self.b.counts_tensor([has, o1, o2])
        gathered = [self.b.counts_tensor([0, 0, 0]) for _ in range(self.world)]
        self.dist.all_gather(gathered, mine, group=self.group)
        halo = (params.halo_prev_obs1, params.halo_prev_obs2)
        for r in range(self.rank):
            h, a, b = (int(x) for x in gathered[r].tolist())
            if h:
                halo = (a, b)
        return halo

    def step(self, params, rec):
        """One distributed graph build.  Leaves this rank's share of the CSR in its HBM;
        returns the local sizes.  `self.last` keeps what fetch_local needs."""
        self.phase_ms = {}
        self._t_last = None
        if hasattr(self.b, "bind_stream"):
            self.b.bind_stream()
        self._mark(None)
        if self.exchange_runs and hasattr(self.b, "group"):
            sizes = self._step_runs(params, rec)
            if sizes is not None:
                return sizes
        return self._step_tuples(params, rec)

    # -- exchange by tuples (fallback: a rank's link stream has no local order) ---------------------------
    def _step_tuples(self, params, rec):
        dist, world = self.dist, self.world
        halo = self.halo(params, rec) if world > 1 else (params.halo_prev_obs1, params.halo_prev_obs2)
        self._mark("halo")
        p = _copy_params(params, halo)
        n_local = self.b.extract(p, rec)
        self._mark("extract")
        send_t, send_o, send_f, tc, fc = self.b.partition(world)
        # bucket sizes: one small all_to_all, then the payload all_to_all
        counts_out = self.b.counts_tensor(np.stack([tc, fc], axis=1).reshape(-1).tolist())
        counts_in = self.b.counts_tensor([0] * (2 * world))
        dist.all_to_all_single(counts_in, counts_out, group=self.group)
        ci = np.asarray(counts_in.tolist(), dtype=np.int64).reshape(world, 2)
        rt, rf = ci[:, 0], ci[:, 1]
        recv_t, recv_o, recv_f = self.b.recv_buffers(int(rt.sum()), int(rf.sum()))
        dist.all_to_all_single(recv_t, send_t, output_split_sizes=rt.tolist(), input_split_sizes=tc.tolist(), group=self.group)
        dist.all_to_all_single(recv_o, send_o, output_split_sizes=rt.tolist(), input_split_sizes=tc.tolist(), group=self.group)
        dist.all_to_all_single(recv_f, send_f, output_split_sizes=rf.tolist(), input_split_sizes=fc.tolist(), group=self.group)
        sizes = self.b.to_graph(p, recv_t, recv_f)
        # exact, order-free partial sums: coverage + counters; the halo slots
        # (last call) are taken from the last rank that made one
        aligned, counters = self.b.partial_tensors()
        n_all = [self.b.counts_tensor([0]) for _ in range(world)]
        dist.all_gather(n_all, self.b.counts_tensor([n_local]), group=self.group)
        ends = counters[[abi.CNT_CALLS, abi.CNT_LAST_OBS1, abi.CNT_LAST_OBS2, abi.CNT_FIRST_OBS1, abi.CNT_FIRST_OBS2]].clone()
        counters[abi.CNT_LAST_OBS1:abi.CNT_FIRST_OBS2 + 1] = 0
        dist.all_reduce(aligned, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=self.group)
        all_ends = [ends.clone() for _ in range(world)]
        dist.all_gather(all_ends, ends, group=self.group)
        E = [[int(x) for x in t.tolist()] for t in all_ends]
        with_calls = [e for e in E if e[0] > 0]
        # every rank starts from its halo, so the last rank's pair already folds in all earlier ranks
        self.last = dict(sizes=sizes, recv_o=recv_o, recv_splits=rt, n_tuples_by_rank=[int(t.item()) for t in n_all],
                         aligned=aligned, counters=counters, last_call=(E[world - 1][1], E[world - 1][2]),
                         first_call=(with_calls[0][3], with_calls[0][4]) if with_calls else (0, 0), halo=halo)
        return sizes

    # -- exchange by runs (default) -------------------------------------------------------------------------
    def _step_runs(self, params, rec):
        """Whole runs (block-grouped links of one edge) are routed by edge hash: 8 bytes per link + 24 per
        run cross NVLink and the receiver starts at the run merge.

        No halo pass: every rank extracts with a PROVISIONAL halo that matches nothing; the one
        all_gather that carries the segment sizes also carries each slice's first and last CreateEdge
        call, so every rank can tell afterwards whether some slice's first call duplicates the call
        before it (CreateGraph.py:835-838).  Only then (rare) the affected ranks extract again with
        their true halo and the sizes are gathered once more.

        Transport: `peer` -- k_runs_pack stores straight into the destination GPU's symmetric-memory
        buffers, two cross-GPU barriers; else three NCCL all_to_all_single calls.
        None: some rank's stream has no local order (tuple-level path)."""
        dist, world, rank = self.dist, self.world, self.rank
        init_halo = (int(params.halo_prev_obs1), int(params.halo_prev_obs2))
        p = _copy_params(params, (NO_MATCH, NO_MATCH))
        need_extract = True
        self.redo_count = 0
        for attempt in range(2):
            if need_extract:
                n_local = self.b.extract(p, rec)
                self._mark("extract")
                if hasattr(self.b, "prepare_exchange"):   # one engine call, one host read
                    calls, last, first, n_runs, lc, rc, send_f, fc = self.b.prepare_exchange(world, n_local)
                    ok = n_runs is not None
                    self._mark("group+route+fishy")
                else:
                    calls, last, first = self.b.call_summary()
                    n_runs = self.b.group()
                    self._mark("group")
                    ok = n_runs is not None
                    lc, rc = self.b.route(world) if ok else (np.zeros(world, np.int64), np.zeros(world, np.int64))
                    send_f, fc = self.b.partition_fishy(world)
                    self._mark("route+fishy")
            mine = self.b.counts_tensor([1 if ok else 0, n_local, 1 if calls > 0 else 0, last[0], last[1], first[0], first[1]]
                                        + lc.tolist() + rc.tolist() + fc.tolist())
            # ONE collective and ONE device->host read for the whole W x (7 + 3W) matrix
            gathered = mine.new_zeros((world, 7 + 3 * world))
            if mine.is_cuda:
                dist.all_gather_into_tensor(gathered, mine, group=self.group)
            else:   # gloo (CPU tests) has no flat all-gather
                dist.all_gather(list(gathered.unbind(0)), mine, group=self.group)
            M = gathered.cpu().numpy().astype(np.int64)
            self._mark("meta")
            if int(M[:, 0].min()) == 0:
                return None
            halos, cur = [], init_halo
            for r in range(world):
                halos.append(cur)
                if M[r, 2]:
                    cur = (int(M[r, 3]), int(M[r, 4]))
            global_last = cur
            firsts = [(int(M[r, 5]), int(M[r, 6])) for r in range(world) if M[r, 2]]
            global_first = firsts[0] if firsts else (0, 0)
            redo = [bool(M[r, 2]) and (int(M[r, 5]), int(M[r, 6])) == halos[r] for r in range(world)]
            if attempt == 0 and any(redo):   # a slice starts with a duplicate of the call before it
                self.redo_count = sum(redo)
                need_extract = redo[rank]
                if need_extract:
                    p = _copy_params(params, halos[rank])
                continue
            break
        # coverage + counters: one allocation, reduced with ONE all-reduce at the end of the step (launching it early on
        # NCCL's own stream was slower: its CTAs spin on the slowest rank's extraction and take SMs from the build)
        span = self.b.partial_span() if hasattr(self.b, "partial_span") else None
        if span is not None:
            span, aligned, counters = span
        else:
            aligned, counters = self.b.partial_tensors()
        n_by_rank = M[:, 1]
        LC, RC, FC = M[:, 7:7 + world], M[:, 7 + world:7 + 2 * world], M[:, 7 + 2 * world:7 + 3 * world]
        rl, rr, rf = LC[:, rank], RC[:, rank], FC[:, rank]
        if int(n_by_rank.sum()) >= 2 ** 32 or int(LC.sum(axis=0).max()) >= 2 ** 30:
            raise ValueError("run-level exchange: more than 2^32 links in the library or 2^30 on one rank")
        w = self.b.obs_words(p) if hasattr(self.b, "obs_words") else 2
        if self.exchange_peer:
            try:
                recv_obs, recv_desc, recv_f = self._transport_peer(LC, RC, FC, send_f, fc, w)
            except _SymmUnavailable as exc:   # no peer mapping on this box: every rank raises alike (collective setup)
                import warnings
                warnings.warn("symmetric memory unavailable (%s): NCCL all-to-all instead" % exc)
                self.exchange_peer = False
        if not self.exchange_peer:
            send_obs, send_desc = self.b.pack(world, int(lc.sum()), int(rc.sum()), w)
            recv_obs, recv_desc, recv_f = self.b.recv_run_buffers(int(rl.sum()), int(rr.sum()), int(rf.sum()), w)
            self._mark("pack")
            dist.all_to_all_single(recv_obs, send_obs, output_split_sizes=rl.tolist(), input_split_sizes=lc.tolist(), group=self.group)
            dist.all_to_all_single(recv_desc, send_desc, output_split_sizes=rr.tolist(), input_split_sizes=rc.tolist(), group=self.group)
            dist.all_to_all_single(recv_f, send_f, output_split_sizes=rf.tolist(), input_split_sizes=fc.tolist(), group=self.group)
            self._mark("all_to_all")
        # what this rank puts on the wire: observations (4 or 8 B per link) + 24 B per run + 8 B per fishy key to OTHER ranks
        others = np.arange(world) != rank
        self.exchange_bytes_out = int(4 * w * LC[rank][others].sum() + 24 * RC[rank][others].sum() + 8 * FC[rank][others].sum())
        max_blocks = int(((n_by_rank + abi.RUN_BLOCK - 1) // abi.RUN_BLOCK).max())
        block_bits = max(1, int(max(max_blocks - 1, 1)).bit_length())
        first_base = np.concatenate([[0], np.cumsum(n_by_rank)[:-1]])
        sizes = self.b.runs_to_graph(p, recv_obs, recv_desc, world, block_bits, rr, rl, first_base, recv_f)
        self._mark("runs_to_graph")
        counters[abi.CNT_LAST_OBS1:abi.CNT_FIRST_OBS2 + 1] = 0   # per-rank slots (last / first call): not sums
        if span is not None:
            dist.all_reduce(span, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(aligned, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=self.group)
        self._mark("all_reduce")
        self.last = dict(sizes=sizes, global_first=True, n_tuples_by_rank=n_by_rank.tolist(), aligned=aligned, counters=counters,
                         last_call=global_last, first_call=global_first, halo=halos[rank])
        return sizes

    def _transport_peer(self, LC, RC, FC, send_f, fc, w):
        """The exchange fused into the pack kernel: every rank owns peer-mapped receive buffers (symmetric
        memory); the W x W count matrix gives every sender its offset inside every receiver."""
        world, rank = self.world, self.rank
        torch = self.b.torch
        try:
            obs_t, obs_h = self.b.symm_buffer("obs", w * int(LC.sum(axis=0).max()) + 2, torch.int32, self.group)
            desc_t, desc_h = self.b.symm_buffer("desc", 6 * int(RC.sum(axis=0).max()) + 6, torch.int32, self.group)
            f_t, f_h = self.b.symm_buffer("fishy", int(FC.sum(axis=0).max()) + 1, torch.int64, self.group)
        except Exception as exc:
            raise _SymmUnavailable(repr(exc))
        link_off = LC[:rank].sum(axis=0)   # my segment's start inside every destination's buffers
        run_off = RC[:rank].sum(axis=0)
        f_off = FC[:rank].sum(axis=0)
        obs_ptrs = [int(obs_h.buffer_ptrs[d]) + 4 * w * int(link_off[d]) for d in range(world)]
        desc_ptrs = [int(desc_h.buffer_ptrs[d]) + 24 * int(run_off[d]) for d in range(world)]
        obs_h.barrier(channel=0)   # every rank is done reading the buffers of the previous step
        self.b.pack_peer(world, obs_ptrs, desc_ptrs)
        fstart = np.concatenate([[0], np.cumsum(fc)])
        for d in range(world):
            if fc[d]:
                f_h.get_buffer(d, (int(fc[d]),), torch.int64, int(f_off[d])).copy_(send_f[int(fstart[d]):int(fstart[d + 1])])
        obs_h.barrier(channel=0)   # all stores have landed
        self._mark("pack+exchange")
        rl, rr, rf = LC[:, rank], RC[:, rank], FC[:, rank]
        return obs_t[:w * int(rl.sum())].view(-1, w), desc_t[:6 * int(rr.sum())].view(-1, 6), f_t[:int(rf.sum())]

    # -- results ------------------------------------------------------------------------------------
    def fetch_local(self, view=False):
        """This rank's edges (GraphResult) with first_idx rewritten to the GLOBAL ordinal of the
        edge's first link (source-rank prefix + ordinal in the source's stream), and the reduced
        aligned_len / counters.  view=True: zero-copy views of the engine's pinned result buffers
        (valid until the next view)."""
        L = self.last
        res = self.b.fetch(L["sizes"], view) if view else self.b.fetch(L["sizes"])
        if res.n_edges and not L.get("global_first"):
            starts = np.concatenate([[0], np.cumsum(L["recv_splits"])])
            src = np.searchsorted(starts, res.first_idx, side="right") - 1
            prefix = np.concatenate([[0], np.cumsum(L["n_tuples_by_rank"])])
            ordinals = np.asarray(L["recv_o"].cpu().numpy()).view(np.uint32).astype(np.int64)
            res.first_idx = prefix[src] + ordinals[res.first_idx]
        res.aligned_len = np.asarray(L["aligned"].cpu().numpy(), dtype=np.int64)
        counters = np.asarray(L["counters"].cpu().numpy(), dtype=np.int64).copy()
        # the globally last CreateEdge call (the next library pass would start from it)
        counters[abi.CNT_LAST_OBS1], counters[abi.CNT_LAST_OBS2] = L["last_call"]
        counters[abi.CNT_FIRST_OBS1], counters[abi.CNT_FIRST_OBS2] = L["first_call"]
        res.counters = counters
        return res

    def fetch_global(self):
        """All ranks' edges merged into one GraphResult (every rank gets it).  Host-side, outside
        the timed path: feeds the order-dependent post-filters of CreateGraph.PE."""
        local = self.fetch_local()
        parts = [None] * self.world
        self.dist.all_gather_object(parts, local, group=self.group)
        return merge_graph_results(parts)


def merge_graph_results(parts):
    """Disjoint edge sets (each sorted by (edge_u, edge_v)) -> one CSR sorted the same way."""
    parts = [p for p in parts if p is not None]
    first = parts[0]
    key = np.concatenate([(p.edge_u.astype(np.int64) << 32) | p.edge_v for p in parts])
    order = np.argsort(key, kind="stable")
    nr = np.concatenate([p.nr_links for p in parts])[order]
    row_ptr = np.concatenate([[0], np.cumsum(nr, dtype=np.int64)])
    # payload: gather each edge's segment
    seg_start = np.concatenate([p.row_ptr[:-1] + off for p, off in zip(parts, np.cumsum([0] + [p.n_links for p in parts[:-1]]))])[order]
    obs_u_all = np.concatenate([p.obs_u for p in parts])
    obs_v_all = np.concatenate([p.obs_v for p in parts])
    idx = np.repeat(seg_start - row_ptr[:-1], nr) + np.arange(int(row_ptr[-1]), dtype=np.int64)
    fields = {}
    for f in ("edge_u", "edge_v", "nr_links", "obs_sum", "obs_sq", "first_idx", "gap", "score", "ks", "sd_obs", "sd_model",
              "fishy", "flags"):
        fields[f] = np.concatenate([getattr(p, f) for p in parts])[order]
    return abi.GraphResult(row_ptr=row_ptr, obs_u=obs_u_all[idx], obs_v=obs_v_all[idx], aligned_len=first.aligned_len,
                           counters=first.counters, **fields)


_host_group = []


def host_group():
    """A gloo group over all ranks for the small Python-object collectives of the ingest (offsets, library metrics): created
    once, collectively, at first use; None (the default group) when the job itself runs on gloo."""
    import torch.distributed as dist
    if not _host_group:
        _host_group.append(None if dist.get_backend() == "gloo" else dist.new_group(backend="gloo"))
    return _host_group[0]


def rank0_value(fn, rank, group=None):
    """fn() evaluated on rank 0, its (picklable) result on every rank"""
    import torch.distributed as dist
    box = [fn() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def ingest_bam_distributed(engine, path, rank, world, group=None, **ingest_kw):
    """Multi-GPU BAM ingest (SURVEY.md 8e + 8f rank 1): rank r inflates and decodes the r-th part of the file on its own
    GPU (besst_bam_ingest_part: the BGZF blocks starting in its byte range, the records starting in those blocks), so a
    library arrives in HBM already range-partitioned in BAM order -- the layout DistributedGraphBuild.step takes -- at
    `world` times the single-GPU inflate rate and without the file's records ever being in one place.

    A part behind the header finds its first record by a plausibility test nothing inside the part can verify.  The ranks
    therefore exchange (first record, landing of the last record) as BGZF virtual offsets -- ONE small all_gather -- and
    every part whose first record is not where the chain of the preceding parts landed is read again from the right
    offset (rare: a false positive needs a byte pattern that passes the record checks inside another record; repeated at
    most once per part).  -> (DeviceRecordBatch of this rank, {"counts": records per rank, "record_base": ordinal of this
    rank's first record in the file, "repeats": parts read again})"""
    import torch.distributed as dist
    dev = engine.ingest_bam(path, part=(rank, world), **ingest_kw)
    forced = {}
    repeats = 0
    for _ in range(world + 2):
        mine = (int(dev.first_voffset), int(dev.landing_voffset), len(dev))
        rows = [None] * world
        if world > 1:
            dist.all_gather_object(rows, mine, group=group)
        else:
            rows = [mine]
        prev, redo = -1, None
        for r, (first, landing, _n) in enumerate(rows):
            if r > 0 and prev >= 0 and first != prev and forced.get(r) != prev and redo is None:
                redo = (r, prev)
            if landing >= 0:
                prev = landing
        if redo is None:
            break
        r, start = redo
        forced[r] = start   # every rank keeps the same book: the loop ends after at most `world` repeats
        repeats += 1
        if r == rank:
            dev = engine.ingest_bam(path, part=(rank, world), start_voffset=start, **ingest_kw)
    else:
        raise RuntimeError("ingest_bam_distributed: the parts' record chains do not meet (%r)" % (rows,))
    counts = [int(n) for _, _, n in rows]
    return dev, {"counts": counts, "record_base": int(sum(counts[:rank])), "repeats": repeats}
