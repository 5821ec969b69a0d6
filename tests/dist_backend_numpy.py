"""TEST INFRASTRUCTURE: a CPU stand-in for besst_b200.dist.CudaBackend so that the host logic of
DistributedGraphBuild (halo, bucket exchange, ordinals, reductions, merge) runs under gloo with
world_size > 1 on a machine without GPUs.  Extraction is the C oracle; bucketing and the
tuples -> CSR step are plain numpy (integers only: scores are covered by the GPU tests)."""
import numpy as np
import torch

import oracle_lib
from besst_b200 import abi
from besst_b200.dist import NO_MATCH, _copy_params


def edge_dest_numpy(u, v, world):
    with np.errstate(over="ignore"):
        x = (u.astype(np.uint64) << np.uint64(32)) | v.astype(np.uint64)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return (x % np.uint64(world)).astype(np.int64)


class NumpyBackend(object):
    def __init__(self, table):
        self.table = table

    def records(self, batch):
        return batch

    def _run(self, params, batch):
        res, tuples, fishy, _ = oracle_lib.graph_build(self.table.rows, self.table.n_scaffolds, params, batch)
        self.tuples = tuples
        self.fishy = np.repeat(np.array(list(fishy.keys()), dtype=np.uint64), list(fishy.values())) if fishy else np.zeros(0, np.uint64)
        self.aligned = torch.from_numpy(res.aligned_len.copy())
        self.counters = torch.from_numpy(res.counters.copy())
        return res

    def tail_last_call(self, params, batch):
        res = self._run(_copy_params(params, (NO_MATCH, NO_MATCH)), batch)
        c = res.counters
        return (1, int(c[abi.CNT_LAST_OBS1]), int(c[abi.CNT_LAST_OBS2])) if c[abi.CNT_CALLS] > 0 else (0, 0, 0)

    def extract(self, params, batch):
        self._run(params, batch)
        return int(self.tuples.shape[0])

    def partition(self, world):
        dest = edge_dest_numpy(self.tuples["u"], self.tuples["v"], world)
        order = np.argsort(dest, kind="stable")
        t = self.tuples[order]
        send_t = torch.from_numpy(np.ascontiguousarray(t).view(np.int32).reshape(-1, 4).copy())
        send_o = torch.from_numpy(order.astype(np.int32))
        fd = edge_dest_numpy((self.fishy >> np.uint64(32)).astype(np.uint32), (self.fishy & np.uint64(0xffffffff)).astype(np.uint32), world)
        forder = np.argsort(fd, kind="stable")
        send_f = torch.from_numpy(self.fishy[forder].view(np.int64).copy())
        return send_t, send_o, send_f, np.bincount(dest, minlength=world).astype(np.int64), np.bincount(fd, minlength=world).astype(np.int64)

    def recv_buffers(self, n_tuples, n_fishy):
        return (torch.zeros((n_tuples, 4), dtype=torch.int32), torch.zeros(n_tuples, dtype=torch.int32),
                torch.zeros(n_fishy, dtype=torch.int64))

    def to_graph(self, params, recv_t, recv_f):
        """tuples -> CSR (integers only), the numpy statement of CreateEdge's upsert (CreateGraph.py:842-862)."""
        t = recv_t.numpy().view(abi.LINK_TUPLE_DTYPE).reshape(-1)
        key = (t["u"].astype(np.int64) << 32) | t["v"]
        order = np.argsort(key, kind="stable")
        ks = key[order]
        heads = np.nonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))[0] if len(ks) else np.zeros(0, np.int64)
        row_ptr = np.concatenate([heads, [len(ks)]]).astype(np.int64)
        E = len(heads)
        ou, ov = t["obs_u"][order], t["obs_v"][order]
        tot = ou.astype(np.int64) + ov
        fk = recv_f.numpy().view(np.uint64)
        fkey = ((fk >> np.uint64(32)).astype(np.int64) << 32) | (fk & np.uint64(0xffffffff)).astype(np.int64)
        ukeys = ks[heads] if E else np.zeros(0, np.int64)
        fs = np.sort(fkey)
        n_large2 = 2 * self.table.n_large_scaffolds
        eu, ev = (ukeys >> 32).astype(np.uint32), (ukeys & 0xffffffff).astype(np.uint32)
        self.result = abi.GraphResult(
            edge_u=eu, edge_v=ev, nr_links=np.diff(row_ptr).astype(np.int32),
            obs_sum=np.add.reduceat(tot, heads) if E else np.zeros(0, np.int64),
            obs_sq=np.add.reduceat(tot * tot, heads) if E else np.zeros(0, np.int64),
            first_idx=order[heads].astype(np.int64) if E else np.zeros(0, np.int64), row_ptr=row_ptr,
            gap=np.zeros(E, np.int32), score=np.full(E, np.nan), ks=np.full(E, np.nan), sd_obs=np.full(E, np.nan),
            sd_model=np.full(E, np.nan),
            fishy=(np.searchsorted(fs, ukeys, side="right") - np.searchsorted(fs, ukeys, side="left")).astype(np.int32),
            flags=((eu < n_large2) & (ev < n_large2)).astype(np.uint8) * abi.EDGE_LL,
            obs_u=ou, obs_v=ov, aligned_len=np.zeros(0, np.int64), counters=np.zeros(abi.N_COUNTERS, np.int64))
        return "sizes"

    # ---- run-level exchange: numpy statement of besst_links_group / runs_route / runs_pack / runs_to_graph ----
    def call_summary(self):
        c = self.counters.numpy()
        return (int(c[abi.CNT_CALLS]), (int(c[abi.CNT_LAST_OBS1]), int(c[abi.CNT_LAST_OBS2])),
                (int(c[abi.CNT_FIRST_OBS1]), int(c[abi.CNT_FIRST_OBS2])))

    def group(self):
        t = self.tuples
        n = t.shape[0]
        key = (t["u"].astype(np.int64) << 32) | t["v"]
        obs = np.stack([t["obs_u"], t["obs_v"]], axis=1).astype(np.int32)
        grouped = np.zeros_like(obs)
        runs = []   # (u, v, count, first, start, block)
        for blk, base in enumerate(range(0, n, abi.RUN_BLOCK)):
            k = key[base:base + abi.RUN_BLOCK]
            order = np.argsort(k, kind="stable")
            grouped[base:base + len(k)] = obs[base:base + len(k)][order]
            ks = k[order]
            heads = np.nonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))[0]
            ends = np.concatenate([heads[1:], [len(ks)]])
            for h, e in zip(heads, ends):
                runs.append((int(ks[h] >> 32), int(ks[h] & 0xffffffff), int(e - h), base + int(order[h]), base + int(h), blk))
        self.grouped = grouped
        self.runs = np.array(runs, dtype=np.int64).reshape(-1, 6)
        return self.runs.shape[0]

    def route(self, world):
        r = self.runs
        self.run_dest = edge_dest_numpy(r[:, 0].astype(np.uint32), r[:, 1].astype(np.uint32), world) if len(r) else np.zeros(0, np.int64)
        lc = np.bincount(self.run_dest, weights=r[:, 2], minlength=world).astype(np.int64) if len(r) else np.zeros(world, np.int64)
        return lc, np.bincount(self.run_dest, minlength=world).astype(np.int64)

    def partition_fishy(self, world):
        fd = edge_dest_numpy((self.fishy >> np.uint64(32)).astype(np.uint32), (self.fishy & np.uint64(0xffffffff)).astype(np.uint32), world)
        forder = np.argsort(fd, kind="stable")
        return torch.from_numpy(self.fishy[forder].view(np.int64).copy()), np.bincount(fd, minlength=world).astype(np.int64)

    def pack(self, world, n_links, n_runs, w=2):
        # any placement inside a destination segment is legal (the device uses atomics): use reverse run order here
        order = np.lexsort((-np.arange(len(self.run_dest)), self.run_dest))
        send_obs = np.zeros((n_links, 2), np.int32)
        send_desc = np.zeros((n_runs, 6), np.int32)
        pos, seg_start, prev = 0, 0, -1
        for slot, r in enumerate(order):
            d = self.run_dest[r]
            if d != prev:
                seg_start, prev = pos, d
            u, v, cnt, first, start, blk = self.runs[r]
            send_obs[pos:pos + cnt] = self.grouped[start:start + cnt]
            send_desc[slot] = np.array([u, v, cnt, first, pos - seg_start, blk], dtype=np.int64).astype(np.uint32).view(np.int32)
            pos += cnt
        return torch.from_numpy(send_obs), torch.from_numpy(send_desc)

    def recv_run_buffers(self, n_links, n_runs, n_fishy, w=2):
        return (torch.zeros((n_links, 2), dtype=torch.int32), torch.zeros((n_runs, 6), dtype=torch.int32),
                torch.zeros(n_fishy, dtype=torch.int64))

    def runs_to_graph(self, params, recv_obs, recv_desc, world, block_bits, src_runs, src_links, src_first, recv_f):
        obs = recv_obs.numpy()
        d = recv_desc.numpy().view(np.uint32).astype(np.int64)
        R = d.shape[0]
        src = np.repeat(np.arange(world), np.asarray(src_runs, dtype=np.int64))
        link_base = np.concatenate([[0], np.cumsum(src_links)[:-1]])
        assert R == 0 or int(d[:, 5].max()) < (1 << block_bits)
        start = link_base[src] + d[:, 4] if R else np.zeros(0, np.int64)
        order = np.lexsort((d[:, 5], src, d[:, 1], d[:, 0])) if R else np.zeros(0, np.int64)
        ekey = (d[:, 0] << 32) | d[:, 1]
        ks = ekey[order]
        heads = np.nonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))[0] if R else np.zeros(0, np.int64)
        E = len(heads)
        idx = np.concatenate([np.arange(start[r], start[r] + d[r, 2]) for r in order]) if R else np.zeros(0, np.int64)
        ou, ov = obs[idx, 0], obs[idx, 1]
        run_off = np.concatenate([[0], np.cumsum(d[order, 2])]) if R else np.zeros(1, np.int64)
        row_ptr = np.concatenate([run_off[heads], [run_off[-1]]]).astype(np.int64)
        tot = ou.astype(np.int64) + ov
        lheads = row_ptr[:-1]
        fk = recv_f.numpy().view(np.uint64)
        fkey = ((fk >> np.uint64(32)).astype(np.int64) << 32) | (fk & np.uint64(0xffffffff)).astype(np.int64)
        ukeys = ks[heads] if E else np.zeros(0, np.int64)
        fs = np.sort(fkey)
        n_large2 = 2 * self.table.n_large_scaffolds
        eu, ev = (ukeys >> 32).astype(np.uint32), (ukeys & 0xffffffff).astype(np.uint32)
        first = (np.asarray(src_first, dtype=np.int64)[src[order[heads]]] + d[order[heads], 3]) if E else np.zeros(0, np.int64)
        self.result = abi.GraphResult(
            edge_u=eu, edge_v=ev, nr_links=np.diff(row_ptr).astype(np.int32),
            obs_sum=np.add.reduceat(tot, lheads) if E else np.zeros(0, np.int64),
            obs_sq=np.add.reduceat(tot * tot, lheads) if E else np.zeros(0, np.int64),
            first_idx=first.astype(np.int64), row_ptr=row_ptr,
            gap=np.zeros(E, np.int32), score=np.full(E, np.nan), ks=np.full(E, np.nan), sd_obs=np.full(E, np.nan),
            sd_model=np.full(E, np.nan),
            fishy=(np.searchsorted(fs, ukeys, side="right") - np.searchsorted(fs, ukeys, side="left")).astype(np.int32),
            flags=((eu < n_large2) & (ev < n_large2)).astype(np.uint8) * abi.EDGE_LL,
            obs_u=ou.astype(np.int32), obs_v=ov.astype(np.int32), aligned_len=np.zeros(0, np.int64), counters=np.zeros(abi.N_COUNTERS, np.int64))
        return "sizes"

    def partial_tensors(self):
        return self.aligned, self.counters

    def counts_tensor(self, values):
        return torch.tensor(values, dtype=torch.int64)

    def fetch(self, sizes, view=False):
        return self.result
