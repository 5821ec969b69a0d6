"""Probe torch symmetric memory (peer-mapped buffers over NVLink) on this box: needed by the fused
pack + exchange kernel.  torchrun --nproc-per-node 2 scripts/symm_probe.py"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 128 << 20   # int32 elements: 512 MB
try:
    t = sm.empty(n, dtype=torch.int32, device=dev)
    hdl = sm.rendezvous(t, group=dist.group.WORLD)
    ptrs = list(hdl.buffer_ptrs)
    print(rank, "rendezvous ok, ptrs", [hex(p) for p in ptrs], flush=True)
    peer = (rank + 1) % world
    src = torch.full((n,), rank + 1, dtype=torch.int32, device=dev)
    pt = hdl.get_buffer(peer, (n,), torch.int32)
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    for _ in range(2):
        t0 = time.perf_counter(); pt.copy_(src); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    ok = bool((t == ((rank - 1) % world) + 1).all().item())
    print(rank, "peer write %.1f GB/s, data correct: %s" % (4 * n / dt / 1e9, ok), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "SYMM_MEM_FAILED", repr(e), flush=True)
dist.barrier(); dist.destroy_process_group()
