"""Diagnostic (not part of the product): concurrent D2H bandwidth per rank under torchrun, pinned-allocator cost."""
import os, time, json, subprocess
import torch, torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
nbytes = 16 * 256 * 256 * 3 * 4
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
res = {"rank": rank, "cpus": len(os.sched_getaffinity(0)), "threads": torch.get_num_threads()}
def bw(fn, n=40):
    fn(); torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return nbytes * n / dt / 1e9
res["d2h_all_GBs"] = bw(lambda: host.copy_(dev, non_blocking=True))
res["h2d_all_GBs"] = bw(lambda: dev.copy_(host, non_blocking=True))
# fresh pinned tensor per copy (what Imitator._to_host does), blocks kept alive like the outputs list keeps them
keep = []
def fresh():
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h.copy_(dev, non_blocking=True); keep.append(h)
res["d2h_fresh_first_GBs"] = bw(fresh, n=20)
keep.clear()
res["d2h_fresh_cached_GBs"] = bw(fresh, n=20)
keep.clear()
# solo: only one rank at a time
solo = {}
for r in range(world):
    barrier()
    if r == rank:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(40): host.copy_(dev, non_blocking=True)
        torch.cuda.synchronize(); solo = nbytes * 40 / (time.perf_counter() - t0) / 1e9
    barrier()
res["d2h_solo_GBs"] = solo
# host memcpy bandwidth of this process (numpy copy of the pinned buffer), all ranks at once
import numpy as np
a = host.numpy(); b = np.empty_like(a)
barrier(); t0 = time.perf_counter()
for _ in range(20): np.copyto(b, a)
res["host_memcpy_all_GBs"] = nbytes * 20 / (time.perf_counter() - t0) / 1e9
out = [None] * world
if world > 1:
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    for r in out: print(json.dumps(r))
    for cmd in ("nvidia-smi topo -m", "lscpu | head -25", "cat /sys/devices/system/node/online", "nproc"):
        try: print(subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout)
        except Exception as e: print(cmd, e)
if world > 1: dist.destroy_process_group()
