"""Dev script (torchrun, one rank per GPU): what the box's host <-> device copies sustain when every rank moves the e2e loop's
bytes at the same time.  pinned = cudaHostAlloc (torch pin_memory), registered = cudaHostRegister on a numpy array (what
pbf_host_register does for caller-owned buffers).  Prints per-rank and aggregate GB/s for H2D alone, D2H alone, both directions at once."""
import os, time, json
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
UP, DOWN = 832_000_000, 960_000_000
dev_u = torch.empty(UP, dtype=torch.uint8, device="cuda"); dev_d = torch.empty(DOWN, dtype=torch.uint8, device="cuda")
def bufs(kind):
    if kind == "pinned":
        return torch.empty(UP, dtype=torch.uint8, pin_memory=True), torch.empty(DOWN, dtype=torch.uint8, pin_memory=True)
    a = np.zeros(UP, dtype=np.uint8); b = np.zeros(DOWN, dtype=np.uint8)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    cudart = torch.cuda.cudart()
    assert int(cudart.cudaHostRegister(ta.data_ptr(), UP, 1)) == 0 and int(cudart.cudaHostRegister(tb.data_ptr(), DOWN, 1)) == 0
    return ta, tb
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
res = {}
s2 = torch.cuda.Stream()
for kind in ("pinned", "registered"):
    hu, hd = bufs(kind)
    for mode in ("h2d", "d2h", "both", "sequential"):
        for rep in range(3):
            barrier(); t0 = time.perf_counter()
            if mode in ("h2d", "both", "sequential"): dev_u.copy_(hu, non_blocking=True)
            if mode == "sequential": torch.cuda.synchronize()
            if mode == "d2h" or mode == "sequential": hd.copy_(dev_d, non_blocking=True)
            if mode == "both":
                with torch.cuda.stream(s2): hd.copy_(dev_d, non_blocking=True)
            barrier(); dt = time.perf_counter() - t0
        nbytes = {"h2d": UP, "d2h": DOWN, "both": UP + DOWN, "sequential": UP + DOWN}[mode]
        res[f"{kind}_{mode}"] = {"ms": dt * 1e3, "gbs_per_rank": nbytes / dt / 1e9, "gbs_aggregate": world * nbytes / dt / 1e9}
if rank == 0: print(json.dumps({"world": world, "bytes_up": UP, "bytes_down": DOWN, "results": res}, indent=1))
if world > 1: dist.destroy_process_group()
