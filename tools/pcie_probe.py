"""Host <-> device copy bandwidth per GPU and in aggregate, with and without NUMA-local pinning.

    python tools/pcie_probe.py                      # one GPU (cuda:0)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py

Every rank copies pinned 256 MiB buffers H2D, D2H and both at once (two streams) for ~1 s each, all ranks at the same
time (barriers around every phase); rank 0 prints one JSON line per phase with per-GPU and aggregate GB/s.  The whole
sequence runs twice: with the process left where the launcher put it, and after `numa.bind_to_device` (CPU affinity =
the GPU's local CPUs, so cudaHostAlloc's pages come from the GPU's NUMA node).  This is the measurement behind the
end-to-end scaling of HostWarpPipeline (0.195 efficiency at 8 GPUs; where is the ceiling?)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pwstablenet_b200 import numa

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MB = 256
n = MB << 20


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def phase(tag, placement):
    h_up = torch.empty(n, dtype=torch.uint8).pin_memory(); h_up.fill_(1)
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory(); h_dn.fill_(2)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev); d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}
    for kind in ("h2d", "d2h", "both"):
        reps = 4
        for attempt in range(2):   # first round calibrates the repetition count for ~1 s
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                if kind in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_up, non_blocking=True)
                if kind in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_dn.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            if attempt == 0:
                reps = max(4, int(reps * 1.0 / max(dt, 1e-3)))
        gbs = reps * n * (2 if kind == "both" else 1) / dt / 1e9
        t = torch.tensor([gbs], device=dev, dtype=torch.float64)
        if world > 1:
            all_t = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(all_t, t)
            vals = [float(x.item()) for x in all_t]
        else:
            vals = [gbs]
        res[kind] = vals
    pl = [placement]
    if world > 1:
        pl = [None] * world
        dist.all_gather_object(pl, placement)
    if rank == 0:
        for kind, vals in res.items():
            print(json.dumps({"probe": "pcie", "pinning": tag, "n_gpus": world, "kind": kind, "buffer_MiB": MB,
                              "per_gpu_GBs": [round(v, 2) for v in vals], "aggregate_GBs": round(sum(vals), 2),
                              "placement": pl if kind == "h2d" else None}), flush=True)
    del h_up, h_dn, d_a, d_b


loc = numa.device_locality(local)
phase("as launched", {"numa_node_of_gpu": loc["node"], "cpus_allowed": len(os.sched_getaffinity(0))})
placement = numa.bind_to_device(local)
phase("numa-local", placement)
if world > 1:
    dist.destroy_process_group()
