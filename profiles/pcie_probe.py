"""Pinned host <-> device copy rates, one process per GPU, all ranks copying at the same time.
  python profiles/pcie_probe.py                                   (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/pcie_probe.py
Per rank: H2D alone, D2H alone, both directions at once (GB/s per direction), 1 GiB buffers, CUDA events.  Rank 0 prints one
line per rank, the aggregate, and the host facts that matter (CPU affinity, NUMA nodes, PCIe topology from nvidia-smi)."""
import os, subprocess, time
import torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 28                                       # 1 GiB of float32
h1 = torch.empty(n, dtype=torch.float32).pin_memory(); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d1 = torch.empty(n, dtype=torch.float32, device=dev); d2 = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def t(fn, reps=3):
    fn(); barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return dt
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
gb = n * 4 / 1e9
mine = torch.tensor([gb / t(h2d), gb / t(d2h), gb / t(both)], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(mine) for _ in range(world)]
if world > 1: dist.all_gather(allr, mine)
else: allr = [mine]
if rank == 0:
    for r, v in enumerate(allr):
        print("rank %d: H2D %.1f GB/s  D2H %.1f GB/s  both directions at once: %.1f GB/s each" % (r, v[0], v[1], v[2]))
    tot = torch.stack(allr).sum(0)
    print("aggregate over %d ranks: H2D %.1f  D2H %.1f  duplex %.1f GB/s each direction (%.1f both)" % (world, tot[0], tot[1], tot[2], 2 * tot[2]))
    print("cpu affinity of rank 0:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "cpus;",
          "numa nodes:", sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) if os.path.isdir("/sys/devices/system/node") else "n/a")
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout)
    except Exception as e:  # noqa: BLE001
        print("nvidia-smi topo failed:", e)
if world > 1: dist.destroy_process_group()
