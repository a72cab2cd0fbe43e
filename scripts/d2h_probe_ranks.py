"""D2H-only probe (developer tool): every rank copies a 1 GiB device buffer to ITS OWN pinned host buffer, all ranks at
once, no kernels -- the host-side ceiling the end-to-end path runs into on many GPUs.  Run under torchrun with
N = 1, 2, 4, 8; rank 0 prints one line.  usage: torchrun ... scripts/d2h_probe_ranks.py [pin_cores 0|1]"""
import os, sys, time
import torch
import torch.distributed as dist

ws = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
pin = len(sys.argv) > 1 and sys.argv[1] == "1"
torch.cuda.set_device(local)
if ws > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if pin:
    cores = sorted(os.sched_getaffinity(0)); per = max(len(cores) // ws, 1)
    os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
n = 1 << 27                                    # 1 GiB of doubles
dev = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
host = torch.empty(n, dtype=torch.float64).pin_memory()
for _ in range(2):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
if ws > 1:
    dist.barrier(); torch.cuda.synchronize()
reps = 8
t0 = time.perf_counter()
for _ in range(reps):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if ws > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    gbs = reps * n * 8 / 1e9 / float(t.item())
    print("D2H only, %d rank(s), cores pinned per rank: %s -> %.1f GB/s per GPU, %.1f GB/s aggregate (1 GiB x %d per rank, "
          "pinned host memory, max over ranks; host: %d cores visible)" % (ws, pin, gbs, gbs * ws, reps, len(os.sched_getaffinity(0)) if not pin else os.cpu_count()))
if ws > 1:
    dist.barrier(); dist.destroy_process_group()
