"""Under torchrun: FMG solve time of `L 8` per rank for L = 7..3 (the coarse part of `7 8` is the `6 8` problem, ...):
marginal cost of each level on N GPUs.  Device time on rank 0's stream between barriers, max over ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgmg_b200.api as api
rank, world = api.init_distributed()
L = api.lib()
import torch, torch.distributed as dist
for log2 in (7, 6, 5, 4, 3):
    H = api.Hierarchy(log2, 8, my_rank=rank, num_ranks=world)
    for _ in range(5):
        H.fmg_solve(0)
    L.hpgmg_b200_sync(); dist.barrier(); L.hpgmg_b200_sync()
    k0 = L.hpgmg_b200_kernel_launches()
    L.hpgmg_b200_bench_mark(0)
    steps = 20
    for _ in range(steps):
        r = H.fmg_solve(0)
    L.hpgmg_b200_bench_mark(1)
    L.hpgmg_b200_sync()
    t = torch.tensor([L.hpgmg_b200_bench_elapsed_ms(0, 1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k = torch.tensor([float(L.hpgmg_b200_kernel_launches() - k0) / steps], dtype=torch.float64, device="cuda")
    ks = [torch.zeros_like(k) for _ in range(world)]
    dist.all_gather(ks, k)
    if rank == 0:
        dims = [H.level(l).contents.dim.i for l in range(H.num_levels)]
        ranks = [H.level(l).contents.num_ranks for l in range(H.num_levels)]
        print(f"world={world} log2_box_dim={log2}: {float(t):.3f} ms/solve  launches/solve per rank {[int(x) for x in ks]}  dims {dims} ranks {ranks} norm {r[0]!r}", flush=True)
    H.close()
L.hpgmg_b200_comm_finalize()
dist.destroy_process_group()
