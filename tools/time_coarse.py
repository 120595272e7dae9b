"""Micro-timing of the coarse end of the cycle: MGVCycle(level) for each level of `hpgmg-fv 7 8`,
coarse kernel on/off, shared-memory residency on/off (CUDA events on the library stream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgmg_b200.api as api

rank, world = api.init_distributed()
L = api.lib()
H = api.Hierarchy(7, 8, my_rank=rank, num_ranks=world, use_graphs=False)
H.fmg_solve(0)


def time_vcycle(level, reps=20):
    for _ in range(3):
        L.MGVCycle(H.mg, api.VECTOR_E, api.VECTOR_R, 0.0, 1.0, level)
    L.hpgmg_b200_bench_mark(0)
    for _ in range(reps):
        L.MGVCycle(H.mg, api.VECTOR_E, api.VECTOR_R, 0.0, 1.0, level)
    L.hpgmg_b200_bench_mark(1)
    L.hpgmg_b200_sync()
    return 1e3 * L.hpgmg_b200_bench_elapsed_ms(0, 1) / reps


for coarse, smem in ((1, 1),):
    L.hpgmg_b200_use_coarse_kernel(coarse)
    L.hpgmg_b200_coarse_levels_in_smem(smem)
    row = []
    for level in range(H.num_levels - 1, 1, -1):
        row.append((H.level(level).contents.dim.i, round(time_vcycle(level), 1)))
    if rank == 0:
        print(f"world={world} p2p={L.hpgmg_b200_p2p_enabled()} coarse_kernel={coarse} smem={smem}  us per MGVCycle(level): {row}", flush=True)


def time_op(name, fn, reps=20):
    for _ in range(3):
        fn()
    L.hpgmg_b200_bench_mark(0)
    for _ in range(reps):
        fn()
    L.hpgmg_b200_bench_mark(1)
    L.hpgmg_b200_sync()
    if rank == 0:
        print(f"   {name}: {1e3 * L.hpgmg_b200_bench_elapsed_ms(0, 1) / reps:.1f} us", flush=True)


for lvl in range(0, H.num_levels - 2):
    lv = H.level(lvl)
    if rank == 0:
        print(f" level {lvl} dim {lv.contents.dim.i} boxes/rank {lv.contents.num_my_boxes}")
    time_op("exchange_boundary(NO_CORNERS)", lambda: L.exchange_boundary(lv, api.VECTOR_E, 2))
    time_op("smooth (6 sweeps)", lambda: L.smooth(lv, api.VECTOR_E, api.VECTOR_R, 0.0, 1.0))
    time_op("residual", lambda: L.residual(lv, api.VECTOR_TEMP, api.VECTOR_E, api.VECTOR_R, 0.0, 1.0))
    lc = H.level(lvl + 1)
    time_op("restriction -> coarser", lambda: L.restriction(lc, api.VECTOR_R, lv, api.VECTOR_TEMP, 0))
    time_op("interpolation_v2 <- coarser", lambda: L.interpolation_vcycle(lv, api.VECTOR_E, 1.0, lc, api.VECTOR_E))
H.close()
if world > 1:
    import torch.distributed as dist
    L.hpgmg_b200_comm_finalize()
    dist.destroy_process_group()
