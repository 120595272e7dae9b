"""Micro-timing of the coarse end of the cycle: MGVCycle(level) for each level of `hpgmg-fv 7 8`,
coarse kernel on/off, shared-memory residency on/off (CUDA events on the library stream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgmg_b200.api as api

api.init(0)
L = api.lib()
H = api.Hierarchy(7, 8, use_graphs=False)
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


for coarse, smem in ((0, 0), (1, 0), (1, 1)):
    L.hpgmg_b200_use_coarse_kernel(coarse)
    L.hpgmg_b200_coarse_levels_in_smem(smem)
    row = []
    for level in range(H.num_levels - 1, 1, -1):
        row.append((H.level(level).contents.dim.i, round(time_vcycle(level), 1)))
    print(f"coarse_kernel={coarse} smem={smem}  us per MGVCycle(level): {row}")
