"""Phase clocks of the single-block coarse kernel (hpgmg_b200_coarse_profile): cycles per category for
MGVCycle started at each coarse-chain level of `hpgmg-fv 7 8`."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgmg_b200.api as api

api.init(0)
L = api.lib()
H = api.Hierarchy(7, 8, use_graphs=False)
H.fmg_solve(0)
names = ["load", "fill", "stencil", "restrict", "zero", "interp", "bottom", "store", "total"]
out = (C.c_longlong * 9)()
L.hpgmg_b200_coarse_profile(1, None)
for level in range(H.num_levels - 1, 3, -1):
    for _ in range(3):
        L.MGVCycle(H.mg, api.VECTOR_E, api.VECTOR_R, 0.0, 1.0, level)
    L.hpgmg_b200_bench_mark(0)
    L.MGVCycle(H.mg, api.VECTOR_E, api.VECTOR_R, 0.0, 1.0, level)
    L.hpgmg_b200_bench_mark(1)
    L.hpgmg_b200_coarse_profile(1, C.cast(out, C.c_void_p))
    us = 1e3 * L.hpgmg_b200_bench_elapsed_ms(0, 1)
    print(f"MGVCycle(level {level}, dim {H.level(level).contents.dim.i}): {us:.1f} us;  kilo-cycles: " +
          ", ".join(f"{n} {out[i] / 1e3:.1f}" for i, n in enumerate(names)), flush=True)
H.close()
