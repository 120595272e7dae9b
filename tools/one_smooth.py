"""profiling aid: one smooth() + one residual() on level `l` of `hpgmg-fv 7 8`, bracketed for `ncu --profile-from-start off`"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgmg_b200.api as api
api.init(0)
L = api.lib()
l = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = api.Hierarchy(7, 8, use_graphs=False)
H.fmg_solve(0)
lv = H.level(l)
for _ in range(3):
    L.smooth(lv, api.VECTOR_U, api.VECTOR_R, 0.0, 1.0)
L.hpgmg_b200_profiler_start()
L.smooth(lv, api.VECTOR_U, api.VECTOR_R, 0.0, 1.0)
L.residual(lv, api.VECTOR_TEMP, api.VECTOR_U, api.VECTOR_R, 0.0, 1.0)
L.hpgmg_b200_profiler_stop()
H.close()
