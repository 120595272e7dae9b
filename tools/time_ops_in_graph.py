"""In-graph cost of the level operators: K back-to-back calls of one operator on one level are recorded into a CUDA
graph (the way a solve runs them) and replayed; prints microseconds per call.  `hpgmg-fv 7 8` on one GPU."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpgmg_b200.api as api

rank, world = api.init_distributed()          # plain `python` (1 GPU) or under torchrun
L = api.lib()
L.hpgmg_graph_begin.restype = C.c_int
L.hpgmg_graph_begin.argtypes = [C.c_void_p, C.c_longlong]
L.hpgmg_graph_end.argtypes = [C.c_void_p, C.c_longlong]
log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 7
H = api.Hierarchy(log2, 8, my_rank=rank, num_ranks=world, use_graphs=True)
H.fmg_solve(0)
U, R, T, E = api.VECTOR_U, api.VECTOR_R, api.VECTOR_TEMP, api.VECTOR_E
key = [1000]


def timed(fn, K=20, reps=5):
    key[0] += 1
    owner = C.c_void_p(key[0])
    if L.hpgmg_graph_begin(owner, key[0]):
        for _ in range(K):
            fn()
        L.hpgmg_graph_end(owner, key[0])
    L.hpgmg_b200_sync()
    L.hpgmg_b200_bench_mark(0)
    for _ in range(reps):
        L.hpgmg_graph_begin(owner, key[0])          # replay
    L.hpgmg_b200_bench_mark(1)
    L.hpgmg_b200_sync()
    return 1e3 * L.hpgmg_b200_bench_elapsed_ms(0, 1) / (K * reps)


for l in range(0, min(H.num_levels - 1, 4 if world > 1 else 99)):
    lv = H.level(l)
    lc = H.level(l + 1)
    c = lv.contents
    row = {
        "smooth": timed(lambda: L.smooth(lv, U, R, 0.0, 1.0)),
        "residual": timed(lambda: L.residual(lv, T, U, R, 0.0, 1.0)),
        "fill(x+bc)": timed(lambda: (L.exchange_boundary(lv, U, 2), L.apply_BCs(lv, U, 2))),
        "restrict": timed(lambda: L.restriction(lc, R, lv, T, 0)),
        "zero": timed(lambda: L.zero_vector(lc, U)),
        "interp_v2": timed(lambda: L.interpolation_vcycle(lv, U, 1.0, lc, U)),
        "interp_v4": timed(lambda: L.interpolation_fcycle(lv, U, 0.0, lc, U)),
        "vcycle": timed(lambda: L.MGVCycle(H.mg, U, R, 0.0, 1.0, l), K=4),
    }
    if rank == 0:
        print(f"world {world} level {l} dim {c.dim.i} box {c.box_dim} boxes {c.num_my_boxes} ranks {c.num_ranks}: " + "  ".join(f"{k} {v:.1f}" for k, v in row.items()), flush=True)
H.close()
if world > 1:
    import torch.distributed as dist
    L.hpgmg_b200_comm_finalize()
    dist.destroy_process_group()
