"""Run under torchrun (one rank per GPU): build `hpgmg-fv <log2> <boxes_per_rank>` across the ranks, run the
driver's Richardson pass and compare with the goldens of the equivalent single-process reference run
(an N-rank run has the same boxes as a 1-rank run with N x the boxes)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hpgmg_b200.api as api

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 5
bpr = int(sys.argv[2]) if len(sys.argv) > 2 else 8
graphs = "nograph" not in sys.argv[3:]
smoother = "cheby" if "cheby" in sys.argv[3:] else "gsrb"
rank, world = api.init_distributed()
L = api.lib()
H = api.Hierarchy(log2, bpr, my_rank=rank, num_ranks=world, use_graphs=graphs, smoother=api.SMOOTHER_CHEBY if smoother == "cheby" else api.SMOOTHER_GSRB)
err, order, norms = H.richardson()
boxes = H.boxes_in_i ** 3
# cell by cell: u of the finest level, every box this rank owns, against the unmodified reference run in ONE process with
# all the boxes (oracle/_ref, shipped prebuilt; each rank runs it on the host for itself)
cells_ok = None
if "cells" in sys.argv[3:]:
    import numpy as np
    import oracle_bindings as ob
    if ob.have_ref(smoother == "cheby"):
        with ob.ref_threads(1):
            R = ob.RefHierarchy(log2, bpr * world, cheby=(smoother == "cheby"))
            R.fmg_solve(0)
        lvl = H.level(0)
        Lc = lvl.contents
        cells_ok = 1
        for b in range(Lc.num_my_boxes):
            gid = Lc.my_boxes[b].global_box_id
            ours = api.interior(lvl, api.download(lvl, b, api.VECTOR_U))
            ref = api.interior(R.level(0), R.array(0, gid, api.VECTOR_U))
            if not np.array_equal(ours, ref):
                cells_ok = 0
        if world > 1:
            import torch
            import torch.distributed as dist
            t = torch.tensor([cells_ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            cells_ok = int(t.item())
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "goldens.json")))["solves"].get(f"{log2} {boxes} {smoother}")
printed = None
if rank == 0:
    if printed is not None:
        ok = ["%.15e" % n[0] for n in norms] == printed[0] and "%.15e" % err == printed[1]
        gold = {"norms": printed[0], "error": printed[1]}
    else:
        ok = gold is not None and [n[0] for n in norms] == gold["norms"] and err == gold["error"]
    print(f"world={world} cfg={log2} {bpr}/rank -> {boxes} boxes, levels={H.num_levels} graphs={graphs} p2p={L.hpgmg_b200_p2p_enabled()}")
    print("  norms", [repr(n[0]) for n in norms], "error", repr(err), "order", round(order, 3))
    print("  golden", gold["norms"] if gold else None, gold["error"] if gold else None)
    if cells_ok is not None:
        print("  cell by cell: u of every box on every rank", "equals" if cells_ok else "DIFFERS FROM", "the single-process reference run")
        ok = bool(cells_ok) if gold is None else (ok and bool(cells_ok))          # no golden for this grid: the cells are the check
    print("  PARITY", ("OK (bit-exact)" if printed is None else "OK (all 16 printed digits of the reference run)") if ok else "MISMATCH")
H.close()
if world > 1:
    import torch.distributed as dist
    L.hpgmg_b200_comm_finalize()
    dist.destroy_process_group()
