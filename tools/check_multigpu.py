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
    print("  PARITY", ("OK (bit-exact)" if printed is None else "OK (all 16 printed digits of the reference run)") if ok else "MISMATCH")
H.close()
if world > 1:
    import torch.distributed as dist
    L.hpgmg_b200_comm_finalize()
    dist.destroy_process_group()
