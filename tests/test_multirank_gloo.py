"""CPU suite, world_size 2 over gloo: the N>1 host logic.  Each rank builds ITS side of the `5 8`
two-rank decomposition (layout-only mode), then the ranks swap their send/recv tables over
torch.distributed and check that every message one rank sends is exactly what the other expects
(same size, for ghost exchange of all three shapes, restriction of all four kinds and
interpolation) -- the property MPI_Irecv/Isend pairs rely on (exchange_boundary.c:33-97)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist
import hpgmg_b200.api as api
import oracle_bindings as ob
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
L = api.lib(); L.hpgmg_b200_set_layout_only(1); L.hpgmg_b200_set_verbose(0)
for agglomerate in (0, 16):
  L.hpgmg_b200_set_agglomeration(agglomerate)
  H = api.Hierarchy(5, 8, my_rank=rank, num_ranks=world, build_operator=False)
  L.MGBuild(H.mg, H.level_h, 0.0, 1.0, 1); H.built = True
  mine = [ob.level_list_summary(H.level(l)) for l in range(H.num_levels)]
  everyone = [None] * world
  dist.all_gather_object(everyone, mine)
  gold = ob.goldens()["decompositions"]["5 8 x%d" % world]
  if agglomerate == 0:
    assert mine == gold[rank], "lists differ from the reference for rank %d" % rank
  else:
    # ownership agglomeration: same boxes on every level, levels with boxes <= 16^3 entirely on rank 0
    for lv, g in zip(mine, gold[rank]):
      assert (lv["dim"], lv["box_dim"], lv["boxes_in"]) == (g["dim"], g["box_dim"], g["boxes_in"])
      if lv["box_dim"] <= 16:
        assert set(lv["rank_of_box"]) == {0} and lv["num_my_boxes"] == (len(lv["rank_of_box"]) if rank == 0 else 0)
      else:
        assert lv["rank_of_box"] == g["rank_of_box"]
  bad = []
  for l in range(len(mine)):
      comms = [("exchange", s) for s in range(3)] + [("restriction", t) for t in range(4)] + [("interpolation", None)]
      for kind, idx in comms:
          get = lambda r, lev: everyone[r][lev][kind] if idx is None else everyone[r][lev][kind][idx]
          # sender side of level l talks to: the same level (exchange), level l+1 (restriction), level l-1 (interpolation)
          lr = l if kind == "exchange" else (l + 1 if kind == "restriction" else l - 1)
          if lr < 0 or lr >= len(mine):
              continue
          for dst, size in get(rank, l)["send"]:
              expect = dict(get(dst, lr)["recv"]).get(rank)
              if expect != size:
                  bad.append((kind, idx, l, rank, dst, size, expect))
  assert not bad, bad
  H.close()
total_boxes = torch.tensor([sum(1 for r in mine[0]["rank_of_box"] if r == rank)])
dist.all_reduce(total_boxes)
assert int(total_boxes) == len(mine[0]["rank_of_box"])
dist.barrier()
if rank == 0:
    print("OK", world)
'''


def test_two_rank_message_tables_agree():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", "-c", WORKER]
    # torchrun has no -c: write the worker to a temp file
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(WORKER)
        path = f.name
    try:
        cmd = cmd[:-2] + [path, ROOT]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert "OK 2" in r.stdout
    finally:
        os.unlink(path)
