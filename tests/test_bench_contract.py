"""CPU suite: the parts of bench.py's contract that need no GPU -- the reference arm (`--impl reference`: the
reference's own CPU build, oracle/_ref/ref_bench, on the host cores) prints exactly one JSON line with the keys the
driver reads, on the process's real stdout, and the algorithmic-bytes model is the one DESIGN.md states."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")


@pytest.mark.skipif(not os.path.exists(REF_BENCH), reason="oracle/_ref/ref_bench not built")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2-box-dim", "4", "--boxes-per-rank", "8",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout + r.stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fmg_dof_per_s" and d["unit"] == "DOF/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # 32^3 as 2^3 boxes of 16^3: the F-cycle norm the reference prints for `hpgmg-fv 4 8` (tests/golden/goldens.json)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "goldens.json")))["solves"]["4 8 gsrb"]["norms"][0]
    assert abs(d["f_cycle_norm"] - gold) <= 1e-6 * abs(gold)      # the arm may run the -Ofast build (5e-8 relative)
    assert d["config"]["workload"].startswith("hpgmg-fv 4 8 per rank on 1 rank(s): fv4 GSRB FMG F-cycle on 32^3")


def test_both_arms_print_the_same_config():
    """The driver compares the `config` dicts of the two arms: they come from one function."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(log2_box_dim=7, boxes_per_rank=8, scaling="weak", global_dim=512, smoother="gsrb")
    for world, dim, boxes in ((1, 256, 8), (2, 256, 8), (4, 384, 27), (8, 512, 64)):
        log2, bpr, d, total, cfg = bench.resolve_workload(a, world)
        assert (log2, bpr, d, total) == (7, 8, dim, boxes) and f"on {dim}^3" in cfg["workload"]
    s = argparse.Namespace(log2_box_dim=7, boxes_per_rank=8, scaling="strong", global_dim=512, smoother="cheby")
    for world in (1, 2, 4, 8):                   # BASELINE config 5: 7 64 / 7 32 / 7 16 / 7 8
        log2, bpr, d, total, cfg = bench.resolve_workload(s, world)
        assert (bpr, d, total) == (64 // world, 512, 64) and cfg["smoother"] == "cheby" and cfg["scaling"] == "strong"
    assert bench.golden_norm(7, 64, "cheby") == 1.5060825298007785e-07
    assert bench.golden_norm(8, 8, "gsrb") == 4.151187798033961e-08


def test_algorithmic_bytes_model():
    """SURVEY.md 8d / DESIGN.md 4: V-cycle visit 2*6*56 + 48 + 9 + 1 + 17 = 747 B per cell of the level, a level m is
    visited m+1 times with 8^-m of the cells, plus 100.6 B once per solve => 1076 B per fine DOF."""
    sys.path.insert(0, ROOT)
    import bench
    visit = 2 * 6 * 56 + 48 + 9 + 1 + 17
    assert visit == 747
    fmg = visit * sum((m + 1) / 8.0 ** m for m in range(40)) + (8 + 16 + 9 * 8 / 7 + 9 * 8 / 7 + 48 + 8)
    assert abs(fmg - bench.ALGORITHMIC_BYTES_PER_DOF) < 1.0
    assert bench.GSRB_SWEEP_BYTES_PER_CELL == 56 and bench.CHEBY_SWEEP_BYTES_PER_CELL == 64
    visit_c = 2 * 6 * 64 + 48 + 9 + 1 + 17
    fmg_c = visit_c * sum((m + 1) / 8.0 ** m for m in range(40)) + (8 + 16 + 9 * 8 / 7 + 9 * 8 / 7 + 48 + 8)
    assert abs(fmg_c - bench.ALGORITHMIC_BYTES_PER_DOF_CHEBY) < 1.0
