#!/usr/bin/env python
"""bench.py -- HPGMG-FV fv4 FMG DOF/s on B200 (BASELINE.json metric) + roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2-box-dim 7] [--boxes-per-rank 8]

A step is what the reference's bench_hpgmg times (hpgmg-fv.c:78-80): zero_vector(U); FMGSolve(...)
on the finest level.  Workload at N=1: `hpgmg-fv 7 8` = 256^3 as 2^3 boxes of 128^3 (configs[1]);
at N>1 the same `7 8` per rank (weak scaling; the reference only builds cubic domains, so 2 ranks
give 256^3, 4 ranks 384^3, 8 ranks 512^3 -- SURVEY.md appendix B) and value counts the global DOF.
Inputs (1.3 GB of level-0 vectors) are far larger than the 126 MB L2, so no explicit flush is used.

`--impl reference` times the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile) on the host's cores with OpenMP through oracle/_ref/ref_bench.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGORITHMIC_BYTES_PER_DOF = 1076.0      # one GSRB F-cycle, per fine-grid DOF
GSRB_SWEEP_BYTES_PER_CELL = 56.0        # x, rhs, Dinv, 3 betas read + x written


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(log2_box_dim, boxes, warmup, steps, threads=None):
    """Time the reference's own FMGSolve on the host cores (oracle/_ref/ref_bench)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
    if not os.path.exists(exe):
        return None
    env = dict(os.environ)
    threads = threads or os.cpu_count() or 1
    env["OMP_NUM_THREADS"] = str(threads)
    out = subprocess.run([exe, str(log2_box_dim), str(boxes), str(warmup), str(steps)], capture_output=True, text=True, env=env).stdout
    m = re.search(r"REF dof=(\d+) seconds_per_solve=([\d.eE+-]+) norm=([\d.eE+-]+) rel=([\d.eE+-]+) threads=(\d+)", out)
    if not m:
        return None
    return {"dof": float(m.group(1)), "seconds": float(m.group(2)), "norm": float(m.group(3)), "threads": int(m.group(5))}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    # the N-rank problem is the 1-process problem with boxes_per_rank*N boxes
    boxes = args.boxes_per_rank * world
    r = run_reference(args.log2_box_dim, boxes, args.warmup, args.steps)
    if r is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_bench missing (reference not built)"})
        return
    value = r["dof"] / r["seconds"]
    line = {"impl": "reference", "metric": "fmg_dof_per_s", "value": value, "unit": "DOF/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (analytic problem.fv.c rhs/beta)",
            "config": {"workload": f"hpgmg-fv {args.log2_box_dim} {args.boxes_per_rank} per rank, fv4 GSRB FMG, {int(round(r['dof'] ** (1 / 3)))}^3"},
            "cpu_baseline": {"value": value, "unit": "DOF/s", "cores": r["threads"], "kind": "reference",
                             "sample": f"{args.steps} FMGSolve after {args.warmup} warm-up, whole workload"},
            "e2e": {"value": value, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "f_cycle_norm": r["norm"]}
    emit(line)


def ours(args):
    import numpy as np
    import hpgmg_b200.api as api
    rank, world = api.init_distributed()
    if world != args.gpus:
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    L = api.lib()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

    H = api.Hierarchy(args.log2_box_dim, args.boxes_per_rank, my_rank=rank, num_ranks=world, verbose=False, use_graphs=not args.no_graphs)
    lvl = H.level(0)
    dof = float(H.dof(0))

    def barrier():
        L.hpgmg_b200_sync()
        if dist is not None:
            dist.barrier()
            L.hpgmg_b200_sync()

    def step():
        return H.fmg_solve(0)

    for _ in range(max(args.warmup, 3)):
        norm_r, rel = step()

    if args.ncu:                     # profiling aid: bracket ONE region for `ncu --profile-from-start off` and leave
        L.hpgmg_b200_profiler_start()
        if args.ncu == "solve":
            step()
        else:                        # two level-0 GSRB sweeps (one per colour)
            for s in range(2):
                L.hpgmg_b200_gsrb_sweep(lvl, api.VECTOR_U if s % 2 == 0 else api.VECTOR_TEMP, api.VECTOR_TEMP if s % 2 == 0 else api.VECTOR_U, api.VECTOR_F, H.a, H.b, s)
        L.hpgmg_b200_profiler_stop()
        H.close()
        return

    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    launches0 = L.hpgmg_b200_kernel_launches()
    barrier()
    t0 = time.perf_counter()
    L.hpgmg_b200_bench_mark(0)
    for _ in range(args.steps):
        norm_r, rel = step()
    L.hpgmg_b200_bench_mark(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = L.hpgmg_b200_bench_elapsed_ms(0, 1)
    launches = L.hpgmg_b200_kernel_launches() - launches0
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = dof / (1e-3 * ms_per_step)

    # ---- end to end through the C-ABI with HOST buffers (pinned): H2D f, zero u, FMGSolve, D2H u ----
    Lc = lvl.contents
    nbytes = Lc.num_my_boxes * Lc.box_volume * 8
    e2e = None
    if nbytes > 0:
        f_host = L.hpgmg_b200_host_alloc_pinned(nbytes)
        u_host = L.hpgmg_b200_host_alloc_pinned(nbytes)
        vol = Lc.box_volume
        for b in range(Lc.num_my_boxes):
            arr = api.download(lvl, b, api.VECTOR_F).reshape(-1)
            C.memmove(f_host + b * vol * 8, arr.ctypes.data, vol * 8)
        for _ in range(2):
            L.hpgmg_fmg_solve_host(H.mg, 0, api.VECTOR_U, api.VECTOR_F, H.a, H.b, 1e-10, f_host, u_host)
        barrier()
        te = time.perf_counter()
        for _ in range(args.steps):
            e2e_norm = L.hpgmg_fmg_solve_host(H.mg, 0, api.VECTOR_U, api.VECTOR_F, H.a, H.b, 1e-10, f_host, u_host)
        barrier()
        e2e_s = (time.perf_counter() - te) / args.steps
        if dist is not None:
            import torch
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        moved = int(L.hpgmg_fmg_solve_host_bytes(H.mg, 0))          # the cells of f in, the cells of u (+ 3 scalars) out
        e2e = {"value": dof / e2e_s, "unit": "DOF/s", "h2d_bytes_per_step": moved, "d2h_bytes_per_step": moved + 24,
               "ms_per_step": 1e3 * e2e_s, "f_cycle_norm": e2e_norm}
        L.hpgmg_b200_host_free_pinned(f_host)
        L.hpgmg_b200_host_free_pinned(u_host)

    # ---- roofline of the dominant kernel: one level-0 GSRB sweep, timed alone with CUDA events ----
    reps = 20
    for s in range(2):
        L.hpgmg_b200_gsrb_sweep(lvl, api.VECTOR_U if s % 2 == 0 else api.VECTOR_TEMP, api.VECTOR_TEMP if s % 2 == 0 else api.VECTOR_U, api.VECTOR_F, H.a, H.b, s)
    L.hpgmg_b200_bench_mark(2)
    for s in range(reps):
        L.hpgmg_b200_gsrb_sweep(lvl, api.VECTOR_U if s % 2 == 0 else api.VECTOR_TEMP, api.VECTOR_TEMP if s % 2 == 0 else api.VECTOR_U, api.VECTOR_F, H.a, H.b, s)
    L.hpgmg_b200_bench_mark(3)
    L.hpgmg_b200_sync()
    sweep_ms = L.hpgmg_b200_bench_elapsed_ms(2, 3) / reps
    local_cells = Lc.num_my_boxes * Lc.box_dim ** 3
    peak, peak_src = measured_peaks()
    achieved = GSRB_SWEEP_BYTES_PER_CELL * local_cells / (1e-3 * sweep_ms) / 1e9
    roofline = {"bound": "hbm", "kernel": "gsrb_sweep(level 0)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "us_per_launch": 1e3 * sweep_ms,
                "algorithmic_bytes_per_launch": GSRB_SWEEP_BYTES_PER_CELL * local_cells,
                "solve_frac_of_hbm_roofline": (value / max(world, 1)) * ALGORITHMIC_BYTES_PER_DOF / (peak * 1e9)}
    ncu = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(ncu):
        with open(ncu) as f:
            roofline["traffic"] = json.load(f).get("gsrb_sweep_dram_bytes_per_launch")

    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        # ---- CPU baseline: the reference itself on this box's cores, bounded sample ----
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = run_reference(args.log2_box_dim, args.boxes_per_rank, 1, 3)
            if r:
                cpu = {"value": r["dof"] / r["seconds"], "unit": "DOF/s", "cores": r["threads"], "kind": "reference",
                       "sample": "3 FMGSolve (after 1 warm-up) of the same 256^3 workload, reference built -O2 -fopenmp",
                       "f_cycle_norm": r["norm"]}
        dim = Lc.dim.i
        line = {"metric": "fmg_dof_per_s", "value": value, "unit": "DOF/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic (analytic problem.fv.c rhs/beta, deterministic)",
                "config": {"workload": f"hpgmg-fv {args.log2_box_dim} {args.boxes_per_rank} per rank: fv4 GSRB FMG F-cycle on {dim}^3 ({Lc.num_my_boxes} boxes of {Lc.box_dim}^3 on rank 0)",
                           "levels": H.num_levels, "l2": "inputs larger than L2 (1.3 GB of level-0 vectors per GPU), no flush",
                           "cuda_graphs": not args.no_graphs, "timing": "CUDA events on the library stream, max over ranks"},
                "gpu_launches": int(launches), "wall_ms_per_step": 1e3 * wall / args.steps,
                "f_cycle_norm": norm_r, "f_cycle_rel": rel,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": sampler.summary()}
        emit(line)
    H.close()
    if dist is not None:
        L.hpgmg_b200_comm_finalize()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries (NCCL's version banner, torchrun) may write to fd 1: keep the real stdout for the JSON line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-box-dim", type=int, default=7)
    ap.add_argument("--boxes-per-rank", type=int, default=8)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu", default="", choices=["", "solve", "sweep"], help="bracket one solve / two GSRB sweeps with cudaProfilerStart/Stop and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
