#!/usr/bin/env python
"""Regenerate tests/golden/goldens.json from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference/finite-volume/source with gcc -O2 -fopenmp, no FMA).

  python tests/golden/make_goldens.py            # only what is missing from goldens.json
  python tests/golden/make_goldens.py --all      # everything again

Recorded per solve configuration (`hpgmg-fv <log2_box_dim> <boxes>` on one rank; an N-rank run has
the same boxes and therefore the same numbers): the F-cycle residual norms of the
driver's three Richardson solves (hpgmg-fv.c:357-365), ||error|| and the observed order
(mg.c:1128-1130), the Gershgorin bound per level (rebuild.c:204) and the bottom-solver iteration
count.  Recorded per decomposition (ranks N, rank r): counts and digests of every block list of
every level plus the send/recv tables -- the "index mapping" the port must reproduce bit for bit.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_bindings as ob  # noqa: E402
import hpgmg_b200.api as api  # noqa: E402

SOLVES = [(4, 1, False), (5, 1, False), (6, 1, False), (4, 8, False), (5, 8, False), (6, 8, False), (7, 8, False),
          (4, 27, False), (5, 27, False), (5, 64, False),
          (7, 27, False), (7, 64, False),      # = `7 8` per GPU on 4 / 8 GPUs (384^3 and 512^3: 4.5 and 10.6 GB on the host)
          (5, 1, True), (5, 8, True), (6, 8, True), (5, 27, True),
          (8, 8, False),                        # BASELINE config 4 on one GPU: 512^3 as 2^3 boxes of 256^3 (10 GB on the host)
          (7, 64, True),                        # BASELINE config 5: Chebyshev, 512^3 as 4^3 boxes of 128^3 (the 1/2/4/8-GPU strong-scaling grid)
          (7, 8, True)]                         # Chebyshev on the headline grid
DECOMPOSITIONS = [(5, 8, 1), (5, 8, 2), (5, 8, 4), (5, 8, 8), (4, 1, 1), (6, 8, 1), (6, 8, 8), (4, 3, 9)]


def solve_record(log2, boxes, cheby):
    H = ob.RefHierarchy(log2, boxes, cheby=cheby)
    H.call("MGResetTimers", H.mg)          # the reference never initialises Krylov_iterations otherwise
    norms = []
    for l in range(3):
        if l > 0:
            H.call("restriction", H.level(l), api.VECTOR_F, H.level(l - 1), api.VECTOR_F, api.RESTRICT_CELL)
        norms.append(H.fmg_solve(l))
    # richardson_error (mg.c:1113-1131) prints only: redo its arithmetic with the same calls
    L1, L2, L0 = H.level(1), H.level(2), H.level(0)
    H.call("restriction", L1, api.VECTOR_TEMP, L0, api.VECTOR_U, api.RESTRICT_CELL)
    H.call("restriction", L2, api.VECTOR_TEMP, L1, api.VECTOR_U, api.RESTRICT_CELL)
    H.call("add_vectors", L1, api.VECTOR_TEMP, 1.0, api.VECTOR_U, -1.0, api.VECTOR_TEMP)
    H.call("add_vectors", L2, api.VECTOR_TEMP, 1.0, api.VECTOR_U, -1.0, api.VECTOR_TEMP)
    e2h, e4h = H.call("norm", L1, api.VECTOR_TEMP), H.call("norm", L2, api.VECTOR_TEMP)
    import math
    rec = {"norms": norms, "error": e2h, "order": math.log(e4h / e2h) / math.log(2),
           "eigs": [H.level(l).contents.dominant_eigenvalue_of_DinvA for l in range(H.num_levels)],
           "dims": [H.level(l).contents.dim.i for l in range(H.num_levels)],
           "box_dims": [H.level(l).contents.box_dim for l in range(H.num_levels)],
           "krylov_iterations_3_solves": H.level(H.num_levels - 1).contents.Krylov_iterations,
           "norm_of_F": H.call("norm", L0, api.VECTOR_F)}
    return rec


# the driver's compile-time variants.  (log2, boxes, periodic, helmholtz)
VARIANTS = [(4, 1, True, False), (4, 8, True, False), (5, 8, True, False), (4, 27, True, False),
            (4, 8, False, True), (5, 8, False, True), (4, 8, True, True), (4, 27, False, True)]


def variant_record(log2, boxes, periodic, helmholtz):
    """hpgmg-fv built with -DUSE_PERIODIC_BC and/or -DUSE_HELMHOLTZ (hpgmg-fv.c:276-302): the three Richardson solves, on ONE
    OpenMP thread (the reference's mean() and dot() sum tile partials in thread order)."""
    import ctypes as C
    lib = None
    a, b, vectors = 0.0, 1.0, None
    if helmholtz:
        lib = api.bind(C.CDLL(os.path.join(ob.REF_DIR, "libhpgmg_ref_helmholtz.so")), {k: api.SIGNATURES[k] for k in ob._REF_SYMBOLS})
        a, b, vectors = 1.0, 1.0, 11
    with ob.ref_threads(1):
        H = ob.RefHierarchy(log2, boxes, bc=api.BC_PERIODIC if periodic else api.BC_DIRICHLET, library=lib, a=a, b=b, vectors=vectors)
        norms = []
        for l in range(3):
            if l > 0:
                H.call("restriction", H.level(l), api.VECTOR_F, H.level(l - 1), api.VECTOR_F, api.RESTRICT_CELL)
            with ob.quiet():
                H.L.zero_vector(H.level(l), api.VECTOR_U)
                H.L.FMGSolve(H.mg, l, api.VECTOR_U, api.VECTOR_F, a, b, 1e-10)
                H.L.residual(H.level(l), api.VECTOR_TEMP, api.VECTOR_U, api.VECTOR_F, a, b)
                norms.append(H.L.norm(H.level(l), api.VECTOR_TEMP))
        L1, L2, L0 = H.level(1), H.level(2), H.level(0)
        H.call("restriction", L1, api.VECTOR_TEMP, L0, api.VECTOR_U, api.RESTRICT_CELL)
        H.call("restriction", L2, api.VECTOR_TEMP, L1, api.VECTOR_U, api.RESTRICT_CELL)
        H.call("add_vectors", L1, api.VECTOR_TEMP, 1.0, api.VECTOR_U, -1.0, api.VECTOR_TEMP)
        H.call("add_vectors", L2, api.VECTOR_TEMP, 1.0, api.VECTOR_U, -1.0, api.VECTOR_TEMP)
        e2h, e4h = H.call("norm", L1, api.VECTOR_TEMP), H.call("norm", L2, api.VECTOR_TEMP)
    import math
    return {"norms": norms, "error": e2h, "order": math.log(e4h / e2h) / math.log(2),
            "eigs": [H.level(l).contents.dominant_eigenvalue_of_DinvA for l in range(H.num_levels)],
            "dims": [H.level(l).contents.dim.i for l in range(H.num_levels)], "a": a, "b": b}


def decomposition_record(log2, boxes_per_rank, ranks):
    out = []
    for r in range(ranks):
        H = ob.RefHierarchy(log2, boxes_per_rank, my_rank=r, num_ranks=ranks, build_operator=False)
        H.build_lists_only()
        out.append([ob.level_list_summary(H.level(l)) for l in range(H.num_levels)])
    return out


def main():
    assert ob.have_ref() and ob.have_ref(True), "build oracle/_ref first: make -C oracle ref"
    G = {"generator": "tests/golden/make_goldens.py", "reference_flags": "gcc -O2 -fopenmp -std=gnu99 -DUSE_BICGSTAB=1 -DUSE_SUBCOMM=1 -DUSE_FCYCLES=1 -DUSE_{GSRB,CHEBY}=1 (x86-64 baseline: no FMA)",
         "solves": {}, "decompositions": {}}
    path = os.path.join(HERE, "goldens.json")
    if "--all" not in sys.argv and os.path.exists(path):
        with open(path) as f:
            G = json.load(f)
    for log2, boxes, cheby in SOLVES:
        key = f"{log2} {boxes} {'cheby' if cheby else 'gsrb'}"
        if key in G["solves"]:
            continue
        print("solve", key, flush=True, file=sys.stderr)
        G["solves"][key] = solve_record(log2, boxes, cheby)
    for log2, bpr, ranks in DECOMPOSITIONS:
        key = f"{log2} {bpr} x{ranks}"
        if key in G["decompositions"]:
            continue
        print("decomposition", key, flush=True, file=sys.stderr)
        G["decompositions"][key] = decomposition_record(log2, bpr, ranks)
    G.setdefault("variants", {})
    for log2, boxes, periodic, helmholtz in VARIANTS:
        key = f"{log2} {boxes} {'periodic' if periodic else 'dirichlet'} {'helmholtz' if helmholtz else 'poisson'}"
        if key in G["variants"]:
            continue
        print("variant", key, flush=True, file=sys.stderr)
        G["variants"][key] = variant_record(log2, boxes, periodic, helmholtz)
    with open(os.path.join(HERE, "goldens.json"), "w") as f:
        json.dump(G, f, indent=0, separators=(",", ":"))
    print("wrote goldens.json", file=sys.stderr)


if __name__ == "__main__":
    main()
