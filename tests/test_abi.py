"""CPU suite: the C-ABI library loads, exports every function include/*.h declares, and its structs
have the reference's layout (no compute calls: there is no GPU here)."""
import ctypes as C
import glob
import os
import re
import subprocess

import hpgmg_b200.api as api
from hpgmg_b200 import _structs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "hpgmg_*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
        text = re.sub(r"typedef\s+\w[\w\s\*]*\(\s*\*\s*\w+\s*\)\s*\([^;]*\)\s*;", "", text)      # function-pointer typedefs
        for m in re.finditer(r"\b([A-Za-z_]\w*)\s*\([^;{}()]*(?:\([^()]*\)[^;{}()]*)*\)\s*(?:__asm__\s*\(\s*\"(\w+)\"\s*\))?\s*;", text):
            names.add(m.group(2) or m.group(1))
    return names - {"__attribute__", "aligned", "defined", "sizeof"}


def test_library_exports_every_declared_symbol():
    assert os.path.exists(api.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(api.LIB_PATH)
    declared = declared_functions()
    assert len(declared) > 70, sorted(declared)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_helmholtz_flavour_exports_the_same_surface():
    """libhpgmg_b200_helmholtz.so = the same sources with -DUSE_HELMHOLTZ (the reference's compile-time switch, defines.h:12-26):
    same exported functions; the vector-id map of that build is in include/hpgmg_defines.h under the same #if."""
    helm = os.path.join(os.path.dirname(api.LIB_PATH), "libhpgmg_b200_helmholtz.so")
    assert os.path.exists(helm), "build first: make -C hpgmg_b200/csrc"
    lib = C.CDLL(helm)
    missing = [n for n in sorted(declared_functions()) if not hasattr(lib, n)]
    assert not missing, missing
    text = open(os.path.join(ROOT, "include", "hpgmg_defines.h")).read()
    assert "VECTOR_ALPHA     9" in text and "VECTOR_L1INV    10" in text and "VECTORS_RESERVED 11" in text      # reference defines.h:24-26


def test_reference_link_surface_is_complete():
    """The 34 external symbols the reference's operators.fv4.o defines plus what mg.o/solvers.o/level.o
    export must all be provided, under the reference's names."""
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True, check=True).stdout
    have = {l.split()[-1] for l in out.splitlines() if l.strip()}
    need = {"stencil_get_radius", "stencil_get_shape", "apply_op", "residual", "smooth", "rebuild_operator",
            "rebuild_operator_blackbox", "restriction", "interpolation_vcycle", "interpolation_fcycle", "interpolation_v2",
            "interpolation_v4", "exchange_boundary", "apply_BCs", "apply_BCs_v1", "apply_BCs_v2", "apply_BCs_v4",
            "extrapolate_betas", "dot", "norm", "mean", "error", "add_vectors", "scale_vector", "zero_vector",
            "shift_vector", "mul_vectors", "invert_vector", "init_vector", "color_vector", "random_vector",
            "initialize_problem", "evaluateBeta", "evaluateF",
            "create_level", "destroy_level", "create_vectors", "reset_level_timers", "qsortInt", "append_block_to_list",
            "MGBuild", "MGSolve", "FMGSolve", "FMGSolve2", "MGPCG", "MGVCycle", "MGDestroy", "MGPrintTiming", "MGResetTimers",
            "richardson_error", "IterativeSolver", "IterativeSolver_NumVectors"}
    assert need <= have, sorted(need - have)


def test_python_signatures_cover_the_bound_surface():
    api.bind(C.CDLL(api.LIB_PATH))            # raises AttributeError on the first missing symbol


def test_struct_layout_equals_reference():
    assert C.sizeof(S.blockCopy_type) == 128 and C.sizeof(S.communicator_type) == 104 and C.sizeof(S.box_type) == 56
    assert C.sizeof(S.level_type) == 1296 and C.sizeof(S.mg_type) == 40
    # the C compiler agrees with ctypes
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "hpgmg_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(blockCopy_type), sizeof(communicator_type), sizeof(box_type),
  sizeof(level_type), sizeof(mg_type), offsetof(level_type, timers)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        got = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert got[:5] == ["128", "104", "56", "1296", "40"], got
    assert int(got[5]) == S.level_type.timers.offset


def test_reference_headers_are_forwarded():
    """Code written against the reference tree (#include "level.h", "mg.h", ...) compiles unchanged."""
    src = '#include "defines.h"\n#include "level.h"\n#include "operators.h"\n#include "mg.h"\n#include "solvers.h"\nint main(void){level_type l; mg_type m; (void)l; (void)m; return VECTOR_F==2 ? 0 : 1;}\n'
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-c", os.path.join(d, "t.c"), "-o", os.path.join(d, "t.o")], check=True)


def test_no_product_file_touches_the_oracle():
    """The product must never route through oracle/ (or any CPU path)."""
    bad = []
    for path in glob.glob(os.path.join(ROOT, "hpgmg_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".c", ".cu", ".cuh", ".h", "Makefile")):
            text = open(path, errors="ignore").read()
            if re.search(r"oracle/|hpgmg_oracle|libhpgmg_ref|oracle_bindings", text):
                bad.append(os.path.relpath(path, ROOT))
    assert not bad, bad
