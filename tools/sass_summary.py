"""Per-kernel instruction census of the built library: `cuobjdump -sass hpgmg_b200/lib/libhpgmg_b200.so` -> one row per kernel
with the mnemonics that prove what the kernel is (TMA bulk-tensor loads, mbarrier ops, FP64 pipe, shared-memory loads,
128-bit global accesses).  Output: profiles/rNN_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hpgmg_b200", "lib", "libhpgmg_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
KEYS = ["UTMALDG", "UBLKCP", "SYNCS", "ACQBULK", "LDS", "STS", "LDG.E.128", "STG.E.128", "LDG", "STG", "DADD", "DMUL", "DFMA", "MUFU.RCP64H", "SHFL", "BAR.SYNC", "ATOM", "RED", "MEMBAR", "total"]
rows = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        rows[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    c = rows[name]
    c["total"] += 1
    for k in KEYS[:-1]:
        if op == k or op.startswith(k + ".") or (k in ("LDG.E.128", "STG.E.128") and k.split(".")[0] in op and ".128" in op):
            c[k] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(rows)} kernels, arch {', '.join(arch)} (cuobjdump -sass; static instruction counts)")
print("# LDG / STG include their .128 forms; DFMA appears only inside IEEE division sequences (the library is built -fmad=false)")
w = max(len(n) for n in rows) if rows else 10
print(f"{'kernel':<{w}} " + " ".join(f"{k:>11}" for k in KEYS))
for n, c in sorted(rows.items(), key=lambda kv: -kv[1]["total"]):
    print(f"{n:<{w}} " + " ".join(f"{c[k]:>11}" for k in KEYS))
