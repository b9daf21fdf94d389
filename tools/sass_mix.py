"""Static instruction mix of the shipped kernels: `cuobjdump -sass` of rcppml_b200/lib/RcppML_gpu.so, opcode counts per
kernel (no GPU needed). Backs the statements of DESIGN.md §4 about the code that actually ships — packed fp32 pairs
(FFMA2 / FADD2), 128-bit global and shared accesses, FP64 tensor-core Gram (DMMA), constant-memory operands (LDC) —
and shows what is NOT there (no UBLKCP / UTMALDG bulk copies, no tcgen05 / UTCxMMA: DESIGN.md §5 says why). Round 2: the
system-scope strong stores of the solve kernels are the NVSwitch multicast stores (PTX multimem.st, kernels_dense.cuh).

    python tools/sass_mix.py [--match half_step_kernel --match normalize_gram_mma] [--out profiles/r01t_sass_mix.json]
"""
import argparse
import collections
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(ROOT, "rcppml_b200", "lib", "RcppML_gpu.so"))
ap.add_argument("--match", action="append", default=[])
ap.add_argument("--out", default="")
a = ap.parse_args()

sass = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
names = {}
try:
    filt = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True)
    for mangled, demangled in zip(re.findall(r"Function : (\S+)", sass), filt.stdout.splitlines()):
        names[mangled] = demangled.replace("(int)", "")
except FileNotFoundError:
    pass

KEYS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "DFMA", "DADD", "DMMA", "HMMA", "MUFU", "LDG", "STG", "LDS", "STS",
        "LDC", "SHFL", "BAR", "ATOM", "RED", "UBLKCP", "UTMALDG", "UTCHMMA", "UTCQMMA", "SYNCS"]
rows = []
for block in sass.split("Function : ")[1:]:
    mangled = block.split()[0]
    name = names.get(mangled, mangled)
    if a.match and not any(m in name for m in a.match):
        continue
    ops = collections.Counter()
    wide = collections.Counter()
    n = 0
    for line in block.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        base = op.split(".")[0]
        ops[base] += 1
        if base in ("LDG", "STG", "LDS", "STS") and ".128" in op:
            wide[base + ".128"] += 1
        # multimem.st.relaxed.sys (NVSwitch multicast store through the multicast alias of a factor) has no mnemonic of its
        # own in this disassembler: it shows as a system-scope strong store
        if base == "STG" and ".STRONG.SYS" in op:
            wide["STG.STRONG.SYS (multimem.st / peer flag stores)"] += 1
    row = {"kernel": re.sub(r"\(.*", "", name)[:140], "sass_instructions": n}
    row.update({k: ops[k] for k in KEYS if ops[k]})
    row.update(wide)
    rows.append(row)
seen, uniq = set(), []
for r in rows:                                   # header-defined kernels are compiled into several objects: keep one
    key = json.dumps(r, sort_keys=True)
    if key not in seen:
        seen.add(key)
        uniq.append(r)
rows = uniq
rows.sort(key=lambda r: -r["sass_instructions"])
for r in rows[:40]:
    print(json.dumps(r))
if a.out:
    with open(a.out, "w") as f:
        json.dump({"library": os.path.relpath(a.lib, ROOT), "how": "cuobjdump -sass, opcode counts per kernel (static)",
                   "kernels": rows}, f, indent=1)
