#!/usr/bin/env python
"""Per-kernel SASS evidence of dune_hdd_b200/libhdd_b200.so, without rebuilding: registers, stack (local memory) and shared
memory from `cuobjdump -res-usage`, and opcode counts from `cuobjdump -sass` for the instructions that show how a kernel
moves its data on sm_100a:

    UBLKCP                 cp.async.bulk (TMA bulk copy engine: global -> shared in the SpMV ring, shared -> global in the
                           write-out of staged assembly rows)
    SYNCS                  mbarrier arrive / expect_tx / try_wait (the bulk copies' completion mechanism)
    LDG.256 / STG.256      256-bit global loads / stores (LDG.E.ENL2.256 / STG.E.ENL2.256, one full 32-byte sector per lane)
    LDG / STG              all global loads / stores
    LDL / STL              local-memory traffic (spills, dynamically indexed per-thread arrays) - should be 0 in hot kernels
    DFMA / DMUL / DADD     fp64 arithmetic
    MUFU                   special-function unit (reciprocal / rsqrt seeds of fp64 division and sqrt)
    BRX / CALL             indexed branches (expression interpreter dispatch) / calls (slow paths of sin / cos / pow)

Usage:  python tools/sass_summary.py [--out profiles/r02_sass_summary] [--all]
writes <out>.json (every kernel) and <out>.md (the hot kernels, or all with --all).
"""
import argparse
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dune_hdd_b200", "libhdd_b200.so")
HOT = ["k_cg_spmv_tma", "k_cg_update", "k_cg_direction", "k_assemble_lhs", "k_assemble_rows", "k_assemble_p1_closed", "k_assemble_q2", "k_indicators", "k_vertex_means",
       "k_mg_pre", "k_mg_post", "k_mg_up", "k_mg_restrict", "k_mg_prolong_add", "k_dg_restrict", "k_dg_prolong_dot",
       "k_freeze", "k_rhs_tensor", "k_fill_csr", "k_error_norms", "k_vertex_galerkin", "k_rap", "k_cube_fill"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def short(name):
    name = re.sub(r"hdd::\(anonymous namespace\)::|\(anonymous namespace\)::|hdd::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sass_summary"))
    ap.add_argument("--all", action="store_true")
    args = ap.parse_args()
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage, cur, arch = {}, None, None
    for line in res.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur and arch and arch.startswith("sm_100"):
            usage[cur] = {"registers": int(m.group(1)), "stack_bytes": int(m.group(2)), "static_smem_bytes": int(m.group(3)),
                          "local_bytes": int(m.group(4))}
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    ops, cur, arch = collections.defaultdict(collections.Counter), None, None
    for line in sass.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1) if arch and arch.startswith("sm_100") else None
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = ops[cur]
        c["instructions"] += 1
        base = op.split(".")[0]
        for key in ("UBLKCP", "SYNCS", "LDG", "STG", "LDL", "STL", "DFMA", "DMUL", "DADD", "MUFU", "BRX", "CALL", "LDS", "STS",
                    "SHFL", "ATOMG", "RED", "LDC"):
            if base == key:
                c[key] += 1
        if base == "LDG" and ".256" in op:
            c["LDG.256"] += 1
        if base == "STG" and ".256" in op:
            c["STG.256"] += 1
        if base == "LDG" and ".128" in op:
            c["LDG.128"] += 1
        if base == "STG" and ".128" in op:
            c["STG.128"] += 1
    names = demangle(sorted(set(usage) | set(ops)))
    table = {}
    for mangled, nice in names.items():
        entry = dict(usage.get(mangled, {}))
        entry.update(ops.get(mangled, {}))
        table[short(nice)] = entry
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".json", "w") as f:
        json.dump({"library": os.path.relpath(SO, ROOT), "arch": "sm_100a", "kernels": table}, f, indent=1, sort_keys=True)
    cols = ["registers", "stack_bytes", "instructions", "UBLKCP", "SYNCS", "LDG", "LDG.256", "LDG.128", "STG", "STG.256", "STG.128",
            "LDL", "STL", "DFMA", "DMUL", "DADD", "MUFU", "BRX", "CALL"]
    lines = ["# SASS summary of `dune_hdd_b200/libhdd_b200.so` (sm_100a), written by `tools/sass_summary.py`", "",
             "`stack_bytes` is the per-thread local-memory frame ptxas reserves (slow paths of `sin` / `cos` / `pow` and the generic",
             "expression interpreter keep a frame even when the hot path never touches it); `LDL` / `STL` are the instructions that",
             "actually access it.", "",
             "| kernel | " + " | ".join(cols) + " |", "|---|" + "---|" * len(cols)]
    for name in sorted(table):
        if not args.all and not any(name.startswith(h) for h in HOT):
            continue
        e = table[name]
        lines.append("| `%s` | " % name + " | ".join(str(e.get(c, 0)) for c in cols) + " |")
    with open(args.out + ".md", "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote %s.json (%d kernels) and %s.md" % (args.out, len(table), args.out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
