#!/usr/bin/env python
"""Joins the SASS-level source page of an .ncu-rep with nvdisasm's line / inlining info and prints, per source function of the
kernel, the executed warp instructions, their share, the average active threads and the stall samples.

    python tools/ncu_by_function.py report.ncu-rep roger_b200/lib/libroger_b200.so svat_api '<mangled kernel substring>' [launch index]

Runs where ncu / cuobjdump / nvdisasm are installed; no GPU needed."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(so, unit, kernel):
    """offset -> inline stack [(file, line), ...] innermost first"""
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.startswith(unit) and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    table, stack, pending, on = {}, [], [], False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            on = kernel in ln
            stack, pending = [], []
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            pending.append((m.group(1), int(m.group(2)), m.group(3), int(m.group(4)) if m.group(4) else None))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if pending:
                # the block lists frames innermost first; the last line (no "inlined at") is the kernel-level frame
                stack = [(f, l) for f, l, _, _ in pending]
                pending = []
            table[int(m.group(1), 16)] = (stack, m.group(2).strip())
    return table


def function_ranges(path):
    """[(first line, name)] of the function definitions of a source file (good enough for this code base's style)"""
    out = []
    for n, ln in enumerate(open(path), 1):
        m = re.match(r"\s*(?:template <[^>]*>\s*)?(?:RD|__device__|__host__ __device__|static|inline|ROGER_POW_FN)[\w\s:<>,*&]*?\b(\w+)\s*\([^;]*$", ln)
        if m and not ln.strip().startswith("//") and m.group(1) not in ("if", "for", "while", "return", "asm"):
            out.append((n, m.group(1)))
    return out


def name_of(ranges, line):
    cur = "?"
    for n, name in ranges:
        if n <= line:
            cur = name
        else:
            break
    return cur


def main():
    rep, so, unit, kernel = sys.argv[1:5]
    launch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    table = line_table(so, unit, kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    s = starts[launch]
    e = starts[launch + 1] if launch + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    body = rows[s + 2:e]
    ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(body[0][ia], 16)
    ranges = {}
    by_fn = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    by_leaf = collections.defaultdict(lambda: [0, 0, 0])
    tot = 0
    for r in body:
        off = int(r[ia], 16) - base
        n, t, smp = int(r[ii] or 0), int(r[it] or 0), int(r[isamp] or 0)
        tot += n
        stack, sass = table.get(off, ([], "?"))
        # the outermost frame inside svat_device.cuh / sas_device.cuh names the process; the innermost frame names the leaf
        top, leaf = "kernel", "kernel"
        for f, l in stack:
            if f not in ranges and os.path.exists(f):
                ranges[f] = function_ranges(f)
            nm = name_of(ranges.get(f, []), l)
            if f.endswith("_device.cuh") and nm not in ("cell_step", "?"):
                top = nm
        if stack:
            f, l = stack[0]
            leaf = os.path.basename(f) + ":" + name_of(ranges.get(f, []), l)
        a = by_fn[top]
        a[0] += n; a[1] += t; a[2] += smp
        for i, h in stall_cols:
            if r[i] not in ("", "0"):
                a[3][h] += int(r[i])
        b = by_leaf[leaf]
        b[0] += n; b[1] += t; b[2] += smp
    print(f"total warp instructions {tot:.4g}")
    tsmp = sum(v[2] for v in by_fn.values()) or 1
    print("%-28s %8s %7s %7s %8s  top stalls" % ("process (outermost frame)", "inst %", "thr/in", "smp %", "inst"))
    for k, v in sorted(by_fn.items(), key=lambda kv: -kv[1][2]):
        st = ", ".join(f"{h[6:]} {100 * c / max(v[2], 1):.0f}%" for h, c in v[3].most_common(3))
        print("%-28s %8.1f %7.1f %7.1f %8.3g  %s" % (k, 100 * v[0] / tot, v[1] / max(v[0], 1), 100 * v[2] / tsmp, v[0], st))
    print()
    print("%-40s %8s %7s %7s" % ("leaf (innermost frame)", "inst %", "thr/in", "smp %"))
    for k, v in sorted(by_leaf.items(), key=lambda kv: -kv[1][2])[:25]:
        print("%-40s %8.1f %7.1f %7.1f" % (k, 100 * v[0] / tot, v[1] / max(v[0], 1), 100 * v[2] / tsmp))


if __name__ == "__main__":
    main()
