#!/usr/bin/env python
"""Per-source-line digest of an ncu SASS page.

ncu's CSV export of the source page carries metrics only for the SASS view, without the line
correlation, so this joins it (by instruction offset inside the kernel) with
`nvdisasm -g` of the same library.

usage: ncu_lines.py <source_sass.csv> <kernel mangled-name substring> [lib.so] [top N]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(lib, kernel_sub):
    """offset -> (file, line) for the first function whose mangled name contains kernel_sub."""
    # the library is linked from several objects whose cubins share one name (cuobjdump -xelf would overwrite
    # them): `lib` may be a directory of .o files (boxer_b200/_native.py's object directory), one cubin each
    files = [os.path.join(lib, f) for f in sorted(os.listdir(lib)) if f.endswith(".o")] if os.path.isdir(lib) else [lib]
    cubins = []
    for f in files:
        tmp = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(f)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubins += [os.path.join(tmp, c) for c in os.listdir(tmp) if c.endswith(".cubin")]
    table = {}
    for f in cubins:
        if kernel_sub.encode() not in open(f, "rb").read():
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", f], capture_output=True, text=True).stdout
        cur, inside, done = None, False, False
        for ln in txt.splitlines():
            if ln.startswith("//----") and ".text." in ln:
                if inside:
                    done = True
                    break
                inside = kernel_sub in ln
                cur = None
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                table[int(m.group(1), 16)] = (cur, m.group(2).strip())
        if done or table:
            break
    return table


def main():
    path, ksub = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.environ.get("TMPDIR", "/tmp"), "boxattn_b200_build")
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(lib, ksub)
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    body = rows[hdr_i + 1:]
    base = int(body[0][0], 16)
    agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot_inst = tot_samp = 0
    mism = 0
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for r in body:
        off = int(r[0], 16) - base
        inst = int(r[col["Instructions Executed"]] or 0)
        samp = int(r[col["# Samples"]] or 0)
        loc, sass = table.get(off, (None, None))
        if sass is None or sass.split()[0].lstrip("@!P0123456789T ") [:3] != r[1].split(";")[0].strip().lstrip("@!P0123456789T ")[:3]:
            mism += 1
        a = agg[loc]
        a[0] += inst
        a[1] += samp
        a[2] += 1
        for s in stall_cols:
            v = int(r[col[s]] or 0)
            if v:
                a[3][s] += v
        tot_inst += inst
        tot_samp += samp
    print(f"# {len(body)} SASS instructions, {tot_inst} warp-instructions executed, {tot_samp} stall samples, "
          f"{mism} offset/opcode mismatches vs local disassembly")
    print(f"# {'file:line':34s} {'inst%':>6s} {'samp%':>6s} {'#sass':>5s}  top stalls")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        name = f"{loc[0]}:{loc[1]}" if loc else "?"
        stalls = ", ".join(f"{k[6:]} {100.0 * v / max(a[1], 1):.0f}%" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
        print(f"{name:36s} {100.0 * a[0] / tot_inst:6.2f} {100.0 * a[1] / tot_samp:6.2f} {a[2]:5d}  {stalls}")


if __name__ == "__main__":
    main()
