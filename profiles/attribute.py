#!/usr/bin/env python
"""Attribute the executed warp-instructions of one kernel in an .ncu-rep to source lines and opcodes.

    python profiles/attribute.py <report.ncu-rep> <object-or-cubin with -lineinfo> <mangled kernel name substring> \
                                 [ncu kernel-name filter, e.g. regex:step_kernel, when the report holds several kernels]

Joins ncu's SASS page (per-instruction `Instructions Executed`) with nvdisasm's line table of the same cubin.
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os


def main(rep, obj, kernel, name_filter=None):
    tmp = tempfile.mkdtemp()
    if obj.endswith(".o") or obj.endswith(".so"):
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    else:
        cubins = [obj]
    instrs = None
    for cubin in cubins:
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        sections = re.split(r"\n//-+ \.text\.", text)
        for sec in sections:
            head = sec.split("\n", 1)[0]
            if kernel in head:
                cur, instrs = None, []
                for ln in sec.split("\n"):
                    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                    if m:
                        cur = (m.group(1).split("/")[-1], int(m.group(2)))
                        continue
                    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                    if m:
                        instrs.append((int(m.group(1), 16), m.group(2), cur))
                break
        if instrs:
            break
    if not instrs:
        sys.exit("kernel not found in " + obj)
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", name_filter] if name_filter else [])
    src = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
    stall_cols = [(c, hdr.index(c)) for c in ("stall_long_sb", "stall_wait", "stall_short_sb", "stall_no_inst",
                                              "stall_barrier", "stall_math") if c in hdr]
    ex, stalls = [], []
    for r in rows[2:]:
        if len(r) <= ie or not r[ia].startswith("0x"):
            if ex:
                break
            continue
        ex.append(int(r[ie]))
        stalls.append([int(r[i] or 0) for _c, i in stall_cols])
    assert len(ex) == len(instrs), (len(ex), len(instrs))
    # stall samples are attributed to the instruction that WAITS (the consumer), so a line's long-scoreboard
    # count says where a load's latency was exposed, not where the load was issued
    for ci, (cname, _i) in enumerate(stall_cols):
        tot_s = sum(st[ci] for st in stalls)
        if not tot_s:
            continue
        per = collections.Counter()
        for st, (_off, op, loc) in zip(stalls, instrs):
            per[loc] += st[ci]
        print("-- %s samples by source line (total %d)" % (cname, tot_s))
        for loc, n in per.most_common(12):
            print("%-28s %8d %5.1f%%" % ("%s:%d" % loc if loc else "?", n, 100.0 * n / tot_s))
    by_line, by_op, tot = collections.Counter(), collections.Counter(), 0
    for n, (_off, op, loc) in zip(ex, instrs):
        by_line[loc] += n
        tot += n
        tok = op.split()
        by_op[(tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]] += n
    print("total warp-instructions: %d" % tot)
    print("-- by source line")
    for loc, n in by_line.most_common(40):
        print("%-28s %11d %5.1f%%" % ("%s:%d" % loc if loc else "?", n, 100.0 * n / tot))
    by_file = collections.Counter()
    for loc, n in by_line.items():
        by_file[loc[0] if loc else "?"] += n
    print("-- by file")
    for f, n in by_file.most_common():
        print("%-28s %11d %5.1f%%" % (f, n, 100.0 * n / tot))
    print("-- by opcode")
    for op, n in by_op.most_common(24):
        print("%-10s %11d %5.1f%%" % (op, n, 100.0 * n / tot))


if __name__ == "__main__":
    main(*sys.argv[1:5])
