#!/usr/bin/env python
"""Static SASS census of one kernel: instructions per source line and per opcode class, split into the
innermost backward-branch loop (the sample loop of step_kernel) and everything outside it.

    python profiles/sass_static.py <cubin> <mangled kernel name substring> [--lines N] [--dump]

No GPU needed (nvdisasm -g -c on a cubin built with -lineinfo); for kernels without lane divergence the
dynamic count per thread is  outside + trip_count * loop.
"""
import collections
import re
import subprocess
import sys


def parse(cubin, kernel):
    text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    for sec in re.split(r"\n//-+ \.text\.", text):
        head = sec.split("\n", 1)[0]
        if kernel not in head:
            continue
        cur, instrs, labels = None, [], {}
        for ln in sec.split("\n"):
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"^(\.L_x_\d+):", ln)
            if m:
                labels[m.group(1)] = len(instrs)
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                instrs.append((int(m.group(1), 16), m.group(2), cur))
        return head, instrs, labels
    sys.exit("kernel not found")


def opclass(op):
    tok = op.split()
    name = (tok[1] if tok[0].startswith("@") else tok[0])
    return name.split(".")[0]


def main():
    cubin, kernel = sys.argv[1], sys.argv[2]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    head, instrs, labels = parse(cubin, kernel)
    # backward branches -> loops [target, branch]
    loops = []
    for i, (_a, op, _l) in enumerate(instrs):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?`\((\.L_x_\d+)\)", op)
        if m and m.group(1) in labels and labels[m.group(1)] <= i:
            loops.append((labels[m.group(1)], i))
    print("kernel:", head.split()[0])
    print("instructions: %d, backward-branch loops: %s" % (len(instrs), [(a, b, b - a + 1) for a, b in loops]))
    if not loops:
        loops = [(0, -1)]
    big = max(loops, key=lambda ab: ab[1] - ab[0])
    inside = [k for k in range(len(instrs)) if big[0] <= k <= big[1]]
    outside = [k for k in range(len(instrs)) if not (big[0] <= k <= big[1])]
    for name, idx in (("LOOP (largest)", inside), ("OUTSIDE", outside)):
        by_op, by_line = collections.Counter(), collections.Counter()
        for k in idx:
            by_op[opclass(instrs[k][1])] += 1
            by_line[instrs[k][2]] += 1
        print("\n== %s: %d instructions" % (name, len(idx)))
        print("  by opcode:", ", ".join("%s %d" % kv for kv in by_op.most_common(30)))
        print("  by line:")
        for loc, n in by_line.most_common(nlines):
            print("    %-28s %5d" % ("%s:%d" % loc if loc else "?", n))
    if "--dump" in sys.argv:
        for k, (a, op, loc) in enumerate(instrs):
            mark = "L" if big[0] <= k <= big[1] else " "
            print("%s %04x %-24s %s" % (mark, a, "%s:%d" % loc if loc else "?", op))


if __name__ == "__main__":
    main()
