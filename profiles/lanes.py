#!/usr/bin/env python
"""Per CUDA source line: executed warp-instructions and average active lanes (divergence) of one kernel, from an
.ncu-rep captured with `--import-source on` (binary built with -lineinfo).  Uses ncu's own source correlation.

    python profiles/lanes.py <report.ncu-rep> [top N | file-name filter] [ncu kernel-name filter]
"""
import collections
import csv
import subprocess
import sys


def load(rep, name_filter=None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if name_filter:
        cmd += ["--kernel-name", name_filter]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    cur, hdr, by = None, None, collections.OrderedDict()
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            ie, it = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            n, t = int(r[ie] or 0), int(r[it] or 0)
            if n:
                key = (cur, int(r[0]))
                old = by.get(key, (0, 0))
                by[key] = (old[0] + n, old[1] + t)
    return by


def main(rep, sel="40", name_filter=None):
    by = load(rep, name_filter)
    tot = sum(v[0] for v in by.values())
    tot_t = sum(v[1] for v in by.values())
    print("warp-instructions %d, thread-instructions %d, average active lanes %.1f" % (tot, tot_t, tot_t / tot))
    if sel.isdigit():
        items = sorted(by.items(), key=lambda kv: -kv[1][0])[:int(sel)]
    else:
        items = sorted((kv for kv in by.items() if sel in kv[0][0]), key=lambda kv: kv[0][1])
    for (f, ln), (n, t) in items:
        print("%-26s %12d %5.1f%%  lanes %4.1f" % ("%s:%d" % (f, ln), n, 100.0 * n / tot, t / max(n, 1)))
    byf = collections.defaultdict(lambda: [0, 0])
    for (f, _ln), (n, t) in by.items():
        byf[f][0] += n
        byf[f][1] += t
    print("-- by file")
    for f, (n, t) in sorted(byf.items(), key=lambda kv: -kv[1][0]):
        print("%-26s %12d %5.1f%%  lanes %4.1f" % (f, n, 100.0 * n / tot, t / max(n, 1)))


if __name__ == "__main__":
    main(*sys.argv[1:4])
