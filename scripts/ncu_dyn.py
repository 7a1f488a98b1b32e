#!/usr/bin/env python
"""Dynamic (executed) instruction accounting of one kernel from an ncu report with --import-source on:
    python scripts/ncu_dyn.py report.ncu-rep [units]      (units = chunks per launch, to print per-chunk numbers)
Prints warp instructions per source line (inlined-from outermost kernel line and innermost line) and per opcode."""
import collections
import csv
import io
import re
import subprocess
import sys


def load(rep, view):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            if hdr is not None:
                break
            continue
        if r and r[0] in ("Address", "#"):
            hdr = r
            continue
        if hdr and len(r) > 6:
            data.append(r)
    return hdr, data


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hdr, data = load(rep, "sass")
    ie, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in data)
    stot = sum(int(r[ist]) for r in data)
    print("total warp instructions %d (%.1f per unit), stall samples %d" % (tot, tot / units, stot))
    ops, st = collections.Counter(), collections.Counter()
    for r in data:
        s = re.sub(r"^@!?U?P\w+\s+", "", r[isrc].strip())
        op = s.split()[0].rstrip(";")
        base = op.split(".")[0]
        if base == "IMAD" and "MOV" in op:
            base = "IMAD.MOV"
        if base in ("F2F", "MUFU"):
            base = op
        ops[base] += int(r[ie])
        st[base] += int(r[ist])
    print("== by opcode")
    for k, v in ops.most_common(40):
        print("%-16s %8.1f per unit %5.1f%%  stall %5.1f%%" % (k, v / units, 100.0 * v / tot, 100.0 * st[k] / max(stot, 1)))
    hdr, data = load(rep, "cuda,sass")
    # the cuda view lists source lines with aggregated counters
    try:
        il, ie, ist = hdr.index("#"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        isrc = hdr.index("Source")
    except ValueError:
        return
    lines = [(int(r[ie]), int(r[ist]), r[il], r[isrc].strip()) for r in data if r[ie].isdigit() and int(r[ie]) > 0]
    lines.sort(reverse=True)
    print("== by source line (innermost)")
    for n, s_, ln, src in lines[:45]:
        print("%8.1f per unit %5.1f%% stall %5.1f%%  %s: %s" % (n / units, 100.0 * n / tot, 100.0 * s_ / max(stot, 1), ln, src[:110]))


if __name__ == "__main__":
    main()
