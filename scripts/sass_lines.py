"""Attribute an ncu source-page (SASS) CSV to CUDA source lines using nvdisasm -gi line info.

usage: python scripts/sass_lines.py <libcastep.so> <ncu_source_page.csv> <kernel substring> [top N]
Prints, per source line: warp instructions executed and stall samples (share of kernel)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

so, src_csv, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = []   # the library holds one cubin per source file: search all of them for the kernel
for cubin in sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis += subprocess.run(["nvdisasm", "-gi", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout.split("\n")
# offset -> (innermost line, outermost line in chain)
off2line = {}
infunc = False
cur = None
chain = []
for ln in dis:
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        infunc = kern in ln
        cur, chain = None, []
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f, l, rest = m.group(1), int(m.group(2)), m.group(3)
        if "inlined at" in rest:
            m2 = re.search(r'inlined at "([^"]+)", line (\d+)', rest)
            chain = [(os.path.basename(f), l), (os.path.basename(m2.group(1)), int(m2.group(2)))]
        else:
            chain = [(os.path.basename(f), l)]
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m:
        off2line[int(m.group(1), 16)] = (list(chain), m.group(2).strip())

rows = list(csv.reader(open(src_csv)))
# find the block for the kernel
start = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name" and kern.replace("ILb1", "").split("E")[0][-10:] in "".join(r[1:2]).replace("::", ""):
        pass
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
blocks = []
for k, hi in enumerate(hdr_idx):
    end = hdr_idx[k + 1] - 1 if k + 1 < len(hdr_idx) else len(rows)
    blocks.append((rows[hi - 1][1] if hi > 0 else "", rows[hi], rows[hi + 1:end]))
name, hdr, data = blocks[0]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(data[0][ia], 16)
per_line = collections.defaultdict(lambda: [0, 0])
per_outer = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r in data:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        continue
    off = int(r[ia], 16) - base
    n, s = int(r[ii] or 0), int(r[isamp] or 0)
    tot_i += n; tot_s += s
    ch, _ = off2line.get(off, ([("?", 0)], ""))
    inner = ch[0] if ch else ("?", 0)
    outer = ch[-1] if ch else ("?", 0)
    per_line[inner][0] += n; per_line[inner][1] += s
    per_outer[outer][0] += n; per_outer[outer][1] += s
print("kernel:", name, " total warp-instructions %d, samples %d" % (tot_i, tot_s))
srcs = {}
def text(f, l):
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
for title, d in (("innermost line", per_line), ("outermost (kernel-level) line", per_outer)):
    print("\n== by %s" % title)
    for (f, l), (n, s) in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-16s %4d  inst %5.1f%%  stall-samples %5.1f%%   %s" % (f, l, 100.0 * n / max(tot_i, 1), 100.0 * s / max(tot_s, 1), text(f, l)))
