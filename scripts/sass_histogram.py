"""Static SASS opcode histogram per kernel of libcastep.so (cuobjdump -sass), written as text.

usage: python scripts/sass_histogram.py [<so>] [kernel substring ...] > profiles/rNN_sass_histogram.txt
Without substrings: the production kernels.  Counts are static instructions of the compiled kernel, not executed ones
(the executed count per launch is in the ncu captures: smsp__inst_executed.sum)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else os.path.join(
    ROOT, "rl_collision_avoidance_b200", "libcastep.so")
subs = [a for a in sys.argv[1:] if not a.endswith(".so")] or [
    "ca_step_kernelILi4ELi7ELb0", "ca_step_kernelILi10ELi5ELb0", "ca_step_stream_kernelILi4ELi7ELb0",
    "ca_world_kernelILb1", "ca_world_kernelILb0", "generate_scenarios_kernel", "ga3c_record_kernel",
    "ga3c_gather_kernel", "predict_kernel", "train_", "plan_scatter_kernel", "lstm_cell_fwd", "lstm_cell_bwd"]

txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
cur = None
hist = collections.OrderedDict()
for ln in txt.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        hist[cur][m.group(1)] += 1

GROUPS = [("fp64", r"^(DFMA|DADD|DMUL|DSETP|DMNMX|MUFU\.(RCP64H|RSQ64H))"), ("fp32", r"^(FFMA|FADD|FMUL|FSETP|FMNMX|FSEL|MUFU)"),
          ("convert", r"^(F2F|F2I|I2F|F2FP|FRND|I2FP)"), ("shuffle/vote", r"^(SHFL|VOTE|REDUX|MATCH)"),
          ("global ld/st", r"^(LDG|STG|LD\.|ST\.|ATOMG|RED)"), ("shared ld/st", r"^(LDS|STS|LDSM|STSM)"),
          ("constant", r"^(LDC|LDCU)"), ("bulk copy / TMA", r"^(UBLKCP|UBLKPF|UTMA|SYNCS|ACQBULK|UBLKRED)"),
          ("tensor (tcgen05)", r"^(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTCCP|UTCATOMSWS)"),
          ("integer/logic/move", r"^(IADD|IMAD|LOP3|SHF|SEL|ISETP|MOV|LEA|PRMT|POPC|FLO|IABS|IMNMX|VIADD|VIMNMX|PLOP3|P2R|R2P|S2R|CS2R|SGXT|BMSK|UMOV|UIADD|ULOP|USHF|UIMAD|ULEA|USEL|UISETP|UPLOP|R2UR|S2UR|UFLO|UPOPC|UPRMT|USGXT|UF2F|UI2F)"),
          ("control", r"^(BRA|EXIT|BSSY|BSYNC|WARPSYNC|BAR|NOP|CALL|RET|BREAK|YIELD|DEPBAR|MEMBAR|FENCE|ERRBAR|CCTL|NANOSLEEP|ELECT|UCGABAR|ACQSHMINIT|UGETNEXTWORKID)")]

for k, h in hist.items():
    if not any(s in k for s in subs):
        continue
    tot = sum(h.values())
    print(f"== {k}  ({tot} static instructions)")
    rest = collections.Counter(h)
    for name, rx in GROUPS:
        ops = {o: c for o, c in rest.items() if re.match(rx, o)}
        for o in ops:
            del rest[o]
        n = sum(ops.values())
        if n:
            detail = ", ".join(f"{o} {c}" for o, c in sorted(ops.items(), key=lambda t: -t[1])[:12])
            print(f"  {name:20s} {n:6d}  {100.0 * n / tot:5.1f} %   {detail}")
    n = sum(rest.values())
    if n:
        print(f"  {'other':20s} {n:6d}  {100.0 * n / tot:5.1f} %   " + ", ".join(f"{o} {c}" for o, c in rest.most_common(12)))
    print()
