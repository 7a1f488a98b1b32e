#!/usr/bin/env python
"""Turns the raw outputs of scripts/gpu_evidence_r02.sh (gpurun_out/) into the tracked round-2 evidence under profiles/.

    python scripts/summarise_r02.py        (in the build container: needs ncu to read the .ncu-rep files, no GPU)
"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
PROF = os.path.join(REPO, "profiles")
ALG = {"phase1": 184 * 262144, "phase2": 352 * 163840, "ragged": 352 * 196604}

KEYS = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg ', 'smsp__inst_executed.sum ',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum ',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tc',
        'sm__pipe_tensor', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__average_warps_issue_stalled',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_summary(rep, units, unit_name):
    rows = ncu_csv(rep, "raw")
    hdr, un, data = rows[0], rows[1], rows[2:]
    lines = []
    d0 = dict(zip(hdr, data[0]))
    lines.append("kernel: %s   grid %s x block %s" % (d0.get("Kernel Name"), d0.get("Grid Size"), d0.get("Block Size")))
    for i, h in enumerate(hdr):
        if any(h.startswith(k.strip()) if k.endswith(' ') and h == k.strip() else (k.strip() in h and not k.endswith(' '))
               for k in KEYS) and "pct_of_peak_sustained_elapsed" not in h.replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ""):
            vals = [r[i] for r in data]
            if all(v in ("", "n/a") for v in vals):
                continue
            lines.append("%-92s %-10s %s" % (h, un[i], " ".join(vals)))
    try:
        inst = float(d0["smsp__inst_executed.sum"])
        lines.append("warp instructions per %s: %.1f (%d %ss per launch)" % (unit_name, inst / units, units, unit_name))
    except Exception:
        pass
    return "\n".join(lines)


def source_summary(rep, units):
    rows = ncu_csv(rep, "source")
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) > 6:
            data.append(r)
    ie, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    tot = sum(int(r[ie] or 0) for r in data)
    stot = sum(int(r[ist] or 0) for r in data)
    ops, st = collections.Counter(), collections.Counter()
    stalls = collections.Counter()
    for r in data:
        s = re.sub(r"^\s*@!?U?P\w+\s+", "", r[isrc].strip())
        op = s.split()[0].rstrip(";") if s.split() else "?"
        base = op.split(".")[0]
        if base == "IMAD" and "MOV" in op:
            base = "IMAD.MOV"
        if base in ("F2F", "MUFU", "SHFL", "LDG", "STG", "STS", "LDS"):
            base = ".".join(op.split(".")[:2])
        ops[base] += int(r[ie] or 0)
        st[base] += int(r[ist] or 0)
        for i in stall_cols:
            if r[i]:
                stalls[hdr[i]] += int(r[i])
    lines = ["executed warp instructions %d (%.1f per unit), stall samples %d" % (tot, tot / units, stot),
             "stall reasons (share of samples): " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, sum(stalls.values())))
                                                               for k, v in stalls.most_common(8)),
             "== executed instructions by opcode (per unit, share, share of stall samples)"]
    for k, v in ops.most_common(32):
        lines.append("%-16s %8.1f %5.1f%%  stall %5.1f%%" % (k, v / units, 100.0 * v / tot, 100.0 * st[k] / max(stot, 1)))
    return "\n".join(lines)


def main():
    os.makedirs(PROF, exist_ok=True)
    # bench lines
    for src, dst in [("r02_bench.json", "r02_bench_line.json"), ("r02_bench_k20.json", "r02_bench_line_steps20.json"),
                     ("r02_bench_ref.json", "r02_bench_line_reference_arm.json")]:
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            line = [l for l in open(p).read().splitlines() if l.startswith("{")][-1]
            with open(os.path.join(PROF, dst), "w") as f:
                f.write(json.dumps(json.loads(line), indent=1) + "\n")
    for src in ["r02_launches_bench.csv", "r02_pytest_gpu.log", "r02_smoke.log"]:
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            shutil.copyfile(p, os.path.join(PROF, src))
    # steady-state traffic
    traffic = {"_comment": "Steady-state DRAM bytes per launch of the step kernel: ncu --cache-control none (no flush between "
                           "launches) over 24 consecutive launches of bench.py's rotation over 6 world sets (scripts/"
                           "gpu_evidence_r02.sh); every byte a step writes is evicted to DRAM during later steps, so reads + "
                           "writes here are what a step really moves.  algorithmic = SURVEY §8(d) figure used for roofline.achieved.",
               "ca_step_kernel": {}}
    for wl in ("phase1", "phase2", "ragged"):
        p = os.path.join(OUT, "r02_traffic_%s.csv" % wl)
        if not os.path.exists(p):
            continue
        rows = [r for r in csv.reader(open(p)) if len(r) > 5]
        hdr = rows[0]
        im, iv = hdr.index("Metric Name"), hdr.index("Metric Value")
        agg = collections.defaultdict(list)
        for r in rows[1:]:
            agg[r[im]].append(float(r[iv].replace(",", "")))
        rd = sum(agg["dram__bytes_read.sum"]) / len(agg["dram__bytes_read.sum"])
        wr = sum(agg["dram__bytes_write.sum"]) / len(agg["dram__bytes_write.sum"])
        traffic["ca_step_kernel"][wl] = {
            "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "l2_bytes_per_launch": sum(agg["lts__t_bytes.sum"]) / max(1, len(agg["lts__t_bytes.sum"])),
            "launches": len(agg["dram__bytes_read.sum"]), "algorithmic_bytes_per_launch": ALG[wl],
            "ratio_to_algorithmic": (rd + wr) / ALG[wl],
            "ncu_us_per_launch_serialised": sum(agg["gpu__time_duration.sum"]) / len(agg["gpu__time_duration.sum"]) / 1e3,
            "source": "profiles/r02_traffic_%s.csv" % wl}
        shutil.copyfile(p, os.path.join(PROF, "r02_traffic_%s.csv" % wl))
    with open(os.path.join(PROF, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    # full captures
    chunks = {"phase1": 8192, "phase2": 5462, "ragged": 10923}
    for wl, n in chunks.items():
        rep = os.path.join(OUT, "r02c_step_%s.ncu-rep" % wl)
        if os.path.exists(rep):
            with open(os.path.join(PROF, "r02_ncu_full_step_%s.txt" % wl), "w") as f:
                f.write("# ncu --set full --clock-control none --import-source on, one launch of the step kernel, bench.py "
                        "--workload %s --streams 1\n" % wl)
                f.write(raw_summary(rep, n, "chunk") + "\n\n" + source_summary(rep, n) + "\n")
    rep = os.path.join(OUT, "r02c_predict.ncu-rep")
    if os.path.exists(rep):
        with open(os.path.join(PROF, "r02_ncu_full_predict_kernel.txt"), "w") as f:
            f.write("# ncu --set full, cap::predict_kernel, 163 840 rows at M = 9 (scripts/predict_probe.py one 9 163840 0)\n")
            f.write(raw_summary(rep, 1280, "tile") + "\n\n" + source_summary(rep, 1280) + "\n")
    # SASS histogram of the shipped library
    hist = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "sass_histogram.py")], stdout=subprocess.PIPE, text=True).stdout
    with open(os.path.join(PROF, "r02_sass_histogram.txt"), "w") as f:
        f.write("# static SASS opcode histogram per kernel of rl_collision_avoidance_b200/libcastep.so (scripts/sass_histogram.py)\n" + hist)
    # timelines
    for src in sorted(os.listdir(OUT)):
        if src.startswith("timeline") and src.endswith(".jsonl"):
            shutil.copyfile(os.path.join(OUT, src), os.path.join(PROF, "r02_" + src))
    print("profiles/ updated")


if __name__ == "__main__":
    main()
