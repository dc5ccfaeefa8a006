#!/usr/bin/env python
"""Turn ncu artefacts brought back from a gpurun call into the text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv          > profiles/rNN_launches.md
    python profiles/summarize.py kernel   gpurun_out/prof.ncu-rep SAMPLES  > profiles/rNN_kernel.md
    python profiles/summarize.py configs  profiles/rNN_config_results.jsonl > table for rNN_config_results.md

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum`; `kernel` reads a `--set full` report
through `ncu -i ... --page raw/source --csv` (no GPU needed) and prints the roofline-relevant counters, the
executed-instruction mix by SASS opcode and the stall reasons by kernel region.
"""
import collections
import csv
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
       "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
       "l1tex__t_requests_pipe_tex_mem_texture.sum", "l1tex__t_sectors_pipe_tex_mem_texture.sum",
       "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
       "l1tex__data_pipe_tex_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__cycles_elapsed.max"]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "selected", "not_selected", "no_instructions", "math_pipe_throttle",
          "mio_throttle", "lg_throttle", "barrier", "dispatch_stall", "branch_resolving"]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"| share | launches | avg us | total us | kernel |\n|---:|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {100 * sum(v) / tot:.1f}% | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e3:.1f} | `{k[:120]}` |")
    print(f"\n{len(rows)} launches, {tot / 1e3:.1f} us of kernel time (ncu-serialised, cold cache: compare shares, not absolutes)")


def ncu_csv(rep, page, extra=()):
    """Rows of one page of a report.  `rep` is a .ncu-rep file, or the prefix of pages already exported on the GPU box
    (`ncu -i X.ncu-rep --page raw --csv > X.raw.csv`, same for `source`: the reports are too big to bring back)."""
    import os
    if not rep.endswith(".ncu-rep") and os.path.exists(f"{rep}.{page}.csv"):
        return list(csv.reader(open(f"{rep}.{page}.csv")))
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def kernel(rep, samples):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## `{name}`\n\n| counter | value | unit |\n|---|---:|---|")
        for w in RAW:
            if w in hdr:
                i = hdr.index(w)
                print(f"| {w} | {r[i]} | {units[i]} |")
        for s in STALLS:
            w = "smsp__pcsamp_warps_issue_stalled_" + s
            if w in hdr:
                print(f"| stall samples: {s} | {r[hdr.index(w)]} | |")
        dr, dw = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic = dr * mult[ur] + dw * mult[uw]
        print(f"\nDRAM traffic {traffic / 1e6:.1f} MB per launch; {traffic / samples:.2f} B per sample ({samples} samples)\n")
        base = re.search(r"(\w+)\s*[<(]", name)
        src = ncu_csv(rep, "source", ("--kernel-name", "regex:" + (base.group(1) if base else ".")))
        if len(src) < 3:
            continue
        h = src[1]
        si, ei, wi = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        ops, tot = collections.Counter(), 0
        for q in src[2:]:
            try:
                n = int(q[ei])
            except (ValueError, IndexError):
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", q[si].strip())
            if m:
                ops[m.group(2).split(".")[0]] += n
                tot += n
        executed = float(r[hdr.index("smsp__inst_executed.sum")])      # the source page can list a kernel twice
        print(f"Executed warp instructions: {executed:.0f} = {executed / samples:.2f} per sample\n\n| opcode | share | per sample |\n|---|---:|---:|")
        for op, n in ops.most_common(16):
            print(f"| {op} | {100 * n / tot:.1f}% | {n / tot * executed / samples:.3f} |")
        print()


def configs(path):
    """Table of benchmarks/run_configs.py JSON lines."""
    import json
    print("| config | ms per call | Gsamples/s | frames/s | B/sample | HBM frac |\n|---|---:|---:|---:|---:|---:|")
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if "gsamples_per_s" in d:
            print(f"| {d['config']} | {d['ms']:.4g} | {d['gsamples_per_s']:.3g} | {d['frames_per_s']:.4g} | {d['bytes_per_sample']} | {d['hbm_frac']:.3f} |")
        else:
            print(f"| {d['config']} | {d['ms']:.4g} | ({d['gvoxels_per_s']:.3g} Gvoxels/s, {d['tflops_layer2']:.3g} TFLOP/s layer 2) | | | |")
        extra = {k: v for k, v in d.items() if k.startswith("mlp_")}
        if extra:
            print("\nInside that step: " + ", ".join(f"{k} = {v:.3g}" for k, v in extra.items()) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "configs":
        configs(sys.argv[2])
    else:
        kernel(sys.argv[2], float(sys.argv[3]))
