#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full, one or a few launches) into the text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep "how it was captured" > profiles/r02_x_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def instruction_mix(rep):
    """Per SASS opcode: warp instructions executed and stall samples (ncu --page source), top 18 by samples."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    mix = {}
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if not hdr or len(r) < len(hdr) - 2 or not r[0].startswith("0x"):
            continue
        d = dict(zip(hdr, r))
        toks = d["Source"].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        op = toks[0].split(".")[0] if toks else "?"
        try:
            n, smp = int(d["Instructions Executed"]), int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        e = mix.setdefault(op, [0, 0])
        e[0] += n
        e[1] += smp
    tot_n = sum(v[0] for v in mix.values()) or 1
    tot_s = sum(v[1] for v in mix.values()) or 1
    print(f"# instruction mix (SASS opcode: share of warp instructions executed, share of stall samples); {tot_n} instructions, {tot_s} samples")
    for op, (n, smp) in sorted(mix.items(), key=lambda kv: -kv[1][1])[:18]:
        print(f"  {op:12s} {100.0 * n / tot_n:6.2f} % of instructions   {100.0 * smp / tot_s:6.2f} % of samples")
    print()


def main():
    rep, how = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none ({how})")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"{'Kernel Name':90s} {d.get('Kernel Name', '')}")
        for k in KEEP:
            if k in d and d[k] != "":
                print(f"{k:90s} {d[k]:>20s} {u.get(k, '')}")
        stalls = sorted(((float(d[k].replace(',', '')), k) for k in hdr if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")
                         and "not_issued" not in k and d.get(k, "") not in ("", "n/a")), reverse=True)
        for v, k in stalls[:8]:
            print(f"{k:90s} {v:20.4f} warps per issue")
        print()
    instruction_mix(rep)


if __name__ == "__main__":
    main()
