"""Summarises an ncu report (read here, no GPU needed) into profiles/: key raw metrics per launch, instruction mix
and top stall lines.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_count [kernel-regex]"""
import csv
import io
import json
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "dram__sectors_read.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    regex = sys.argv[3] if len(sys.argv) > 3 else None
    extra = ["--kernel-name", f"regex:{regex}"] if regex else []
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"] + extra))))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = {"value": r[hdr.index(k)], "unit": units[hdr.index(k)]}
        launches.append(d)
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"] + extra))))
    mix, stalls, total_inst, total_samples = Counter(), [], 0, 0
    h = src[1] if len(src) > 2 else []
    smp = next((c for c in h if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"), None)
    if len(src) > 2 and smp and "Source" in h and "Instructions Executed" in h:
        i_src, i_ex, i_smp = h.index("Source"), h.index("Instructions Executed"), h.index(smp)
        for r in src[2:]:
            if r and r[0] == "Kernel Name":
                break
            if len(r) <= i_ex or not r[i_ex].isdigit():
                continue
            text = r[i_src].strip()
            op = (text.split()[1] if text.startswith("@") else text.split()[0]).split(".")[0]
            mix[op] += int(r[i_ex])
            total_inst += int(r[i_ex])
            total_samples += int(r[i_smp])
            stalls.append((int(r[i_smp]), int(r[i_ex]), text))
    stalls.sort(reverse=True)
    summary = {"report": rep, "launches": launches,
               "instruction_mix_pct": {op: round(100 * n / max(1, total_inst), 2) for op, n in mix.most_common(16)},
               "warp_instructions_first_launch": total_inst,
               "top_stall_lines": [{"samples_pct": round(100 * s / max(1, total_samples), 2), "executed": e, "sass": t}
                                   for s, e, t in stalls[:12]]}
    first = launches[0]
    try:
        rd, wr = float(first["dram__bytes_read.sum"]["value"]), float(first["dram__bytes_write.sum"]["value"])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        summary["dram_bytes_per_launch"] = rd * scale[first["dram__bytes_read.sum"]["unit"]] + wr * scale[first["dram__bytes_write.sum"]["unit"]]
    except Exception:
        pass
    json.dump(summary, open(out + ".json", "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "launches"}, indent=1))
    for l in launches:
        print({k: v["value"] + " " + v["unit"] for k, v in l.items() if k != "kernel"})


if __name__ == "__main__":
    main()
