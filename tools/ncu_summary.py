"""Summarise .ncu-rep captures (ncu --set full) as a markdown table and, for the render kernel, the JSON bench.py reads
(profiles/render_kernel_traffic.json).   usage: python tools/ncu_summary.py report.ncu-rep [--json out.json] [--source "cmd"]"""
import csv
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "kernel duration under ncu", "gpu_time"),
    ("smsp__inst_executed.sum", "warp instructions", "warp_instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (of 32)", "thread_inst_per_warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %", "warps_active_pct"),
    ("launch__registers_per_thread", "registers / thread", "registers_per_thread"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %", "pipe_xu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %", "pipe_lsu_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %", "pipe_fp64_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX throughput %", "l1tex_throughput_pct"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %", "l1tex_hit_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %", "lts_throughput_pct"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %", "lts_hit_pct"),
    ("dram__bytes_read.sum", "DRAM bytes read", "dram_bytes_read"),
    ("dram__bytes_write.sum", "DRAM bytes written", "dram_bytes_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak", "dram_throughput_pct"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["wait", "long_scoreboard", "not_selected", "no_instruction", "branch_resolving", "short_scoreboard", "math_pipe_throttle",
               "barrier", "dispatch_stall", "lg_throttle", "mio_throttle", "imc_miss"]
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    js, per_kernel = {}, []
    for vals in rows[2:]:
        js = {}
        name = vals[hdr.index("Kernel Name")]
        print(f"## {name.split('(')[0]}\n\n| metric | value |\n|---|---|")
        for m, label, key in METRICS:
            if m not in hdr:
                continue
            i = hdr.index(m)
            print(f"| {label} (`{m}`) | {vals[i]} {units[i]} |")
            try:
                x = float(vals[i].replace(",", ""))
                js[key] = x * SCALE.get(units[i], 1.0) if key.startswith("dram_bytes") else x
                if key == "gpu_time":          # always microseconds in the JSON
                    js[key] = x * {"second": 1e6, "s": 1e6, "msecond": 1e3, "ms": 1e3, "usecond": 1.0, "us": 1.0, "nsecond": 1e-3, "ns": 1e-3}.get(units[i], 1.0)
            except ValueError:
                pass
        st = []
        for s in STALL_NAMES:
            m = STALLS % s
            if m in hdr:
                st.append((float(vals[hdr.index(m)].replace(",", "")), s))
        tot = sum(v for v, _ in st) + 1.0
        print("| stall cycles per issued instruction (warp-cycles; `selected` = 1) | " +
              ", ".join(f"{s} {v:.2f}" for v, s in sorted(st, reverse=True) if v >= 0.05) + f" (sum incl. selected {tot:.2f}) |")
        print()
        js["kernel"] = name.split("(")[0]
        per_kernel.append(js)
    if "--json" in sys.argv:
        # several kernels in one capture = the kernels of ONE frame (config 5: wave_primary + wave_shade): the JSON describes the
        # longest one and carries the frame's total DRAM bytes and every kernel's share
        js = max(per_kernel, key=lambda k: k.get("gpu_time", 0.0))
        js["dram_bytes_per_launch"] = int(sum(k.get("dram_bytes_read", 0) + k.get("dram_bytes_write", 0) for k in per_kernel))
        if len(per_kernel) > 1:
            js["frame_kernels"] = [{"kernel": k["kernel"], "gpu_time_us_under_ncu": k.get("gpu_time"),
                                    "dram_bytes": int(k.get("dram_bytes_read", 0) + k.get("dram_bytes_write", 0)),
                                    "issue_active_pct": k.get("issue_active_pct")} for k in per_kernel]
        if "gpu_time" in js:
            js["gpu_time_us_under_ncu"] = js.pop("gpu_time")
        if "--source" in sys.argv:
            js["source"] = sys.argv[sys.argv.index("--source") + 1]
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(js, f, indent=1)


if __name__ == "__main__":
    main()
