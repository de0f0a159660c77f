"""Summarise an ncu report of an event kernel into profiles/:

    python tools/ncu_summary.py <report.ncu-rep> <tag> [events per launch] [ecmc_kernel_name string]

Writes profiles/<tag>_ncu_summary.json: DRAM bytes per launch, fp64 pipe utilisation, issue utilisation, registers,
stalls, instruction mix per event. With the fourth argument (what Engine.kernel_name() returned for the captured run) the
summary carries "kernel_name", and bench.py quotes it -- labelled as a committed capture -- when a run launches the same
kernel."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu_csv(report, page, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    report, tag = sys.argv[1], sys.argv[2]
    events_per_launch = float(sys.argv[3]) if len(sys.argv) > 3 else None
    kernel_name = sys.argv[4] if len(sys.argv) > 4 else None
    raw = ncu_csv(report, "raw")
    header, units, values = raw[0], raw[1], raw[2]
    metric = {h: (v, u) for h, u, v in zip(header, units, values)}

    def get(name, scale=1.0):
        value, unit = metric[name]
        value = float(value.replace(",", ""))
        factor = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(unit, 1.0)
        return value * factor * scale

    summary = {
        "report": os.path.basename(report),
        "kernel": metric["Kernel Name"][0],
        "duration_ms": get("gpu__time_duration.sum") * 1e3,
        "dram_bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
        "dram_throughput_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "fp64_pipe_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": get("launch__registers_per_thread"),
        "waves_per_sm": get("launch__waves_per_multiprocessor"),
        "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
        "warp_instructions": get("smsp__inst_executed.sum"),
        "local_load_instructions": get("smsp__sass_inst_executed_op_local_ld.sum"),
        "local_store_instructions": get("smsp__sass_inst_executed_op_local_st.sum"),
        "stall_per_issue": {name.split("issue_stalled_")[1].split("_per_issue")[0]: float(v[0])
                            for name, v in metric.items() if name.startswith("smsp__average_warps_issue_stalled_")
                            and name.endswith("_per_issue_active.ratio") and float(v[0]) >= 0.05},
    }
    if kernel_name:
        summary["kernel_name"] = kernel_name
    if events_per_launch:
        summary["events_per_launch"] = events_per_launch
        summary["warp_instructions_per_event"] = summary["warp_instructions"] / events_per_launch
        summary["dram_bytes_per_event"] = summary["dram_bytes_per_launch"] / events_per_launch
    source = ncu_csv(report, "source")
    head = source[1]
    i_src, i_exec = head.index("Source"), head.index("Instructions Executed")
    mix = collections.Counter()
    for row in source[2:]:
        try:
            count = int(row[i_exec])
        except (ValueError, IndexError):
            continue
        match = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", row[i_src])
        mix[match.group(2) if match else "?"] += count
    if events_per_launch:
        summary["instruction_mix_per_event"] = {op: round(n / events_per_launch, 1) for op, n in mix.most_common(24)}
    folder = os.environ.get("NCU_SUMMARY_DIR", os.path.join(ROOT, "profiles"))  # on the GPU box: gpurun_out
    os.makedirs(folder, exist_ok=True)
    for name in (f"{tag}_ncu_summary.json",):
        with open(os.path.join(folder, name), "w") as handle:
            json.dump(summary, handle, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
