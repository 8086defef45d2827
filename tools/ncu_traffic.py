#!/usr/bin/env python
"""Extract the per-launch DRAM traffic (and the headline counters) of a kernel from an `ncu --set full` capture and record
them where bench.py picks them up (profiles/stream_traffic.json) -- so that `roofline.traffic` is read from a capture, never
typed in.

    python tools/ncu_traffic.py <capture.ncu-rep> <workload>/<precision> [--summary profiles/<name>.md]

The capture is made on the GPU box with the command in tools/profile_stream.sh; this script runs wherever `ncu` is on PATH
(reading a report needs no GPU).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read_report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[-1]
    out = {}
    for n, u, v in zip(names, units, vals):
        out[n] = (u, v)
    return out


def main():
    rep, key = sys.argv[1], sys.argv[2]
    summary = sys.argv[sys.argv.index("--summary") + 1] if "--summary" in sys.argv else None
    m = read_report(rep)
    def bytes_of(name):
        u, v = m[name]
        return float(v.replace(",", "")) * SCALE[u]
    traffic = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
    path = os.path.join(ROOT, "profiles", "stream_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[key] = {"dram_bytes": traffic, "dram_bytes_read": bytes_of("dram__bytes_read.sum"), "dram_bytes_write": bytes_of("dram__bytes_write.sum"),
                  "kernel": m.get("Kernel Name", ("", ""))[1], "duration": " ".join(reversed(m["gpu__time_duration.sum"])),
                  "source": "ncu --set full capture %s, read by tools/ncu_traffic.py" % os.path.basename(rep)}
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print("%s: %.3f GB per launch -> %s" % (key, traffic / 1e9, path))
    if summary:
        with open(summary, "w") as f:
            f.write("| metric | value | unit |\n|---|---|---|\n")
            for n in WANT:
                if n in m:
                    f.write("| %s | %s | %s |\n" % (n, m[n][1], m[n][0]))
        print("summary ->", summary)


if __name__ == "__main__":
    main()
