"""Summarise ncu outputs (read here, without a GPU):
    python tools/ncu_summary.py launches gpurun_out/launches_c2.csv       # per-kernel share of the launch list
    python tools/ncu_summary.py full gpurun_out/prof_scan_c2.ncu-rep      # key metrics of a --set full capture
    python tools/ncu_summary.py profile c2 gpurun_out/prof_scan_c2.ncu-rep profiles/scan_profile.json
        # the facts bench.py's `roofline` block quotes for the dominant kernel (DRAM bytes per launch, pipe utilisation, the unit
        # that binds), merged into profiles/scan_profile.json under the workload's key
"""
import collections
import csv
import io
import subprocess
import sys

KEYS_EXTRA_TENSOR = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                     "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
                     "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, n = collections.OrderedDict(), 0
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"{path}: {n} launches, {tot:.1f} us total (ncu per-launch times are cold-cache and serialised: compare SHARES)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:44s} n={c:4d} total={t:10.1f} us avg={t / c:9.1f} us share={100 * t / tot:5.1f}%")


def full(path, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        print("--", row[hdr.index("Kernel Name")][:100], "grid", row[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for k in list(KEYS) + KEYS_EXTRA_TENSOR + list(extra):
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:90s} {row[i]:>16s} {units[i]}")


def profile(workload, path, out_json):
    import json
    import os

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, row = rows[0], rows[-1]

    def val(k):
        return float(row[hdr.index(k)].replace(",", "")) if k in hdr and row[hdr.index(k)] not in ("", "n/a") else None

    def byt(k):
        v = val(k)
        u = rows[1][hdr.index(k)] if k in hdr else ""
        return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    pipes = {"alu": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
             "fma": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
             "xu": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
             "lsu": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
             "tensor_imma": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
             "issue_active": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "dram": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
             "warps_active": val("sm__warps_active.avg.pct_of_peak_sustained_active")}
    units = {k: v for k, v in pipes.items() if k in ("alu", "fma", "xu", "lsu", "tensor_imma", "dram") and v is not None}
    top = max(units, key=units.get)
    entry = {"kernel": row[hdr.index("Kernel Name")].split("(")[0], "source": f"ncu --set full, {os.path.basename(path)} (one launch of the main scan round)",
             "duration_us_under_ncu": val("gpu__time_duration.sum"),
             "dram_bytes_per_launch": (byt("dram__bytes_read.sum") or 0) + (byt("dram__bytes_write.sum") or 0),
             "pipes_pct": {k: (None if v is None else round(v, 1)) for k, v in pipes.items()},
             "registers_per_thread": val("launch__registers_per_thread"),
             "bound": {"alu": "alu (integer / logic pipe: A-fragment expansion + epilogue)", "fma": "fma", "xu": "xu",
                       "lsu": "lsu (shared-memory fragment loads)", "tensor_imma": "tensor (IMMA)", "dram": "hbm"}[top] +
                      f"; issue slots active {pipes['issue_active']:.0f} % -- latency-bound below every pipe's peak"}
    data = {}
    if os.path.exists(out_json):
        data = json.load(open(out_json))
    data[workload] = entry
    json.dump(data, open(out_json, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "profile":
        profile(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3:])
