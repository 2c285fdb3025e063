#!/usr/bin/env python
"""Condense an .ncu-rep (or an ncu launch-list CSV) into the short text summaries kept under profiles/.

    python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep  > profiles/r1_closest.txt
    python tools/ncu_summary.py list gpurun_out/launches.csv   > profiles/r1_launches.txt
    python tools/ncu_summary.py json <workload> gpurun_out/prof.ncu-rep   merges the counters bench.py reads (roofline.traffic,
        warp instructions per ray) into profiles/kernel_counters.json, keyed by the hash of the kernel source as it is NOW --
        run it right after the capture, before editing kd_kernels.cuh
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {d.get('Kernel Name', '?')}  (id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:86s} {d[k]:>16s} {u.get(k, '')}")


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    unit = rows[1][ui]
    total = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum launch list: {path} ({sum(len(v) for v in agg.values())} launches; per-launch times are cold-cache and serialised)")
    print(f"{'kernel':70s} {'launches':>8s} {'mean ' + unit:>14s} {'sum ' + unit:>16s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:70]:70s} {len(v):8d} {sum(v) / len(v):14.1f} {sum(v):16.1f} {100 * sum(v) / total:6.1f}%")


def to_json(workload, path):
    import json, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    dest = bench.COUNTERS_JSON
    try:
        rec = json.load(open(dest))
    except Exception:
        rec = {}
    sha = bench.kernel_source_hash()
    if rec.get("kernel_source_sha") != sha:
        rec = {"kernel_source_sha": sha, "workloads": {}}
    rec["source"] = "ncu --set full --clock-control none captures summarised by tools/ncu_summary.py json; per launch"
    kinds = {"0": "closest", "1": "shadow", "2": "tshadow"}
    w = rec["workloads"].setdefault(workload, {})
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        m = re.search(r"(traceKernel|setupKernel)<(?:\(int\))?(\d)", name)
        if not m:
            continue
        kind, is_setup = kinds[m.group(2)], m.group(1) == "setupKernel"
        if (is_setup and "setup" in w.get(kind, {})) or (not is_setup and "kernel" in w.get(kind, {})):
            continue  # the first launch of each kernel
        f = lambda k: float(d[k].replace(",", "")) if d.get(k) not in (None, "") else None
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        u = dict(zip(hdr, rows[1]))
        dram = sum(f(k) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t_unit = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]]
        entry = {
            "kernel": d["Kernel Name"].split("(")[0].replace("void ", ""), "capture": os.path.basename(path),
            "duration_ms_under_ncu": f("gpu__time_duration.sum") * t_unit, "dram_bytes": dram,
            "dram_bytes_read": f("dram__bytes_read.sum") * scale[u["dram__bytes_read.sum"]],
            "warp_inst": f("smsp__inst_executed.sum"), "lanes_per_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l2_sector_bytes": 32.0 * (f("lts__t_sectors.sum") or ((f("lts__t_sectors_op_read.sum") or 0.0) + (f("lts__t_sectors_op_write.sum") or 0.0))),
            "l1_hit_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": f("lts__t_sector_hit_rate.pct"),
            "long_scoreboard_per_issue": f("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            "registers": f("launch__registers_per_thread"),
        }
        # a two-pass batch is setupKernel<Q> + traceKernel<Q, ., true>: the traversal pass is the entry, the setup pass hangs below it
        if is_setup:
            w.setdefault(kind, {})["setup"] = entry
        else:
            entry.update({k: v for k, v in w.get(kind, {}).items() if k == "setup"})
            w[kind] = entry
    json.dump(rec, open(dest, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec["workloads"][workload], indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "json":
        to_json(sys.argv[2], sys.argv[3])
    else:
        (rep if sys.argv[1] == "rep" else launch_list)(sys.argv[2])
