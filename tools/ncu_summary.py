"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  > profiles/rNN_<kernel>.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum.per_cycle_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "gpc__cycles_elapsed.max",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    agg = defaultdict(lambda: [0, 0.0, "", ""])
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"]
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg[short]
        a[0] += 1
        a[1] += float(d["Metric Value"].replace(",", "")) / (1e3 if d["Metric Unit"] == "ns" else 1.0)
        a[2], a[3] = d["Grid Size"], d["Block Size"]
    tot = sum(v[1] for v in agg.values())
    print("| launches | total us | share | grid | block | kernel |")
    print("|---:|---:|---:|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[2]} | {v[3]} | `{k[:90]}` |")
    print(f"\ntotal device time of listed launches: {tot / 1e3:.2f} ms (cold-cache, serialised by ncu: compare shares)")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, zip(units, r)))
        print(f"### {d['Kernel Name'][1][:160]}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in KEEP:
            if k in d:
                print(f"| {k} | {d[k][1]} | {d[k][0]} |")
        for k in sorted(d):
            if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
                v = d[k][1]
                try:
                    if float(v.replace(",", "")) >= 0.03:
                        print(f"| stall: {k[len(STALL):-len('_per_issue_active.ratio')]} (warps per issue-active) | {v} | |")
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
