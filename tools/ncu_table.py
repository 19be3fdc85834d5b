"""One markdown table (a row per kernel) from an ncu --set full report: duration, DRAM bytes and rate, pipe / issue
utilisation, top stall reasons.

    python tools/ncu_table.py gpurun_out/x.ncu-rep [peak_GBps] > profiles/rNN_x.md
"""
import csv
import io
import re
import subprocess
import sys


def main(path, peak=6547.8):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]

    def val(r, name, scale=1.0):
        if name not in col or r[col[name]] in ("", "n/a"):
            return float("nan")
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]]
        if u in ("ns", "nsecond"):
            v *= 1e-3
        elif u in ("ms", "msecond"):
            v *= 1e3
        elif u == "Gbyte":
            v *= 1e3
        elif u == "Kbyte":
            v *= 1e-3
        elif u == "byte":
            v *= 1e-6
        return v * scale

    print(f"| kernel | us | regs | blocks/SM | DRAM rd MB | DRAM wr MB | DRAM GB/s | of {peak:.0f} | issue % | LSU data pipe % | FMA pipe % | "
          "warps eligible | smem conflicts / wavefronts | local ld | top stalls (warps per issue) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for r in rows[2:]:
        name = re.sub(r"pal::|palhost::|f2h::|fft2::|void |\(.*", "", r[col["Kernel Name"]])
        us = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / us * 1e3 / 1e3 if us == us and us > 0 else float("nan")      # MB / us = TB/s -> GB/s
        gbs = (rd + wr) / us * 1e3
        occ = min(val(r, "launch__occupancy_limit_registers"), val(r, "launch__occupancy_limit_shared_mem"),
                  val(r, "launch__occupancy_limit_warps"))
        conf = val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        wf = val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        st = sorted(((float(r[col[h]].replace(",", "") or 0), h) for h in stall), reverse=True)[:4]
        sts = ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.1f}" for v, h in st)
        print(f"| `{name[:80]}` | {us:.1f} | {val(r, 'launch__registers_per_thread'):.0f} | {occ:.0f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | "
              f"{100 * gbs / peak:.0f}% | {val(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{val(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{val(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{val(r, 'smsp__warps_eligible.avg.per_cycle_active'):.2f} | {conf / 1e6:.2f}M / {wf / 1e6:.1f}M | "
              f"{val(r, 'sass__inst_executed_local_loads'):.0f} | {sts} |")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 6547.8)
