#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: tools/ncu_summary.py <file.ncu-rep> <out.txt> [frames_per_launch]"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu --set full --clock-control none summary of %s" % rep.split("/")[-1]]
    for r in rows[2:]:
        lines.append("")
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("  %-80s %s %s" % (k, r[i], units[i]))
                vals[k] = (r[i], units[i])
        try:
            rd = float(vals["dram__bytes_read.sum"][0]) * UNIT_SCALE[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0]) * UNIT_SCALE[vals["dram__bytes_write.sum"][1]]
            lines.append("  dram traffic per launch (read+write)                                             %.0f bytes" % (rd + wr))
            if frames:
                lines.append("  dram traffic per frame                                                           %.0f bytes" % ((rd + wr) / frames))
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
