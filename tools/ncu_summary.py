"""Key metrics of an .ncu-rep (read with `ncu -i ... --page raw --csv`), one column per launch."""
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active",
        "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu", "smsp__warp_issue_stalled",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__sass_inst_executed_op_shared",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or KEYS
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            print("%-78s %-10s %s" % (h[:78], units[i][:10], "  ".join(r[i][:40] for r in data)))


main()
