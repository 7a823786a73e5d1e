"""Key raw metrics of every kernel in an ncu report: python tools/ncu_raw.py rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64pipe%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("lts__t_sectors_op_read.sum", "l2_rd_sectors"), ("lts__t_sectors_op_write.sum", "l2_wr_sectors")]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("-" * 100)
    for k, lab in want:
        if k in idx:
            print(f"  {lab:16s} {r[idx[k]]} {units[idx[k]]}")
    st = sorted(((float(r[idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)[:6]
    print("  stalls/issue    ", ", ".join(f"{n}={v:.2f}" for v, n in st))
