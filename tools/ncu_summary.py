"""profiles/ helper: key metrics of one kernel from an .ncu-rep (raw page) as text."""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else None      # substring of the kernel name; default: last launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
if want:
    k = hdr.index("Kernel Name")
    vals = [r for r in rows[2:] if want in r[k]][-1]
    print(f"kernel: {vals[k]}")
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum dram__bytes_read.sum.per_second
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__cycles_elapsed.max sm__cycles_elapsed.max.per_second
smsp__cycles_active.avg sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__grid_size launch__block_size smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed.avg.per_cycle_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed smsp__inst_executed_op_shared_ld.sum
l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct lts__t_sectors_srcunit_tex_op_read.sum
lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warp_latency_per_inst_issued.ratio""".split()
for k in keys:
    if k in d:
        print(f"{k:88s} {d[k][0]:>18s} {d[k][1]}")
