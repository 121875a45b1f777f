"""Condense `ncu -i <rep> --page raw --csv --print-units base` into the columns the design notes quote.
Usage: python profiles/ncu_summary.py raw.csv out.csv"""
import csv, sys

KEEP = """ID
Kernel Name
gpu__time_duration.sum
dram__bytes_read.sum
dram__bytes_write.sum
dram__throughput.avg.pct_of_peak_sustained_elapsed
lts__t_bytes.sum
lts__t_sector_hit_rate.pct
l1tex__t_sector_hit_rate.pct
sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active
sm__inst_executed.sum
smsp__inst_executed.sum
smsp__thread_inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_issued.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__warp_issue_stalled_barrier_per_warp_active.pct
smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct
smsp__warp_issue_stalled_membar_per_warp_active.pct
smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct
smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct
smsp__warp_issue_stalled_wait_per_warp_active.pct
launch__registers_per_thread
launch__shared_mem_per_block_static
launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem
launch__grid_size
launch__block_size
sm__maximum_warps_per_active_cycle_pct
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
lts__t_sectors_op_atom.sum
lts__t_sectors_op_red.sum
smsp__sass_inst_executed_op_global_atom.sum
local_load_store
smsp__inst_executed_op_local_ld.sum
smsp__inst_executed_op_local_st.sum""".split("\n")


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = [i for i, n in enumerate(names) if n in KEEP]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([names[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[hdr + 2:]:
            if len(r) >= len(names):
                w.writerow([r[i][:90] for i in idx])
    print("kept", len(idx), "columns of", len(names))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
