"""Summarise an ncu report (``ncu -i X.ncu-rep --page raw --csv`` piped or given as a file) into a markdown table of
the metrics the roofline discussion uses.  Usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py [title]"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
]


def main():
    title = sys.argv[1] if len(sys.argv) > 1 else "ncu summary"
    rows = [r for r in csv.reader(sys.stdin) if r]
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    for r in rows[2:]:
        print(f"## {r[idx['Kernel Name']][:110]}  (launch id {r[idx['ID']]})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in idx and r[idx[w]] != "":
                print(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
        print()


if __name__ == "__main__":
    main()
