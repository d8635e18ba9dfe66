"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into a small markdown table of the metrics the roofline uses."""
import csv, subprocess, sys, re
KEYS = [
 ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
 ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
 ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
 ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
 ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem tensor-operand wavefronts %"),
 ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
 ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
 ("smsp__inst_executed.sum", "warp instructions"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
 ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
 ("launch__grid_size", "grid"), ("launch__block_size", "block"),
 ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
 ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
 ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
 ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
 ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
 ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
 ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
]
def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    print("| kernel | " + " | ".join(k[1] for k in KEYS if k[0] in col) + " |")
    print("|---|" + "---|" * sum(1 for k in KEYS if k[0] in col))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])[:60]
        vals = []
        for k, _ in KEYS:
            if k in col:
                vals.append((r[col[k]] + " " + u[col[k]]).strip())
        print("| " + name + " | " + " | ".join(vals) + " |")
if __name__ == "__main__":
    main(sys.argv[1])
