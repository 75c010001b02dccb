"""Condense `ncu -i report.ncu-rep --page raw --csv` into the handful of metrics that decide what bounds a kernel:
duration, DRAM bytes, pipe utilisation (tensor / alu / fma / xu / lsu), issue activity, occupancy limiters and the warp
stall reasons per issued instruction, one block per captured launch.
  ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_key_metrics.py [substring of kernel name]"""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "sm__cycles_elapsed.max.per_second",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else ""
    rows = list(csv.reader(sys.stdin))
    if len(rows) < 3:
        raise SystemExit("no launches in the input (expected the --page raw --csv dump of an .ncu-rep)")
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if want not in name:
            continue
        print(f"== {name[:110]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k in KEEP:
            if k in idx and r[idx[k]] != "":
                print(f"   {k:86s} {r[idx[k]]:>18s} {units[idx[k]]}")
        stalls = sorted(((float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for h, i in idx.items()
                         if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")),
                        reverse=True)
        print("   stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]))


if __name__ == "__main__":
    main()
