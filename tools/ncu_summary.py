"""Text summary of an .ncu-rep for profiles/: the metrics /opt/skills/guides/B200_PROFILING.md names, one block per launch.

    python tools/ncu_summary.py gpurun_out/xline_r2b.ncu-rep "header line" > profiles/xline_r2.summary.txt
"""
import csv
import subprocess
import sys

METRICS = [
    "Kernel Name", "Grid Size", "Block Size",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_red.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main(path, header, labels):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# " + header)
    for li, r in enumerate(rows[2:]):
        tag = labels[li] if li < len(labels) else ""
        print(f"\n== launch {li}" + (f": {tag}" if tag else ""))
        for m in METRICS:
            if m in idx:
                print(f"{m:100s} {units[idx[m]]:16s} {r[idx[m]][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", sys.argv[3:])
