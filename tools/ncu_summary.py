"""Condense an `ncu --set full` report into the handful of counters DESIGN.md / bench.py cite.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
(reads the report with `ncu -i ... --page raw --csv`; no GPU needed)."""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__cluster", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_uniform", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled",
    "sm__sass_inst_executed_op_shared", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("=== kernel:", d.get("Kernel Name"), "| id", d.get("ID"))
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(k) for k in KEEP) and "TriageCompute" not in h or h.endswith(
                    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
                print(f"{h:90s} {v:>18s} {u}")
        print()


if __name__ == "__main__":
    main()
