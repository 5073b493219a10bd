"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV for profiles/:
   python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_summary.csv"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "lts__t_bytes.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in WANT or h.startswith("sm__pipe_tensor") and h.endswith("pct_of_peak_sustained_active")]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for i in cols])
    print(f"{out}: {len(rows) - 2} kernels, {len(cols)} metrics")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
