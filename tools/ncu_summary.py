"""Summarise an ncu report (`ncu --set full ... -o x`) as the per-launch table kept under profiles/.

    python tools/ncu_summary.py x.ncu-rep ["header line" ...] > profiles/rNN_ncu_<what>.txt

One column per captured launch, one row per metric; the metrics are the ones DESIGN.md argues with (duration, occupancy,
issue slots, pipes, L1/TEX data stage shared by texture and shared-memory wavefronts, hit rates, DRAM bytes, stall reasons)."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tex.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__f_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def fmt(v):
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    return str(int(f)) if f == int(f) and abs(f) < 1e15 else f"{f:.6g}"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    for line in sys.argv[2:]:
        print(line)
    if sys.argv[2:]:
        print()
    kcol = head.index("Kernel Name")
    w = 100
    print("metric [unit]".ljust(w) + "".join(f"launch{i}".rjust(16) for i in range(len(data))))
    print("kernel".ljust(w) + "".join(r[kcol].split("<")[-1].split(">")[0][:14].rjust(16) for r in data))
    for m in METRICS:
        cols = [i for i, c in enumerate(head) if c == m]
        if not cols:
            continue
        i = cols[0]
        name = f"{m} [{units[i]}]" if units[i] else m
        print(name.ljust(w) + "".join(fmt(r[i]).rjust(16) for r in data))


if __name__ == "__main__":
    main()
