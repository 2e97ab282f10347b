"""Summarise an `ncu --set full` report: one row per metric, one column per kernel.

    python tools/ncu_summarize.py report.ncu-rep out.csv [traffic.json "source description"]

With a third argument also writes the DRAM bytes per launch of every captured kernel (what
bench.py quotes as roofline.traffic) together with the commit the kernels were built from.
Runs in the build container (ncu reads the report without a GPU)."""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = """launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__shared_mem_per_block_static launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__occupancy_limit_warps gpu__time_duration.sum sm__cycles_elapsed.max
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum dram__bytes_read.sum dram__bytes_write.sum
sm__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio""".split()

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [re.sub(r"\(int\)", "", r[ix["Kernel Name"]]).split("(")[0] for r in data]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in METRICS:
            if m in ix:
                w.writerow([m, units[ix[m]]] + [r[ix[m]] for r in data])
    if len(sys.argv) > 3:
        def val(r, m):
            return float(r[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1.0)
        commit = subprocess.run(["git", "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
        kern = {}
        for r, n in zip(data, names):
            short = n.replace("void ", "").split("<")[0]
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            kern[short] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr,
                           "duration_us": val(r, "gpu__time_duration.sum")}
        with open(sys.argv[3], "w") as f:
            json.dump({"source": sys.argv[4] if len(sys.argv) > 4 else rep, "commit": commit, "kernels": kern}, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
