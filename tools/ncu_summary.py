"""Per-kernel key counters and DRAM traffic from an `ncu --set full` report of tools/prof_run.py.
usage: python tools/ncu_summary.py report.ncu-rep frames out_kernels.csv out_traffic.json
(reads the report with `ncu -i ... --page raw --csv`; keeps the first launch of every kernel)"""
import csv
import json
import subprocess
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def main():
    rep, frames, out_csv, out_json = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in idx]
    seen, keep = set(), []
    for r in data:
        name = r[idx["Kernel Name"]]
        short = name.replace("<unnamed>::", "").split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
        short = {"k_squares4": "k_squares"}.get(short, short)      # same PROF label in the library (cvb_grid.cu)
        if short in seen:
            continue
        seen.add(short)
        keep.append((short, r))
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols); w.writerow([units[idx[c]] for c in cols])
        for _, r in keep:
            w.writerow([r[idx[c]] for c in cols])
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = {}
    for short, r in keep:
        tot = 0.0
        for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[c]]) * scale[units[idx[c]]]
        traffic[short] = int(tot / frames)
    # utilisation of the pipes a kernel can be bound by (percent of peak): bench.py reports the busiest one as `roofline.bound`
    pipe_cols = {"dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                 "lsu": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                 "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
                 "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                 "alu": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
                 "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"}
    pipes = {short: {k: round(float(r[idx[c]]), 1) for k, c in pipe_cols.items() if c in idx} for short, r in keep}
    conflicts = {}
    for short, r in keep:
        a, b = "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
        if a in idx and b in idx and float(r[idx[a]]) > 0:
            conflicts[short] = round(float(r[idx[b]]) / float(r[idx[a]]), 3)
    json.dump({"source": "%s: ncu --set full, tools/prof_run.py %d 1 (%d x 1080p board frames per launch), dram__bytes_read.sum + "
                         "dram__bytes_write.sum of the first launch of each kernel / %d" % (out_csv, frames, frames, frames),
               "dram_bytes_per_frame": traffic, "pipes_pct": pipes, "shared_conflict_share": conflicts},
              open(out_json, "w"), indent=1)
    for short, r in keep:
        print("%-14s %8.3f ms  issue %5.1f%%  warps %5.1f%%  dram/frame %.2f MB" % (
            short, float(r[idx["gpu__time_duration.sum"]]), float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
            float(r[idx["sm__warps_active.avg.pct_of_peak_sustained_active"]]), traffic[short] / 1e6))


if __name__ == "__main__":
    main()
