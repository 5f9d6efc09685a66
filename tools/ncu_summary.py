"""Summarise an ncu report for profiles/: python tools/ncu_summary.py <rep> <out.txt> [workload-key]
Writes the key metrics as text and (with a workload key) records DRAM bytes per launch in profiles/ncu_traffic.json."""
import csv, io, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__cluster_max_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
lines = [f"# ncu --set full summary of {os.path.basename(rep)} (one launch per row below)"]
traffic = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    lines.append("")
    for w in want:
        if w in d:
            lines.append(f"{w:85s} {d[w]} {units[hdr.index(w)]}")
    def num(k):
        v = float(d[k].replace(",", "")); u = units[hdr.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    lines.append(f"{'dram traffic (read+write) per launch':85s} {traffic / 1e6:.3f} MB")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if key and traffic is not None:
    path = os.path.join(os.path.dirname(out), "ncu_traffic.json")
    tj = json.load(open(path)) if os.path.exists(path) else {}
    tj[key] = traffic
    json.dump(tj, open(path, "w"), indent=1)
