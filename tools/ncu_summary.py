"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (needs `ncu` on PATH; no GPU required)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed_pipe_xu.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    print(f"== {name}  grid {r[idx['launch__grid_size']]} x {r[idx['launch__block_size']]}")
    for w in WANT:
        if w in idx and not w.startswith("launch__grid") and not w.startswith("launch__block"):
            print(f"   {w:75s} {r[idx[w]]:>14s} {units[idx[w]]}")
