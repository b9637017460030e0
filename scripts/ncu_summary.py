"""Summarise an .ncu-rep (read here, no GPU): python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
print(f"# ncu --set full summary of `{rep}` (per launch; cold-cache, serialised: compare shares, not absolutes)\n")
for r in rows[2:]:
    print(f"## {r[idx['Kernel Name']][:110]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in idx:
            print(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    print()
