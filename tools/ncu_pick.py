"""Print selected metrics of every kernel in an .ncu-rep (via `ncu -i ... --page raw --csv`).  Usage:
python tools/ncu_pick.py report.ncu-rep [substring ...]   (default: a fixed list of headline metrics)"""
import csv, subprocess, sys
DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "launch__shared_mem_per_block_dynamic"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
pick = sys.argv[2:]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:110], "grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if (pick and any(p in h for p in pick)) or (not pick and h in DEFAULT):
            print(f"   {h} [{u}] = {v}")
