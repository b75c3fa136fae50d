"""Summarise an .ncu-rep (read on the CPU box): key roofline / pipe / stall metrics per captured kernel.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [more-metric-regex]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
idx = {h: i for i, h in enumerate(hdr)}
for r in data:
    print("=" * 100)
    print(r[idx["Kernel Name"]][:160])
    for w in want:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
    stalls = [(h, r[i]) for h, i in idx.items() if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    stalls = sorted(((float(v.replace(",", "")), h) for h, v in stalls if v not in ("", "n/a")), reverse=True)[:8]
    for v, h in stalls:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.3f}")
    if extra:
        for h, i in idx.items():
            if re.search(extra, h):
                print(f"  + {h:70s} {r[i]:>18s} {units[i]}")
