"""Text summary of an .ncu-rep (one block per kernel launch): python tools/ncu_summary.py rep.ncu-rep "<command that was profiled>" [kernel-regex]
Key launch / throughput / pipe metrics plus the warp-stall mix (stalled warps per issue-active cycle), in the layout of profiles/r01_*_ncu_summary.txt."""
import csv, io, re, subprocess, sys

rep, cmd = sys.argv[1], sys.argv[2]
kre = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio"]
col = {h: i for i, h in enumerate(hdr)}
print(f"ncu --set full --clock-control none --import-source on : {cmd}")
print("(B200; numbers under ncu are not bench values)\n")
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if kre and not kre.search(name):
        continue
    print(f"{'Kernel Name':90s} {name}")
    for k in KEYS:
        if k in col and r[col[k]] != "":
            print(f"{k:90s} {r[col[k]]:>16s} {units[col[k]]}")
    stalls = []
    for h, i in col.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
        if m and r[i] != "":
            stalls.append((float(r[i].replace(",", "")), m.group(1)))
    if stalls:
        print("\nwarp stall reasons (warps stalled per issue-active cycle):")
        for v, nm in sorted(stalls, reverse=True)[:12]:
            print(f"  {nm:28s} {v:.3f}")
    print()
