"""Turn an .ncu-rep (ncu --set full) into a compact text summary for profiles/ (run on the CPU box):
python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_global_red.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = [h for h in hdr if "average_warps_issue_stalled" in h]
units = rows[1]
for r in rows[2:]:
    print("=" * 100)
    print(r[idx["Kernel Name"]])
    for k in KEYS:
        if k in idx:
            print(f"  {k:75s} {r[idx[k]]:>18s} {units[idx[k]]}")
    print("  -- warp stall reasons (average warps stalled per issue-active cycle), >= 0.2 only")
    for k in STALL:
        try:
            v = float(r[idx[k]])
        except ValueError:
            continue
        if v >= 0.2:
            print(f"  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):75s} {v:18.3f}")
