"""Condense `ncu -i X.ncu-rep --page raw --csv` into the metrics the roofline discussion uses.
usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr (of 32)"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "SM issue-slot utilisation % (inst issued, of peak)"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (warp instructions / active cycle / SM)"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA (FP32) pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall imc_miss / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch / issue"),
    ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex_throttle / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping / issue"),
    ("smsp__average_warps_issue_stalled_misc_per_issue_active.ratio", "stall misc / issue"),
    ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "stall selected / issue"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (warp instr)"),
    ("smsp__inst_executed_op_local_st.sum", "local stores (warp instr)"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.avg", "SM elapsed cycles"),
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if not l.startswith("==")))
hdr, units = rows[0], rows[1]
print(f"# {sys.argv[1]} (ncu --set full --clock-control none; per-launch, cold-cache, serialised)\n")
for r in rows[2:]:
    print(f"## {r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')}  (launch id {r[0]})\n")
    print("| metric | value |\n|---|---|")
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
    print()
