import csv,sys,subprocess
rep=sys.argv[1]; N=int(sys.argv[2]) if len(sys.argv)>2 else 32768
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__shared_mem_per_block_dynamic',
'smsp__inst_executed.sum','sm__issue_active.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active',
'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
'dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']
def tobytes(v,un):
    return float(v)*{'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}[un]
for r in rows[2:]:
    d=dict(zip(hdr,r)); u=dict(zip(hdr,units))
    print("== "+d['Kernel Name'])
    for k in keys:
        if k in d: print(f"  {k:95s} {d[k]:>16s} {u[k]}")
    inst=float(d['smsp__inst_executed.sum']); wf=float(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']); bc=float(d['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'])
    rd=tobytes(d['dram__bytes_read.sum'],u['dram__bytes_read.sum']); wr=tobytes(d['dram__bytes_write.sum'],u['dram__bytes_write.sum'])
    unit = "per frame" if N > 1 else "per launch"
    print(f"  {unit} ({N} frames): {inst/N:.0f} warp-instructions, {wf/N:.0f} shared-memory wavefronts ({bc/N:.0f} from bank conflicts), DRAM {rd/N:.0f} B read + {wr/N:.0f} B written")
    print()
