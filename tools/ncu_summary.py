#!/usr/bin/env python
"""Key metrics of an .ncu-rep capture as text:  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
keys=["gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__cycles_elapsed.avg","sm__cycles_active.avg","lts__t_bytes.sum","lts__t_sectors_srcunit_tex_op_read.sum","smsp__cycles_active.avg","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__inst_executed.sum","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","launch__grid_size","launch__block_size","smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct","smsp__warp_issue_stalled_barrier_per_warp_active.pct","smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct","smsp__warp_issue_stalled_wait_per_warp_active.pct","smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct","smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct","smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct","smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct","smsp__warp_issue_stalled_membar_per_warp_active.pct","smsp__warp_issue_stalled_sleeping_per_warp_active.pct","smsp__warp_issue_stalled_not_selected_per_warp_active.pct","smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct","smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct","smsp__warp_issue_stalled_no_instruction_per_warp_active.pct","smsp__average_warp_latency_issue_stalled"]
d={h:(v,u) for h,u,v in zip(hdr,units,vals)}
print("kernel:", d.get("Kernel Name",("?",""))[0][:90])
for k in keys:
    for h in d:
        if h==k or h.endswith(k):
            print(f"{h:92s} {d[h][0]} {d[h][1]}")
            break
