"""Key raw metrics (duration, DRAM bytes, occupancy limits, instruction/LSU/L1/L2 counters, top stall reasons) of every
kernel in an .ncu-rep:  python tools/ncu_raw_summary.py report.ncu-rep"""
import subprocess, csv, io, sys
rep=sys.argv[1]
want = sys.argv[2:] or None
txt=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(txt)))
hdr=rows[0]; units=rows[1]
keys=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread","launch__grid_size","launch__block_size","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__warps_active.avg.pct_of_peak_sustained_active","sm__throughput.avg.pct_of_peak_sustained_elapsed","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__inst_executed.sum","sm__inst_executed_pipe_fp64.sum","smsp__inst_executed_pipe_fp64.sum","sm__inst_executed_pipe_lsu.sum","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","lts__t_sectors_srcunit_tex.sum","lts__t_bytes.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__cycles_active.avg","smsp__cycles_active.avg","l1tex__lsu_writeback_active_mem_lgds.sum","smsp__average_warp_latency_issue_stalled_long_scoreboard","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print("=====")
    for k in keys:
        if k in d: print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    # stall reasons
    st=[(float(d[k].replace(',','')),k) for k in hdr if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio") and d[k]]
    for v,k in sorted(st,reverse=True)[:8]: print(f"     stall {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')} {v:.2f}")
