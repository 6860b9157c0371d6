import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__inst_executed.sum','launch__grid_size','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_read.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','launch__shared_mem_config_size','sm__maximum_warps_per_active_cycle_pct']
for vals in rows[2:]:
    print('==',vals[hdr.index('Kernel Name')][:80])
    for i,h in enumerate(hdr):
        if h in want: print('  ',h,'=',vals[i],rows[1][i])
    st={h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):float(vals[i]) for i,h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h}
    print('   stalls/issue:',', '.join('%s %.2f'%(k,v) for k,v in sorted(st.items(),key=lambda kv:-kv[1])[:7]))
