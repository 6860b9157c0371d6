"""Summarise `ncu --set full` reports of the NMS3D kernels: per launch duration, SM / L2 / DRAM throughput, issue-active, stall mix."""
import csv, subprocess, sys
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.sum', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum']
short = {'gpu__time_duration.sum': 'us', 'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'SM%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'L2%',
         'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'L1%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'DRAM%',
         'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%', 'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps%',
         'launch__grid_size': 'grid', 'launch__block_size': 'block', 'launch__registers_per_thread': 'regs', 'smsp__inst_executed.sum': 'inst',
         'sm__inst_executed_pipe_fma.sum': 'fma_inst', 'lts__t_bytes.sum': 'L2_bytes', 'dram__bytes_read.sum': 'dram_rd', 'dram__bytes_write.sum': 'dram_wr'}
for rep in sys.argv[1:]:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print('## ' + rep.split('/')[-1])
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')].split('(')[0].replace('roi3d::', '').replace('void ', '')
        d = {short[h]: (vals[i], units[i]) for i, h in enumerate(hdr) if h in short}
        def f(k):
            v, u = d.get(k, ('', ''))
            try:
                x = float(v.replace(',', ''))
            except ValueError:
                return v
            if u in ('ns',): return '%.1f' % (x / 1e3)
            if u in ('us',): return '%.1f' % x
            if u == 'ms': return '%.1f' % (x * 1e3)
            return ('%.1f' % x) if u == '%' else ('%.3g %s' % (x, u)).strip()
        st = {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): float(vals[i]) for i, h in enumerate(hdr)
              if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h}
        top = ', '.join('%s %.2f' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:4])
        print('  %-28s %7s us  grid %-6s SM %5s%%  L2 %5s%%  L1 %5s%%  DRAM %5s%%  issue %5s%%  warps %5s%%  inst %s  L2 bytes %s | stalls/issue: %s' % (
            name, f('us'), f('grid'), f('SM%'), f('L2%'), f('L1%'), f('DRAM%'), f('issue%'), f('warps%'), f('inst'), f('L2_bytes'), top))
