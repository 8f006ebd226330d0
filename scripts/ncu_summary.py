"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline discussion needs.
usage: python scripts/ncu_summary.py file.ncu-rep [more...]"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r)); u = dict(zip(h, units))
        print('==', path, d.get('Kernel Name', '')[:90])
        for k in KEYS:
            if k in d: print('  %-62s %s %s' % (k, d[k], u[k]))
        stalls = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                         if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and v not in ('', 'n/a')), reverse=True)[:7]
        for v, k in stalls: print('  stall %-56s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
