"""Print the handful of raw ncu metrics the optimisation notes quote.  usage: ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_keys.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]
want = ['gpu__time_duration.sum', 'launch__occupancy_limit', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'launch__shared_mem_config_size', 'launch__shared_mem_per_block_dynamic', 'l1tex__t_sector_hit_rate.pct',
        'issue_stalled', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct',
        'sm__inst_executed_pipe_lsu.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print(r[h.index('Kernel Name')] if 'Kernel Name' in h else '')
    for i, n in enumerate(h):
        if any(w in n for w in want) and r[i] not in ('', '0'):
            print('  %-95s %-14s %s' % (n, rows[1][i], r[i]))
