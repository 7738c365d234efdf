"""Print the handful of ncu metrics we steer by from a .ncu-rep (one block per captured launch)."""
import csv, io, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'launch__grid_size', 'launch__block_size', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_lsu.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_xu.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_fmaheavy.sum', 'smsp__inst_executed_pipe_fmalite.sum']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')])
    for w in WANT:
        if w in hdr:
            print('  ', w, r[hdr.index(w)], units[hdr.index(w)])
    st = [(float(r[i]), h.split('issue_stalled_')[1].split('_per_')[0]) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    print('   stalls:', ', '.join(f'{n} {v:.2f}' for v, n in sorted(st, reverse=True)[:8]))
