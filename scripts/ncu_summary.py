"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python scripts/ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print('==', f, '|', d.get('Kernel Name', '')[:70])
        for k in KEYS:
            if k in d:
                print('  %-72s %s %s' % (k, d[k], u[k]))
        st = [(h.replace('smsp__pcsamp_warps_issue_stalled_', ''), float(v.replace(',', ''))) for h, v in d.items()
              if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and v not in ('', 'n/a')]
        tot = sum(v for _, v in st) or 1
        print('  stalls: ' + ', '.join('%s %.0f%%' % (h, 100 * v / tot) for h, v in sorted(st, key=lambda x: -x[1])[:8]))
