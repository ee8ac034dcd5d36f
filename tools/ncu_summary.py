"""Summarise an .ncu-rep: key metrics of each profiled kernel (reads `ncu --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed.sum.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'smsp__inst_executed_op_global_red.sum']


def main():
    path = sys.argv[1]
    text = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                          text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('kernel:', d.get('Kernel Name'), ' grid', d.get('Grid Size'), ' block', d.get('Block Size'))
        for k in KEYS:
            if k in d:
                print('  %-75s %s %s' % (k, d[k], units[hdr.index(k)]))
        for k in hdr:
            if 'warps_issue_stalled' in k and k.endswith('per_issue_active.ratio'):
                print('  %-75s %s' % (k.replace('smsp__average_warps_issue_stalled_', 'stall '), d[k]))


if __name__ == '__main__':
    main()
