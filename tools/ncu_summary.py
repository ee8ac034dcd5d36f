"""Summarise an .ncu-rep: key metrics of each profiled kernel (reads `ncu --page raw --csv`).

    python tools/ncu_summary.py report.ncu-rep                      # text summary
    python tools/ncu_summary.py report.ncu-rep --json out.json --units N [--capture "what ran"]

--json writes the per-unit costs bench.py's roofline reads (profiles/r2_*_ncu.json): N is the
number of units (member-steps, events ...) the captured launch processed."""
import csv
import json
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


def number(text):
    try:
        return float(text.replace(',', ''))
    except ValueError:
        return None


def write_json(path, d, units_of, out, n_units, capture):
    def get(key):
        return number(d.get(key, ''))

    def scaled(key, scale):
        value = get(key)
        unit = units_of.get(key, '')
        for prefix, factor in (('K', 1e3), ('M', 1e6), ('G', 1e9)):
            if unit.startswith(prefix) and scale == 'bytes':
                return value * factor
        return value
    ti, tt = get('thread_inst_executed'), get('thread_inst_executed_true')
    wi = get('smsp__inst_executed.sum')
    dur = get('gpu__time_duration.sum')
    if units_of.get('gpu__time_duration.sum', '').startswith('us'):
        dur /= 1e3
    elif units_of.get('gpu__time_duration.sum', '').startswith('ns'):
        dur /= 1e6
    elif units_of.get('gpu__time_duration.sum', '').startswith('s'):
        dur *= 1e3
    dram = scaled('dram__bytes_read.sum', 'bytes') + scaled('dram__bytes_write.sum', 'bytes')
    issue = get('smsp__issue_active.avg.pct_of_peak_sustained_active')
    lanes = get('smsp__thread_inst_executed_per_inst_executed.ratio')
    doc = {
        'source': path, 'capture': capture, 'kernel': d.get('Kernel Name'),
        'units_in_launch': n_units, 'duration_ms': dur,
        'warp_inst_executed': wi, 'thread_inst_executed': ti, 'thread_inst_executed_true': tt,
        'warp_inst_per_unit': wi / n_units * 1.0,
        'thread_inst_executed_per_member_step': ti / n_units,
        'thread_inst_executed_true_per_member_step': tt / n_units,
        'dram_bytes': dram, 'dram_bytes_per_member_step': dram / n_units,
        'issue_active_pct': issue, 'lanes_per_inst': lanes,
        'frac_issue_x_lanes': issue / 100.0 * lanes / 32.0,
        'pipe_fmaheavy_pct': get('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'),
        'pipe_alu_pct': get('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
        'pipe_fp64_pct': get('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
        'registers_per_thread': get('launch__registers_per_thread'),
        'shared_bank_conflicts': get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
        'shared_wavefronts': get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
    }
    with open(out, 'w') as handle:
        json.dump(doc, handle, indent=1)
        handle.write('\n')


def main():
    path = sys.argv[1]
    text = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                          text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    if '--json' in sys.argv:
        out = sys.argv[sys.argv.index('--json') + 1]
        n_units = float(sys.argv[sys.argv.index('--units') + 1])
        capture = sys.argv[sys.argv.index('--capture') + 1] if '--capture' in sys.argv else ''
        want = sys.argv[sys.argv.index('--kernel') + 1] if '--kernel' in sys.argv else None
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            if want is None or want in d.get('Kernel Name', ''):
                write_json(path, d, dict(zip(hdr, units)), out, n_units, capture)
                print('wrote', out)
                return
        sys.exit('no kernel matched')
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
