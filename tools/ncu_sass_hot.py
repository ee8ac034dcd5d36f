"""Per-instruction executed counts from an .ncu-rep source page: prints the SASS with executed
warp-instruction counts per warp-step, so hot regions can be read top to bottom.
usage: python tools/ncu_sass_hot.py report.ncu-rep WARP_STEPS [min_count]"""
import csv
import subprocess
import sys


def main():
    path, warp_steps = sys.argv[1], float(sys.argv[2])
    thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    text = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr = rows[1]
    i_src, i_ex, i_thr, i_smp = (hdr.index('Source'), hdr.index('Instructions Executed'),
                                 hdr.index('Avg. Threads Executed'), hdr.index('# Samples'))
    total = 0.0
    for k, r in enumerate(rows[2:]):
        if len(r) <= i_ex:
            continue
        ex = float(r[i_ex]) / warp_steps
        total += ex
        if ex >= thresh:
            print('%5d %8.2f %5s %6s  %s' % (k, ex, r[i_thr], r[i_smp], r[i_src].strip()))
    print('total warp-instructions per warp-step: %.1f' % total)


if __name__ == '__main__':
    main()
