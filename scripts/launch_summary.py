"""Summarise an ncu launch list (csv from `ncu --metrics ... --csv --log-file X`) per kernel and grid.
Usage: python scripts/launch_summary.py gpurun_out/launches.csv [last_n_launches] [--each]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r['ID']), {'name': r['Kernel Name'], 'grid': r['Grid Size'], 'block': r['Block Size']})
        try:
            d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
        except ValueError:
            pass
    return list(rows.values())


def short(name):
    name = re.sub(r'\((const |unsigned |int|float|void|long|bool|\w+ \*).*', '', name)
    return name.replace('void ', '').replace('<unnamed>::', '')[:58]


def main():
    rows = load(sys.argv[1])
    args = [a for a in sys.argv[2:] if not a.startswith('--')]
    if args:
        rows = rows[-int(args[0]):]
    T = sum(r.get('gpu__time_duration.sum', 0) for r in rows) / 1e3
    mets = [('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tens%'),
            ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
            ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
            ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%')]
    if '--each' in sys.argv:
        for r in rows:
            t = r.get('gpu__time_duration.sum', 0) / 1e3
            print('%-58s %-16s %9.1f us ' % (short(r['name']), r['grid'], t) +
                  ' '.join('%s %5.1f' % (lbl, r[m]) for m, lbl in mets if m in r))
        print('total %.1f us over %d launches' % (T, len(rows)))
        return
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r['name']), {'n': 0, 't': 0.0, 'w': collections.Counter()})
        t = r.get('gpu__time_duration.sum', 0) / 1e3
        a['n'] += 1
        a['t'] += t
        for m, lbl in mets:
            if m in r:
                a['w'][lbl] += r[m] * t
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
        print('%-58s %4d %9.1f us %5.1f%%  ' % (k, a['n'], a['t'], 100 * a['t'] / T) +
              ' '.join('%s %5.1f' % (lbl, a['w'][lbl] / max(a['t'], 1e-9)) for _, lbl in mets if lbl in a['w']))
    print('total %.1f us over %d launches' % (T, len(rows)))


if __name__ == '__main__':
    main()
