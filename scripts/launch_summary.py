"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) -> per-kernel table
    python scripts/launch_summary.py gpurun_out/rNN_launches.csv "header comment" > profiles/rNN_launches.txt"""
import csv
import collections
import re
import sys


def main(path, title=''):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = None
    agg = collections.OrderedDict()
    n = 0
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'^\(anonymous namespace\)::|^<unnamed>::', '', d['Kernel Name'])
        name = re.sub(r'\(.*$', '', name)
        val = float(d['Metric Value'].replace(',', ''))
        unit = d.get('Metric Unit', 'ns')
        ms = val / 1e6 if unit in ('ns', 'nsecond') else (val / 1e3 if unit in ('us', 'usecond') else val)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        n += 1
    total = sum(a[1] for a in agg.values())
    if title:
        print('# ' + title)
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)')
    print('# launches %d, total %.2f ms' % (n, total))
    print('%-46s %6s %10s %6s' % ('kernel', 'calls', 'ms', 'share'))
    for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-46s %6d %10.3f %5.1f%%' % (name[:46], c, ms, 100 * ms / total))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
