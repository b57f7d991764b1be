"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list per (kernel, grid)."""
import collections, csv, sys

def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1000 if unit == 'ns' else v * 1000 if unit == 'ms' else v
        key = (row['Kernel Name'].split('(')[0][-48:], row.get('Grid Size', ''))
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    print('%-50s %-18s %4s %10s %9s %6s' % ('kernel', 'grid', 'n', 'total_us', 'avg_us', 'share'))
    for (name, grid), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-50s %-18s %4d %10.1f %9.2f %5.1f%%' % (name, grid, c, t, t / c, 100 * t / tot))
    print('total %.1f us over %d launches' % (tot, sum(c for c, _ in agg.values())))

if __name__ == '__main__':
    main(sys.argv[1])
