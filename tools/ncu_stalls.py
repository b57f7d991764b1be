"""Stall / issue summary of one kernel in an .ncu-rep captured with --set full --import-source on (needs the ncu CLI,
no GPU): key metrics, warp-state breakdown (smsp__average_warp*_per_issue_active), instruction mix, hottest SASS lines."""
import collections, csv, subprocess, sys


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[-1])), dict(zip(rows[0], rows[1]))


def sass(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[1], rows[2:]


def main(path):
    v, u = raw(path)
    print('==', path)
    print('kernel:', v.get('Kernel Name'), ' grid', v.get('launch__grid_size'), ' block', v.get('launch__block_size'))
    for k in ('gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
              'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_active',
              'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
              'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
              'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
              'smsp__average_warp_latency_per_inst_issued.ratio'):
        if k in v:
            print('  %-66s %s %s' % (k, v[k], u.get(k, '')))
    st = sorted(((float(x.replace(',', '')), k) for k, x in v.items()
                 if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio') and x),
                reverse=True)
    print('  warp cycles per issued instruction, by stall reason (top 8):')
    for x, k in st[:8]:
        print('    %-40s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), x))
    hdr, data = sass(path)
    isrc, iex, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    mix = collections.Counter()
    for r in data:
        op = r[isrc].strip().split()
        if op:
            o = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
            mix[o.split('.')[0]] += int(r[iex] or 0)
    tot = sum(mix.values())
    print('  warp instructions executed: %d; mix: %s' % (tot, ', '.join('%s %.1f%%' % (k, 100.0 * c / tot)
                                                                       for k, c in mix.most_common(12))))
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    ts = sum(int(r[isamp] or 0) for r in data)
    print('  hottest SASS lines by sampled warp cycles (%d samples):' % ts)
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:10]:
        top = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[0]
        print('    %5.1f%%  %-56s %s' % (100.0 * int(r[isamp]) / max(ts, 1), r[isrc].strip()[:56], top[1]))


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
