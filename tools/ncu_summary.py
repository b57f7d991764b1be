"""Key metrics of every kernel in an .ncu-rep (needs the ncu CLI; no GPU)."""
import csv, subprocess, sys, json
WANT = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = r[i] + ((' ' + units[i]) if units[i] else '')
        res.append(d)
    return res
if __name__ == '__main__':
    for p in sys.argv[1:]:
        print('==', p)
        for d in main(p):
            print(json.dumps(d))
