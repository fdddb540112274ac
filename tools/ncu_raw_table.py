#!/usr/bin/env python
"""`ncu -i rep --page raw --csv` (one row per profiled launch) -> a compact per-launch table of the numbers the
roofline discussion uses.   python tools/ncu_raw_table.py gpurun_out/x.raw.csv [...] > profiles/xx.txt"""
import csv
import re
import sys

COLS = [
    ('time us', 'gpu__time_duration.sum', 1.0),
    ('tensor %', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 1.0),
    ('smem-tc %', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 1.0),
    ('lts %', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 1.0),
    ('dram %', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1.0),
    ('rd MB', 'dram__bytes_read.sum', 1.0),
    ('wr MB', 'dram__bytes_write.sum', 1.0),
    ('regs', 'launch__registers_per_thread', 1.0),
    ('smem KB', 'launch__shared_mem_per_block_dynamic', 1.0),
    ('GHz', 'smsp__cycles_elapsed.avg.per_second', 1.0),
]
SCALE = {'byte': 1e-6, 'kbyte': 1e-3, 'mbyte': 1.0, 'gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0,
         'msecond': 1e3, 'second': 1e6, 'hz': 1e-9, 'khz': 1e-6, 'mhz': 1e-3, 'ghz': 1.0, 'kbyte/block': 1.0, 'byte/block': 1e-3}


def short(name):
    m = re.match(r'(?:void )?(?:r2n2::)?(\w+)(<.*>)?', name)
    base = m.group(1)
    tp = (m.group(2) or '').replace('__nv_bfloat16', 'bf16')
    return base + tp


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        print('# %s' % path)
        print('%-4s %-62s %-12s ' % ('id', 'kernel', 'grid') + ' '.join('%9s' % c[0] for c in COLS))
        for r in rows[2:]:
            vals = []
            for label, key, _ in COLS:
                if key not in ix or r[ix[key]] == '':
                    vals.append('%9s' % '-')
                    continue
                v = float(r[ix[key]].replace(',', ''))
                u = units[ix[key]].lower()
                if label in ('time us', 'rd MB', 'wr MB', 'GHz', 'smem KB'):
                    v *= SCALE.get(u, 1.0)
                vals.append('%9.1f' % v if label != 'GHz' else '%9.2f' % v)
            print('%-4s %-62s %-12s ' % (r[ix['ID']], short(r[ix['Kernel Name']])[:62], r[ix['Grid Size']].replace(' ', '')) + ' '.join(vals))
        print()


if __name__ == '__main__':
    main()
