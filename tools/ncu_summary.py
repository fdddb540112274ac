#!/usr/bin/env python
"""Condense `ncu --set full` reports (.ncu-rep) into the few numbers the roofline discussion needs and
write profiles/<name>_summary.txt + profiles/ncu_traffic.json (read by bench.py for roofline.traffic).

  python tools/ncu_summary.py gpurun_out/a.ncu-rep[:key] ... --out profiles/r01   (needs the ncu CLI, no GPU)
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sectors_srcunit_tex.sum',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_uniform.sum',
    'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
    'smsp__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
    'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__clocks.avg.per_second'.replace('clocks', 'cycles_elapsed'),
]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    u = unit.lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}.get(u, 1)


def main():
    args = sys.argv[1:]
    out = 'profiles/ncu'
    if '--out' in args:
        i = args.index('--out')
        out = args[i + 1]
        args = args[:i] + args[i + 2:]
    traffic_path = os.path.join(os.path.dirname(out) or '.', 'ncu_traffic.json')
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for a in args:
        rep, _, key = a.partition(':')
        txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        name = os.path.splitext(os.path.basename(rep))[0]
        lines = ['# %s (ncu --set full --clock-control none; cold-cache, serialised replays)' % rep]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            lines.append('kernel: %s grid %s block %s' % (d.get('Kernel Name'), d.get('Grid Size'), d.get('Block Size')))
            for k in KEYS:
                if k in d and d[k] != '':
                    lines.append('  %-85s %s %s' % (k, d[k], u[k]))
            rd = to_bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum'])
            wr = to_bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
            lines.append('  dram traffic per launch: %.1f MB (read %.1f + write %.1f)' % ((rd + wr) / 1e6, rd / 1e6, wr / 1e6))
            if key:
                traffic[key] = {'dram_bytes_per_launch': rd + wr, 'kernel': d.get('Kernel Name'),
                                'duration_us': d.get('gpu__time_duration.sum'), 'report': os.path.basename(rep)}
        with open('%s_%s_summary.txt' % (out, name), 'w') as f:
            f.write('\n'.join(lines) + '\n')
        print('\n'.join(lines))
    with open(traffic_path, 'w') as f:
        json.dump(traffic, f, indent=1)


if __name__ == '__main__':
    main()
