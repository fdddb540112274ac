#!/usr/bin/env python
"""Aggregate an ncu per-launch CSV of ONE training step (tools/one_step.py under
`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
per kernel family -> profiles/ncu_traffic.json (read by bench.py for roofline.traffic) and a text table.

  python tools/ncu_traffic.py gpurun_out/r02_step_traffic.csv [--calls gpurun_out/calls.json] --out profiles/r02_step_traffic
"""
import csv
import json
import re
import sys

FAMILIES = [
    ('conv_fwd', r'conv_fwd_umma_persistent|conv_fwd3_umma_kernel|conv_fwd_umma_kernel|conv_fwd_simt_kernel|conv_fwd2cta'),
    ('conv_wgrad', r'conv_wgrad'),
    ('maxpool', r'maxpool_'),
    ('adam', r'adam_kernel'),
    ('gru_gates', r'gru_|lstm_'),
    ('softmax_ce', r'softmax_ce3d'),
    ('bias_grad', r'bias_grad'),
    ('lrelu', r'lrelu_'),
    ('other', r'.'),
]

UNIT = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0,
        'usecond': 1e-6, 'nsecond': 1e-9, 'msecond': 1e-3, 'second': 1.0}


def parse(path):
    """long-format ncu csv (one row per launch x metric) -> list of dict(id, kernel, metrics...)"""
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rd = csv.DictReader(lines)
    launches = {}
    for r in rd:
        i = int(r['ID'])
        d = launches.setdefault(i, {'id': i, 'kernel': r['Kernel Name']})
        unit = (r.get('Metric Unit') or '').lower()
        try:
            v = float(r['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        d[r['Metric Name']] = v * UNIT.get(unit, 1.0)
    return [launches[k] for k in sorted(launches)]


def main():
    args = sys.argv[1:]
    out = 'profiles/step_traffic'
    calls = None
    if '--out' in args:
        i = args.index('--out'); out = args[i + 1]; args = args[:i] + args[i + 2:]
    if '--calls' in args:
        i = args.index('--calls'); calls = json.load(open(args[i + 1])); args = args[:i] + args[i + 2:]
    launches = parse(args[0])
    fam = {}
    per_kernel = {}
    for L in launches:
        name = L['kernel']
        short = re.sub(r'\(.*', '', name).replace('void ', '').replace('r2n2::', '')
        key = next(k for k, pat in FAMILIES if re.search(pat, short))
        b = L.get('dram__bytes_read.sum', 0.0) + L.get('dram__bytes_write.sum', 0.0)
        t = L.get('gpu__time_duration.sum', 0.0)
        for table, k in ((fam, key), (per_kernel, short)):
            d = table.setdefault(k, dict(launches=0, dram_bytes_total=0.0, time_s=0.0))
            d['launches'] += 1
            d['dram_bytes_total'] += b
            d['time_s'] += t
    tot_t = sum(d['time_s'] for d in fam.values())
    lines = ['# %s: one training step, %d launches, ncu serialised time %.3f ms' % (args[0], len(launches), tot_t * 1e3),
             '%-34s %8s %12s %10s %8s %12s' % ('family / kernel', 'launches', 'dram MB', 'time ms', 'share', 'GB/s (ncu)')]
    for table in (fam, per_kernel):
        for k, d in sorted(table.items(), key=lambda kv: -kv[1]['time_s']):
            lines.append('%-34s %8d %12.1f %10.4f %8.3f %12.0f' % (k[:34], d['launches'], d['dram_bytes_total'] / 1e6,
                                                                  d['time_s'] * 1e3, d['time_s'] / tot_t,
                                                                  d['dram_bytes_total'] / max(d['time_s'], 1e-12) / 1e9))
        lines.append('')
    algo = None
    if calls:
        algo = {}
        for c in calls['calls']:
            k = 'conv_fwd' if c['name'].startswith('r2n2_conv_fwd') else ('conv_wgrad' if 'wgrad' in c['name'] else c['name'])
            a = algo.setdefault(k, dict(calls=0, bytes=0.0, flops=0.0, ms=0.0))
            a['calls'] += 1; a['bytes'] += c.get('bytes', 0.0); a['flops'] += c.get('flops', 0.0); a['ms'] += c['ms']
        lines.append('# algorithmic operand bytes / flops per C-ABI entry point (CUDA-event time of an un-profiled step)')
        for k, a in sorted(algo.items(), key=lambda kv: -kv[1]['ms']):
            lines.append('%-34s %8d %12.1f MB %10.4f ms %10.1f GFLOP' % (k[:34], a['calls'], a['bytes'] / 1e6, a['ms'], a['flops'] / 1e9))
    txt = '\n'.join(lines) + '\n'
    open(out + '.txt', 'w').write(txt)
    print(txt)
    import os
    tp = os.path.join(os.path.dirname(out) or '.', 'ncu_traffic.json')
    traffic = json.load(open(tp)) if os.path.exists(tp) else {}
    f = fam.get('conv_fwd')
    if f:
        traffic['conv_fwd_step'] = dict(launches=f['launches'], dram_bytes_total=f['dram_bytes_total'],
                                        ncu_time_ms=f['time_s'] * 1e3, share_of_step_ncu=f['time_s'] / tot_t,
                                        algorithmic_bytes=None if not algo else algo.get('conv_fwd', {}).get('bytes'),
                                        source=os.path.basename(args[0]))
    traffic['step_families'] = {k: dict(launches=d['launches'], dram_bytes_total=d['dram_bytes_total'],
                                        ncu_time_ms=d['time_s'] * 1e3) for k, d in fam.items()}
    json.dump(traffic, open(tp, 'w'), indent=1)


if __name__ == '__main__':
    main()
