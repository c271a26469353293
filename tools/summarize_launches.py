"""Summarise an ncu launch list (tools/gpu_launchlist.sh): one training step, per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches_resnet50.csv [out.txt]"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    cur = {}
    for r in rd:
        key = r['ID']
        if key not in cur:
            cur[key] = dict(id=int(key), name=r['Kernel Name'], grid=r.get('Grid Size', ''), block=r.get('Block Size', ''))
            rows.append(cur[key])
        v = float(r['Metric Value'].replace(',', '')) if r['Metric Value'] not in ('', 'n/a') else 0.0
        unit = r['Metric Unit']
        m = r['Metric Name']
        if m == 'gpu__time_duration.sum':
            cur[key]['us'] = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
        else:
            scale = dict(byte=1, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9).get(unit, 1)
            cur[key][m] = v * scale
    marks = [i for i, r in enumerate(rows) if 'softmax_ce_rows' in r['name']]
    if len(marks) >= 2:
        rows = rows[marks[0]:marks[1]]
    short = lambda n: re.sub(r'\(.*', '', re.sub(r'^void ', '', re.sub(r'\(anonymous namespace\)::', '', n)))
    agg = OrderedDict()
    for r in rows:
        k = short(r['name'])
        a = agg.setdefault(k, defaultdict(float))
        a['n'] += 1
        a['us'] += r.get('us', 0.0)
        a['rd'] += r.get('dram__bytes_read.sum', 0.0)
        a['wr'] += r.get('dram__bytes_write.sum', 0.0)
    tot = sum(a['us'] for a in agg.values())
    out = ['one training step: %d launches, %.2f ms serialised device time' % (len(rows), tot / 1e3),
           '%-52s %6s %10s %7s %10s %10s' % ('kernel', 'n', 'ms', 'share', 'dram rd GB', 'dram wr GB')]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        out.append('%-52s %6d %10.3f %6.1f%% %10.3f %10.3f' % (k[:52], a['n'], a['us'] / 1e3, 100 * a['us'] / tot,
                                                               a['rd'] / 1e9, a['wr'] / 1e9))
    out.append('')
    out.append('top 25 single launches:')
    for r in sorted(rows, key=lambda r: -r.get('us', 0.0))[:25]:
        out.append('  %-44s grid %-18s %9.1f us  rd %7.1f MB  wr %7.1f MB' % (
            short(r['name'])[:44], r['grid'], r.get('us', 0.0), r.get('dram__bytes_read.sum', 0) / 1e6,
            r.get('dram__bytes_write.sum', 0) / 1e6))
    conv = [a for k, a in agg.items() if 'conv_tc' in k or 'tc_splitk' in k or 'tc_prep' in k or 'im2col' in k or 'split_bf16' in k]
    conv_bytes = sum(a['rd'] + a['wr'] for a in conv)
    out.append('')
    out.append('conv stack (conv_tc_* + prep/split/reduce helpers): %.3f ms, %.2f GB DRAM traffic per step' % (
        sum(a['us'] for a in conv) / 1e3, conv_bytes / 1e9))
    if len(sys.argv) > 3:
        import json
        json.dump({'workload': 'resnet50_uq8_dst_b256', 'batch': 256, 'conv_dram_bytes_per_step': conv_bytes,
                   'conv_serialised_ms': sum(a['us'] for a in conv) / 1e3,
                   'source': 'ncu launch list (tools/gpu_launchlist.sh), one training step'}, open(sys.argv[3], 'w'))
    text = '\n'.join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')


if __name__ == '__main__':
    main()
