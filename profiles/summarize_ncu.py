#!/usr/bin/env python
"""Turn the two ncu outputs of a round into the markdown tables kept under profiles/.

    python profiles/summarize_ncu.py <launch list .csv> <raw page .csv of the --set full capture> [passes]

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python bench.py ...
  raw page    : ncu -i X.ncu-rep --page raw --csv > Y.csv          (run in the build container)
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'^void ', '', name)
    m = re.match(r'skb_foreach_kernel<(\w+)>', name)
    if m:
        return f'skb_foreach_kernel<{m.group(1)}>'
    return name.split('(')[0]


def launch_table(path, passes):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg = OrderedDict()
    for r in rows:
        n = short(r['Kernel Name'])
        t = float(r['Metric Value']) / (1e3 if r['Metric Unit'] in ('ns', 'nsecond') else 1.0)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    lib = {k: v for k, v in agg.items() if k.startswith('skb_')}
    other = sum(v[1] for k, v in agg.items() if not k.startswith('skb_'))
    tot = sum(v[1] for v in lib.values())
    out = ['| kernel | launches | total us | share |', '|---|---|---|---|']
    for k, v in sorted(lib.items(), key=lambda kv: -kv[1][1]):
        out.append(f'| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |')
    out.append('')
    out.append(f'Library kernels: {tot:.0f} us over {passes} passes = {tot / passes:.0f} us per pass; other kernels in the '
               f'capture (torch fill / index kernels of the input generation): {other:.0f} us.')
    return '\n'.join(out)


def full_table(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
            ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
            ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
            ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
            ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active lanes'),
            ('launch__registers_per_thread', 'regs'), ('lts__t_sector_hit_rate.pct', 'L2 hit %')]
    idx = [(hdr.index(c), lbl) for c, lbl in cols if c in hdr]
    name_i = hdr.index('Kernel Name')
    best = OrderedDict()
    t_i = hdr.index('gpu__time_duration.sum')
    for d in data:   # the longest launch of each kernel
        n = short(d[name_i])
        if n not in best or float(d[t_i]) > float(best[n][t_i]):
            best[n] = d
    out = ['| kernel | ' + ' | '.join(l for _, l in idx) + ' |', '|---|' + '---|' * len(idx)]
    for n, d in sorted(best.items(), key=lambda kv: -float(kv[1][t_i])):
        cells = []
        for i, _ in idx:
            v, u = d[i], units[i]
            try:
                cells.append(f'{float(v):.4g} {u}'.strip())
            except ValueError:
                cells.append(v)
        out.append(f'| `{n}` | ' + ' | '.join(cells) + ' |')
    return '\n'.join(out)


if __name__ == '__main__':
    passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    print('## Launch list (share of the library\'s device time; cold-cache, serialised: compare SHARES)\n')
    print(launch_table(sys.argv[1], passes))
    print('\n## `ncu --set full`, the longest launch of each kernel in one pass\n')
    print(full_table(sys.argv[2]))
