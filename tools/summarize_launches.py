#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share."""
import collections
import csv
import io
import sys


def main(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows:
        name = r['Kernel Name']
        if 'distribution' in name:
            short = 'torch random fill (input generation)'
        elif 'at::' in name:
            short = 'torch: ' + name.split('<')[0][-50:]
        else:
            short = name.split('(')[0]
        tot[short] += float(r['Metric Value'])
        cnt[short] += 1
    total = sum(tot.values())
    print(f'# {path}: {len(rows)} launches, {total / 1e6:.3f} ms total (cold-cache, serialised: compare shares)')
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f'{v / 1e6:10.3f} ms {100 * v / total:6.2f}%  x{cnt[k]:<4d} {k}')


if __name__ == '__main__':
    main(sys.argv[1])
