#!/usr/bin/env python
"""Print the per-kernel durations of one build + walk out of an `ncu --metrics gpu__time_duration.sum --csv` launch list
(the last step in the file).  usage: python tools/launch_list.py gpurun_out/x.csv [first-kernel-substring]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
seq = [(r[ki].split('(')[0], float(r[vi].replace(',', ''))) for r in rows[1:]]
first = sys.argv[2] if len(sys.argv) > 2 else 'top_rank'
idx = [i for i, (n, _) in enumerate(seq) if first in n]
start = idx[-1] if idx else 0
tot = 0.0
for n, v in seq[start:]:
    print(f"{n[:64]:64s} {v / 1000:9.1f} us"); tot += v
    if 'walk_kernel' in n: break
print(f"{'sum (serialised, cold caches)':64s} {tot / 1000:9.1f} us")
