#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS, in program order) with
`nvdisasm --print-line-info` of the same function.  usage: ncu_lines.py sass.csv nvdisasm.txt [top] [section]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if len(r) > 5 and 'Source' in r]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # which kernel of the report (several launches: several sections)
hi = his[which]
rows = rows[:his[which + 1] - 1] if which + 1 < len(his) else rows
hdr = rows[hi]
si = hdr.index('Source'); wi = hdr.index('Warp Stall Sampling (All Samples)'); ii = hdr.index('Instructions Executed'); ti = hdr.index('Thread Instructions Executed')
sass = [(r[si].strip(), int(r[wi] or 0), int(r[ii] or 0), int(r[ti] or 0)) for r in rows[hi + 1:] if len(r) > max(si, wi, ii)]
line = None; fname = None; dis = []
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: fname, line = m.group(1).split('/')[-1], int(m.group(2)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: dis.append((fname, line, m.group(2)))
n = min(len(sass), len(dis))
if len(sass) != len(dis): print('# warning: %d SASS rows vs %d disassembled instructions' % (len(sass), len(dis)))
agg = {}
for k in range(n):
    key = (dis[k][0], dis[k][1]); a = agg.setdefault(key, [0, 0, 0]); a[0] += sass[k][1]; a[1] += sass[k][2]; a[2] += sass[k][3]
tot = sum(a[0] for a in agg.values()) or 1; toti = sum(a[1] for a in agg.values()) or 1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = {}
print('# samples %d, warp instructions %d' % (tot, toti))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        try: src[f] = open('/root/repo/starflood_b200/csrc/' + f).read().split('\n')
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:110] if ln and ln - 1 < len(src[f]) else ''
    print('%5.1f%% stall %5.1f%% inst  lanes %4.1f  %s:%d  %s' % (100 * a[0] / tot, 100 * a[1] / toti, a[2] / max(a[1], 1), f, ln or 0, text))
