#!/usr/bin/env python
"""Opcode histogram of the hot kernels out of the shipped library: cuobjdump -sass per cubin, grouped by kernel.
usage: python tools/sass_hist.py starflood_b200/libsfengine.so kernel_substring [kernel_substring ...]"""
import collections, re, subprocess, sys, tempfile, os
so = os.path.abspath(sys.argv[1]); pats = sys.argv[2:]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
for cubin in sorted(os.listdir(tmp)):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    fn = None; hist = None
    def flush():
        if fn and hist and any(p in fn for p in pats):
            demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            tot = sum(hist.values())
            print("## %s  (%s, %d instructions)" % (demangled[:150], cubin, tot))
            groups = collections.Counter()
            for op, n in hist.items(): groups[op.split('.')[0]] += n
            print("   " + "  ".join("%s %d" % kv for kv in groups.most_common(28)))
            tma = {k: v for k, v in hist.items() if k.startswith(("UBLKCP", "SYNCS", "UTMA", "UTC", "LDG.E.ENL2.256", "STG.E.ENL2.256"))}
            if tma: print("   bulk copy / mbarrier / 256-bit: " + "  ".join("%s %d" % kv for kv in sorted(tma.items())))
    for line in txt.split("\n"):
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            flush(); fn = m.group(1); hist = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and hist is not None: hist[m.group(1)] += 1
    flush()
