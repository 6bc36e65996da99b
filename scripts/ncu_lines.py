"""Per-source-line totals of an .ncu-rep source page (instructions executed, stall samples), joined with the line table of
the kernel's cubin (nvdisasm -g) by instruction order.
    python scripts/ncu_lines.py <file.ncu-rep> <object.o> <kernel-symbol-substring> [top]"""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile

rep, obj, sym = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ix, ss = hdr.index('Instructions Executed'), hdr.index('# Samples')
ins = [(r[1].strip(), int(r[ix] or 0), int(r[ss] or 0)) for r in rows[2:] if len(r) > ix]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for ln in dis:
    if ln.startswith('\t.section') or ln.startswith('//---'):
        on = sym in ln if '.text.' in ln else (on if not ln.startswith('//---') else on)
        if '.text.' in ln:
            on = sym in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
if len(lines) != len(ins):
    print('warning: %d disassembled instructions vs %d profiled' % (len(lines), len(ins)))
tot_i, tot_s = collections.Counter(), collections.Counter()
for (txt, n, s), loc in zip(ins, lines):
    tot_i[loc] += n
    tot_s[loc] += s
ti, ts = sum(tot_i.values()), sum(tot_s.values())
src = {}
print('total warp-instructions %d, samples %d' % (ti, ts))
for loc, n in tot_i.most_common(top):
    if loc and loc[0] not in src:
        for root in ('hamiltorch_b200/csrc', '.'):
            p = os.path.join(root, loc[0])
            if os.path.exists(p):
                src[loc[0]] = open(p).read().splitlines()
                break
    text = src.get(loc[0], [''] * 100000)[loc[1] - 1].strip()[:90] if loc else ''
    print('%5.1f%% inst %5.1f%% samples  %s:%s  %s' % (100.0 * n / ti, 100.0 * tot_s[loc] / max(ts, 1), loc[0] if loc else '?', loc[1] if loc else '?', text))
