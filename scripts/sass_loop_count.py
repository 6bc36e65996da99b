"""Static instruction count of a kernel's main loop from an object file (nvdisasm with line info), weighting the innermost
backward-branch loop by a trip count.  Usage: python scripts/sass_loop_count.py <obj.o> <mangled kernel substring> [inner trips]"""
import collections, os, pickle, re, subprocess, sys, tempfile

obj, pat = sys.argv[1], sys.argv[2]
trips = int(sys.argv[3]) if len(sys.argv) > 3 else 5
tmp = tempfile.mkdtemp()
subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
txt = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(txt) if l.startswith('.text.') and pat in l)
insts, labels, pend, cur = [], {}, [], None
for l in txt[start + 1:]:
    if l.startswith('\t.section') or l.startswith('.text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*(\.L_x_\d+):', l)
    if m:
        pend.append(m.group(1))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(\S.*?);', l)
    if m:
        a = int(m.group(1), 16)
        for p in pend:
            labels[p] = a
        pend = []
        insts.append((a, m.group(2), cur))
back = []
for a, t, c in insts:
    m = re.search(r'BRA.*`\((\.L_x_\d+)\)', t)
    if m and labels.get(m.group(1), 1 << 30) < a:
        back.append((labels[m.group(1)], a))
back.sort(key=lambda b: b[1] - b[0])
print('backward branches (start, end, static size):', [(hex(s), hex(e), sum(1 for a, _, _ in insts if s <= a <= e)) for s, e in back])
outer = max(back, key=lambda b: b[1] - b[0])
inner = [b for b in back if b != outer and outer[0] <= b[0] and b[1] <= outer[1]]
w = collections.Counter()
ops = collections.Counter()
for a, t, c in insts:
    if outer[0] <= a <= outer[1]:
        k = trips if any(s <= a <= e for s, e in inner) else 1
        w[c] += k
        ops[t.split()[1] if t.startswith('@') else t.split()[0]] += k
print('weighted loop instructions (all paths):', sum(w.values()))
print(sorted(ops.items(), key=lambda kv: -kv[1])[:25])
