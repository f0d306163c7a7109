"""Per-source-line stall samples / executed instructions of one kernel in an ncu report.

    python scripts/ncu_lines.py <report.ncu-rep> <object.o> <kernel-substring> [top]

ncu's CSV source page is per SASS instruction; nvdisasm --print-line-info gives the source line of each SASS
instruction of the same cubin in the same order.  Joined by instruction index.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# lines of the wanted function
lines_of = []
cur_line, active = None, False
for ln in dis.splitlines():
    if ln.startswith('.text.'):
        active = kname in ln
        continue
    if ln.startswith('//---') or ln.lstrip().startswith('.section'):
        if '.text.' in ln and kname not in ln:
            active = False
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        inl = 'inlined' in ln
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
        lines_of.append(cur_line)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5 and r[0].startswith('0x')]
iN, iE = hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
if len(data) != len(lines_of):
    print('warning: %d SASS rows in report vs %d in disassembly (different build?)' % (len(data), len(lines_of)))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r, l in zip(data, lines_of):
    a = agg[l]
    a[0] += int(r[iN])
    a[1] += int(r[iE])
    for i, h in stall_cols:
        if r[i] and r[i] != '0':
            a[2][h[6:]] += int(r[i])
ts, te = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
print('# %s: %d samples, %d warp-instructions' % (kname, ts, te))
print('# samples%  exec%   file:line   top stalls')
for l, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ', '.join('%s %d' % kv for kv in a[2].most_common(3))
    print('%7.1f %7.1f   %s:%s   %s' % (100.0 * a[0] / ts, 100.0 * a[1] / te, l[0] if l else '?', l[1] if l else '?', st))
