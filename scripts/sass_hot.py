"""Hot spots of an ncu source page (sass csv): per-opcode instruction counts and top stall lines.
Usage: ncu -i X.ncu-rep --page source --csv --print-source sass > f.csv; python scripts/sass_hot.py f.csv [kernel_index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
# split per kernel
kernels = []
cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        kernels.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
k = kernels[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr = k['rows'][0]
body = [r for r in k['rows'][1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
print(k['name'][:100])
ops = collections.Counter()
samples = collections.Counter()
tot_inst = tot_samp = 0
for r in body:
    op = r[ci['Source']].split()
    op = [t for t in op if not t.startswith('@')]
    name = op[0].split('.')[0] if op else '?'
    n = int(r[ci['Instructions Executed']] or 0)
    s = int(r[ci['# Samples']] or 0)
    ops[name] += n
    samples[name] += s
    tot_inst += n
    tot_samp += s
print('instructions executed (warp-level): %d, samples %d' % (tot_inst, tot_samp))
for name, n in ops.most_common(22):
    print('  %-10s %10d %5.1f%%   samples %5.1f%%' % (name, n, 100.0 * n / tot_inst, 100.0 * samples[name] / max(tot_samp, 1)))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in body:
    for h in stall_cols:
        agg[h] += int(r[ci[h]] or 0)
print('stall reasons:', ', '.join('%s %.1f%%' % (h[6:], 100.0 * v / max(tot_samp, 1)) for h, v in agg.most_common(8)))
print('top lines by samples:')
for r in sorted(body, key=lambda r: -int(r[ci['# Samples']] or 0))[:14]:
    print('  %6s %5.1f%%  %s' % (r[ci['# Samples']], 100.0 * int(r[ci['# Samples']] or 0) / max(tot_samp, 1), r[ci['Source']][:90]))
