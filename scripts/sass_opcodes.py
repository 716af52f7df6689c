"""Per-kernel counts of the SASS opcodes that prove (or disprove) a Blackwell-native kernel
(B200_PROFILING.md "What proves a Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld / st -> LDTM / STTM,
TMA -> UTMALDG / UTMASTG (tensor maps) / UBLKCP (bulk copy), cp.async -> LDGSTS, mma.sync -> HMMA, plus FFMA for scale.
Usage: python scripts/sass_opcodes.py [lcr-net_b200/liblcr_b200.so] > profiles/rN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, 'lcr-net_b200', 'liblcr_b200.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'HMMA', 'SYNCS',
       'USETMAXREG', 'FFMA', 'MUFU', 'SHFL', 'BAR', 'ATOM', 'RED']
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        raw = m.group(1)
        dem = subprocess.run(['c++filt', raw], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::', '', dem)
        name = re.sub(r'\(.*', '', name)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        counts[name]['_total'] += 1
        for o in OPS:
            if op == o or op.startswith(o + '.') or (o == 'HMMA' and op.startswith('HMMA')):
                counts[name][o] += 1
print('# SASS opcode counts per kernel of %s (cuobjdump -sass, sm_100a)' % os.path.relpath(lib, REPO))
print('# %-72s %7s ' % ('kernel', 'instr') + ' '.join('%8s' % o for o in OPS))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    print('%-74s %7d ' % (k[:74], c['_total']) + ' '.join('%8d' % c[o] for o in OPS))
print('%-74s %7d ' % ('TOTAL', tot['_total']) + ' '.join('%8d' % tot[o] for o in OPS))
