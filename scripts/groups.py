"""Print the kernel-group table of bench.py JSON lines.  Usage: python scripts/groups.py LOG..."""
import json
import sys
for f in sys.argv[1:]:
    for ln in open(f):
        if ln.startswith('{'):
            d = json.loads(ln)
            kg = (d.get('roofline') or {}).get('kernel_groups') or d.get('kernel_groups')
            e2e = d.get('e2e') or {}
            print('%s: %.1f %s, %.3f ms/step, e2e %.1f, launches %s' % (f, d['value'], d['unit'], d['ms_per_step'],
                                                                   e2e.get('value', float('nan')), d.get('gpu_launches')))
            for k, v in sorted((kg or {}).items(), key=lambda kv: -kv[1]['ms']):
                print('    %-22s %8.4f ms %4d launches' % (k, v['ms'], v['launches']))
