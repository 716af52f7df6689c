# Diagnostic: repeat the registration workload and print the per-step times (looking for sporadic slow steps)
for i in 1 2 3 4 5 6 7 8 9 10; do timeout 300 python bench.py --workload pairs --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); w=r['workloads']['pairs'] if 'workloads' in r else r
print(round(w['value'],1), w['ms_steps'], w['ms_steps_e2e'])"; done
