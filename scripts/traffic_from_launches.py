"""DRAM traffic per kernel group from an ncu launch list that carries dram__bytes_read.sum and
dram__bytes_write.sum (same pass as the launch-time list; see profiles/README or DESIGN.md 6).
Writes the json bench.py reads for `roofline.traffic`.

Usage: python scripts/traffic_from_launches.py gpurun_out/launches.csv LAST_N_LAUNCHES profiles/rN_traffic.json
"""
import json
import sys

sys.path.insert(0, __file__.rsplit('/', 1)[0])
import launch_summary as ls  # noqa: E402

# kernel-name prefix -> LcrProfScope group (csrc/*.cu)
GROUPS = [('gemm_tf32x3', 'gemm_tf32x3'), ('gemm_kernel', 'gemm_f32'), ('kpconv_gather', 'kpconv_gather'),
          ('kpconv_c1', 'kpconv_c1'), ('query_kernel', 'radius_query'), ('query_self_kernel', 'radius_query'), ('spill_kernel', 'radius_query'),
          ('gn_partial', 'group_norm_stats'), ('gn_finalize', 'group_norm_stats'), ('gn_apply', 'group_norm_apply'),
          ('maxpool', 'maxpool'), ('hidden_partial', 'netvlad_hidden'), ('attention_tc', 'attention_tc'),
          ('sinkhorn', 'sinkhorn'), ('l2_topk', 'l2_topk')]



def main():
    path, last_n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rows = ls.load(path)[-last_n:]
    agg = {}
    for r in rows:
        name = ls.short(r['name'])
        for prefix, group in GROUPS:
            if name.startswith(prefix):
                a = agg.setdefault(group, {'launches': 0, 'dram_bytes': 0.0, 'ns': 0.0})
                a['launches'] += 1
                a['dram_bytes'] += r.get('dram__bytes_read.sum', 0.0) + r.get('dram__bytes_write.sum', 0.0)
                a['ns'] += r.get('gpu__time_duration.sum', 0.0)
                break
    for a in agg.values():
        a['dram_bytes_per_launch'] = a['dram_bytes'] / max(a['launches'], 1)
    json.dump({'source': path.rsplit('/', 1)[-1], 'launches_in_step': last_n, 'groups': agg}, open(out, 'w'), indent=1)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['ns']):
        print('%-20s %4d launches %9.1f us %10.1f MB' % (k, a['launches'], a['ns'] / 1e3, a['dram_bytes'] / 1e6))


if __name__ == '__main__':
    main()
