"""Generates tests/golden/metrics_golden.npz by running the UNMODIFIED reference evaluator
(experiments/lcrnet/loss_reg.py:278-334 ``Evaluator``, modules/registration/metrics.py) on seeded inputs.
Run in the build container (needs /root/reference):  python tests/golden/make_metrics_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_import  # noqa: E402


def make_case(seed):
    g = torch.Generator().manual_seed(seed)
    def rot(scale):
        a = torch.randn(3, 3, generator=g, dtype=torch.float64) * scale
        return torch.linalg.matrix_exp(a - a.t())
    gt = torch.eye(4, dtype=torch.float64)
    gt[:3, :3] = rot(0.5)
    gt[:3, 3] = torch.randn(3, generator=g, dtype=torch.float64) * 5
    est = torch.eye(4, dtype=torch.float64)
    est[:3, :3] = gt[:3, :3] @ rot(0.02 * (seed + 1))
    est[:3, 3] = gt[:3, 3] + torch.randn(3, generator=g, dtype=torch.float64) * 0.5 * (seed + 1)
    n = 200
    anc = torch.randn(n, 3, generator=g) * 20
    pos = anc @ gt[:3, :3].float().t() + gt[:3, 3].float() + torch.randn(n, 3, generator=g) * 0.8
    npos, nanc, c = 40, 37, 60
    pi = torch.randint(0, npos, (c,), generator=g)
    ai = torch.randint(0, nanc, (c,), generator=g)
    gti = torch.stack([torch.randint(0, npos, (90,), generator=g), torch.randint(0, nanc, (90,), generator=g)], 1)
    gti[:30, 0], gti[:30, 1] = pi[:30], ai[:30]
    gto = torch.rand(90, generator=g) - 0.2
    return {'gt': gt.float(), 'est': est.float(), 'pos_corr': pos, 'anc_corr': anc, 'pos_idx': pi, 'anc_idx': ai,
            'gt_idx': gti, 'gt_ov': gto, 'npos': npos, 'nanc': nanc}


def main():
    ref_import.install()
    from experiments.lcrnet.loss_reg import Evaluator
    cfg = ref_import._EasyDict({'eval': {'acceptance_overlap': 0.0, 'acceptance_radius': 1.0, 'rre_threshold': 5.0,
                                         'rte_threshold': 2.0}})
    ev = Evaluator(cfg)
    out = {}
    for seed in range(4):
        c = make_case(seed)
        od = {'pos_points_c': torch.zeros(c['npos'], 3), 'anc_points_c': torch.zeros(c['nanc'], 3),
              'gt_node_corr_overlaps': c['gt_ov'], 'gt_node_corr_indices': c['gt_idx'],
              'pos_node_corr_indices': c['pos_idx'], 'anc_node_corr_indices': c['anc_idx'],
              'pos_corr_points': c['pos_corr'], 'anc_corr_points': c['anc_corr'], 'estimated_transform': c['est']}
        res = ev(od, {'transform': c['gt']})
        for k, v in c.items():
            out['in%d_%s' % (seed, k)] = np.asarray(v)
        for k, v in res.items():
            out['out%d_%s' % (seed, k)] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'metrics_golden.npz'), **out)
    print({k: float(v) for k, v in out.items() if k.startswith('out')})


if __name__ == '__main__':
    main()
