"""Registration evaluator metrics (lcrnet_b200.metrics) against the reference's own ``Evaluator``: the committed
fixture tests/golden/metrics_golden.npz was produced by running the unmodified reference
(tests/golden/make_metrics_golden.py); when /root/reference is present the reference is also run live."""
import os

import numpy as np
import pytest
import torch

from lcrnet_b200 import metrics as M
from util import GOLDEN, REF_PRESENT

G = np.load(os.path.join(GOLDEN, 'metrics_golden.npz'))


def _case(seed):
    t = lambda k: torch.from_numpy(G['in%d_%s' % (seed, k)])
    od = {'pos_points_c': torch.zeros(int(G['in%d_npos' % seed]), 3), 'anc_points_c': torch.zeros(int(G['in%d_nanc' % seed]), 3),
          'pos_node_corr_indices': t('pos_idx'), 'anc_node_corr_indices': t('anc_idx'),
          'pos_corr_points': t('pos_corr'), 'anc_corr_points': t('anc_corr'), 'estimated_transform': t('est')}
    return od, t('gt'), t('gt_idx'), t('gt_ov')


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_evaluator_matches_reference_fixture(seed):
    od, gt, gti, gto = _case(seed)
    res = M.evaluate(od, gt, gti, gto)
    for k in ('PIR', 'IR', 'RRE', 'RTE', 'RR'):
        want = float(G['out%d_%s' % (seed, k)])
        assert abs(float(res[k]) - want) <= 1e-6 * max(1.0, abs(want)), (k, float(res[k]), want)


def test_fixture_covers_both_recall_outcomes():
    assert {float(G['out%d_RR' % s]) for s in range(4)} == {0.0, 1.0}


def test_error_measures_basic_properties():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(3, 3, generator=g, dtype=torch.float64)
    R = torch.linalg.matrix_exp(a - a.t())
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3], T[:3, 3] = R, torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    rre, rte = M.isotropic_transform_error(T, T)
    assert float(rre) < 1e-5 and float(rte) == 0.0
    ang = 0.3
    Rz = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float64)
    T2 = T.clone()
    T2[:3, :3] = R @ Rz
    T2[:3, 3] += torch.tensor([0.0, 3.0, 4.0], dtype=torch.float64)
    rre, rte = M.isotropic_transform_error(T, T2)
    assert abs(float(rre) - np.degrees(ang)) < 1e-6 and abs(float(rte) - 5.0) < 1e-12
    assert float(M.registration_recall(torch.tensor(4.9), torch.tensor(1.9))) == 1.0
    assert float(M.registration_recall(torch.tensor(5.0), torch.tensor(1.9))) == 0.0      # strict <
    stacked = M.isotropic_transform_error(torch.stack([T, T]), torch.stack([T, T2]), reduction='none')
    assert stacked[0].shape == (2,) and float(stacked[1][1]) == pytest.approx(5.0)


@pytest.mark.skipif(not REF_PRESENT, reason='reference tree not present')
def test_evaluator_matches_reference_live():
    import sys
    sys.path.insert(0, GOLDEN)
    import make_metrics_golden as mk
    import ref_import
    ref_import.install()
    from experiments.lcrnet.loss_reg import Evaluator
    cfg = ref_import._EasyDict({'eval': {'acceptance_overlap': 0.0, 'acceptance_radius': 1.0, 'rre_threshold': 5.0,
                                         'rte_threshold': 2.0}})
    ev = Evaluator(cfg)
    for seed in (7, 8):
        c = mk.make_case(seed)
        od = {'pos_points_c': torch.zeros(c['npos'], 3), 'anc_points_c': torch.zeros(c['nanc'], 3),
              'gt_node_corr_overlaps': c['gt_ov'], 'gt_node_corr_indices': c['gt_idx'],
              'pos_node_corr_indices': c['pos_idx'], 'anc_node_corr_indices': c['anc_idx'],
              'pos_corr_points': c['pos_corr'], 'anc_corr_points': c['anc_corr'], 'estimated_transform': c['est']}
        want = ev(od, {'transform': c['gt']})
        got = M.evaluate(od, c['gt'], c['gt_idx'], c['gt_ov'])
        for k in want:
            assert abs(float(got[k]) - float(want[k])) <= 1e-6 * max(1.0, abs(float(want[k]))), k


@pytest.mark.gpu
def test_evaluate_on_model_output():
    """Evaluator metrics on a real LCRNet output dict (synthetic pair with known ground truth; random weights, so the
    values are not meaningful -- shapes, ranges and device handling are)."""
    from lcrnet_b200 import demo, synth
    pos, anc, gt = synth.make_pair(1, 11)
    out = demo.run_pair(np.ascontiguousarray(pos[::4]), np.ascontiguousarray(anc[::4]), pre_voxel=0.3)
    res = M.evaluate(out, gt)
    assert 0.0 <= float(res['IR']) <= 1.0 and 0.0 <= float(res['RRE']) <= 180.0 and float(res['RTE']) >= 0.0
    assert float(res['RR']) in (0.0, 1.0)
    same = dict(out, estimated_transform=torch.as_tensor(gt, dtype=torch.float32, device=out['estimated_transform'].device))
    res = M.evaluate(same, gt)
    assert float(res['RRE']) < 0.05 and float(res['RTE']) < 1e-5 and float(res['RR']) == 1.0
