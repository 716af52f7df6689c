"""Runs the UNMODIFIED reference ``LCRNet`` (experiments/lcrnet/model_family/LCRNet.py) on a seeded
synthetic scan pair with the seeded weights of ``checkpoint.random_state_dict('lcrnet')`` and
writes tests/golden/pair_golden.npz.  Build container only (needs /root/reference)."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

LIMITS = [30, 30, 30, 30]
CASE = (3, 7351, 6)      # scene seed, noise seed, raw-point stride


def make_pair_data(scene, seed, stride):
    from lcrnet_b200 import synth
    ref, src, T = synth.make_pair(scene, seed)
    return np.ascontiguousarray(ref[::stride]), np.ascontiguousarray(src[::stride]), T


def reference_pair_forward(raw_ref, raw_src, sd, limits=LIMITS):
    import ref_import
    ref_import.install()
    from experiments.lcrnet.data import precompute_data_stack_mode
    from experiments.lcrnet.modules.ops import grid_subsample
    from experiments.lcrnet.model_family.LCRNet import create_model
    cfg = ref_import.model_cfg(limits, tempfile.mkdtemp())
    pts = torch.from_numpy(np.concatenate([raw_ref, raw_src], 0))
    lens = torch.tensor([len(raw_ref), len(raw_src)], dtype=torch.int64)
    p0, l0 = grid_subsample(pts, lens, voxel_size=0.3)
    data = precompute_data_stack_mode(p0, l0, cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                      cfg.backbone.init_radius, limits)
    data['features'] = torch.ones(p0.shape[0], 1)
    data['batch_size'] = 1
    data = {k: ([x.contiguous() for x in v] if isinstance(v, list) else v) for k, v in data.items()}
    torch.manual_seed(7351)
    np.random.seed(7351)
    model = create_model(cfg).eval()
    model.load_state_dict(sd, strict=True)
    taps = {}
    tap = lambda name: (lambda mod, i, o: taps.__setitem__(name, o))
    hooks = [model.transformer.register_forward_hook(tap('transformer')),
             model.node_optimal_transport.register_forward_hook(tap('node_ot')),
             model.optimal_transport.register_forward_hook(tap('point_ot')),
             model.kpdecoder.register_forward_hook(tap('kpdecoder')),
             model.encoder.register_forward_hook(lambda mod, i, o: taps.__setitem__('feats_c', o[-1].clone()))]
    with torch.no_grad():
        out = model(data)
    for h in hooks:
        h.remove()
    return data, out, taps


def make_pair_golden():
    from lcrnet_b200 import checkpoint
    sd = checkpoint.random_state_dict('lcrnet', seed=7351)
    raw_ref, raw_src, T = make_pair_data(*CASE)
    data, out, taps = reference_pair_forward(raw_ref, raw_src, sd)
    g = {'case': np.array(CASE), 'limits': np.array(LIMITS), 'weight_seed': 7351,
         'lengths': np.stack([l.numpy() for l in data['lengths']]),
         'estimated_transform': out['estimated_transform'].numpy(),
         'pos_feature_global': out['pos_feature_global'].numpy(), 'anc_feature_global': out['anc_feature_global'].numpy(),
         'node_counts': np.array([int(x) for x in out['length']]),
         'pos_points_c': out['pos_points_c'].numpy(), 'anc_points_c': out['anc_points_c'].numpy(),
         'pos_node_corr_indices': out['pos_node_corr_indices'].numpy(),
         'anc_node_corr_indices': out['anc_node_corr_indices'].numpy(),
         'n_corr': out['corr_scores'].shape[0], 'corr_scores_sum': float(out['corr_scores'].double().sum()),
         'enhanced_pos_head': taps['transformer'][0][0, :8].numpy(), 'enhanced_anc_head': taps['transformer'][1][0, :8].numpy(),
         'feats_f_head': taps['kpdecoder'][0][:8].numpy(), 'node_ot_diag': taps['node_ot'][0].diagonal().numpy()}
    path = os.path.join(HERE, 'pair_golden.npz')
    np.savez_compressed(path, **g)
    print('nodes', g['node_counts'], 'node pairs', len(g['pos_node_corr_indices']), 'corr', g['n_corr'])
    print(g['estimated_transform'])
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    make_pair_golden()
