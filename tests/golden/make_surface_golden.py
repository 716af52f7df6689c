"""Records the call-surface contract of the UNMODIFIED reference models on the pair of pair_golden.npz:
output_dict keys, python types, dtypes and shapes of ``LCRNet.forward`` (model_family/LCRNet.py:274-321),
``LCRNet_Matching.forward`` (LCRNet_Matching_infer.py:261-287) and ``LCRNet_GlobalDescrition.forward``
(LCRNet_GlobalDescrition.py:60-74), plus the state_dict key lists.  Writes tests/golden/surface_golden.json.
Build container only (needs /root/reference)."""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)


def describe(v):
    if torch.is_tensor(v):
        return {'type': 'Tensor', 'dtype': str(v.dtype).replace('torch.', ''), 'shape': list(v.shape)}
    if isinstance(v, (tuple, list)):
        return {'type': type(v).__name__, 'items': [describe(x) for x in v]}
    return {'type': type(v).__name__}


def main():
    import ref_import
    from make_pair_golden import CASE, LIMITS, make_pair_data
    from lcrnet_b200 import checkpoint
    ref_import.install()
    from experiments.lcrnet.data import precompute_data_stack_mode
    from experiments.lcrnet.model_family import LCRNet as m_full
    from experiments.lcrnet.model_family import LCRNet_GlobalDescrition as m_glob
    from experiments.lcrnet.model_family import LCRNet_Matching_infer as m_match
    from experiments.lcrnet.modules.ops import grid_subsample
    cfg = ref_import.model_cfg(LIMITS, tempfile.mkdtemp())
    raw_ref, raw_src, _ = make_pair_data(*CASE)
    pts = torch.from_numpy(np.concatenate([raw_ref, raw_src], 0))
    lens = torch.tensor([len(raw_ref), len(raw_src)], dtype=torch.int64)
    p0, l0 = grid_subsample(pts, lens, voxel_size=0.3)
    data = precompute_data_stack_mode(p0, l0, cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                      cfg.backbone.init_radius, LIMITS)
    data['features'] = torch.ones(p0.shape[0], 1)
    data['batch_size'] = 1
    data = {k: ([x.contiguous() for x in v] if isinstance(v, list) else v) for k, v in data.items()}
    sd = checkpoint.random_state_dict('lcrnet', seed=7351)
    out = {}
    torch.manual_seed(7351)
    full = m_full.create_model(cfg).eval()
    full.load_state_dict(sd, strict=True)
    match = m_match.create_model(cfg).eval()
    missing = match.load_state_dict(sd, strict=False)
    with torch.no_grad():
        o_full, o_match = full(data), match(data)
    out['lcrnet'] = {k: describe(v) for k, v in o_full.items()}
    out['matching'] = {k: describe(v) for k, v in o_match.items()}
    out['matching_state_dict_keys'] = sorted(match.state_dict().keys())
    out['matching_unexpected_from_lcrnet'] = sorted(missing.unexpected_keys)
    out['matching_missing_from_lcrnet'] = sorted(missing.missing_keys)
    out['matching_equals_lcrnet'] = {k: bool(torch.equal(o_full[k], o_match[k])) for k in ('estimated_transform', 'corr_scores')}
    # single-scan descriptor model
    one = precompute_data_stack_mode(p0[:l0[0]].contiguous(), l0[:1].contiguous(), cfg.backbone.num_stages,
                                     cfg.backbone.init_voxel_size, cfg.backbone.init_radius, LIMITS)
    one['features'] = torch.ones(int(l0[0]), 1)
    one['batch_size'] = 1
    one = {k: ([x.contiguous() for x in v] if isinstance(v, list) else v) for k, v in one.items()}
    glob = m_glob.create_model(cfg).eval()
    glob.load_state_dict(checkpoint.random_state_dict('global_descriptor', seed=7351), strict=True)
    with torch.no_grad():
        o_glob = glob(one)
    out['global_descriptor'] = {k: describe(v) for k, v in o_glob.items()}
    path = os.path.join(HERE, 'surface_golden.json')
    json.dump(out, open(path, 'w'), indent=1, sort_keys=True)
    print('lcrnet keys', len(out['lcrnet']), 'matching keys', len(out['matching']), 'global', list(out['global_descriptor']))
    print('matching state dict', len(out['matching_state_dict_keys']), 'missing', out['matching_missing_from_lcrnet'],
          'unexpected', len(out['matching_unexpected_from_lcrnet']), out['matching_equals_lcrnet'])
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
