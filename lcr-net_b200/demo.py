"""Two-scan demo on the B200 backend: the flow of the reference's ``demo/demo.py:16-116`` (load two
0.3 m-voxelised scans, collate them as one registration pair, run ``LCRNet``, print the global
descriptor L2 distance and the estimated 4x4 transform, demo.py:67-81).

    python -m lcrnet_b200.demo --pos 003854.npy --anc 000958.npy --weights best-model-mixed.tar
    python -m lcrnet_b200.demo --synthetic            # seeded synthetic pair, random weights
"""
import argparse

import numpy as np
import torch

from . import checkpoint, lcrnet, synth
from . import data as gdata

NUM_STAGES, VOXEL, RADIUS = 4, 0.3, 4.25 * 0.3


def load_scan(path):
    a = np.load(path) if path.endswith('.npy') else np.fromfile(path, dtype=np.float32).reshape(-1, 4)
    return np.ascontiguousarray(a[:, :3], dtype=np.float32)


def run_pair(pos, anc, state_dict=None, pre_voxel=None, device='cuda'):
    """pos = reference scan, anc = source scan (estimated_transform maps anc -> pos)."""
    limits = gdata.calibrate_neighbors_scans([pos, anc], NUM_STAGES, VOXEL, RADIUS, pre_voxel=pre_voxel,
                                             device=device, scans_per_sample=2)
    net = lcrnet.create_model(lcrnet.default_cfg(limits)).eval()
    sd = state_dict if state_dict is not None else checkpoint.random_state_dict('lcrnet', 7351)
    sd = {k[7:] if k.startswith('module.') else k: v for k, v in sd.items()}      # base_tester.py:115-119
    net.load_state_dict(sd, strict=False)
    net = net.to(device)
    d = gdata.scans_collate_fn_stack_mode([pos, anc], NUM_STAGES, VOXEL, RADIUS, limits, pre_voxel=pre_voxel,
                                          stack_size=2, int32=True, upsampling=True, device=device)
    return net(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pos')
    ap.add_argument('--anc')
    ap.add_argument('--weights')
    ap.add_argument('--synthetic', action='store_true')
    a = ap.parse_args()
    if a.synthetic or not (a.pos and a.anc):
        pos, anc, gt = synth.make_pair(0, 7351)
        pre_voxel = VOXEL
        print('ground-truth transform (anc -> pos):\n', gt)
    else:
        pos, anc, pre_voxel = load_scan(a.pos), load_scan(a.anc), None
    sd = torch.load(a.weights, map_location='cpu', weights_only=True)['model'] if a.weights else None
    out = run_pair(pos, anc, sd, pre_voxel)
    dist = torch.norm(out['pos_feature_global'] - out['anc_feature_global'], dim=1)
    print('L2 feature distance: %.6f' % float(dist))
    print('Estimated transformation:\n', out['estimated_transform'].cpu().numpy())
    print('correspondences: %d' % out['corr_scores'].shape[0])


if __name__ == '__main__':
    main()
