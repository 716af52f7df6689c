"""Generates tests/golden/model_golden.npz by running the UNMODIFIED reference Python model
(``experiments.lcrnet.model_family.LCRNet_GlobalDescrition`` through tests/golden/ref_import.py,
with the reference C++ operators compiled into oracle/_ref) on seeded synthetic scans with the
seeded weights of ``lcrnet_b200.checkpoint.random_state_dict``.  Build container only."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

LIMITS = [40, 40, 40, 40]
CASES = [('s0', 0, 7351, 4), ('s1', 1, 7352, 2)]  # (name, scene_seed, noise seed, raw-point stride)
ROWS = 8                                           # rows of every block output kept in the fixture


def reference_forward(raw, sd, limits=LIMITS):
    """Reference pyramid (data.py:10-74) + LCRNet_GlobalDescrition.forward on ONE scan (B = 1)."""
    import ref_import
    ref_import.install()
    from experiments.lcrnet.data import precompute_data_stack_mode
    from experiments.lcrnet.modules.ops import grid_subsample
    from experiments.lcrnet.model_family.LCRNet_GlobalDescrition import create_model
    cfg = ref_import.model_cfg(limits, tempfile.mkdtemp())
    pts = torch.from_numpy(raw)
    lens = torch.tensor([len(raw)], dtype=torch.int64)
    p0, l0 = grid_subsample(pts, lens, voxel_size=0.3)          # stands in for the offline 0.3 m pre-pass
    data = precompute_data_stack_mode(p0, l0, cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                      cfg.backbone.init_radius, limits)
    data['features'] = torch.ones(p0.shape[0], 1)
    data = {k: ([x.contiguous() for x in v] if isinstance(v, list) else v) for k, v in data.items()}
    torch.manual_seed(7351)
    np.random.seed(7351)
    model = create_model(cfg).eval()
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    blocks = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, n=n: blocks.__setitem__(n, o.detach().clone()))
             for n, m in model.encoder.named_children()]
    with torch.no_grad():
        out = model(data)
    for h in hooks:
        h.remove()
    return data, blocks, out['anc_global']


def make_model_golden():
    from lcrnet_b200 import checkpoint, synth
    sd = checkpoint.random_state_dict('global_descriptor', seed=7351)
    out = {'limits': np.array(LIMITS), 'rows': ROWS, 'weight_seed': 7351}
    for name, scene, seed, stride in CASES:
        raw = np.ascontiguousarray(synth.make_scan(scene, seed)[::stride])
        data, blocks, desc = reference_forward(raw, sd)
        out[name + '_case'] = np.array([scene, seed, stride])
        out[name + '_lengths'] = np.array([int(l[0]) for l in data['lengths']])
        out[name + '_widths'] = np.array([t.shape[1] for t in data['neighbors']])
        out[name + '_descriptor'] = desc.numpy()
        for bn, t in blocks.items():
            t = t.numpy()
            out['%s_%s_head' % (name, bn)] = t[:ROWS].copy()
            out['%s_%s_stats' % (name, bn)] = np.array([t.mean(dtype=np.float64), np.abs(t).mean(dtype=np.float64),
                                                        (t.astype(np.float64) ** 2).sum()])
        print(name, out[name + '_lengths'], out[name + '_widths'], float(np.linalg.norm(desc.numpy())))
    path = os.path.join(HERE, 'model_golden.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    make_model_golden()
