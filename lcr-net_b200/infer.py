"""GT-free inference entry points on the B200 backend: the flows of the reference's
``experiments/inference/*.py`` with the dataset loaders replaced by plain lists of scan files.

  generate_descriptors   infer_loop_detection_descriptor_generation.py:49-72 -> ``{seq}_{idx}.npz`` records
  find_loops             infer_loop_detection_find_top1.py:46-120 -> ``predicted_des_L2_dis.npz`` + top-1 text
  register_pairs         infer_registration.py:51-80 -> ``<seq>_pose`` text lines

    python -m lcrnet_b200.infer descriptors --scans DIR --seq 8 --out FEATURES [--weights CKPT]
    python -m lcrnet_b200.infer loops --features FEATURES --seq 8 --root DATASET_ROOT --thres 0.11
    python -m lcrnet_b200.infer register --scans DIR --pairs top1.txt --seq 8 --out POSES [--weights CKPT]

Scans are ``.bin`` (raw velodyne; the 0.3 m pre-voxel runs on the GPU as the first stage) or ``.npy``
(already pre-voxelised ``downsampled_xyzi``).  Neighbour limits are calibrated on the device like the
reference's loaders do (data.py:408-433).  All compute runs in liblcr_b200.so; there is no CPU path.
"""
import argparse
import glob
import os
import os.path as osp

import numpy as np
import torch

from . import checkpoint, formats, lcrnet, model, pipeline, retrieval
from . import data as gdata

NUM_STAGES, VOXEL, RADIUS = 4, 0.3, 4.25 * 0.3


def _strip_module(sd):
    return {k[7:] if k.startswith('module.') else k: v for k, v in sd.items()}       # base_tester.py:115-119


def load_weights(path):
    return _strip_module(torch.load(path, map_location='cpu', weights_only=True)['model'])


# parameters the inference forward never reads (LCRNet.py:60: only the training loss uses it)
_UNUSED_AT_INFERENCE = ('proj_node_overlap_score.',)


def _load_checked(net, state_dict, kind):
    """``load_state_dict`` as base_tester.py:111-122 does it (strict=False, 'module.' stripped), but a checkpoint
    that leaves parameters of the inference path uninitialised is an ERROR here: this package zero-initialises
    its parameter holders, so a mismatched checkpoint would silently produce descriptors / poses from zero
    weights.  ``state_dict=None`` selects seeded random weights (smoke runs) with a loud warning."""
    import warnings
    if state_dict is None:
        warnings.warn('lcrnet_b200.infer: no checkpoint given -- running with SEEDED RANDOM weights '
                      '(checkpoint.random_state_dict(%r, 7351)); outputs are not meaningful' % kind, stacklevel=3)
        state_dict = checkpoint.random_state_dict(kind, 7351)
    sd = _strip_module(state_dict)
    own = net.state_dict()
    res = net.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    missing = [k for k in res.missing_keys if not k.startswith(_UNUSED_AT_INFERENCE)
               and not k.endswith('num_batches_tracked')]
    if missing:
        raise RuntimeError('checkpoint does not cover %d tensor(s) of the %s model (first: %s): refusing to run '
                           'with uninitialised weights' % (len(missing), kind, ', '.join(missing[:5])))
    return net


def _scan_list(scans):
    """list of arrays / file names, or a directory of .bin / .npy files in frame order."""
    if isinstance(scans, str):
        names = sorted(glob.glob(osp.join(scans, '*.bin')) + glob.glob(osp.join(scans, '*.npy')),
                       key=lambda x: int(osp.splitext(osp.basename(x))[0]))
        return names
    return list(scans)


def _load(scan):
    return formats.read_scan(scan) if isinstance(scan, str) else np.ascontiguousarray(scan, dtype=np.float32)


def _dev(device):
    d = torch.device(device)
    return torch.device('cuda', torch.cuda.current_device()) if d.type == 'cuda' and d.index is None else d


def _needs_prevoxel(scan):
    return isinstance(scan, str) and scan.endswith('.bin')


def generate_descriptors(scans, out_dir, seq, state_dict=None, batch_scans=32, pre_voxel=None, device='cuda',
                         indices=None, neighbor_limits=None):
    """One ``{seq}_{idx}.npz`` (key ``anc_global`` [1, 256]) per scan; returns the [N, 256] database (device).
    ``pre_voxel``: None = decide per input (raw ``.bin`` files are pre-voxelised at 0.3 m).
    ``neighbor_limits``: None = calibrate on the first scans (data.py:408-433)."""
    scans = _scan_list(scans)
    device = _dev(device)
    os.makedirs(out_dir, exist_ok=True)
    if pre_voxel is None:
        pre_voxel = VOXEL if scans and _needs_prevoxel(scans[0]) else 0
    net = model.create_model(model.default_cfg()).eval()
    net = _load_checked(net, state_dict, 'global_descriptor').to(device)
    limits = neighbor_limits
    if limits is None:
        limits = gdata.calibrate_neighbors_scans([_load(s) for s in scans[:4]], NUM_STAGES, VOXEL, RADIUS,
                                                      pre_voxel=pre_voxel or None, device=device)
    pipe = pipeline.DescriptorPipeline(net, limits, NUM_STAGES, VOXEL, RADIUS, pre_voxel=pre_voxel or None,
                                       n_streams=2, device=device)
    out = []
    for b0 in range(0, len(scans), batch_scans):
        batch = [_load(s) for s in scans[b0:b0 + batch_scans]]
        pts = torch.from_numpy(np.concatenate(batch, 0)).pin_memory()
        desc = pipe(pts, [len(b) for b in batch])
        out.append(desc)
        host = desc.cpu().numpy()
        for i, d in enumerate(host):
            idx = b0 + i if indices is None else indices[b0 + i]
            retrieval.save_descriptor_npz(osp.join(out_dir, '%s_%s.npz' % (seq, idx)), d)
    pipe.close()
    return torch.cat(out, 0) if out else torch.zeros((0, 256), device=device)


def find_loops(features_root, seq, dataset_root=None, thres=0.11, k=50, gap=100, device='cuda'):
    """Candidate rows (i, j, d2) of every query against its causal database (rows re-normalised first, as
    infer_loop_detection_find_top1.py:75 does), stored as ``predicted_des_L2_dis.npz``; with ``dataset_root``
    also the top-1 text file.  Returns (rows float64 [P, 3], top-1 file name or None)."""
    device = _dev(device)
    db = formats.load_descriptors(features_root, seq, normalize=True)
    rows = retrieval.loop_candidates(torch.from_numpy(db).to(device), k=k, gap=gap)
    formats.save_candidate_rows(osp.join(features_root, 'predicted_des_L2_dis'), rows)
    name = formats.write_top1(dataset_root, int(seq), rows, db.shape[0], thres) if dataset_root else None
    return rows, name


def register_pairs(pairs, out_dir, seq_id, state_dict=None, pre_voxel=None, device='cuda'):
    """pairs: iterable of (pos_idx, anc_idx, pos_scan, anc_scan) (arrays or file names); appends one pose line
    per pair to ``out_dir/<seq_id>_pose`` (estimated_transform maps anc -> pos) and returns the transforms."""
    pairs = list(pairs)
    device = _dev(device)
    if not pairs:
        return []
    first = [_load(pairs[0][2]), _load(pairs[0][3])]
    if pre_voxel is None:
        pre_voxel = VOXEL if _needs_prevoxel(pairs[0][2]) else 0
    limits = gdata.calibrate_neighbors_scans(first, NUM_STAGES, VOXEL, RADIUS, pre_voxel=pre_voxel or None,
                                                  device=device)
    net = lcrnet.create_model(lcrnet.default_cfg(limits)).eval()
    net = _load_checked(net, state_dict, 'lcrnet').to(device)
    out = []
    for pos_idx, anc_idx, pos, anc in pairs:
        d = gdata.scans_collate_fn_stack_mode([_load(pos), _load(anc)], NUM_STAGES, VOXEL, RADIUS, limits,
                                              pre_voxel=pre_voxel or None, stack_size=2, int32=True, upsampling=True,
                                              device=device)
        T = net(d)['estimated_transform'].cpu().numpy()
        formats.append_pose(out_dir, seq_id, pos_idx, anc_idx, T)
        out.append(T)
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    sub = ap.add_subparsers(dest='cmd', required=True)
    a = sub.add_parser('descriptors')
    a.add_argument('--scans', required=True)
    a.add_argument('--seq', required=True)
    a.add_argument('--out', required=True)
    a.add_argument('--weights')
    b = sub.add_parser('loops')
    b.add_argument('--features', required=True)
    b.add_argument('--seq', type=int, required=True)
    b.add_argument('--root')
    b.add_argument('--thres', type=float, default=0.11)
    c = sub.add_parser('register')
    c.add_argument('--scans', required=True)
    c.add_argument('--pairs', required=True, help='text file with "pos_idx anc_idx ..." per line (the top-1 file)')
    c.add_argument('--seq', required=True)
    c.add_argument('--out', required=True)
    c.add_argument('--weights')
    args = ap.parse_args()
    if args.cmd == 'descriptors':
        sd = load_weights(args.weights) if args.weights else None
        db = generate_descriptors(args.scans, args.out, args.seq, sd)
        print('%d descriptors written to %s' % (db.shape[0], args.out))
    elif args.cmd == 'loops':
        rows, name = find_loops(args.features, args.seq, args.root, args.thres)
        print('%d candidate rows; top-1 file: %s' % (len(rows), name))
    else:
        names = _scan_list(args.scans)
        by_idx = {int(osp.splitext(osp.basename(n))[0]): n for n in names}
        pairs = []
        for line in open(args.pairs):
            f = line.split()
            if len(f) >= 2:
                pairs.append((int(f[0]), int(f[1]), by_idx[int(f[0])], by_idx[int(f[1])]))
        sd = load_weights(args.weights) if args.weights else None
        Ts = register_pairs(pairs, args.out, args.seq, sd)
        print('%d poses appended to %s' % (len(Ts), osp.join(args.out, '%s_pose' % args.seq)))


if __name__ == '__main__':
    main()
