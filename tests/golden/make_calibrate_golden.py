"""Runs the UNMODIFIED reference collates and neighbour-limit calibration
(experiments/lcrnet/data.py:77-127 registration_collate_fn_stack_mode, :350-406
test_loop_detection_collate_fn_stack_mode_online, :408-433 calibrate_neighbors_stack_mode) with the
reference's own C++ operators on seeded synthetic scans and writes tests/golden/calibrate_golden.npz.
Build container only (needs /root/reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

NUM_STAGES, VOXEL, RADIUS = 4, 0.3, 4.25 * 0.3
STRIDE = 3            # raw-point stride: keeps the hist_n-wide tables of the reference small
N_PAIRS = 3


def make_samples():
    """Registration samples (dicts with the keys dataset_demo.py:66-78 returns) and loop-detection samples."""
    from lcrnet_b200 import synth
    from oracle import native as on
    reg, ld = [], []
    for i in range(N_PAIRS):
        ref, src, T = synth.make_pair(40 + i, 7351 + i)
        pv = []
        for raw in (ref[::STRIDE], src[::STRIDE]):
            raw = np.ascontiguousarray(raw)
            p, _ = on.grid_subsample(raw, np.array([len(raw)], dtype=np.int64), VOXEL)   # the offline 0.3 m pre-pass
            pv.append(p)
        reg.append({'seq_id': 0, 'ref_frame': 2 * i, 'src_frame': 2 * i + 1, 'ref_points': pv[0], 'src_points': pv[1],
                    'ref_feats': np.ones((len(pv[0]), 1), np.float32), 'src_feats': np.ones((len(pv[1]), 1), np.float32),
                    'transform': T.astype(np.float32)})
        ld.append({'seq_id': 0, 'anc_idx': i, 'anc_points': pv[0], 'anc_feats': np.ones((len(pv[0]), 1), np.float32)})
    return reg, ld


def main():
    import ref_import
    ref_import.install()
    from experiments.lcrnet import data as rdata
    reg, ld = make_samples()
    lim_reg = rdata.calibrate_neighbors_stack_mode(reg, rdata.registration_collate_fn_stack_mode, NUM_STAGES, VOXEL,
                                                   RADIUS)
    lim_ld = rdata.calibrate_neighbors_stack_mode(ld, rdata.test_loop_detection_collate_fn_stack_mode_online,
                                                  NUM_STAGES, VOXEL, RADIUS)
    lim_reg_t500 = rdata.calibrate_neighbors_stack_mode(reg, rdata.registration_collate_fn_stack_mode, NUM_STAGES,
                                                        VOXEL, RADIUS, keep_ratio=0.6, sample_threshold=500)
    limits = [int(x) for x in lim_reg]
    d_reg = rdata.registration_collate_fn_stack_mode(reg[:1], NUM_STAGES, VOXEL, RADIUS, limits)
    d_reg2 = rdata.registration_collate_fn_stack_mode(reg[:2], NUM_STAGES, VOXEL, RADIUS, limits)
    d_ld = rdata.test_loop_detection_collate_fn_stack_mode_online(ld[:1], NUM_STAGES, VOXEL, RADIUS, limits)
    d_raw = rdata.registration_collate_fn_stack_mode(reg[:1], NUM_STAGES, VOXEL, RADIUS, limits, precompute_data=False)
    g = {'stride': STRIDE, 'n_pairs': N_PAIRS, 'limits_registration': np.array(lim_reg),
         'limits_loop_detection': np.array(lim_ld), 'limits_registration_q60_t500': np.array(lim_reg_t500),
         'keys_registration': np.array(sorted(d_reg.keys())), 'keys_registration_b2': np.array(sorted(d_reg2.keys())),
         'keys_loop_detection': np.array(sorted(d_ld.keys())), 'keys_raw': np.array(sorted(d_raw.keys())),
         'lengths_registration': np.stack([l.numpy() for l in d_reg['lengths']]),
         'lengths_registration_b2': np.stack([l.numpy() for l in d_reg2['lengths']]),
         'lengths_loop_detection': np.stack([l.numpy() for l in d_ld['lengths']]),
         'features_shape_registration_b2': np.array(d_reg2['features'].shape),
         'features_shape_loop_detection': np.array(d_ld['features'].shape),
         'points3_registration': d_reg['points'][3].numpy(),
         'neighbors3_registration': d_reg['neighbors'][3].numpy().astype(np.int32),
         'subsampling2_loop_detection': d_ld['subsampling'][2].numpy().astype(np.int32),
         'upsampling2_loop_detection': d_ld['upsampling'][2].numpy().astype(np.int32),
         'b2_types': np.array([type(d_reg2[k]).__name__ for k in sorted(d_reg2.keys())]),
         'b1_types': np.array([type(d_reg[k]).__name__ for k in sorted(d_reg.keys())])}
    path = os.path.join(HERE, 'calibrate_golden.npz')
    np.savez_compressed(path, **g)
    print('registration limits', lim_reg, 'loop detection limits', lim_ld, 'q60/t500', lim_reg_t500)
    print('keys', g['keys_registration'], g['b1_types'])
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
