"""Formats either side of the hot path (lcrnet_b200.formats) against the literal restatement of the reference's
loops (oracle/formats_oracle.py).  CPU only, except the pre-voxel stage."""
import os

import numpy as np
import pytest

from lcrnet_b200 import formats as F
from oracle import formats_oracle as FO


def _case(seed, n=260, k=6):
    """Synthetic candidate rows (k nearest-first rows per query i in [101, n-2]) and ground truth lists."""
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(101, n - 1):
        d = np.sort(rng.uniform(0.0, 0.6, k))
        j = rng.integers(0, i - 100, k)
        rows += [(i, int(jj), float(dd)) for jj, dd in zip(j, d)]
    rows = np.array(rows, dtype=np.float64)
    gt = np.empty(n, dtype=object)
    for i in range(n):
        if i > 120 and rng.random() < 0.5:
            own = rows[rows[:, 0] == i]
            pick = own[rng.integers(0, len(own)), 1] if len(own) and rng.random() < 0.7 else rng.integers(1, max(2, i - 100))
            gt[i] = np.array([int(pick), int(rng.integers(1, max(2, i - 100)))])
        else:
            gt[i] = np.array([], dtype=np.int64)
    return rows, gt


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_pr_ap_f1_recall_match_reference_loops(seed):
    rows, gt = _case(seed)
    pair = np.asarray(rows, dtype='float32').reshape(-1, 3)
    p_ref, r_ref = FO.compute_pr_overlap(pair, gt)
    p, r = F.compute_pr(rows, gt)
    assert p == p_ref and r == r_ref
    assert F.compute_ap(p, r) == FO.compute_ap(p_ref, r_ref)
    f1, idx = F.compute_f1(p, r)
    pr, rr = np.asarray(p_ref, float), np.asarray(r_ref, float)
    with np.errstate(invalid='ignore', divide='ignore'):
        ref_f1 = 2 * pr * rr / (pr + rr)
    assert f1 == np.nanmax(ref_f1) and idx == int(np.nanargmax(ref_f1))
    for topn in (1, 3, 6):
        assert F.recall_at_n(rows, gt, topn) == FO.compute_topn(rows, gt, topn)


def test_top1_and_pose_text(tmp_path):
    rows, _ = _case(3)
    pair = np.asarray(rows, dtype='float32').reshape(-1, 3)
    for thres in (0.05, 0.11, 0.3):
        assert F.top1_lines(rows, 260, thres) == FO.find_top1_lines(pair, 260, thres)
    name = F.write_top1(str(tmp_path), 2, rows, 260, 0.11)
    assert name.endswith('result/top1_with_thres_0.11/02.txt')
    assert open(name).read() == ''.join(FO.find_top1_lines(pair, 260, 0.11))
    T = np.random.default_rng(0).standard_normal((4, 4)).astype(np.float32)
    assert F.pose_line(12, 345, T) == FO.pose_line(12, 345, T)
    F.append_pose(str(tmp_path / 'out'), '08', 12, 345, T)
    F.append_pose(str(tmp_path / 'out'), '08', 13, 346, T)
    lines = open(tmp_path / 'out' / '08_pose').read().splitlines(keepends=True)
    assert lines == [FO.pose_line(12, 345, T), FO.pose_line(13, 346, T)]


def test_descriptor_and_candidate_files(tmp_path):
    from lcrnet_b200 import retrieval
    rng = np.random.default_rng(1)
    db = rng.standard_normal((12, 256)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    order = rng.permutation(12)
    for i in order:                                           # written out of order, named {seq}_{idx}.npz
        retrieval.save_descriptor_npz(str(tmp_path / ('8_%d.npz' % i)), db[i])
    retrieval.save_descriptor_npz(str(tmp_path / '9_0.npz'), db[0])          # another sequence
    got = F.load_descriptors(str(tmp_path), seq=8)
    assert np.array_equal(got, db)
    assert np.load(tmp_path / '8_3.npz')['anc_global'].shape == (1, 256)
    rows = np.array([(101, 0, 0.25), (101, 1, 0.5), (102, 0, 0.125)])
    F.save_candidate_rows(str(tmp_path / 'predicted_des_L2_dis'), rows)
    raw = np.load(tmp_path / 'predicted_des_L2_dis.npz')['arr_0']
    assert raw.shape == (3, 1, 3) and raw.dtype == np.float64              # np.array(row_list) of [1,3] rows
    assert np.array_equal(F.load_candidate_rows(str(tmp_path / 'predicted_des_L2_dis.npz')), rows.astype(np.float32))
    assert 'predicted_des_L2_dis.npz' not in [os.path.basename(f) for f in F.descriptor_files(str(tmp_path))]


def test_raw_scan_files(tmp_path):
    rng = np.random.default_rng(2)
    xyzi = rng.standard_normal((1000, 4)).astype(np.float32)
    xyzi.tofile(tmp_path / '000000.bin')
    assert np.array_equal(F.read_kitti_bin(str(tmp_path / '000000.bin')), xyzi)
    assert np.array_equal(F.read_scan(str(tmp_path / '000000.bin')), xyzi[:, :3])
    F.save_downsampled(str(tmp_path / '000000.npy'), xyzi[:, :3], xyzi[:, 3])
    assert np.array_equal(np.load(tmp_path / '000000.npy'), xyzi)
    assert np.array_equal(F.read_scan(str(tmp_path / '000000.npy')), xyzi[:, :3])
    (tmp_path / 'bad.bin').write_bytes(b'\0' * 10)
    with pytest.raises(ValueError):
        F.read_kitti_bin(str(tmp_path / 'bad.bin'))


def test_loop_rows_oracle_matches_numpy_topk():
    """The brute-force restatement of the faiss loop agrees with oracle.model_oracle.l2_topk (the a15 oracle)."""
    from oracle import model_oracle as mo
    rng = np.random.default_rng(4)
    emb = rng.standard_normal((140, 256)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    rows = FO.loop_rows_bruteforce(emb, k=5, gap=100)
    q = np.arange(101, 139)
    d2, idx = mo.l2_topk(emb[q], emb, 5, valid_counts=np.maximum(q - 100, 0).astype(np.int32))
    flat = [(int(i), int(j)) for i, jr in zip(q, idx) for j in jr if j >= 0]
    assert [(int(r[0]), int(r[1])) for r in rows] == flat


@pytest.mark.gpu
def test_prevoxel_stage_matches_reference_operator():
    """Raw 64k-point scan -> L0 on the GPU == the reference's grid_subsampling(0.3) (C oracle, bit-exact)."""
    import torch
    from lcrnet_b200 import synth
    from oracle import native as on
    raw = synth.make_scan(3)
    got = F.prevoxel(raw, 0.3).cpu().numpy()
    want, _ = on.grid_subsample(raw, np.array([len(raw)], dtype=np.int64), 0.3)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_inference_entry_points_end_to_end(tmp_path):
    """raw .bin scans -> descriptor records -> candidate rows / top-1 text -> pose text, all on the GPU backend
    (lcrnet_b200.infer), files in the reference's formats."""
    import torch
    from lcrnet_b200 import infer, synth
    scan_dir = tmp_path / 'velodyne'
    scan_dir.mkdir()
    scans = []
    for i in range(6):
        xyz = synth.make_scan(i // 2, 7351 + i)[::4]
        scans.append(xyz)
        np.concatenate([xyz, np.zeros((len(xyz), 1), np.float32)], 1).tofile(scan_dir / ('%06d.bin' % i))
    feat = tmp_path / 'features'
    lim = [40, 40, 40, 40]
    db = infer.generate_descriptors(str(scan_dir), str(feat), 8, batch_scans=4, neighbor_limits=lim)
    assert db.shape == (6, 256)
    assert torch.allclose(db.norm(dim=1), torch.ones(6, device=db.device), atol=1e-5)
    files = F.descriptor_files(str(feat), 8)
    assert [os.path.basename(f) for f in files] == ['8_%d.npz' % i for i in range(6)]
    assert np.allclose(F.load_descriptors(str(feat), 8), db.cpu().numpy(), atol=1e-7)
    # batches of 4 + 2 scans == one scan at a time
    single = infer.generate_descriptors([str(scan_dir / '000004.bin')], str(tmp_path / 'f2'), 8, indices=[4],
                                         neighbor_limits=lim)
    assert float((single[0] - db[4]).norm()) < 1e-5
    rows, name = infer.find_loops(str(feat), 8, str(tmp_path / 'root'), thres=10.0, k=3, gap=2)
    want = FO.loop_rows_bruteforce(F.load_descriptors(str(feat), 8, normalize=True), k=3, gap=2)
    assert rows.shape == want.shape and np.array_equal(rows[:, :2], want[:, :2])
    assert np.allclose(rows[:, 2], want[:, 2], atol=1e-5)
    assert np.array_equal(F.load_candidate_rows(str(feat / 'predicted_des_L2_dis.npz')), rows.astype(np.float32))
    assert open(name).read() == ''.join(F.top1_lines(rows, 6, 10.0))
    Ts = infer.register_pairs([(0, 1, str(scan_dir / '000000.bin'), str(scan_dir / '000001.bin')),
                               (2, 3, scans[2], scans[3])], str(tmp_path / 'poses'), '08', pre_voxel=0.3)
    lines = open(tmp_path / 'poses' / '08_pose').read().splitlines(keepends=True)
    assert lines == [FO.pose_line(0, 1, Ts[0]), FO.pose_line(2, 3, Ts[1])]
    for T in Ts:
        R = T[:3, :3].astype(np.float64)
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-4 and abs(np.linalg.det(R) - 1) < 1e-4
