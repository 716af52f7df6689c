"""Data formats on either side of the hot path: what the reference's scripts
read before the path starts and write after it ends.  Host-side glue, numpy only; the device work
(0.3 m pre-voxel, descriptors, top-k, registration) is the C-ABI library's.

  raw scans        KITTI ``velodyne/*.bin`` (float32 x, y, z, intensity), the offline 0.3 m pre-pass of
                   data/Kitti/downsample_pcd.py:21-42 -> ``downsampled_xyzi/<seq>/<frame>.npy`` (float32 [N, 4])
  descriptors      ``{seq}_{idx}.npz`` / ``{idx}.npz`` with key ``anc_global`` float32 [1, 256]
                   (test_loop_detection.py:60-69, read back by eval_loop_detection_overlap_dataset.py:162-176)
  candidate rows   ``predicted_des_L2_dis.npz``: ``arr_0`` float64 [P, 1, 3] rows (query i, match j, squared L2)
                   (eval_loop_detection_overlap_dataset.py:183-214; readers reshape to [P, 3])
  top-1 text       ``result/top1_with_thres_%.2f/%02d.txt`` lines ``i j d  \\n`` for rows with d < thres
                   (infer_loop_detection_find_top1.py:14-43)
  pose text        ``<seq>_pose`` lines ``pos anc r11 r12 r13 t1 r21 ... t3 \\n`` with %.6f
                   (infer_registration.py:69-80)
  evaluation       precision / recall sweep, AP, max F1 and recall@N of
                   eval_loop_detection_overlap_dataset.py:14-121
"""
import glob
import os
import os.path as osp

import numpy as np


# ----------------------------------------------------------------------------- raw scans
def read_kitti_bin(path):
    """velodyne .bin -> float32 [N, 4] (x, y, z, intensity) (downsample_pcd.py:29)."""
    raw = np.fromfile(path, dtype=np.float32)
    if raw.size % 4 != 0:
        raise ValueError('%s: not a float32 [N, 4] velodyne file (%d values)' % (path, raw.size))
    return raw.reshape(-1, 4)


def read_scan(path):
    """xyz float32 [N, 3] from a raw ``.bin`` or a pre-voxelised ``.npy`` (dataset_demo.py reads the
    first three columns of ``downsampled_xyzi``)."""
    if path.endswith('.bin'):
        pts = read_kitti_bin(path)
    elif path.endswith('.npy'):
        pts = np.load(path)
    else:
        raise ValueError('unsupported scan file: %s' % path)
    return np.ascontiguousarray(pts[:, :3], dtype=np.float32)


def prevoxel(points, voxel=0.3, device=None):
    """The offline 0.3 m pre-pass as the first on-GPU stage: raw xyz -> L0 on the device.

    Uses the reference's own barycentre voxel subsampling (the a1 kernel, bit-exact with
    ``grid_subsampling``) in place of open3d's ``voxel_down_sample`` (open3d is not part of the reference
    tree; its voxel origin is ``min_bound - voxel/2`` instead of ``min_bound``, so the two pre-passes produce
    different -- equally valid -- L0 clouds; every downstream parity statement starts from L0)."""
    import torch
    from . import ext
    pts = torch.as_tensor(points, dtype=torch.float32)
    pts = pts[:, :3].contiguous()
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    pts = pts.to(device)
    lens = torch.tensor([pts.shape[0]], dtype=torch.int64, device=device)
    out, _ = ext.grid_subsampling(pts, lens, float(voxel))
    return out


def save_downsampled(path, xyz, intensity=None):
    """``downsampled_xyzi`` record: float32 [N, 4] .npy (downsample_pcd.py:36-46); intensity 0 if not given."""
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    inten = np.zeros((xyz.shape[0], 1), np.float32) if intensity is None else \
        np.asarray(intensity, dtype=np.float32).reshape(-1, 1)
    np.save(path, np.concatenate([xyz, inten], axis=1))


# ----------------------------------------------------------------------------- descriptors
def descriptor_files(features_root, seq=None):
    """Descriptor records of a sequence in frame order.  Keys follow the two reference readers: plain
    ``{idx}.npz`` sorted by int (eval_...:162-165) or ``{seq}_{idx}.npz`` sorted by (seq, idx)
    (infer_..._find_top1.py:58-61)."""
    pattern = '*.npz' if seq is None else '%d*.npz' % seq
    names = [f for f in glob.glob(osp.join(features_root, pattern))
             if osp.basename(f) != 'predicted_des_L2_dis.npz']
    return sorted(names, key=lambda x: [int(i) for i in osp.splitext(osp.basename(x))[0].split('_')])


def load_descriptors(features_root, seq=None, normalize=False):
    """[N, 256] float32 database from the per-scan records.  ``normalize`` re-normalises the rows as the
    inference variant does (infer_loop_detection_find_top1.py:75)."""
    rows = [np.load(f)['anc_global'].astype(np.float32).reshape(-1, 256) for f in descriptor_files(features_root, seq)]
    if not rows:
        return np.zeros((0, 256), np.float32)
    db = np.concatenate(rows)
    if normalize:
        db = db / np.linalg.norm(db, axis=1, keepdims=True)
    return db


def save_candidate_rows(path, rows):
    """``predicted_des_L2_dis.npz`` exactly as the reference stores it: ``np.array(row_list)`` of [1, 3] rows
    -> ``arr_0`` float64 [P, 1, 3] (eval_...:209-211)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 1, 3)
    np.savez_compressed(path, rows)


def load_candidate_rows(path):
    """float32 [P, 3] as every reader reshapes it (eval_...:213-218)."""
    rows = np.asarray(np.load(path)['arr_0'], dtype='float32')
    return rows.reshape((len(rows), 3))


# ----------------------------------------------------------------------------- top-1 / pose text
def top1_rows(rows, n_frames, thres=0.11):
    """Rows with distance below ``thres``, grouped by ascending query index (find_top1, :14-27)."""
    rows = np.asarray(rows, dtype=np.float32).reshape(-1, 3)      # the readers hold the rows as float32
    keep = (rows[:, 2] < thres) & (rows[:, 0] >= 0) & (rows[:, 0] < n_frames - 1)
    sel = rows[keep]
    order = np.argsort(sel[:, 0].astype(np.int64), kind='stable')      # queries ascending, hits in stored order
    return sel[order]


def top1_lines(rows, n_frames, thres=0.11):
    """The text lines of ``top1_with_thres_*/NN.txt`` (:36-40), including the two trailing blanks."""
    # the distance is an np.float32 formatted by an f-string, i.e. through float(): the digits of the DOUBLE
    return ['%d %d %s  \n' % (int(r[0]), int(r[1]), format(r[2], '')) for r in top1_rows(rows, n_frames, thres)]


def write_top1(dataset_root, seq, rows, n_frames, thres=0.11):
    path = '%s/result/top1_with_thres_%.2f' % (dataset_root, thres)
    os.makedirs(path, exist_ok=True)
    name = '%s/%02d.txt' % (path, seq)
    with open(name, 'a') as f:
        f.writelines(top1_lines(rows, n_frames, thres))
    return name


def pose_line(pos_idx, anc_idx, transform):
    """``pos anc r11 .. t3`` (first 12 entries of the row-major 4x4, %.6f) (infer_registration.py:75-77)."""
    m = np.asarray(transform, dtype=np.float64).reshape(-1)[:12]
    return '%s %s %s \n' % (pos_idx, anc_idx, ' '.join('%.6f' % v for v in m))


def append_pose(output_dir, seq_id, pos_idx, anc_idx, transform):
    os.makedirs(output_dir, exist_ok=True)
    with open(osp.join(output_dir, '%s_pose' % seq_id), 'a') as f:
        f.write(pose_line(pos_idx, anc_idx, transform))


# ----------------------------------------------------------------------------- evaluation
def _first_hit(rows, n):
    """Per query i < n: (has_row, match j, distance) of its FIRST stored row (rows are stored nearest first)."""
    rows = np.asarray(rows, dtype=np.float32).reshape(-1, 3)
    q = rows[:, 0].astype(np.int64)
    has = np.zeros(n, bool)
    j = np.full(n, -1, np.int64)
    d = np.full(n, np.inf, np.float32)
    ok = (q >= 0) & (q < n)
    first = np.unique(q[ok], return_index=True)
    idx = np.flatnonzero(ok)[first[1]]
    has[first[0]] = True
    j[first[0]] = rows[idx, 1].astype(np.int64)
    d[first[0]] = rows[idx, 2]
    return has, j, d


def compute_pr(rows, ground_truth, thre_range=(0.0, 1.0), interval=0.01, start=150):
    """Precision / recall sweep of compute_PR_overlap (eval_...:66-121): query idx in [start, len(gt) - 1),
    decision on its nearest candidate; stops after the first threshold with recall 1.
    ``ground_truth``: sequence of index arrays (the ``loop_gt_*`` npz objects)."""
    n = len(ground_truth)
    has, j, d = _first_hit(rows, n)
    ids = np.arange(start, n - 1)
    if len(ids) and not has[ids].all():
        raise IndexError('query %d has no candidate row (the reference indexes [0] of an empty selection)'
                         % int(ids[~has[ids]][0]))
    gt_any = np.array([np.asarray(ground_truth[i]).any() for i in ids], bool)
    correct = np.array([j[i] in np.asarray(ground_truth[i]) for i in ids], bool)
    precisions, recalls = [], []
    for thres in np.arange(thre_range[0], thre_range[1], interval):
        reject = d[ids] > thres
        fns = int((reject & gt_any).sum())
        tps = int((~reject & correct).sum())
        fps = int((~reject & ~correct).sum())
        precision = 1 if fps == 0 else float(tps) / (float(tps) + float(fps))
        recall = 1 if fns == 0 else float(tps) / (float(tps) + float(fns))
        precisions.append(precision)
        recalls.append(recall)
        if recall == 1:
            break
    return precisions, recalls


def compute_ap(precision, recall):
    """eval_...:14-18."""
    ap = 0.
    for i in range(1, len(precision)):
        ap += (recall[i] - recall[i - 1]) * precision[i]
    return ap


def compute_f1(precision, recall):
    """Max F1 and its index (eval_...:20-27)."""
    precision, recall = np.asarray(precision, dtype=np.float64), np.asarray(recall, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        f1 = 2 * precision * recall / (precision + recall)
    return float(np.nanmax(f1)), int(np.nanargmax(f1))


def recall_at_n(rows, ground_truth, topn):
    """compute_topN (eval_...:29-62): fraction of queries with ground truth whose first ``topn`` candidates
    contain a true loop."""
    rows = np.asarray(rows, dtype=np.float32).reshape(-1, 3)
    q = rows[:, 0].astype(np.int64)
    order = np.argsort(q, kind='stable')
    q_sorted, rows_sorted = q[order], rows[order]
    starts = np.searchsorted(q_sorted, np.arange(len(ground_truth)), side='left')
    ends = np.searchsorted(q_sorted, np.arange(len(ground_truth)), side='right')
    have, tps = 0, 0
    for idx in range(0, len(ground_truth) - 1):
        gt = np.asarray(ground_truth[idx])
        if not gt.any():
            continue
        have += 1
        cand = rows_sorted[starts[idx]:ends[idx], 1]
        if np.isin(cand[:topn], gt).any():
            tps += 1
        elif len(cand) < topn:        # the reference indexes row t of the selection and raises here
            raise IndexError('query %d has fewer than %d candidate rows' % (idx, topn))
    return tps / have if have else 0.0
