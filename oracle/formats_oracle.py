"""TEST INFRASTRUCTURE (oracle): literal restatement of the reference's evaluation / text-output loops, used
only by tests/ to check lcrnet_b200.formats.  Plain python loops, small inputs.

Follows experiments/loop_detection/eval_loop_detection_overlap_dataset.py:14-121 (compute_AP, compute_F1,
compute_topN, compute_PR_overlap) and experiments/inference/infer_loop_detection_find_top1.py:14-43 (find_top1),
experiments/inference/infer_registration.py:69-77 (pose line).  Parity pin: these ARE the reference's loops with
the file I/O and printing removed; the reference has no golden vectors for them.
"""
import numpy as np


def compute_ap(precision, recall):                       # eval_...:14-18
    ap = 0.
    for i in range(1, len(precision)):
        ap += (recall[i] - recall[i - 1]) * precision[i]
    return ap


def compute_topn(des_dists, ground_truth, topn):         # eval_...:29-62
    des_dists = np.asarray(des_dists, dtype='float32')
    des_dists = des_dists.reshape((len(des_dists), 3))
    all_have_gt = 0
    tps = 0
    for idx in range(0, len(ground_truth) - 1):
        gt_idxes = ground_truth[int(idx)]
        if not gt_idxes.any():
            continue
        all_have_gt += 1
        for t in range(topn):
            if des_dists[des_dists[:, 0] == int(idx), :][t, 1] in gt_idxes:
                tps += 1
                break
    return tps / all_have_gt


def compute_pr_overlap(pair_dist, ground_truth, thre_range=(0, 1), interval=0.01, start=150):   # eval_...:66-121
    precisions, recalls = [], []
    for thres in np.arange(thre_range[0], thre_range[1], interval):
        tps = fps = tns = fns = 0
        for idx in range(start, len(ground_truth) - 1):
            gt_idxes = ground_truth[int(idx)]
            reject_flag = False
            if pair_dist[pair_dist[:, 0] == int(idx), 2][0] > thres:
                reject_flag = True
            if reject_flag:
                if not gt_idxes.any():
                    tns += 1
                else:
                    fns += 1
            else:
                if pair_dist[pair_dist[:, 0] == int(idx), 1][0] in gt_idxes:
                    tps += 1
                else:
                    fps += 1
        precision = 1 if fps == 0 else float(tps) / (float(tps) + float(fps))
        recall = 1 if fns == 0 else float(tps) / (float(tps) + float(fns))
        precisions.append(precision)
        recalls.append(recall)
        if recall == 1:
            break
    return precisions, recalls


def find_top1_lines(des_dists, n, thres=0.11):           # infer_loop_detection_find_top1.py:14-40
    top1_with_thre = []
    for idx in range(0, n - 1):
        dist = des_dists[des_dists[:, 0] == int(idx)]
        if dist.shape[0] == 0:
            continue
        if dist[dist[:, 2] < thres].shape[0] == 0:
            continue
        else:
            top1_with_thre.append(dist[dist[:, 2] < thres])
    lines = []
    for i in range(len(top1_with_thre)):
        for j in range(top1_with_thre[i].shape[0]):
            lines.append(f'{int(top1_with_thre[i][j][0])} {int(top1_with_thre[i][j][1])} {(top1_with_thre[i][j][2])}  \n')
    return lines


def pose_line(positive_idx, anchor_idx, estimated_transform):      # infer_registration.py:75-77
    M2 = estimated_transform.reshape(-1)[:12]
    return (f'{positive_idx} {anchor_idx} {M2[0]:.6f} {M2[1]:.6f} {M2[2]:.6f} {M2[3]:.6f} {M2[4]:.6f} {M2[5]:.6f} '
            f'{M2[6]:.6f} {M2[7]:.6f} {M2[8]:.6f} {M2[9]:.6f} {M2[10]:.6f} {M2[11]:.6f} \n')


def loop_rows_bruteforce(emb, k=50, gap=100):            # eval_...:183-207 with faiss replaced by exact L2
    rows = []
    for i in range(gap + 1, emb.shape[0] - 1):
        db = emb[:i - gap]
        d = ((db - emb[i]) ** 2).sum(1)
        order = np.lexsort((np.arange(len(d)), d))[:k]
        for j in order:
            rows.append((i, int(j), float(d[j])))
    return np.array(rows, dtype=np.float64).reshape(-1, 3)
