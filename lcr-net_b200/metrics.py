"""Registration evaluator metrics: what ``Evaluator`` (experiments/lcrnet/loss_reg.py:278-334)
computes from an ``LCRNet`` output dict and the ground-truth transform, plus the underlying error measures of
experiments/lcrnet/modules/registration/metrics.py:47-110.  Evaluation bookkeeping on a handful of small tensors
(host glue, any device); the hot path is untouched.  The RANSAC alternative of utils/utils/open3d.py:145-173 is
open3d's (third-party, absent here) and is not reproduced: LGR (a13) is the shipped estimator."""
import math

import torch


def _rt(transform):
    transform = torch.as_tensor(transform)
    return transform[..., :3, :3], transform[..., :3, 3]


def relative_rotation_error(gt_rotations, rotations):
    """RRE = acos((trace(R^T . R_gt) - 1) / 2) in degrees (metrics.py:47-65)."""
    mat = torch.matmul(torch.as_tensor(rotations).transpose(-1, -2), torch.as_tensor(gt_rotations))
    trace = mat[..., 0, 0] + mat[..., 1, 1] + mat[..., 2, 2]
    x = (0.5 * (trace - 1.0)).clamp(min=-1.0, max=1.0)
    return 180.0 * torch.arccos(x) / math.pi


def relative_translation_error(gt_translations, translations):
    """RTE = ||t_gt - t|| (metrics.py:68-81)."""
    return torch.linalg.norm(torch.as_tensor(gt_translations) - torch.as_tensor(translations), dim=-1)


def isotropic_transform_error(gt_transforms, transforms, reduction='mean'):
    """(RRE, RTE) of 4x4 transforms (metrics.py:84-110)."""
    assert reduction in ('mean', 'sum', 'none')
    gt_r, gt_t = _rt(gt_transforms)
    r, t = _rt(transforms)
    rre, rte = relative_rotation_error(gt_r, r), relative_translation_error(gt_t, t)
    if reduction == 'mean':
        return rre.mean(), rte.mean()
    if reduction == 'sum':
        return rre.sum(), rte.sum()
    return rre, rte


def apply_transform(points, transform):
    """points [N, 3], transform [4, 4] -> R p + t (ops/transformation.py:7-61, the 2-D case)."""
    r, t = _rt(transform)
    return torch.matmul(points, r.transpose(-1, -2)) + t


def registration_recall(rre, rte, rre_threshold=5.0, rte_threshold=2.0):
    """RR of one pair: 1 if RRE < 5 deg and RTE < 2 m (loss_reg.py:318-324, config_ld.py:51-52)."""
    return torch.logical_and(torch.lt(rre, rre_threshold), torch.lt(rte, rte_threshold)).float()


def inlier_ratio(pos_corr_points, anc_corr_points, gt_transform, acceptance_radius=1.0):
    """IR: fraction of fine correspondences within ``acceptance_radius`` under the ground truth (loss_reg.py:308-316)."""
    d = torch.linalg.norm(pos_corr_points - apply_transform(anc_corr_points, gt_transform), dim=1)
    return torch.lt(d, acceptance_radius).float().mean()


def coarse_precision(pos_node_corr_indices, anc_node_corr_indices, gt_node_corr_indices, gt_node_corr_overlaps,
                     n_pos, n_anc, acceptance_overlap=0.0):
    """PIR: fraction of predicted node correspondences that are ground-truth overlapping patches
    (loss_reg.py:286-306)."""
    keep = torch.gt(gt_node_corr_overlaps, acceptance_overlap)
    gt = gt_node_corr_indices[keep]
    gt_map = torch.zeros((n_pos, n_anc), device=gt.device)
    gt_map[gt[:, 0], gt[:, 1]] = 1.0
    return gt_map[pos_node_corr_indices, anc_node_corr_indices].mean()


def evaluate(output_dict, gt_transform, gt_node_corr_indices=None, gt_node_corr_overlaps=None,
             acceptance_radius=1.0, rre_threshold=5.0, rte_threshold=2.0, acceptance_overlap=0.0):
    """The ``Evaluator.forward`` result dict (loss_reg.py:326-334): IR, RRE, RTE, RR (+ PIR when the ground-truth
    node correspondences are given)."""
    gt = torch.as_tensor(gt_transform, dtype=torch.float32, device=output_dict['estimated_transform'].device)
    rre, rte = isotropic_transform_error(gt, output_dict['estimated_transform'])
    res = {'IR': inlier_ratio(output_dict['pos_corr_points'], output_dict['anc_corr_points'], gt, acceptance_radius),
           'RRE': rre, 'RTE': rte, 'RR': registration_recall(rre, rte, rre_threshold, rte_threshold)}
    if gt_node_corr_indices is not None:
        res['PIR'] = coarse_precision(output_dict['pos_node_corr_indices'], output_dict['anc_node_corr_indices'],
                                      gt_node_corr_indices, gt_node_corr_overlaps,
                                      output_dict['pos_points_c'].shape[0], output_dict['anc_points_c'].shape[0],
                                      acceptance_overlap)
    return res
