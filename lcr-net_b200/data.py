"""Stack-mode pyramid construction on the GPU: the reference's collate hot loop
(experiments/lcrnet/data.py:10-74 ``precompute_data_stack_mode`` and the collate functions
:77-127, :350-406) with both native operators running as CUDA kernels in the main process.

The reference runs this in DataLoader worker processes on the CPU; a CUDA replacement runs in
the main process on the current stream (the reference supports worker collates that only stack
raw points via ``precompute_data=False``, data.py:119-124).
"""
import numpy as np
import torch

from . import ops


def precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits, int32=False,
                               upsampling=True, order='reference'):
    """Same arguments / result dict as data.py:10-74.  Extra keys: ``lengths_host`` (list of
    python int lists, saves the model a device sync).  ``int32`` keeps the tables in the kernels'
    native index width; ``upsampling=False`` skips the three tables only the registration decoder
    reads (the descriptor path never touches them).  With ``int32`` the searches are queued back to back
    (no per-table read-back: tables keep ``neighbor_limits[i]`` columns, pads = number of support rows) and their
    status words ride on the single device->host read of the level lengths."""
    assert num_stages == len(neighbor_limits)
    from . import ext
    defer = [] if int32 and all(l and l > 0 for l in neighbor_limits) else None
    points_list, lengths_list, lengths_host = [], [], []
    neighbors_list, subsampling_list, upsampling_list = [], [], []
    for i in range(num_stages):
        if i > 0:
            points, lengths = ops.grid_subsample(points, lengths, voxel_size=voxel_size, order=order)
        points_list.append(points)
        lengths_list.append(lengths)
        voxel_size *= 2
    for i in range(num_stages):
        cur_points, cur_lengths = points_list[i], lengths_list[i]
        neighbors_list.append(ops.radius_search(cur_points, cur_points, cur_lengths, cur_lengths, radius,
                                                neighbor_limits[i], int32=int32, defer=defer))
        if i < num_stages - 1:
            sub_points, sub_lengths = points_list[i + 1], lengths_list[i + 1]
            subsampling_list.append(ops.radius_search(sub_points, cur_points, sub_lengths, cur_lengths, radius,
                                                      neighbor_limits[i], int32=int32, defer=defer))
            if upsampling:
                upsampling_list.append(ops.radius_search(cur_points, sub_points, cur_lengths, sub_lengths, radius * 2,
                                                         neighbor_limits[i + 1], int32=int32, defer=defer))
        radius *= 2
    if defer:
        n_len = lengths_list[0].numel() * num_stages
        both = torch.cat([torch.stack(lengths_list).reshape(-1)] + [m.to(torch.int64) for m in defer]).cpu()
        lengths_host = both[:n_len].reshape(num_stages, -1).tolist()
        ext.check_deferred(both[n_len:].reshape(-1, 2).tolist())
    else:
        lengths_host = torch.stack(lengths_list).cpu().tolist()
    return {
        'points': points_list,
        'lengths': lengths_list,
        'lengths_host': lengths_host,
        'neighbors': neighbors_list,
        'subsampling': subsampling_list,
        'upsampling': upsampling_list,
    }


def scans_collate_fn_stack_mode(scans, num_stages, voxel_size, search_radius, neighbor_limits, pre_voxel=None,
                                stack_size=1, int32=True, upsampling=False, device='cuda'):
    """Batch of raw scans -> one data_dict (the shape of
    ``test_loop_detection_collate_fn_stack_mode_online``, data.py:350-406, for many scans at once).
    scans: list of float32 [Ni, 3] (numpy or torch, host or device).  ``pre_voxel`` applies the
    offline 0.3 m voxel pre-pass (data/Kitti/downsample_pcd.py:29 stands for it) on the GPU first.
    ``stack_size`` consecutive scans form one stack (1: descriptor path, 2: a registration pair)."""
    ts = [torch.as_tensor(s, dtype=torch.float32) for s in scans]
    lengths = torch.tensor([t.shape[0] for t in ts], dtype=torch.int64)
    points = torch.cat(ts, 0).contiguous()
    points, lengths = points.to(device, non_blocking=True), lengths.to(device, non_blocking=True)
    return device_collate(points, lengths, num_stages, voxel_size, search_radius, neighbor_limits, pre_voxel,
                          stack_size, int32, upsampling)


def device_collate(points, lengths, num_stages, voxel_size, search_radius, neighbor_limits, pre_voxel=None,
                   stack_size=1, int32=True, upsampling=False):
    """As scans_collate_fn_stack_mode for points/lengths that are already resident on the GPU."""
    if pre_voxel:
        points, lengths = ops.grid_subsample(points, lengths, pre_voxel)
    d = precompute_data_stack_mode(points, lengths, num_stages, voxel_size, search_radius, neighbor_limits,
                                   int32=int32, upsampling=upsampling)
    d['features'] = torch.ones((d['points'][0].shape[0], 1), dtype=torch.float32, device=points.device)
    d['batch_size'] = len(d['lengths_host'][0]) // stack_size
    d['stack_size'] = stack_size
    return d


def calibrate_neighbors_stack_mode(scans, num_stages, voxel_size, search_radius, keep_ratio=0.8,
                                   sample_threshold=2000, pre_voxel=None, device='cuda'):
    """data.py:408-433: histogram of neighbourhood sizes per level (table limit 'hist_n' = the
    number of points in a ball of the search radius at unit voxel density), keep the
    ``keep_ratio`` quantile.  Runs the counting pass of the radius kernel only."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (search_radius / voxel_size + 1) ** 3))
    hists = np.zeros((num_stages, hist_n), dtype=np.int64)
    max_limits = [hist_n] * num_stages
    for scan in scans:
        d = scans_collate_fn_stack_mode([scan], num_stages, voxel_size, search_radius, max_limits,
                                        pre_voxel=pre_voxel, int32=True, upsampling=False, device=device)
        counts = [(nb < nb.shape[0]).sum(1).cpu().numpy() for nb in d['neighbors']]
        hists += np.stack([np.bincount(c, minlength=hist_n)[:hist_n] for c in counts])
        if np.min(np.sum(hists, axis=1)) > sample_threshold:
            break
    cum = np.cumsum(hists.T, axis=0)
    return [int(x) for x in np.sum(cum < (keep_ratio * cum[hist_n - 1, :]), axis=0)]
