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
    reads (the descriptor path never touches them); ``upsampling='nearest'`` keeps only their column 0 (shape [N, 1]: all
    the decoder's nearest_upsample reads), found without the sort.  With ``int32`` the searches are queued back to back
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
    # one support grid per level (supports = level i, radius r_i): the self table, the subsampling table of level
    # i + 1 and the upsampling table of level i - 1 (radius 2 r_{i-1} = r_i) all search it
    radii = [radius * 2 ** i for i in range(num_stages)]
    rows = [p.shape[0] for p in points_list]
    grids = [ext.SupportGrid(points_list[i], lengths_list[i], radii[i], max(rows[max(i - 1, 0):i + 2]))
             if points_list[i].is_cuda else None for i in range(num_stages)]
    upsampling_list = [None] * (num_stages - 1)
    for i in range(num_stages):
        cur_points, cur_lengths = points_list[i], lengths_list[i]
        neighbors_list.append(ops.radius_search(cur_points, cur_points, cur_lengths, cur_lengths, radii[i],
                                                neighbor_limits[i], int32=int32, defer=defer, grid=grids[i]))
        if i < num_stages - 1:
            sub_points, sub_lengths = points_list[i + 1], lengths_list[i + 1]
            subsampling_list.append(ops.radius_search(sub_points, cur_points, sub_lengths, cur_lengths, radii[i],
                                                      neighbor_limits[i], int32=int32, defer=defer, grid=grids[i]))
        if i > 0 and upsampling:
            fine_points, fine_lengths = points_list[i - 1], lengths_list[i - 1]
            upsampling_list[i - 1] = ops.radius_search(fine_points, cur_points, fine_lengths, cur_lengths, radii[i],
                                                       neighbor_limits[i], int32=int32, defer=defer, grid=grids[i],
                                                       nearest=(upsampling == 'nearest'))
    if not upsampling:
        upsampling_list = []
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


def _merge_samples(data_dicts):
    """data.py:96-105 / :377-385: values with the same key gathered into lists, numpy -> torch."""
    collated = {}
    for data_dict in data_dicts:
        for key, value in data_dict.items():
            if isinstance(value, np.ndarray):
                value = torch.from_numpy(value)
            collated.setdefault(key, []).append(value)
    return collated


def _finish_collate(collated, feats, points_list, batch_size, num_stages, voxel_size, search_radius, neighbor_limits,
                    precompute_data, device, stack_size):
    lengths = torch.LongTensor([p.shape[0] for p in points_list])
    points = torch.cat([torch.as_tensor(p, dtype=torch.float32) for p in points_list], dim=0)
    if batch_size == 1:
        for key, value in collated.items():           # remove wrapping brackets if batch_size is 1
            collated[key] = value[0]
    collated['features'] = feats
    if precompute_data:
        from . import _lib
        _lib.require_cuda()
        dev = torch.device(device)
        d = precompute_data_stack_mode(points.contiguous().to(dev, non_blocking=True), lengths.to(dev), num_stages,
                                       voxel_size, search_radius, neighbor_limits)
        collated.update(d)
        collated['features'] = feats.to(dev)
        if stack_size is not None:
            collated['stack_size'] = stack_size
    else:
        collated['points'] = points
        collated['lengths'] = lengths
    collated['batch_size'] = batch_size
    return collated


def registration_collate_fn_stack_mode(data_dicts, num_stages, voxel_size, search_radius, neighbor_limits,
                                       precompute_data=True, device='cuda'):
    """data.py:77-127, same arguments, same dict: points ordered [ref_1..ref_B, src_1..src_B], ``features`` the
    concatenated ref/src features, the pyramid keys of precompute_data_stack_mode (int64 tables, as the reference
    returns them), ``batch_size``; every other key passed through (unwrapped when batch_size == 1).  The pyramid is
    built on ``device`` -- this collate runs in the main process (DataLoader ``num_workers=0``); worker processes
    may still stack raw points with ``precompute_data=False`` (data.py:119-124)."""
    batch_size = len(data_dicts)
    collated = _merge_samples(data_dicts)
    feats = torch.cat(collated.pop('ref_feats') + collated.pop('src_feats'), dim=0)
    points_list = collated.pop('ref_points') + collated.pop('src_points')
    # the model treats the whole collated sample as ONE stack (reference GroupNorm semantics): stack_size None
    return _finish_collate(collated, feats, points_list, batch_size, num_stages, voxel_size, search_radius,
                           neighbor_limits, precompute_data, device, None)


def test_loop_detection_collate_fn_stack_mode_online(data_dicts, num_stages, voxel_size, search_radius,
                                                     neighbor_limits, precompute_data=True, device='cuda'):
    """data.py:350-406, same arguments, same dict (``features`` = the first sample's ``anc_feats``, like the
    reference)."""
    batch_size = len(data_dicts)
    collated = _merge_samples(data_dicts)
    feats = collated.pop('anc_feats')
    points_list = collated.pop('anc_points')
    return _finish_collate(collated, feats[0], points_list, batch_size, num_stages, voxel_size, search_radius,
                           neighbor_limits, precompute_data, device, None)


test_loop_detection_collate_fn_stack_mode_online.__test__ = False      # a collate function, not a pytest test


def calibrate_neighbors_stack_mode(dataset, collate_fn, num_stages, voxel_size, search_radius, keep_ratio=0.8,
                                   sample_threshold=2000):
    """data.py:408-433, same arguments: histogram of neighbourhood sizes per level over the samples of
    ``dataset`` (collated ONE SAMPLE at a time with a table limit of 'hist_n' = the number of points in a ball of
    the search radius at unit voxel density; a sample may hold several clouds, e.g. a registration pair), stop
    once every level has seen more than ``sample_threshold`` points, keep the ``keep_ratio`` quantile.  A row with
    hist_n or more neighbours falls off the histogram exactly like in the reference (``bincount(...)[:hist_n]``)."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (search_radius / voxel_size + 1) ** 3))
    hists = np.zeros((num_stages, hist_n), dtype=np.int64)
    max_limits = [hist_n] * num_stages
    for i in range(len(dataset)):
        d = collate_fn([dataset[i]], num_stages, voxel_size, search_radius, max_limits, precompute_data=True)
        counts = [(nb < nb.shape[0]).sum(1).cpu().numpy() for nb in d['neighbors']]
        hists += np.stack([np.bincount(c, minlength=hist_n)[:hist_n] for c in counts])
        if np.min(np.sum(hists, axis=1)) > sample_threshold:
            break
    cum = np.cumsum(hists.T, axis=0)
    return [int(x) for x in np.sum(cum < (keep_ratio * cum[hist_n - 1, :]), axis=0)]


def calibrate_neighbors_scans(scans, num_stages, voxel_size, search_radius, keep_ratio=0.8, sample_threshold=2000,
                              pre_voxel=None, device='cuda', scans_per_sample=1):
    """calibrate_neighbors_stack_mode for a plain list of raw scans: ``scans_per_sample`` consecutive scans form
    one sample (2 = the reference's registration samples, a stacked pair), optional 0.3 m ``pre_voxel``."""
    samples = [scans[i:i + scans_per_sample] for i in range(0, len(scans), scans_per_sample)]

    def collate_fn(sample, num_stages, voxel_size, search_radius, limits, precompute_data=True):
        return scans_collate_fn_stack_mode(sample[0], num_stages, voxel_size, search_radius, limits,
                                           pre_voxel=pre_voxel, int32=True, upsampling=False, device=device)
    return calibrate_neighbors_stack_mode(samples, collate_fn, num_stages, voxel_size, search_radius, keep_ratio,
                                          sample_threshold)


def build_dataloader_stack_mode(dataset, collate_fn, num_stages, voxel_size, search_radius, neighbor_limits,
                                batch_size=1, num_workers=0, shuffle=False, drop_last=False, distributed=False,
                                precompute_data=True, pin_memory=False):
    """data.py:436-468 (+ utils/utils/torch.py:53-80).  ``num_workers`` must be 0 when ``precompute_data`` is set:
    the pyramid is built by CUDA kernels in the main process (threading row)."""
    from functools import partial
    if precompute_data and num_workers != 0:
        raise RuntimeError('the B200 collate builds the pyramid on the GPU in the main process: use num_workers=0 '
                           '(or precompute_data=False in the workers)')
    sampler = torch.utils.data.DistributedSampler(dataset) if distributed else None
    return torch.utils.data.DataLoader(
        dataset, batch_size=batch_size, num_workers=num_workers, shuffle=False if distributed else shuffle,
        sampler=sampler, drop_last=drop_last, pin_memory=pin_memory,
        collate_fn=partial(collate_fn, num_stages=num_stages, voxel_size=voxel_size, search_radius=search_radius,
                           neighbor_limits=neighbor_limits, precompute_data=precompute_data))
