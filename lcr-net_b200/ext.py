"""Drop-in for the reference's native operator module ``utils.ext``
(utils/extensions/pybind.cpp:7-24): same function names, argument order, dtypes, shapes and
padding conventions.  ``install()`` registers it as ``sys.modules['utils.ext']`` so the
reference's ``ops/grid_subsample.py:4`` / ``ops/radius_search.py:4`` resolve to it.

Differences from the reference module (all additive):
* tensors may live on the GPU (results are returned on the input's device); CPU tensors are
  copied to the current CUDA device and the result copied back, so DataLoader-style callers work;
* extra keyword arguments (``order``, ``limit``, ``int32``) used by the fused pipeline.
Errors: wrong dtype / shape / contiguity raise RuntimeError like the reference's TORCH_CHECKs
(radius_neighbors.cpp:12-23, grid_subsampling.cpp:10-15).
"""
import sys
import types

import torch

from . import _lib


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _to_dev(t):
    return t if t.is_cuda else t.cuda(non_blocking=True)


def grid_subsampling(points, lengths, voxel_size, order='reference'):
    """``utils.ext.grid_subsampling`` (grid_subsampling.cpp:5-62): returns [s_points, s_lengths]."""
    _check(points.dtype == torch.float32, 'points must be a float tensor')
    _check(lengths.dtype == torch.int64, 'lengths must be an long tensor')
    _check(points.is_contiguous(), 'points must be contiguous')
    _check(lengths.is_contiguous(), 'lengths must be contiguous')
    _check(points.dim() == 2 and points.shape[1] == 3, 'points must be (N, 3)')
    if not torch.cuda.is_available():
        raise RuntimeError('lcrnet_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    on_cpu = not points.is_cuda
    p, l = _to_dev(points), _to_dev(lengths)
    n, b = p.shape[0], l.shape[0]
    L = _lib.lib()
    out = torch.empty_like(p)
    out_len = torch.empty_like(l)
    meta = torch.empty(2, dtype=torch.int64, device=p.device)  # [total, status(i32 in low half)]
    status = meta[1:].view(torch.int32)
    ws_bytes = L.lcr_grid_subsample_ws_bytes(n, b)
    ws = _lib.workspace.get(ws_bytes, p.device)
    _lib.check(L.lcr_grid_subsample(_lib.ptr(p), n, _lib.ptr(l), b, float(voxel_size), 1 if order == 'reference' else 0,
                                    _lib.ptr(out), _lib.ptr(out_len), _lib.ptr(meta), _lib.ptr(status),
                                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr(p.device)))
    host = torch.cat([meta, out_len]).cpu()  # the single D2H sync of this operator
    total = int(host[0])
    if int(host[1]) & 0xFFFFFFFF:
        raise RuntimeError('grid_subsampling: voxel key overflow (cloud extent / voxel_size too large)')
    s_points = out[:total]
    s_lengths = out_len
    if on_cpu:
        return [s_points.cpu(), host[2:].clone()]
    return [s_points, s_lengths]


class SupportGrid:
    """The uniform grid of one support set at one radius, kept in its own workspace so that several searches
    (self / subsampling / upsampling tables of a pyramid level, data.py:28-66) build it once."""

    def __init__(self, s_points, s_lengths, radius, max_queries):
        self.key = (s_points.data_ptr(), s_points.shape[0], s_lengths.data_ptr(), float(radius))
        nbytes = _lib.lib().lcr_radius_neighbors_ws_bytes(int(max_queries), s_points.shape[0], s_lengths.shape[0])
        self.ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=s_points.device)
        self.max_queries = int(max_queries)
        self.built = False


def radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius, limit=0, int32=False, defer=None, grid=None,
                     nearest=False):
    """``utils.ext.radius_neighbors`` (radius_neighbors.cpp:5-68): (Nq, max_count) int64 table,
    padded with Ns.  ``limit`` > 0 fuses the ``[:, :limit]`` cut of ops/radius_search.py:25-26.
    ``defer`` (a list, with ``limit`` > 0): do not read the [max_count, status] words back -- the table keeps
    ``limit`` columns (pads = Ns) and the device words are appended to the list for ONE later check
    (``check_deferred``): the pyramid builder queues its 7-10 searches without draining the stream.
    ``grid`` (SupportGrid of these supports and this radius): built by the first search that uses it, reused after.
    ``nearest``: only column 0 of the table (Nq, 1) -- the closest support inside the radius -- without the sort."""
    for name, t in (('q_points', q_points), ('s_points', s_points)):
        _check(t.dtype == torch.float32, '%s must be a float tensor' % name)
        _check(t.is_contiguous(), '%s must be contiguous' % name)
        _check(t.dim() == 2 and t.shape[1] == 3, '%s must be (N, 3)' % name)
    for name, t in (('q_lengths', q_lengths), ('s_lengths', s_lengths)):
        _check(t.dtype == torch.int64, '%s must be an long tensor' % name)
        _check(t.is_contiguous(), '%s must be contiguous' % name)
    _check(q_lengths.shape[0] == s_lengths.shape[0], 'q_lengths and s_lengths must have the same batch size')
    if not torch.cuda.is_available():
        raise RuntimeError('lcrnet_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    on_cpu = not q_points.is_cuda
    q, s, ql, sl = _to_dev(q_points), _to_dev(s_points), _to_dev(q_lengths), _to_dev(s_lengths)
    nq, ns, b = q.shape[0], s.shape[0], ql.shape[0]
    L = _lib.lib()
    dev = q.device
    meta = torch.zeros(2, dtype=torch.int32, device=dev)  # [max_count, status]
    if grid is not None and not on_cpu:
        _check(grid.key == (s.data_ptr(), ns, sl.data_ptr(), float(radius)) and nq <= grid.max_queries,
               'radius_neighbors: the SupportGrid belongs to other supports / radius')
        ws = grid.ws
    else:
        grid = None
        ws_bytes = L.lcr_radius_neighbors_ws_bytes(nq, ns, b)
        ws = _lib.workspace.get(ws_bytes, dev)
    stream = _lib.stream_ptr(dev)
    dtype = torch.int32 if int32 else torch.int64

    def run(width, out):
        reuse = (1 if (grid is not None and grid.built) else 0) | (2 if nearest else 0)
        _lib.check(L.lcr_radius_neighbors_ex(_lib.ptr(q), nq, _lib.ptr(s), ns, _lib.ptr(ql), _lib.ptr(sl), b,
                                             float(radius), width, _lib.ptr(out), 0 if int32 else 1, None,
                                             _lib.ptr(meta), _lib.ptr(meta[1:]), _lib.ptr(ws), ws.numel(), reuse,
                                             stream))
        if grid is not None:
            grid.built = True

    if nearest:
        out = torch.empty((nq, 1), dtype=dtype, device=dev)
        run(1, out)
        if defer is not None and not on_cpu:
            defer.append(meta)
            return out
        st = int(meta[1])
    elif limit and limit > 0:
        out = torch.empty((nq, limit), dtype=dtype, device=dev)
        run(limit, out)
        if defer is not None and not on_cpu:
            defer.append(meta)
            return out
        mc, st = meta.tolist()
        if mc < limit:
            out = out[:, :mc]
    else:
        run(0, None)  # counting pass: the reference width is the maximum count over all queries
        mc, st = meta.tolist()
        out = torch.empty((nq, mc), dtype=dtype, device=dev)
        if mc > 0 and st == 0:
            _check(mc <= 8192, 'radius_neighbors: more than 8192 neighbours per query')
            run(mc, out)
            st = int(meta[1])
    if st != 0:
        raise RuntimeError('radius_neighbors: capacity exceeded (cloud spans > 16384 cells per axis or '
                           '> 8192 neighbours for one query)')
    return out.cpu() if on_cpu else out


def check_deferred(status_words):
    """Raise if any deferred search overflowed; ``status_words`` = host list of [max_count, status] pairs."""
    for mc, st in status_words:
        if st != 0:
            raise RuntimeError('radius_neighbors: capacity exceeded (cloud spans > 16384 cells per axis or '
                               '> 8192 neighbours for one query)')


def radius_filter(*args, **kwargs):
    """Exported by the reference (pybind.cpp:19-23) but never called on the inference path
    (only a commented-out call at modules/vote/vote.py:91); greedy NMS replaces it."""
    raise NotImplementedError('utils.ext.radius_filter is not on the LCR-Net inference path')


def install():
    """Register this module as ``utils.ext`` (and a ``utils`` namespace if none is importable)."""
    if 'utils' not in sys.modules:
        try:
            __import__('utils')
        except ImportError:
            pkg = types.ModuleType('utils')
            pkg.__path__ = []
            sys.modules['utils'] = pkg
    mod = sys.modules[__name__]
    sys.modules['utils.ext'] = mod
    setattr(sys.modules['utils'], 'ext', mod)
    return mod
