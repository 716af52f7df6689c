"""Batch drivers for the descriptor path: a batch of scans is cut into contiguous chunks, each
chunk runs collate + encoder + NetVLAD on its own CUDA stream, driven by its own host thread
(ctypes and torch release the GIL while a call is in flight).

Why: one chunk's pipeline is a chain of ~290 dependent launches with 11 device->host size
read-backs (the subsample / radius operators size their outputs on the host, as the reference's
operators do); on a single stream every read-back drains the GPU and every kernel's tail leaves
SMs idle.  With two or more chunks in flight the tails and drains of one chunk are filled by
kernels of another; the host->device copy of a chunk overlaps the compute of its predecessor.
Scans are independent units, so results are identical to the single-stream path.
"""
from concurrent.futures import ThreadPoolExecutor

import torch

from . import data as gdata
from . import ops


def _record_stream(obj, stream):
    """Marks every tensor reachable from ``obj`` as in use on ``stream`` (torch's caching allocator then keeps
    the block away from the side stream that allocated it until the consumer's queued work has run)."""
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream(v, stream)


class DescriptorPipeline:
    """``pipe(points, lengths)`` -> descriptors f32[n_scans, 256] on the device.

    points  float32 [sum(lengths), 3], pinned host memory or device memory
    lengths python list / int64 host tensor with the number of points of every scan
    """

    def __init__(self, net, neighbor_limits, num_stages=4, voxel_size=0.3, search_radius=1.275, pre_voxel=0.3,
                 n_streams=2, device=None):
        self.net, self.limits = net, list(neighbor_limits)
        self.num_stages, self.voxel, self.radius, self.pre_voxel = num_stages, voxel_size, search_radius, pre_voxel
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_streams = max(1, int(n_streams))
        self.streams = [torch.cuda.Stream(self.device) for _ in range(self.n_streams)]
        self.pool = ThreadPoolExecutor(self.n_streams) if self.n_streams > 1 else None
        # every derived weight is built once, here, and the device is synchronised: the chunks then only READ the
        # caches (the entries also carry events, ops.Derived, should a weight be updated in place later)
        with torch.cuda.device(self.device):
            ops.prepare(net)

    def _chunk(self, points, lens, lo, hi, row_lo, row_hi, out, stream, ready):
        torch.cuda.set_device(self.device)
        with torch.cuda.stream(stream):
            stream.wait_event(ready)
            p = points[row_lo:row_hi]
            if not p.is_cuda:
                p = p.to(self.device, non_blocking=True)
            l = torch.tensor(lens[lo:hi], dtype=torch.int64).to(self.device, non_blocking=True)
            d = gdata.device_collate(p, l, self.num_stages, self.voxel, self.radius, self.limits,
                                     pre_voxel=self.pre_voxel, stack_size=1, int32=True, upsampling=False)
            out[lo:hi] = self.net(d)['anc_global']
            done = torch.cuda.Event()
            done.record(stream)
        return done

    def __call__(self, points, lengths):
        lens = [int(x) for x in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
        n = len(lens)
        out = torch.empty((n, 256), dtype=torch.float32, device=self.device)
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        k = min(self.n_streams, n)
        bounds = [n * i // k for i in range(k + 1)]
        rows = [0]
        for x in lens:
            rows.append(rows[-1] + x)
        jobs = [(points, lens, bounds[i], bounds[i + 1], rows[bounds[i]], rows[bounds[i + 1]], out, self.streams[i],
                 ready) for i in range(k)]
        if self.pool is None or k == 1:
            events = [self._chunk(*j) for j in jobs]
        else:
            events = [f.result() for f in [self.pool.submit(self._chunk, *j) for j in jobs]]
        for e in events:
            main.wait_event(e)
        return out

    def close(self):
        if self.pool is not None:
            self.pool.shutdown()
            self.pool = None


class PairPipeline:
    """``pipe(points, lengths)`` -> list of per-pair output dicts of ``LCRNet`` (registration path).

    The batch of pairs (2 scans each, consecutive) is cut into ``n_streams`` chunks, each collated and
    registered on its own CUDA stream from its own host thread.  The forward is batched over the pairs of a chunk
    (one launch per matching stage, two device->host size read-backs per forward), so ONE chunk is the fastest
    setting (measured at 32 pairs per step: 1 / 2 / 3 / 4 chunks = 455 / 416 / 366 / 344 pairs/s); more chunks only
    help when the caller needs the first results early.  Pairs are independent units: results equal the
    single-stream path."""

    def __init__(self, net, neighbor_limits, num_stages=4, voxel_size=0.3, search_radius=1.275, pre_voxel=0.3,
                 n_streams=1, device=None):
        self.net, self.limits = net, list(neighbor_limits)
        self.num_stages, self.voxel, self.radius, self.pre_voxel = num_stages, voxel_size, search_radius, pre_voxel
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_streams = max(1, int(n_streams))
        self.streams = [torch.cuda.Stream(self.device) for _ in range(self.n_streams)]
        self.pool = ThreadPoolExecutor(self.n_streams) if self.n_streams > 1 else None
        with torch.cuda.device(self.device):
            ops.prepare(net)

    def _chunk(self, points, lens, lo, hi, row_lo, row_hi, stream, ready):
        torch.cuda.set_device(self.device)
        with torch.cuda.stream(stream):
            stream.wait_event(ready)
            p = points[row_lo:row_hi]
            if not p.is_cuda:
                p = p.to(self.device, non_blocking=True)
            l = torch.tensor(lens[lo:hi], dtype=torch.int64).to(self.device, non_blocking=True)
            d = gdata.device_collate(p, l, self.num_stages, self.voxel, self.radius, self.limits,
                                     pre_voxel=self.pre_voxel, stack_size=2, int32=True, upsampling='nearest')
            out = self.net(d)
            done = torch.cuda.Event()
            done.record(stream)
        n_pairs = (hi - lo) // 2
        if n_pairs == 1:
            return [out], done
        keys = list(out.keys())
        return [{k: out[k][i] for k in keys} for i in range(n_pairs)], done

    def __call__(self, points, lengths):
        lens = [int(x) for x in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
        assert len(lens) % 2 == 0, 'scans come in (reference, source) pairs'
        n_pairs = len(lens) // 2
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        k = min(self.n_streams, n_pairs)
        bounds = [2 * (n_pairs * i // k) for i in range(k + 1)]
        rows = [0]
        for x in lens:
            rows.append(rows[-1] + x)
        jobs = [(points, lens, bounds[i], bounds[i + 1], rows[bounds[i]], rows[bounds[i + 1]], self.streams[i], ready)
                for i in range(k)]
        if self.pool is None or k == 1:
            res = [self._chunk(*j) for j in jobs]
        else:
            res = [f.result() for f in [self.pool.submit(self._chunk, *j) for j in jobs]]
        outs = []
        for o, e in res:
            main.wait_event(e)
            _record_stream(o, main)       # allocated on a side stream, consumed by the caller on ``main``
            outs += o
        return outs

    def close(self):
        if self.pool is not None:
            self.pool.shutdown()
            self.pool = None
