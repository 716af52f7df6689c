#!/usr/bin/env python
"""bench.py -- scan-pairs/sec of the LCR-Net descriptor hot path on B200.

Workload (BASELINE.json configs[1], "single-scan encoder+global-descriptor forward, 64k
synthetic pts"): every step processes a batch of P = 32 scan pairs (2P synthetic 65 536-point
KITTI-shaped scans) through raw points -> 0.3 m voxel pre-pass -> 3-level voxel pyramid ->
7 radius-neighbour tables -> 11-block KPConv encoder -> NetVLAD descriptor, one descriptor per
scan, plus the squared-L2 descriptor distance of every pair (the loop-detection score).

  value : pairs/s with the raw scans already resident in HBM (CUDA events, L2 flushed
          between timed steps, max over ranks).
  e2e   : the same through the public API from pinned HOST buffers (H2D of the raw scans and
          D2H of descriptors + distances inside the timed region).
  roofline / cpu_baseline : see DESIGN.md "Measurement".

Launch: `python bench.py --gpus 1` or
`python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N` (weak scaling: every
rank runs its own P pairs, no data-path collective: independent units).
`--impl reference` times the CPU path (reference C++ operators from oracle/_ref when shipped +
the torch-CPU oracle port of the model) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

NUM_STAGES, VOXEL, RADIUS = 4, 0.3, 4.25 * 0.3
METRIC, UNIT = 'scan_pairs_per_sec_64k_descriptor_path', 'pairs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--pairs', type=int, default=32, help='scan pairs per step per GPU')
    ap.add_argument('--cpu-pairs', type=int, default=1, help='pairs in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ncu', action='store_true', help='one warm-up step + one step only (for ncu captures)')
    ap.add_argument('--streams', type=int, default=2,
                    help='descriptor workload: chunks of the batch in flight on separate CUDA streams (pipeline.py)')
    ap.add_argument('--workload', default='descriptor', choices=['descriptor', 'db', 'pairs'],
                    help='descriptor: the headline line (configs[1]); db: descriptor-database build + all-gather + '
                         'top-25 (configs[3]/[4]); pairs: full registration of scan pairs (configs[2])')
    ap.add_argument('--db-scans', type=int, default=4000, help='db workload: scans per GPU')
    return ap.parse_args()


def make_scans(n_pairs, rank=0):
    from lcrnet_b200 import synth
    scans = []
    for i in range(n_pairs):
        ref, src, _ = synth.make_pair(scene_seed=1000 * rank + i, seed=7351 + i)
        scans += [ref, src]
    return scans


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ----------------------------------------------------------------------------- CPU path
def cpu_pair_pipeline(scans, sd, limits, use_ref_ops):
    """The reference's CPU path on the same scans: collate (C++ operators) + torch-CPU model."""
    import torch
    from oracle import model_oracle as mo
    from oracle import native as on
    descs = []
    for s in scans:
        lens = np.array([len(s)], dtype=np.int64)
        if use_ref_ops:
            p0, l0 = on.ref_grid_subsample(s, lens, VOXEL)
            pts, ls = [p0], [l0]
            v = VOXEL
            for _ in range(1, NUM_STAGES):
                v *= 2
                p, l = on.ref_grid_subsample(pts[-1], ls[-1], v)
                pts.append(p)
                ls.append(l)
            nb, sub, r = [], [], RADIUS
            for i in range(NUM_STAGES):
                nb.append(on.ref_radius_neighbors(pts[i], pts[i], ls[i], ls[i], r, limits[i]))
                if i < NUM_STAGES - 1:
                    sub.append(on.ref_radius_neighbors(pts[i + 1], pts[i], ls[i + 1], ls[i], r, limits[i]))
                r *= 2
            t = torch.from_numpy
            data = {'points': [t(x) for x in pts], 'neighbors': [t(x) for x in nb], 'subsampling': [t(x) for x in sub]}
        else:
            p0, l0 = on.grid_subsample(s, lens, VOXEL)
            data = mo.precompute_pyramid(p0, l0, NUM_STAGES, VOXEL, RADIUS, limits)
        with torch.no_grad():
            descs.append(mo.global_descriptor(sd, data))
    d = torch.cat(descs)
    return d, ((d[0::2] - d[1::2]) ** 2).sum(1)


def time_cpu(scans, sd, limits, steps=1, warmup=0):
    import torch
    from oracle import native as on
    on.build()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    use_ref = on.ref_lib() is not None
    for _ in range(warmup):
        cpu_pair_pipeline(scans, sd, limits, use_ref)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_pair_pipeline(scans, sd, limits, use_ref)
    dt = (time.perf_counter() - t0) / steps
    n_pairs = len(scans) // 2
    kind_ops = 'reference C++ operators (oracle/_ref)' if use_ref else 'C oracle operators'
    return {'value': n_pairs / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d pair(s) = %d scans of 65536 pts per step; collate single-threaded (%s), model = torch-CPU '
                      'oracle port with %d threads; %.2f s per step' % (n_pairs, len(scans), kind_ops, cores, dt)}, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from lcrnet_b200 import checkpoint
    sd = checkpoint.random_state_dict('global_descriptor', 7351)
    scans = make_scans(args.cpu_pairs)
    limits = calibrated_limits_cpu(scans[:2])
    base, dt = time_cpu(scans, sd, limits, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': min(args.warmup, 1), 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args.pairs, limits, args.streams), reference_sample_pairs=args.cpu_pairs),
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def calibrated_limits_cpu(scans):
    """calibrate_neighbors_stack_mode (data.py:408-433) with the C oracle: 80 % quantile."""
    from oracle import model_oracle as mo
    from oracle import native as on
    pre = [on.grid_subsample(s, np.array([len(s)], dtype=np.int64), VOXEL)[0] for s in scans]
    return mo.calibrate_limits([[p] for p in pre], NUM_STAGES, VOXEL, RADIUS)


def workload_config(pairs, limits, streams=1):
    return {'workload': 'configs[1]: descriptor path (0.3 m pre-voxel -> pyramid -> radius tables -> KPConv encoder '
                        '-> NetVLAD) on synthetic 64k-point scans, %d pairs (%d scans) per step per GPU, + pair '
                        'descriptor distance' % (pairs, 2 * pairs),
            'pairs_per_step_per_gpu': pairs, 'points_per_scan': 65536, 'neighbor_limits': limits,
            'weights': 'seeded random (checkpoint.random_state_dict(7351))', 'cache': 'L2 flushed between timed steps',
            'streams': streams}


# ----------------------------------------------------------------------------- GPU path
def run_b200(args):
    import torch
    import torch.distributed as dist
    from lcrnet_b200 import _lib, checkpoint, model
    from lcrnet_b200 import data as gdata
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L = _lib.lib()
    net = model.create_model(model.default_cfg()).eval()
    sd = checkpoint.random_state_dict('global_descriptor', 7351)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)

    scans = make_scans(args.pairs, rank)
    limits = gdata.calibrate_neighbors_scans(scans[:2], NUM_STAGES, VOXEL, RADIUS, pre_voxel=VOXEL, device=dev)
    host_pts = torch.from_numpy(np.concatenate(scans, 0)).pin_memory()
    host_len = torch.tensor([len(s) for s in scans], dtype=torch.int64).pin_memory()
    dev_pts, dev_len = host_pts.to(dev), host_len.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_desc = torch.empty((2 * args.pairs, 256), dtype=torch.float32).pin_memory()
    out_dist = torch.empty(args.pairs, dtype=torch.float32).pin_memory()

    from lcrnet_b200 import pipeline
    lens_list = host_len.tolist()
    pipe1 = pipeline.DescriptorPipeline(net, limits, NUM_STAGES, VOXEL, RADIUS, pre_voxel=VOXEL, n_streams=1)
    pipe = pipe1 if args.streams <= 1 or args.ncu else pipeline.DescriptorPipeline(
        net, limits, NUM_STAGES, VOXEL, RADIUS, pre_voxel=VOXEL, n_streams=args.streams)

    def step(points, lengths, p=None):
        desc = (p or pipe)(points, lens_list)
        diff = desc[0::2] - desc[1::2]
        return desc, (diff * diff).sum(1)

    def step_e2e():
        desc, dd = step(host_pts, None)           # every chunk copies its own slice of the pinned host buffer
        out_desc.copy_(desc, non_blocking=True)
        out_dist.copy_(dd, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    if args.ncu:
        step(dev_pts, dev_len)
        torch.cuda.synchronize()
        step(dev_pts, dev_len)
        torch.cuda.synchronize()
        return
    for _ in range(max(args.warmup, 3)):
        step(dev_pts, dev_len)
        step_e2e()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = L.lcr_launch_count()
    barrier()
    ms = timed(lambda: step(dev_pts, dev_len), args.steps)
    barrier()
    launches = (L.lcr_launch_count() - n0) // args.steps
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    clocks = sampler.stop() if sampler else None

    total, total_e2e = sum(ms), sum(ms_e2e)
    if world > 1:
        t = torch.tensor([total, total_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total, total_e2e = t.tolist()
    pairs_all = args.pairs * world * args.steps
    value = pairs_all / (total * 1e-3)
    e2e_value = pairs_all / (total_e2e * 1e-3)

    # parity spot check of the benchmarked configuration is in tests/; here only sanity
    desc, dd = step(dev_pts, dev_len)
    assert torch.isfinite(desc).all() and abs(float(desc.norm(dim=1).mean()) - 1.0) < 1e-4

    # per-kernel event timing needs the launches serialised: the single-stream pipeline
    roof = dominant_kernel_roofline(lambda a, b: step(a, b, pipe1), dev_pts, dev_len, flush) if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.pairs, limits, args.streams),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(host_pts.numel() * 4 + host_len.numel() * 8),
                    'd2h_bytes_per_step': int(out_desc.numel() * 4 + out_dist.numel() * 4),
                    'ms_per_step': total_e2e / args.steps},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof}
    if world == 1 and not args.no_cpu_baseline:
        cpu_scans = scans[:2 * args.cpu_pairs]
        base, _ = time_cpu(cpu_scans, sd, limits)
        line['cpu_baseline'] = base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel_roofline(step, pts, lens, flush):
    """Roofline entry for the dominant kernel of the step (the fp32 tile GEMM, gemm.cu, which
    carries the KPConv contraction and the unary layers): algorithmic FLOPs of all its launches in
    one step / their summed device time, measured live with CUDA events around every launch via
    the library's own per-kernel timing hook.  See DESIGN.md "Measurement"."""
    import torch
    from lcrnet_b200 import _lib
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    L = _lib.lib()
    if not hasattr(L, 'lcr_profile_begin'):
        return None
    L.lcr_profile_begin()
    flush.fill_(1)
    step(pts, lens)
    torch.cuda.synchronize()
    import ctypes
    n = L.lcr_profile_end()
    rows = []
    for i in range(n):
        name = ctypes.create_string_buffer(64)
        ms, flops, bytes_ = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        L.lcr_profile_get(i, name, 64, ctypes.byref(ms), ctypes.byref(flops), ctypes.byref(bytes_))
        rows.append((name.value.decode(), ms.value, flops.value, bytes_.value))
    agg = {}
    for name, ms, fl, by in rows:
        a = agg.setdefault(name, [0.0, 0.0, 0.0, 0])
        a[0] += ms
        a[1] += fl
        a[2] += by
        a[3] += 1
    if not agg:
        return None
    leaf = {k: v for k, v in agg.items() if not k.endswith('_total')}
    nested = ('radius_query', 'netvlad_hidden')          # already inside a *_total group
    total_ms = sum(v[0] for k, v in agg.items() if k not in nested)
    top = max(leaf.items(), key=lambda kv: kv[1][0])
    name, (ms, fl, by, cnt) = top
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    tens_peak = peaks.get('bf16_tflops', 1590.0)
    src = 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback (B200_PROFILING.md)'
    groups = {k: {'ms': round(v[0], 4), 'launches': v[3], 'share': round(v[0] / total_ms, 4),
                  'gflops': round(v[1] / 1e9, 3), 'mbytes': round(v[2] / 1e6, 3)} for k, v in agg.items()}
    traffic, traffic_src = ncu_traffic(name, cnt)
    if fl > 0 and name.startswith('gemm'):
        ach = fl / (ms * 1e-3) / 1e12
        note = ('achieved = fp32-equivalent algorithmic flops (2MNK) of all launches of the group in one step / their '
                'summed event-timed duration; peak = dense bf16 tensor peak.  The tcgen05 kernel issues 3 TF32 MMAs per '
                'fp32 product (3xTF32 split, DESIGN.md 5) and TF32 runs at half the bf16 rate, so the tensor pipe sees '
                '6x this fraction' if name == 'gemm_tf32x3' else
                'fp32 SIMT GEMM; fraction is against the dense bf16 tensor peak')
        return {'kernel': name, 'bound': 'tensor', 'achieved': ach, 'peak': tens_peak, 'unit': 'TFLOP/s',
                'frac': ach / tens_peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': src,
                'note': note, 'tf32_mma_tflops': 3 * ach if name == 'gemm_tf32x3' else None,
                'launches_per_step': cnt, 'ms_per_step': ms, 'share_of_step': ms / total_ms, 'kernel_groups': groups}
    ach = by / (ms * 1e-3) / 1e9
    return {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
            'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': src, 'launches_per_step': cnt,
            'ms_per_step': ms, 'share_of_step': ms / total_ms, 'kernel_groups': groups}


def ncu_traffic(group, launches):
    """dram__bytes_read.sum + dram__bytes_write.sum of the group's launches in one step, from the committed ncu
    capture of this same command (profiles/*_traffic.json, written by scripts/traffic_from_launches.py); bytes per
    step over all launches of the group, like `achieved`.  None when the capture does not match this run."""
    import glob
    for path in sorted(glob.glob(os.path.join(REPO, 'profiles', '*_traffic.json')), reverse=True):
        try:
            g = json.load(open(path))['groups'].get(group)
        except Exception:
            continue
        if g and g['launches'] == launches:
            return g['dram_bytes'], os.path.relpath(path, REPO)
    return None, None


def run_db(args):
    """configs[3]/[4]: every rank builds the descriptors of its contiguous shard of `db_scans`
    scans (a pool of 64 distinct synthetic scans cycled with a small per-batch offset: descriptor cost
    does not depend on the content; batches of 64 scans through the 2-stream DescriptorPipeline), ONE
    all-gather assembles the database, then each rank answers its own shard's queries with the
    brute-force L2 top-25 kernel.  One step = the whole build + search."""
    import torch
    import torch.distributed as dist
    from lcrnet_b200 import _lib, checkpoint, model, retrieval
    from lcrnet_b200 import data as gdata
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    net = model.create_model(model.default_cfg()).eval()
    net.load_state_dict(checkpoint.random_state_dict('global_descriptor', 7351), strict=True)
    net = net.to(dev)
    from lcrnet_b200 import pipeline
    pool = make_scans(32, rank)
    limits = gdata.calibrate_neighbors_scans(pool[:2], NUM_STAGES, VOXEL, RADIUS, pre_voxel=VOXEL, device=dev)
    pts = torch.from_numpy(np.concatenate(pool, 0)).to(dev)
    lens_list = [len(s) for s in pool]
    n_local, batch = args.db_scans, len(pool)
    jitter = torch.linspace(0.0, 0.05, steps=(n_local + batch - 1) // batch, device=dev)
    pipe = pipeline.DescriptorPipeline(net, limits, NUM_STAGES, VOXEL, RADIUS, pre_voxel=VOXEL, n_streams=args.streams)

    def step():
        out = []
        for b in range((n_local + batch - 1) // batch):
            out.append(pipe(pts + jitter[b], lens_list))
        local_db = torch.cat(out)[:n_local].contiguous()
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        db = retrieval.all_gather_descriptors(local_db)
        ev2 = torch.cuda.Event(enable_timing=True)
        ev2.record()
        d2, idx = retrieval.search(local_db, db, k=25)
        return db, idx, (ev, ev2)

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times, gather_ms, search_ms = [], [], []
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        db, idx, (g0, g1) = step()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
        gather_ms.append(g0.elapsed_time(g1))
        search_ms.append(g1.elapsed_time(b))
    # every query's nearest database row is itself (distance 0): index check across the gather
    start, _ = retrieval.shard_range(n_local * world, rank, world)
    ok = bool((idx[:, 0].cpu() == torch.arange(start, start + n_local)).float().mean() > 0.99)
    total = sum(times)
    if world > 1:
        t = torch.tensor([total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t)
    if rank == 0:
        n_all = n_local * world * args.steps
        print(json.dumps({'metric': 'db_build_scans_per_sec', 'value': n_all / (total * 1e-3), 'unit': 'scans/s',
                          'n_gpus': world, 'steps': args.steps, 'warmup': 1, 'ms_per_step': total / args.steps,
                          'higher_is_better': True, 'scaling': 'weak', 'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': 'configs[%d]: %d-scan descriptor DB build sharded %d x B200, one '
                                                 'all-gather, brute-force L2 top-25' % (3 if world == 1 else 4,
                                                                                          n_local * world, world),
                                     'scans_per_gpu': n_local, 'db_rows': n_local * world, 'k': 25},
                          'all_gather_ms': float(np.mean(gather_ms)), 'topk_ms': float(np.mean(search_ms)),
                          'all_gather_bytes': int(n_local * world * 256 * 4),
                          'queries_per_sec': n_local * world / (float(np.mean(search_ms)) * 1e-3),
                          'self_match_ok': ok}))
    if world > 1:
        dist.destroy_process_group()


def run_pairs(args):
    """configs[2]: full registration (LCRNet: encoder, 3D-RoFormer, vote, matching, LGR) of a batch of
    scan pairs per step; pairs/s."""
    import torch
    from lcrnet_b200 import checkpoint, lcrnet
    from lcrnet_b200 import data as gdata
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    scans = make_scans(args.pairs)
    limits = gdata.calibrate_neighbors_scans(scans[:2], NUM_STAGES, VOXEL, RADIUS, pre_voxel=VOXEL, device=dev)
    net = lcrnet.create_model(lcrnet.default_cfg(limits)).eval()
    net.load_state_dict(checkpoint.random_state_dict('lcrnet', 7351), strict=True)
    net = net.to(dev)
    pts = torch.from_numpy(np.concatenate(scans, 0)).to(dev)
    lens = torch.tensor([len(s) for s in scans], dtype=torch.int64, device=dev)

    from lcrnet_b200 import pipeline
    lens_list = [len(s) for s in scans]
    pipe = pipeline.PairPipeline(net, limits, NUM_STAGES, VOXEL, RADIUS, pre_voxel=VOXEL, n_streams=args.streams)
    pipe1 = pipeline.PairPipeline(net, limits, NUM_STAGES, VOXEL, RADIUS, pre_voxel=VOXEL, n_streams=1)

    def step(p=None):
        return (p or pipe)(pts, lens_list)

    for _ in range(max(1, min(args.warmup, 2))):
        out = step()
    torch.cuda.synchronize()
    times = []
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = step()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    roof = dominant_kernel_roofline(lambda p, l: step(pipe1), pts, lens, torch.empty(1 << 20, dtype=torch.uint8, device=dev))
    T = torch.stack([o['estimated_transform'] for o in out])
    n_corr = [int(o['corr_scores'].shape[0]) for o in out]
    print(json.dumps({'metric': 'registration_pairs_per_sec', 'value': args.pairs * args.steps / (sum(times) * 1e-3),
                      'unit': 'pairs/s', 'n_gpus': 1, 'steps': args.steps, 'ms_per_step': sum(times) / args.steps,
                      'higher_is_better': True, 'dtype': 'f32', 'data': 'synthetic',
                      'config': {'workload': 'configs[2]: scan-pair registration (LCRNet full forward), batch %d pairs'
                                             % args.pairs, 'neighbor_limits': limits,
                                 'weights': 'seeded random: correspondences are not meaningful', 'streams': args.streams},
                      'mean_correspondences': float(np.mean(n_corr)), 'finite': bool(torch.isfinite(T).all()),
                      'kernel_groups': roof['kernel_groups'] if roof else None}))


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    elif a.workload == 'db':
        run_db(a)
    elif a.workload == 'pairs':
        run_pairs(a)
    else:
        run_b200(a)
