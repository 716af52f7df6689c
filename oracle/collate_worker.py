"""Process-pool worker for the CPU baseline's collate figure (TEST / BENCH INFRASTRUCTURE ONLY).

The reference builds its pyramid in DataLoader worker processes (``num_workers=8``, config_reg.py:55;
utils/utils/torch.py:65-75); this module is what one such worker does for one scan: 0.3 m pre-voxel stand-in,
3-level voxel pyramid and the 7 (+3 upsampling) radius tables with the reference's own C++ operators
(oracle/_ref) when shipped, the C restatement otherwise.  Imports numpy only (cheap to spawn)."""
import time

import numpy as np

from . import native as on

NUM_STAGES, VOXEL, RADIUS = 4, 0.3, 4.25 * 0.3


def collate_scan(args):
    scan, limits, upsampling = args
    use_ref = on.ref_lib() is not None
    sub = on.ref_grid_subsample if use_ref else on.grid_subsample
    rad = on.ref_radius_neighbors if use_ref else on.radius_neighbors
    t0 = time.perf_counter()
    lens = np.array([len(scan)], dtype=np.int64)
    p, l = sub(scan, lens, VOXEL)
    pts, ls, v = [p], [l], VOXEL
    for _ in range(1, NUM_STAGES):
        v *= 2
        p, l = sub(pts[-1], ls[-1], v)
        pts.append(p)
        ls.append(l)
    r, rows = RADIUS, 0
    for i in range(NUM_STAGES):
        rows += rad(pts[i], pts[i], ls[i], ls[i], r, limits[i]).shape[0]
        if i < NUM_STAGES - 1:
            rows += rad(pts[i + 1], pts[i], ls[i + 1], ls[i], r, limits[i]).shape[0]
            if upsampling:
                rows += rad(pts[i], pts[i + 1], ls[i], ls[i + 1], r * 2, limits[i + 1]).shape[0]
        r *= 2
    return time.perf_counter() - t0, rows, use_ref


def collate_rate(scans, limits, processes=8, upsampling=False, timeout=300):
    """scans/s of ``processes`` worker PROCESSES collating ``scans`` concurrently (the reference's DataLoader
    workers), and of one process.  Workers are plain subprocesses of this module's CLI (no fork of a CUDA parent,
    killed on timeout); the rate is scans / (last worker's end - first worker's start), library loading excluded."""
    import json
    import os
    import subprocess
    import sys
    import tempfile
    t0 = time.perf_counter()
    one = [collate_scan((s, list(limits), upsampling)) for s in scans[:2]]
    t_one = (time.perf_counter() - t0) / len(one)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        shares = [scans[i::processes] for i in range(processes)]
        procs = []
        for i, share in enumerate(shares):
            if not share:
                continue
            path = os.path.join(tmp, 'share%d.npz' % i)
            np.savez(path, *share)
            procs.append(subprocess.Popen([sys.executable, '-m', 'oracle.collate_worker', path, json.dumps(list(limits)),
                                           '1' if upsampling else '0'], cwd=root, stdout=subprocess.PIPE, text=True))
        spans = []
        for pr in procs:
            try:
                out, _ = pr.communicate(timeout=timeout)
                spans.append(json.loads(out.strip().splitlines()[-1]))
            except Exception:
                pr.kill()
                return {'processes': processes, 'error': 'worker failed or timed out'}
    dt = max(s['end'] for s in spans) - min(s['start'] for s in spans)
    return {'processes': len(procs), 'scans': len(scans), 'scans_per_s': len(scans) / dt,
            'one_process_scans_per_s': 1.0 / t_one, 'upsampling_tables': bool(upsampling),
            'operators': 'reference C++ (oracle/_ref)' if one[0][2] else 'C oracle'}


if __name__ == '__main__':
    import json
    import sys
    data = np.load(sys.argv[1])
    lim, ups = json.loads(sys.argv[2]), sys.argv[3] == '1'
    share = [data[k] for k in data.files]
    collate_scan((share[0][:2048].copy(), lim, ups))      # load the libraries outside the timed span
    start = time.time()
    for s in share:
        collate_scan((s, lim, ups))
    print(json.dumps({'start': start, 'end': time.time(), 'scans': len(share)}))
