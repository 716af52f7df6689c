"""Diagnostic: the registration bench loop with a sampling thread that records where the main thread is every 2 ms;
prints the stack histogram of any step that takes more than 1.2x the median."""
import collections
import gc
import os
import sys
import threading
import time
import traceback
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from lcrnet_b200 import checkpoint, lcrnet, pipeline, synth
from lcrnet_b200 import data as gdata

n_pairs = 32
scans = []
for i in range(n_pairs):
    ref, src, _ = synth.make_pair(i, 7351 + i)
    scans += [ref, src]
limits = [57, 58, 59, 54]
net = lcrnet.create_model(lcrnet.default_cfg(limits)).eval()
net.load_state_dict(checkpoint.random_state_dict('lcrnet', 7351), strict=True)
net = net.cuda()
pipe = pipeline.PairPipeline(net, limits, 4, 0.3, 1.275, pre_voxel=0.3, n_streams=1)
host = torch.from_numpy(np.concatenate(scans, 0)).pin_memory()
dev = host.cuda()
lens = [len(s) for s in scans]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
main_id = threading.get_ident()
samples, stop = [], False


def sampler():
    while not stop:
        fr = sys._current_frames().get(main_id)
        if fr is not None:
            st = traceback.extract_stack(fr, limit=6)
            samples.append((time.perf_counter(), ' < '.join('%s:%d' % (os.path.basename(f.filename), f.lineno) for f in reversed(st))))
        time.sleep(0.002)


gc.disable()
for _ in range(3):
    pipe(dev, lens)
    pipe(host, lens)
torch.cuda.synchronize()
threading.Thread(target=sampler, daemon=True).start()
recs = []
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    gc.collect()
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    outs = pipe(dev, lens)
    b.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    recs.append((a.elapsed_time(b), t0, t1))
stop = True
med = float(np.median([r[0] for r in recs]))
print('step ms:', [round(r[0], 1) for r in recs], 'host wall ms:', [round((r[2] - r[1]) * 1e3, 1) for r in recs])
for k, (ms, t0, t1) in enumerate(recs):
    if ms > 1.2 * med:
        hist = collections.Counter(s for t, s in samples if t0 <= t <= t1)
        print('--- slow step %d (%.1f ms): main-thread samples' % (k, ms))
        for s, n in hist.most_common(12):
            print('   %4d  %s' % (n, s))
