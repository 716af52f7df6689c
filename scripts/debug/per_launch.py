"""Per-launch profile scopes of one descriptor step (64 scans, single stream)."""
import sys, os, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from lcrnet_b200 import _lib, checkpoint, model, pipeline
class A: pairs = 32
ctx = bench.Ctx()
inp = bench.Inputs(ctx, 32)
net = model.create_model(model.default_cfg()).eval()
net.load_state_dict(checkpoint.random_state_dict('global_descriptor', 7351), strict=True)
net = net.cuda()
pipe = pipeline.DescriptorPipeline(net, inp.limits, n_streams=1)
for _ in range(2):
    pipe(inp.dev_pts, inp.lens)
torch.cuda.synchronize()
L = _lib.lib()
L.lcr_profile_begin()
pipe(inp.dev_pts, inp.lens)
torch.cuda.synchronize()
n = L.lcr_profile_end()
for i in range(n):
    name = ctypes.create_string_buffer(64)
    ms, fl, by = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    L.lcr_profile_get(i, name, 64, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by))
    nm = name.value.decode()
    if nm in ('kpconv_gather', 'gemm_tf32x3', 'radius_query', 'maxpool', 'group_norm_apply', 'kpconv_c1', 'grid_subsample'):
        print('%-18s %8.3f ms  %8.2f GFLOP  %8.1f MB  -> %6.1f TFLOP/s %6.0f GB/s' % (
            nm, ms.value, fl.value / 1e9, by.value / 1e6, fl.value / max(ms.value, 1e-9) / 1e9, by.value / max(ms.value, 1e-9) / 1e6))
