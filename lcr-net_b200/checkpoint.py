"""Checkpoint layout of the reference model family and deterministic random weights.

The reference saves ``torch.save({'epoch', 'iteration', 'model': state_dict})``
(utils/engine/base_trainer.py:111-127) and loads with ``strict=False`` after stripping a DDP
``module.`` prefix (base_tester.py:111-122).  ``state_dict_spec`` lists the exact tensor names and
shapes of ``LCRNet_GlobalDescrition`` (170 tensors) as produced by the reference classes
(backbone4.py:11-58, kpconv/modules.py, netvlad/NetVlad.py:12-47,165-187); pre-trained weights are
not distributed with the reference, so benchmarks and tests use ``random_state_dict``.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

INIT_DIM, KERNEL_SIZE, GROUPS = 64, 15, 32
INIT_RADIUS, INIT_SIGMA = 4.25 * 0.3, 2.0 * 0.3


def encoder_blocks(init_dim=INIT_DIM):
    """(name, kind, c_in, c_out, stage, strided) for the 11 encoder blocks (backbone4.py:15-58).
    stage s uses radius INIT_RADIUS * 2**s and sigma INIT_SIGMA * 2**s."""
    d = init_dim
    return [
        ('encoder1_1', 'conv', 1, d, 0, False), ('encoder1_2', 'res', d, 2 * d, 0, False),
        ('encoder2_1', 'res', 2 * d, 2 * d, 0, True), ('encoder2_2', 'res', 2 * d, 4 * d, 1, False),
        ('encoder2_3', 'res', 4 * d, 4 * d, 1, False),
        ('encoder3_1', 'res', 4 * d, 4 * d, 1, True), ('encoder3_2', 'res', 4 * d, 8 * d, 2, False),
        ('encoder3_3', 'res', 8 * d, 8 * d, 2, False),
        ('encoder4_1', 'res', 8 * d, 8 * d, 2, True), ('encoder4_2', 'res', 8 * d, 16 * d, 3, False),
        ('encoder4_3', 'res', 16 * d, 16 * d, 3, False),
    ]


def _unary_spec(prefix, cin, cout):
    return [(prefix + 'mlp.weight', (cout, cin)), (prefix + 'mlp.bias', (cout,)),
            (prefix + 'norm.norm.weight', (cout,)), (prefix + 'norm.norm.bias', (cout,))]


def _kpconv_spec(prefix, cin, cout):
    return [(prefix + 'weights', (KERNEL_SIZE, cin, cout)), (prefix + 'bias', (cout,)),
            (prefix + 'kernel_points', (KERNEL_SIZE, 3))]


def encoder_spec(prefix='encoder.'):
    spec = []
    for name, kind, cin, cout, _, _ in encoder_blocks():
        p = prefix + name + '.'
        if kind == 'conv':
            spec += _kpconv_spec(p + 'KPConv.', cin, cout)
            spec += [(p + 'norm.norm.weight', (cout,)), (p + 'norm.norm.bias', (cout,))]
        else:
            mid = cout // 4
            if cin != mid:
                spec += _unary_spec(p + 'unary1.', cin, mid)
            spec += _kpconv_spec(p + 'KPConv.', mid, mid)
            spec += [(p + 'norm_conv.norm.weight', (mid,)), (p + 'norm_conv.norm.bias', (mid,))]
            spec += _unary_spec(p + 'unary2.', mid, cout)
            if cin != cout:
                spec += _unary_spec(p + 'unary_shortcut.', cin, cout)
    return spec


def _bn_spec(prefix, c):
    return [(prefix + 'weight', (c,)), (prefix + 'bias', (c,)), (prefix + 'running_mean', (c,)),
            (prefix + 'running_var', (c,)), (prefix + 'num_batches_tracked', ())]


def netvlad_spec(prefix='netvlad.', feature=1024, clusters=64, out=256):
    return ([(prefix + 'cluster_weights', (feature, clusters)), (prefix + 'cluster_weights2', (1, feature, clusters)),
             (prefix + 'hidden1_weights', (clusters * feature, out))]
            + _bn_spec(prefix + 'bn1.', clusters) + _bn_spec(prefix + 'bn2.', out)
            + [(prefix + 'context_gating.gating_weights', (out, out))] + _bn_spec(prefix + 'context_gating.bn1.', out))


def state_dict_spec(kind='global_descriptor'):
    if kind == 'global_descriptor':
        return encoder_spec() + netvlad_spec()
    raise ValueError(kind)


def default_kernel_points(radius, rng):
    """A centre point plus 14 points spread on a sphere (stand-in for the reference's optimised
    disposition file, which real checkpoints carry as the ``kernel_points`` buffer, kpconv.py:65)."""
    n = KERNEL_SIZE - 1
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + 5 ** 0.5) * i
    pts = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], 1) * 0.66
    pts = np.concatenate([np.zeros((1, 3)), pts], 0) + rng.normal(scale=0.01, size=(KERNEL_SIZE, 3))
    return (pts * radius).astype(np.float32)


def random_state_dict(kind='global_descriptor', seed=7351):
    """Deterministic (numpy PCG64) weights with the reference's names/shapes/dtypes."""
    rng = np.random.default_rng(seed)
    radius = {}
    for name, _, _, _, stage, _ in encoder_blocks():
        radius['encoder.' + name + '.KPConv.kernel_points'] = INIT_RADIUS * 2 ** stage
    sd = OrderedDict()
    for name, shape in state_dict_spec(kind):
        leaf = name.rsplit('.', 1)[-1]
        if leaf == 'num_batches_tracked':
            t = torch.tensor(1000, dtype=torch.int64)
        elif leaf == 'kernel_points':
            t = torch.from_numpy(default_kernel_points(radius[name], rng))
        elif leaf == 'running_var':
            t = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif leaf == 'running_mean':
            t = torch.from_numpy((0.1 * rng.standard_normal(shape)).astype(np.float32))
        elif name.endswith('norm.weight') or (leaf == 'weight' and len(shape) == 1):
            t = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32))
        elif leaf == 'bias':
            t = torch.from_numpy((0.05 * rng.standard_normal(shape)).astype(np.float32))
        elif leaf == 'weights':  # KPConv [K, Cin, Cout]
            bound = math.sqrt(3.0 / (shape[0] * shape[1])) * 2.0
            t = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
        elif leaf in ('cluster_weights', 'cluster_weights2', 'hidden1_weights', 'gating_weights'):
            fan = 1024 if leaf != 'gating_weights' else 256
            t = torch.from_numpy((rng.standard_normal(shape) / math.sqrt(fan)).astype(np.float32))
        else:  # Linear weight [Cout, Cin]
            bound = math.sqrt(3.0 / shape[1]) * 1.5
            t = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
        sd[name] = t
    return sd
