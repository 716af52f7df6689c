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


def _res_spec(p, cin, cout):
    mid = cout // 4
    spec = []
    if cin != mid:
        spec += _unary_spec(p + 'unary1.', cin, mid)
    spec += _kpconv_spec(p + 'KPConv.', mid, mid)
    spec += [(p + 'norm_conv.norm.weight', (mid,)), (p + 'norm_conv.norm.bias', (mid,))]
    spec += _unary_spec(p + 'unary2.', mid, cout)
    if cin != cout:
        spec += _unary_spec(p + 'unary_shortcut.', cin, cout)
    return spec


def _linear_spec(p, cin, cout):
    return [(p + 'weight', (cout, cin)), (p + 'bias', (cout,))]


def _ln_spec(p, c):
    return [(p + 'weight', (c,)), (p + 'bias', (c,))]


def vote_encoder_spec(prefix='vote_encoder.', d=INIT_DIM):
    """backbone4.py:92-119 + modules/vote/vote.py:112-147 (input_feats_dim 256, MLP 512/256)."""
    v = prefix + 'vote.'
    spec = (_linear_spec(v + 'mlp_modules.0.', 256, 512) + _ln_spec(v + 'mlp_modules.1.', 512)
            + _linear_spec(v + 'mlp_modules.3.', 512, 256) + _ln_spec(v + 'mlp_modules.4.', 256)
            + _linear_spec(v + 'ctr_reg.', 256, 3))
    spec += _res_spec(prefix + 'encoder6_1.', 4 * d, 4 * d)
    spec += _res_spec(prefix + 'encoder6_2.', 4 * d, 8 * d)
    spec += _res_spec(prefix + 'encoder6_3.', 8 * d, 8 * d)
    return spec


def transformer_spec(prefix='transformer.', d_in=1024, d=128, d_out=256, layers=8):
    """thdroformer_linear.py:12-48, rpetransformer.py:57-220, vanilla_transformer.py:13-144."""
    spec = (_linear_spec(prefix + 'embedding.encoder.', 3, d) + _linear_spec(prefix + 'embedding.encoder2.', d, d // 2)
            + _linear_spec(prefix + 'in_proj.', d_in, d))
    for i in range(layers):
        p = prefix + 'transformer.layers.%d.' % i
        for n in ('proj_q.', 'proj_k.', 'proj_v.'):
            spec += _linear_spec(p + 'attention.attention.' + n, d, d)
        spec += _linear_spec(p + 'attention.linear.', d, d) + _ln_spec(p + 'attention.norm.', d)
        spec += (_linear_spec(p + 'output.expand.', d, 2 * d) + _linear_spec(p + 'output.squeeze.', 2 * d, d)
                 + _ln_spec(p + 'output.norm.', d))
    return spec + _linear_spec(prefix + 'out_proj.', d, d_out)


def kpdecoder_spec(prefix='kpdecoder.', d=INIT_DIM):
    return (_unary_spec(prefix + 'decoder3.', 12 * d, 8 * d) + _unary_spec(prefix + 'decoder2.', 12 * d, 4 * d)
            + _linear_spec(prefix + 'decoder1.mlp.', 6 * d, 2 * d))


def state_dict_spec(kind='global_descriptor'):
    if kind == 'global_descriptor':
        return encoder_spec() + netvlad_spec()
    if kind == 'lcrnet':  # model_family/LCRNet.py:25-112 (373 tensors, 25 093 190 parameters)
        return (encoder_spec() + vote_encoder_spec() + _linear_spec('proj_node_overlap_score.', 512, 1)
                + transformer_spec() + kpdecoder_spec()
                + [('node_optimal_transport.alpha', ()), ('optimal_transport.alpha', ())] + netvlad_spec())
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
    for name, stage in (('encoder6_1', 3), ('encoder6_2', 4), ('encoder6_3', 4)):
        radius['vote_encoder.' + name + '.KPConv.kernel_points'] = INIT_RADIUS * 2 ** stage
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
        elif leaf == 'alpha':
            t = torch.tensor(1.0 + 0.25 * rng.standard_normal(), dtype=torch.float32)
        elif name.startswith('transformer.embedding.'):
            # angles theta = A p + b must stay O(1) over a +-80 m scene
            scale = {'encoder.weight': 0.05, 'encoder.bias': 0.1, 'encoder2.weight': 0.06, 'encoder2.bias': 0.1}
            t = torch.from_numpy((scale[name.split('embedding.')[1]] * rng.standard_normal(shape)).astype(np.float32))
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
