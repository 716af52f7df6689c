"""Model family on the B200 kernels, with the reference's class names, constructor arguments,
``forward(data_dict)`` contract and ``state_dict`` layout:

  KPConv / GroupNorm / UnaryBlock / ConvBlock / ResidualBlock   modules/kpconv/{kpconv,modules}.py
  KPEncoder                                                     backbone4.py:11-89
  NetVLADLoupe2 / GatingContext                                 modules/netvlad/NetVlad.py
  LCRNet_GlobalDescrition, create_model                         model_family/LCRNet_GlobalDescrition.py

The modules only hold parameters (so ``load_state_dict`` of a reference checkpoint works
unchanged, base_tester.py:111-122); all arithmetic is in liblcr_b200.so.  Inference only
(``eval()`` semantics: BatchNorm uses running statistics).

Extension over the reference: a data_dict may carry ``stack_size`` (clouds per stack, see
data.py); the model then evaluates many independent stacks in one pass -- GroupNorm statistics
per stack and one descriptor per cloud -- instead of the reference's one stack per forward.
"""
import torch
import torch.nn as nn

from . import ops
from .checkpoint import encoder_blocks


class KPConv(nn.Module):
    """kpconv.py:12-122 (parameters ``weights`` [K, Cin, Cout], ``bias``; buffer ``kernel_points``)."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, bias=False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.radius, self.sigma = radius, sigma
        self.weights = nn.Parameter(torch.zeros(kernel_size, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter('bias', None)
        self.register_buffer('kernel_points', torch.zeros(kernel_size, 3))
        self._w_nk = None
        self._kp_host = None

    def kernel_points_host(self):
        """CPU copy of the kernel points (kernel arguments of the fast gather; cached; the device->host copy
        is synchronous, so the value is complete when this returns)."""
        kp = self.kernel_points
        key = (kp.data_ptr(), kp._version)
        if self._kp_host is None or self._kp_host[0] != key:
            self._kp_host = (key, kp.detach().to('cpu', torch.float32).contiguous())
        return self._kp_host[1]

    def weights_nk(self):
        """[c_out, 15 * c_in] copy of the weights for the tensor-core contraction (cached, stream-safe)."""
        w = self.weights
        return ops.derived(self, '_w_nk', (w.data_ptr(), w._version),
                           lambda: w.detach().reshape(-1, w.shape[2]).t().contiguous())

    def prepare_b200(self):
        self.kernel_points_host()
        if self.in_channels > 1 and self.weights.is_cuda and ops.use_tensor_cores():
            ops.tf32_split(self.weights_nk())

    def forward(self, s_feats, q_points, s_points, neighbor_indices, s_flags=None, gn=None):
        """``gn`` = (stacks, eps, groups) of the GroupNorm that follows: returns (out, stats), the statistics
        coming out of the GEMM epilogue when the shapes allow (ops.gn_fusable)."""
        return ops.kpconv(s_feats, q_points, s_points, neighbor_indices, self.kernel_points, self.sigma,
                          self.weights, self.bias, s_flags, self.weights_nk() if self.in_channels > 1 else None,
                          self.kernel_points_host(), gn=gn)


class GroupNorm(nn.Module):
    """modules.py:33-50: holds ``norm.weight`` / ``norm.bias``; statistics over all rows of a stack."""

    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups, self.num_channels = num_groups, num_channels
        self.norm = nn.GroupNorm(num_groups, num_channels)

    def stats(self, x, stacks):
        return ops.group_norm_stats(x, stacks, self.norm.eps, self.num_groups)

    def spec(self, stacks):
        """(stacks, eps, groups): what a fused producer needs to emit this norm's statistics."""
        return (stacks, self.norm.eps, self.num_groups)

    def forward(self, x, stacks, leaky=False, want_flags=False, stats=None):
        return ops.group_norm_apply(x, self.stats(x, stacks) if stats is None else stats, self.norm.weight,
                                    self.norm.bias, stacks, leaky=leaky, want_flags=want_flags, groups=self.num_groups)


class UnaryBlock(nn.Module):
    """modules.py:53-83: Linear -> GroupNorm -> LeakyReLU(0.1)."""

    def __init__(self, in_channels, out_channels, group_norm, has_relu=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.has_relu = in_channels, out_channels, has_relu
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)
        self.norm = GroupNorm(group_norm, out_channels)
        self._wt = None

    def weight_t(self):
        w = self.mlp.weight
        return ops.derived(self, '_wt', (w.data_ptr(), w._version), lambda: w.detach().t().contiguous())

    def prepare_b200(self):
        self.weight_t()
        w = self.mlp.weight
        if w.is_cuda and w.shape[1] % 32 == 0 and w.shape[0] % 4 == 0 and ops.use_tensor_cores():
            ops.tf32_split(w)

    def linear(self, x):
        return ops.linear(x, self.weight_t(), self.mlp.bias, self.mlp.weight)

    def linear_stats(self, x, stacks):
        """Linear output and the statistics of the GroupNorm that follows it (fused when possible)."""
        return ops.linear(x, self.weight_t(), self.mlp.bias, self.mlp.weight, gn=self.norm.spec(stacks))

    def forward(self, x, stacks, want_flags=False):
        y, stats = self.linear_stats(x, stacks)
        return self.norm(y, stacks, leaky=self.has_relu, want_flags=want_flags, stats=stats)


class ConvBlock(nn.Module):
    """modules.py:104-146: KPConv -> GroupNorm -> LeakyReLU."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, group_norm, bias=True):
        super().__init__()
        self.KPConv = KPConv(in_channels, out_channels, kernel_size, radius, sigma, bias=bias)
        self.norm = GroupNorm(group_norm, out_channels)

    def forward(self, s_feats, q_points, s_points, neighbor_indices, stacks):
        if self.KPConv.in_channels > 1:
            x, stats = self.KPConv(s_feats, q_points, s_points, neighbor_indices, gn=self.norm.spec(stacks))
            return self.norm(x, stacks, leaky=True, stats=stats)
        x = self.KPConv(s_feats, q_points, s_points, neighbor_indices)
        return self.norm(x, stacks, leaky=True)


class ResidualBlock(nn.Module):
    """modules.py:149-225 (bottleneck: unary1 -> KPConv -> GN/LeakyReLU -> unary2, + shortcut)."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, group_norm, strided=False, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.strided = in_channels, out_channels, strided
        mid = out_channels // 4
        self.unary1 = UnaryBlock(in_channels, mid, group_norm, bias=bias) if in_channels != mid else nn.Identity()
        self.KPConv = KPConv(mid, mid, kernel_size, radius, sigma, bias=bias)
        self.norm_conv = GroupNorm(group_norm, mid)
        self.unary2 = UnaryBlock(mid, out_channels, group_norm, has_relu=False, bias=bias)
        if in_channels != out_channels:
            self.unary_shortcut = UnaryBlock(in_channels, out_channels, group_norm, has_relu=False, bias=bias)
        else:
            self.unary_shortcut = nn.Identity()

    def forward(self, s_feats, q_points, s_points, neighbor_indices, q_stacks, s_stacks):
        if isinstance(self.unary1, nn.Identity):
            x, flags = s_feats, None
        else:
            x, flags = self.unary1(s_feats, s_stacks, want_flags=True)
        x, cstats = self.KPConv(x, q_points, s_points, neighbor_indices, flags, gn=self.norm_conv.spec(q_stacks))
        x = self.norm_conv(x, q_stacks, leaky=True, stats=cstats)
        x, stats = self.unary2.linear_stats(x, q_stacks)
        shortcut = ops.maxpool(s_feats, neighbor_indices) if self.strided else s_feats
        n2 = self.unary2.norm.norm
        if isinstance(self.unary_shortcut, nn.Identity):
            return ops.group_norm_apply(x, stats, n2.weight, n2.bias, q_stacks, leaky=True, other=shortcut)
        sc, sstats = self.unary_shortcut.linear_stats(shortcut, q_stacks)
        ns = self.unary_shortcut.norm
        return ops.group_norm_apply(x, stats, n2.weight, n2.bias, q_stacks, leaky=True, other=sc,
                                    other_norm=(sstats, ns.norm.weight, ns.norm.bias))


def make_stacks(data_dict, device):
    """Per-level ops.Stacks from ``lengths`` and the optional ``stack_size`` key."""
    lh = data_dict.get('lengths_host')
    if lh is None:
        lh = torch.stack([l.reshape(-1) for l in data_dict['lengths']]).cpu().tolist()
    size = data_dict.get('stack_size')
    out = []
    for lens in lh:
        if size is None:
            groups = [sum(lens)]                     # reference semantics: the whole batch is one stack
        else:
            groups = [sum(lens[i:i + size]) for i in range(0, len(lens), size)]
        out.append(ops.Stacks(groups, device))
    return out, lh


class KPEncoder(nn.Module):
    """backbone4.py:11-89."""

    def __init__(self, input_dim, init_dim, kernel_size, init_radius, init_sigma, group_norm):
        super().__init__()
        self.block_names = []
        for name, kind, cin, cout, stage, strided in encoder_blocks(init_dim):
            cin = input_dim if name == 'encoder1_1' else cin
            r, s = init_radius * 2 ** stage, init_sigma * 2 ** stage
            if kind == 'conv':
                blk = ConvBlock(cin, cout, kernel_size, r, s, group_norm)
            else:
                blk = ResidualBlock(cin, cout, kernel_size, r, s, group_norm, strided=strided)
            setattr(self, name, blk)
            self.block_names.append(name)

    def forward(self, feats, data_dict, stacks=None, return_blocks=False):
        p, nb, sub = data_dict['points'], data_dict['neighbors'], data_dict['subsampling']
        if stacks is None:
            stacks, _ = make_stacks(data_dict, feats.device)
        blocks = {}
        x = self.encoder1_1(feats, p[0], p[0], nb[0], stacks[0])
        blocks['encoder1_1'] = x
        x = self.encoder1_2(x, p[0], p[0], nb[0], stacks[0], stacks[0])
        blocks['encoder1_2'] = x
        feats_list = [x]
        for s in (1, 2, 3):
            for j, (idx, sp, ss) in enumerate(((sub[s - 1], p[s - 1], stacks[s - 1]), (nb[s], p[s], stacks[s]),
                                               (nb[s], p[s], stacks[s]))):
                name = 'encoder%d_%d' % (s + 1, j + 1)
                x = getattr(self, name)(x, p[s], sp, idx, stacks[s], ss)
                blocks[name] = x
            feats_list.append(x)
        return (feats_list, blocks) if return_blocks else feats_list


class GatingContext(nn.Module):
    """NetVlad.py:165-201 (parameter holder)."""

    def __init__(self, dim, add_batch_norm=True):
        super().__init__()
        self.dim = dim
        self.gating_weights = nn.Parameter(torch.zeros(dim, dim))
        self.bn1 = nn.BatchNorm1d(dim)


class NetVLADLoupe2(nn.Module):
    """NetVlad.py:12-87 (eval mode, gating=True, add_norm=True, batch normalisation)."""

    def __init__(self, feature_size, cluster_size, output_dim, gating=True, add_norm=True, is_training=False):
        super().__init__()
        assert feature_size == 1024 and cluster_size == 64 and output_dim == 256 and gating and add_norm, \
            'the B200 NetVLAD kernels are specialised to the LCR-Net head (1024 x 64 -> 256, gated)'
        self.feature_size, self.cluster_size, self.output_dim = feature_size, cluster_size, output_dim
        self.cluster_weights = nn.Parameter(torch.zeros(feature_size, cluster_size))
        self.cluster_weights2 = nn.Parameter(torch.zeros(1, feature_size, cluster_size))
        self.hidden1_weights = nn.Parameter(torch.zeros(cluster_size * feature_size, output_dim))
        self.bn1 = nn.BatchNorm1d(cluster_size)
        self.bn2 = nn.BatchNorm1d(output_dim)
        self.context_gating = GatingContext(output_dim)
        self._packed = None

    @staticmethod
    def _pack(bn):
        assert abs(bn.eps - 1e-5) < 1e-12
        return torch.cat([bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var]).contiguous()

    def _bn_params(self):
        bns = (self.bn1, self.bn2, self.context_gating.bn1)
        key = tuple((b.weight.data_ptr(), b.weight._version, b.running_mean._version) for b in bns)
        return ops.derived(self, '_packed', key, lambda: tuple(self._pack(b) for b in bns))

    def prepare_b200(self):
        self._bn_params()

    def forward(self, feats, scan_off, n_scans):
        """feats [rows, 1024] (un-normalised encoder output), scan_off int64 [n_scans+1] on device.
        Includes the F.normalize before and after the head (LCRNet_GlobalDescrition.py:36-38)."""
        assert not self.training, 'inference only (BatchNorm running statistics)'
        bn1, bn2, bng = self._bn_params()
        return ops.netvlad(feats, scan_off, n_scans, self.cluster_weights, self.cluster_weights2,
                           self.hidden1_weights, bn1, bn2, self.context_gating.gating_weights, bng)


class LCRNet_GlobalDescrition(nn.Module):
    """model_family/LCRNet_GlobalDescrition.py:10-108 (eval branch :60-74)."""

    def __init__(self, cfg):
        super().__init__()
        b = cfg.backbone
        self.encoder = KPEncoder(b.input_dim, b.init_dim, b.kernel_size, b.init_radius, b.init_sigma, b.group_norm)
        self.netvlad = NetVLADLoupe2(feature_size=1024, cluster_size=64, output_dim=256, gating=True, add_norm=True,
                                     is_training=False)

    @torch.no_grad()
    def forward(self, data_dict):
        assert not self.training, 'lcrnet_b200 implements inference only: call model.eval()'
        feats = data_dict['features'].detach()
        stacks, lengths_host = make_stacks(data_dict, feats.device)
        feats_c = self.encoder(feats, data_dict, stacks)[-1]
        if data_dict.get('stack_size') is None:
            scan = ops.Stacks([feats_c.shape[0]], feats.device)       # one descriptor for the whole input
        else:
            scan = ops.Stacks(lengths_host[-1], feats.device)          # one descriptor per cloud
        return {'anc_global': self.netvlad(feats_c, scan.off, scan.n)}


def create_model(cfg):
    return LCRNet_GlobalDescrition(cfg)


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def default_cfg():
    """The values of experiments/lcrnet/config_model.py:33-43 that the descriptor path reads."""
    return _Cfg(backbone=_Cfg(num_stages=4, init_voxel_size=0.3, kernel_size=15, base_radius=4.25, base_sigma=2.0,
                              init_radius=4.25 * 0.3, init_sigma=2.0 * 0.3, group_norm=32, input_dim=1, init_dim=64,
                              output_dim=256))
