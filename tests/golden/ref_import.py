"""Import the UNMODIFIED reference Python (``/root/reference``) on torch-CPU.

Used ONLY by tests/golden/make_golden.py (fixture generation, in the build container) and by
the optional in-container cross-check tests.  Never imported by the product and never on the
GPU box (``/root/reference`` does not exist there).

What the shim does (none of it edits the reference):
* ``utils.ext`` := a module whose ``grid_subsampling`` / ``radius_neighbors`` call the reference
  C++ (``oracle/_ref/libref_ext.so`` = utils/extensions/cpu/** compiled where they lie).
* stubs for pip modules that are absent here (easydict, IPython, ipdb, open3d, ...).
* ``np.int = int`` (rpetransformer.py:48 uses the alias numpy removed).
* ``Tensor.cuda`` / ``Module.cuda`` become CPU no-ops (72 hard-wired ``.cuda()`` calls).
"""
import logging
import os
import struct
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get('LCR_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'experiments', 'lcrnet'))


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)
        super().__setattr__(k, v)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = dict(e or {}, **f)
        for k in d:
            setattr(self, k, d[k])


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _read_ply_xyz(path):
    with open(path, 'rb') as f:
        data = f.read()
    end = data.index(b'end_header\n') + len(b'end_header\n')
    n = int([ln for ln in data[:end].split(b'\n') if ln.startswith(b'element vertex')][0].split()[-1])
    vals = struct.unpack('<%dd' % (3 * n), data[end:end + 24 * n])
    return np.array(vals, dtype=np.float64).reshape(n, 3)


def _make_ext_module():
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from oracle import native as on

    def grid_subsampling(points, lengths, voxel_size):
        p, l = on.ref_grid_subsample(points.numpy(), lengths.numpy(), voxel_size)
        return [torch.from_numpy(p), torch.from_numpy(l)]

    def radius_neighbors(q, s, ql, sl, radius):
        return torch.from_numpy(on.ref_radius_neighbors(q.numpy(), s.numpy(), ql.numpy(), sl.numpy(), radius))

    def radius_filter(*a, **k):
        raise NotImplementedError

    return _stub('utils.ext', grid_subsampling=grid_subsampling, radius_neighbors=radius_neighbors,
                 radius_filter=radius_filter)


_installed = False


def install():
    """Idempotent.  After this, ``import experiments.lcrnet...`` works on torch-CPU."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    _installed = True
    sys.path.insert(0, REF_ROOT)
    np.int = int
    torch.Tensor.cuda = lambda self, *a, **k: self.contiguous()
    torch.nn.Module.cuda = lambda self, *a, **k: self

    _stub('easydict', EasyDict=_EasyDict)
    _stub('IPython', embed=lambda *a, **k: None)
    _stub('ipdb', set_trace=lambda *a, **k: None)
    _stub('coloredlogs', ColoredFormatter=logging.Formatter)
    mpl = _stub('matplotlib', use=lambda *a, **k: None)
    mpl.pyplot = _stub('matplotlib.pyplot')
    _stub('mpl_toolkits')
    _stub('mpl_toolkits.mplot3d', Axes3D=object)
    pml = _stub('pytorch_metric_learning')
    pml.distances = _stub('pytorch_metric_learning.distances', LpDistance=object, CosineSimilarity=object)
    pml.losses = _stub('pytorch_metric_learning.losses')
    pml.miners = _stub('pytorch_metric_learning.miners')
    pml.reducers = _stub('pytorch_metric_learning.reducers')
    _stub('tensorboardX', SummaryWriter=object)

    class _PC:
        def __init__(self):
            self.points = None

    o3d = _stub('open3d')
    o3d.io = _stub('open3d.io', read_point_cloud=lambda p: types.SimpleNamespace(points=_read_ply_xyz(p)),
                   write_point_cloud=lambda *a, **k: None)
    o3d.geometry = _stub('open3d.geometry', PointCloud=_PC, LineSet=object, TriangleMesh=object)
    o3d.utility = _stub('open3d.utility', Vector3dVector=lambda x: x, Vector2iVector=lambda x: x)
    o3d.visualization = _stub('open3d.visualization')
    o3d.pipelines = _stub('open3d.pipelines')
    o3d.pipelines.registration = _stub('open3d.pipelines.registration')

    import utils  # the reference's top-level package (namespace for utils.ext)
    utils.ext = _make_ext_module()


def model_cfg(neighbor_limits, tmp_root):
    """The easydict of experiments/lcrnet/config_model.py (config_reg.py / config_ld.py only add
    output dirs and optimiser settings and mkdir at import time, so they are not imported)."""
    install()
    from experiments.lcrnet import config_model
    config_model._C.output_root = tmp_root
    cfg = config_model.make_cfg()
    cfg.neighbor_limits = list(int(x) for x in neighbor_limits)
    cfg.vis = False
    return cfg
