"""Call-surface conformance: the reference's wrappers and scripts must be able to swap
this package in.

CPU (here, with /root/reference): after ``ext.install()`` the reference's own ``ops/grid_subsample.py`` /
``ops/radius_search.py`` and its collate resolve to this package's operators (run in a subprocess so the import
caches of other tests do not interfere; without a GPU the routed call must raise THIS package's "no CPU fallback"
error).  GPU: output dicts of the three models match the reference's key sets, python types, dtypes and shapes
(tests/golden/surface_golden.json, recorded from the unmodified reference models), the reference-named collates
and ``calibrate_neighbors_stack_mode`` reproduce the reference's limits, key sets and tables
(tests/golden/calibrate_golden.npz)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from util import GOLDEN, REF_PRESENT, canonical_rows

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S = json.load(open(os.path.join(GOLDEN, 'surface_golden.json')))
sys.path.insert(0, GOLDEN)

_WIRING = r'''
import sys, types
sys.path.insert(0, %(repo)r); sys.path.insert(0, %(golden)r)
import numpy as np, torch
import ref_import
ref_import.install()                                   # stubs for the pip modules the reference imports
from lcrnet_b200 import ext
ext.install()                                          # ... then this package becomes utils.ext
from experiments.lcrnet.modules.ops import grid_subsample as gs_mod, radius_search as rs_mod
gs_file = sys.modules['experiments.lcrnet.modules.ops.grid_subsample']
rs_file = sys.modules['experiments.lcrnet.modules.ops.radius_search']
assert gs_file.ext_module is ext and rs_file.ext_module is ext, 'reference wrappers did not bind to lcrnet_b200.ext'
pts = torch.rand(64, 3); lens = torch.tensor([64])
from experiments.lcrnet import data as rdata
calls = []
for fn in (lambda: gs_mod(pts, lens, voxel_size=0.3), lambda: rs_mod(pts, pts, lens, lens, 0.5, 8),
           lambda: rdata.precompute_data_stack_mode(pts, lens, 2, 0.3, 0.6, [8, 8])):
    try:
        out = fn()
        calls.append('ran')
    except RuntimeError as e:
        calls.append('no-cpu-fallback' if 'no CPU fallback' in str(e) else 'other: %%s' %% e)
print('RESULT', torch.cuda.is_available(), calls)
'''


@pytest.mark.skipif(not REF_PRESENT, reason='reference tree not present')
def test_reference_wrappers_resolve_to_this_package():
    code = _WIRING % {'repo': REPO, 'golden': GOLDEN}
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('RESULT')][-1]
    has_gpu = 'True' in line.split()[1]
    expect = 'ran' if has_gpu else 'no-cpu-fallback'
    assert line.count(expect) == 3, line


def test_matching_model_state_dict_layout():
    from lcrnet_b200 import lcrnet
    m = lcrnet.create_matching_model(lcrnet.default_cfg([30] * 4))
    assert sorted(m.state_dict().keys()) == S['matching_state_dict_keys']
    full = lcrnet.create_model(lcrnet.default_cfg([30] * 4))
    extra = sorted(set(full.state_dict()) - set(m.state_dict()))
    assert extra == S['matching_unexpected_from_lcrnet'] and all(k.startswith('netvlad.') for k in extra)


def _check(got, want, key, exact_shape=True):
    assert want['type'] == ('Tensor' if torch.is_tensor(got) else type(got).__name__), key
    if want['type'] == 'Tensor':
        assert str(got.dtype).replace('torch.', '') == want['dtype'], (key, got.dtype, want['dtype'])
        if exact_shape:
            assert list(got.shape) == want['shape'], (key, list(got.shape), want['shape'])
        else:
            assert got.dim() == len(want['shape']) and list(got.shape[1:]) == want['shape'][1:], key
    elif want['type'] in ('tuple', 'list'):
        assert len(got) == len(want['items']), key
        for g, w in zip(got, want['items']):
            _check(g, w, key, exact_shape)


@pytest.fixture(scope='module')
def pair_dict():
    from lcrnet_b200 import data as gdata
    from make_pair_golden import CASE, LIMITS, make_pair_data
    raw_ref, raw_src, _ = make_pair_data(*CASE)
    d = gdata.scans_collate_fn_stack_mode([raw_ref, raw_src], 4, 0.3, 1.275, LIMITS, pre_voxel=0.3, stack_size=2,
                                          int32=False, upsampling=True)
    return d, LIMITS


@pytest.mark.gpu
@pytest.mark.parametrize('kind', ['lcrnet', 'matching'])
def test_output_dict_contract(pair_dict, kind):
    """Key set == the reference's (private ``_``-prefixed debugging keys aside); every value has the reference's
    python type, dtype and shape (the discrete stages of this pair are identical to the reference's, pair_golden)."""
    from lcrnet_b200 import checkpoint, lcrnet
    d, limits = pair_dict
    make = lcrnet.create_model if kind == 'lcrnet' else lcrnet.create_matching_model
    net = make(lcrnet.default_cfg(limits)).eval()
    sd = checkpoint.random_state_dict('lcrnet', 7351)
    net.load_state_dict({k: v for k, v in sd.items() if k in net.state_dict()}, strict=True)
    out = net.cuda()(d)
    want = S[kind]
    public = {k for k in out if not k.startswith('_')}
    assert public == set(want), (sorted(public - set(want)), sorted(set(want) - public))
    data_dependent = {'pos_corr_points', 'anc_corr_points', 'corr_scores'}       # a near-tie may move the count by one
    for k in sorted(want):
        _check(out[k], want[k], k, exact_shape=k not in data_dependent)
    assert abs(out['corr_scores'].shape[0] - want['corr_scores']['shape'][0]) <= 2


@pytest.mark.gpu
def test_global_descriptor_contract(pair_dict):
    from lcrnet_b200 import checkpoint, model
    from lcrnet_b200 import data as gdata
    from make_pair_golden import CASE, LIMITS, make_pair_data
    raw_ref, _, _ = make_pair_data(*CASE)
    d = gdata.scans_collate_fn_stack_mode([raw_ref], 4, 0.3, 1.275, LIMITS, pre_voxel=0.3, int32=False)
    d.pop('stack_size')                                   # reference semantics: the whole input is one stack
    net = model.create_model(model.default_cfg()).eval()
    net.load_state_dict(checkpoint.random_state_dict('global_descriptor', 7351), strict=True)
    out = net.cuda()(d)
    assert set(out) == set(S['global_descriptor'])
    _check(out['anc_global'], S['global_descriptor']['anc_global'], 'anc_global')


@pytest.mark.gpu
def test_reference_named_collates_and_calibration():
    """data.registration_collate_fn_stack_mode / test_loop_detection_collate_fn_stack_mode_online /
    calibrate_neighbors_stack_mode with the reference's argument lists against the reference's own results."""
    from lcrnet_b200 import data as gdata
    from make_calibrate_golden import make_samples
    from oracle import native
    G = np.load(os.path.join(GOLDEN, 'calibrate_golden.npz'))
    reg, ld = make_samples()
    lim = gdata.calibrate_neighbors_stack_mode(reg, gdata.registration_collate_fn_stack_mode, 4, 0.3, 4.25 * 0.3)
    assert lim == list(G['limits_registration'])
    lim_ld = gdata.calibrate_neighbors_stack_mode(ld, gdata.test_loop_detection_collate_fn_stack_mode_online, 4, 0.3,
                                                  4.25 * 0.3)
    assert lim_ld == list(G['limits_loop_detection'])
    lim_q = gdata.calibrate_neighbors_stack_mode(reg, gdata.registration_collate_fn_stack_mode, 4, 0.3, 4.25 * 0.3,
                                                 keep_ratio=0.6, sample_threshold=500)
    assert lim_q == list(G['limits_registration_q60_t500'])
    limits = [int(x) for x in G['limits_registration']]
    d1 = gdata.registration_collate_fn_stack_mode(reg[:1], 4, 0.3, 4.25 * 0.3, limits)
    d2 = gdata.registration_collate_fn_stack_mode(reg[:2], 4, 0.3, 4.25 * 0.3, limits)
    dl = gdata.test_loop_detection_collate_fn_stack_mode_online(ld[:1], 4, 0.3, 4.25 * 0.3, limits)
    dr = gdata.registration_collate_fn_stack_mode(reg[:1], 4, 0.3, 4.25 * 0.3, limits, precompute_data=False)
    own = {'lengths_host'}                                # this package's extra key (saves the model a device sync)
    assert sorted(set(d1) - own) == list(G['keys_registration'])
    assert sorted(set(d2) - own) == list(G['keys_registration_b2'])
    assert sorted(set(dl) - own) == list(G['keys_loop_detection'])
    assert sorted(dr) == list(G['keys_raw'])
    for d, names, types in ((d1, G['keys_registration'], G['b1_types']), (d2, G['keys_registration_b2'], G['b2_types'])):
        for k, t in zip(names, types):
            assert type(d[k]).__name__ == t, (k, type(d[k]).__name__, t)
    st = lambda d: torch.stack([l.cpu() for l in d['lengths']]).numpy()
    assert np.array_equal(st(d1), G['lengths_registration']) and np.array_equal(st(d2), G['lengths_registration_b2'])
    assert np.array_equal(st(dl), G['lengths_loop_detection'])
    assert list(d2['features'].shape) == list(G['features_shape_registration_b2'])
    assert list(dl['features'].shape) == list(G['features_shape_loop_detection'])
    assert d1['neighbors'][0].dtype == torch.int64                       # the reference's table dtype at this surface
    P3 = d1['points'][3].cpu().numpy()
    assert np.array_equal(P3, G['points3_registration'])                 # bit-exact, reference order
    ref = G['neighbors3_registration'].astype(np.int64)
    assert np.array_equal(canonical_rows(ref, native.neighbor_d2(P3, P3, ref))[0], d1['neighbors'][3].cpu().numpy())
    Q2, Q3 = dl['points'][2].cpu().numpy(), dl['points'][3].cpu().numpy()
    for key, q, sp, tab in (('subsampling2_loop_detection', Q3, Q2, dl['subsampling'][2]),
                            ('upsampling2_loop_detection', Q2, Q3, dl['upsampling'][2])):
        ref = G[key].astype(np.int64)
        assert np.array_equal(canonical_rows(ref, native.neighbor_d2(q, sp, ref))[0], tab.cpu().numpy()), key
