"""CPU: the torch-CPU model oracle (oracle/model_oracle.py) against the fixture generated from
the UNMODIFIED reference Python model (tests/golden/model_golden.npz), and -- when
/root/reference is present -- against the reference run live."""
import os

import numpy as np
import pytest
import torch

from lcrnet_b200 import checkpoint, synth
from oracle import model_oracle as mo
from util import GOLDEN, REF_PRESENT

G = np.load(os.path.join(GOLDEN, 'model_golden.npz'))
LIMITS = [int(x) for x in G['limits']]
REL_TOL = 1e-4  # north_star: descriptors within 1e-4 relative


def _case(name):
    scene, seed, stride = (int(x) for x in G[name + '_case'])
    raw = np.ascontiguousarray(synth.make_scan(scene, seed)[::stride])
    from oracle import native
    p0, l0 = native.grid_subsample(raw, np.array([len(raw)], dtype=np.int64), 0.3)
    return mo.precompute_pyramid(p0, l0, limits=LIMITS)


@pytest.fixture(scope='module')
def sd():
    return checkpoint.random_state_dict('global_descriptor', int(G['weight_seed']))


def test_state_dict_layout():
    spec = checkpoint.state_dict_spec('global_descriptor')
    assert len(spec) == 170
    n_param = sum(int(np.prod(s)) for n, s in spec
                  if not n.endswith(('num_batches_tracked', 'kernel_points', 'running_mean', 'running_var')))
    assert n_param == 21999744  # SURVEY.md C.3 (probe of the reference model)


@pytest.mark.parametrize('name', ['s0', 's1'])
def test_oracle_matches_reference_fixture(name, sd):
    data = _case(name)
    assert [int(l[0]) for l in data['lengths']] == list(G[name + '_lengths'])
    assert [t.shape[1] for t in data['neighbors']] == list(G[name + '_widths'])
    feats = torch.ones(data['points'][0].shape[0], 1)
    with torch.no_grad():
        feats_list, blocks = mo.kpencoder(sd, feats, data, return_all=True)
        desc = mo.netvlad(sd, feats_list[-1])
    rows = int(G['rows'])
    for bn, t in blocks.items():
        head = G['%s_encoder%s_head' % (name, bn)]
        got = t[:rows].numpy()
        assert np.abs(got - head).max() <= REL_TOL * max(1.0, np.abs(head).max()), bn
        stats = G['%s_encoder%s_stats' % (name, bn)]
        assert abs(float((t.double() ** 2).sum()) - stats[2]) <= 1e-4 * stats[2], bn
    ref = G[name + '_descriptor']
    rel = np.linalg.norm(desc.numpy() - ref) / np.linalg.norm(ref)
    assert rel < REL_TOL


def test_l2_topk_oracle_is_brute_force():
    rng = np.random.default_rng(0)
    db = rng.standard_normal((300, 256)).astype(np.float32)
    q = db[200:220] + 0.01 * rng.standard_normal((20, 256)).astype(np.float32)
    d2, idx = mo.l2_topk(q, db, 5)
    assert (idx[:, 0] == np.arange(200, 220)).all()
    assert (np.diff(d2, axis=1) >= 0).all()
    d2c, idxc = mo.l2_topk(q, db, 5, valid_counts=np.arange(100, 120))
    assert (idxc < np.arange(100, 120)[:, None]).all()


@pytest.mark.skipif(not REF_PRESENT, reason='reference tree not present')
def test_oracle_matches_reference_live(sd):
    import sys
    sys.path.insert(0, GOLDEN)
    from make_model_golden import reference_forward
    raw = np.ascontiguousarray(synth.make_scan(5, 99)[::8])
    full_sd = checkpoint.random_state_dict('global_descriptor', 7351)
    data, blocks, desc = reference_forward(raw, full_sd, [30, 30, 30, 30])
    from oracle import native
    p0, l0 = native.grid_subsample(raw, np.array([len(raw)], dtype=np.int64), 0.3)
    mine = mo.precompute_pyramid(p0, l0, limits=[30, 30, 30, 30])
    for a, b in zip(mine['points'], data['points']):
        assert torch.equal(a, b)
    with torch.no_grad():
        got = mo.global_descriptor(sd, mine)
    rel = float(torch.linalg.norm(got - desc) / torch.linalg.norm(desc))
    assert rel < REL_TOL
