"""GPU parity of the model kernels (through the C ABI) against the torch-CPU oracle and the
reference fixture.  Floating point: tolerance 1e-4 relative (north_star), written per test."""
import os

import numpy as np
import pytest
import torch

from lcrnet_b200 import checkpoint, synth
from oracle import model_oracle as mo
from oracle import native as on
from util import GOLDEN

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, 'model_golden.npz'))
LIMITS = [int(x) for x in G['limits']]
REL = 1e-4


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def max_err(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope='module')
def sd():
    return checkpoint.random_state_dict('global_descriptor', int(G['weight_seed']))


@pytest.fixture(scope='module')
def net(sd):
    from lcrnet_b200 import model
    m = model.create_model(model.default_cfg()).eval()
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def _pyramid(scene, seed, stride, limits=LIMITS):
    raw = np.ascontiguousarray(synth.make_scan(scene, seed)[::stride])
    p0, l0 = on.grid_subsample(raw, np.array([len(raw)], dtype=np.int64), 0.3)
    return mo.precompute_pyramid(p0, l0, limits=limits)


def _to_cuda(data):
    out = {}
    for k, v in data.items():
        out[k] = [t.cuda() for t in v] if isinstance(v, list) else v
    return out


@pytest.mark.parametrize('c_in,c_out', [(1, 64), (32, 32), (64, 64), (128, 128), (256, 256)])
def test_kpconv_vs_oracle(c_in, c_out):
    from lcrnet_b200 import ops
    rng = np.random.default_rng(c_in)
    s = rng.uniform(-6, 6, (900, 3)).astype(np.float32)
    q = s[::3].copy()
    ql, sl = np.array([len(q)], dtype=np.int64), np.array([len(s)], dtype=np.int64)
    idx = on.radius_neighbors(q, s, ql, sl, 2.5, limit=37)
    feats = rng.standard_normal((len(s), c_in)).astype(np.float32)
    if c_in == 1:
        feats = np.abs(feats)
    feats[::7] = -np.abs(feats[::7])          # rows with a non-positive sum: exercise the neighbour_num quirk
    w = (rng.standard_normal((15, c_in, c_out)) * 0.1).astype(np.float32)
    b = rng.standard_normal(c_out).astype(np.float32)
    kp = checkpoint.default_kernel_points(2.5, rng)
    t = torch.from_numpy
    sdd = {'KPConv.kernel_points': t(kp), 'KPConv.weights': t(w), 'KPConv.bias': t(b)}
    ref = mo.kpconv(sdd, 'KPConv.', t(feats), t(q), t(s), t(idx), 1.2).numpy()
    got = ops.kpconv(t(feats).cuda(), t(q).cuda(), t(s).cuda(), t(idx).cuda(), t(kp).cuda(), 1.2, t(w).cuda(),
                     t(b).cuda()).cpu().numpy()
    assert max_err(got, ref) < REL                                   # fp32 SIMT contraction (gemm.cu)
    if c_in > 1:                                                     # tcgen05 3xTF32 contraction (gemm_tc.cu)
        w_nk = t(w).reshape(-1, c_out).t().contiguous().cuda()
        got_tc = ops.kpconv(t(feats).cuda(), t(q).cuda(), t(s).cuda(), t(idx).cuda(), t(kp).cuda(), 1.2, t(w).cuda(),
                            t(b).cuda(), weights_nk=w_nk).cpu().numpy()
        assert max_err(got_tc, ref) < REL
        assert max_err(got_tc, got) < 1e-5
        # gather variants with the kernel points as kernel arguments: fast dense loop, sparse lists, auto,
        # non-zero mask dispatch, packed FFMA2 loop, warp-level mma.sync (3xTF32)
        from lcrnet_b200 import _lib
        try:
            for mode in (1, 2, 3, 4, 5, 6, 7):
                _lib.lib().lcr_set_gather_mode(mode)
                got_v = ops.kpconv(t(feats).cuda(), t(q).cuda(), t(s).cuda(), t(idx).cuda(), t(kp).cuda(), 1.2,
                                   t(w).cuda(), t(b).cuda(), weights_nk=w_nk,
                                   kernel_points_host=t(kp).contiguous()).cpu().numpy()
                assert max_err(got_v, ref) < REL, mode
                assert max_err(got_v, got_tc) < 1e-5, mode
        finally:
            _lib.lib().lcr_set_gather_mode(3)
    else:
        got_fast = ops.kpconv(t(feats).cuda(), t(q).cuda(), t(s).cuda(), t(idx).cuda(), t(kp).cuda(), 1.2, t(w).cuda(),
                              t(b).cuda(), kernel_points_host=t(kp).contiguous()).cpu().numpy()
        assert max_err(got_fast, ref) < REL
        assert max_err(got_fast, got) < 1e-5


def test_kpconv_sparse_wide_table():
    """Sparse gather with H > 64 (two super-chunks) and rows made only of pads."""
    from lcrnet_b200 import ops
    rng = np.random.default_rng(5)
    s = rng.uniform(-3, 3, (700, 3)).astype(np.float32)
    q = np.concatenate([s[::5], np.full((3, 3), 50.0, np.float32)])      # last queries have no neighbours
    ql, sl = np.array([len(q)], dtype=np.int64), np.array([len(s)], dtype=np.int64)
    idx = on.radius_neighbors(q, s, ql, sl, 2.5, limit=150)
    assert idx.shape[1] > 64
    feats = rng.standard_normal((len(s), 64)).astype(np.float32)
    w = (rng.standard_normal((15, 64, 64)) * 0.1).astype(np.float32)
    kp = checkpoint.default_kernel_points(2.5, rng)
    t = torch.from_numpy
    sdd = {'KPConv.kernel_points': t(kp), 'KPConv.weights': t(w)}
    ref = mo.kpconv(sdd, 'KPConv.', t(feats), t(q), t(s), t(idx), 1.2).numpy()
    w_nk = t(w).reshape(-1, 64).t().contiguous().cuda()
    from lcrnet_b200 import _lib
    try:
        for mode in (1, 2, 4, 5, 6, 7):
            _lib.lib().lcr_set_gather_mode(mode)
            got = ops.kpconv(t(feats).cuda(), t(q).cuda(), t(s).cuda(), t(idx).cuda(), t(kp).cuda(), 1.2, t(w).cuda(),
                             None, weights_nk=w_nk, kernel_points_host=t(kp).contiguous()).cpu().numpy()
            assert max_err(got, ref) < REL, mode
    finally:
        _lib.lib().lcr_set_gather_mode(3)


@pytest.mark.parametrize('c', [32, 64, 128, 256, 512, 1024])
def test_unary_groupnorm_vs_oracle(c):
    from lcrnet_b200 import ops
    rng = np.random.default_rng(c)
    rows = [300, 2, 517]
    x = (rng.standard_normal((sum(rows), 64)) * 2 + 0.5).astype(np.float32)
    w = (rng.standard_normal((c, 64)) * 0.2).astype(np.float32)
    b = rng.standard_normal(c).astype(np.float32)
    gam, bet = rng.standard_normal(c).astype(np.float32), rng.standard_normal(c).astype(np.float32)
    t = torch.from_numpy
    sdd = {'u.mlp.weight': t(w), 'u.mlp.bias': t(b), 'u.norm.norm.weight': t(gam), 'u.norm.norm.bias': t(bet)}
    stacks = ops.Stacks(rows, 'cuda')
    y = ops.linear(t(x).cuda(), t(w).t().contiguous().cuda(), t(b).cuda())
    st = ops.group_norm_stats(y, stacks)
    got, flags = ops.group_norm_apply(y, st, t(gam).cuda(), t(bet).cuda(), stacks, leaky=True, want_flags=True)
    o = 0
    for n in rows:  # the reference normalises one stack per forward
        ref = mo.unary(sdd, 'u.', t(x[o:o + n])).numpy()
        assert max_err(got[o:o + n].cpu().numpy(), ref) < REL
        o += n
    rs = got.sum(1).cpu().numpy()
    clear = np.abs(rs) > 1e-3
    assert (flags.cpu().numpy().astype(bool) == (rs > 0))[clear].all()


def test_maxpool_vs_oracle():
    from lcrnet_b200 import ops
    rng = np.random.default_rng(3)
    x = rng.standard_normal((500, 128)).astype(np.float32)
    idx = rng.integers(0, 501, (200, 23)).astype(np.int64)   # 500 = pad row (zeros)
    ref = mo.maxpool(torch.from_numpy(x), torch.from_numpy(idx)).numpy()
    got = ops.maxpool(torch.from_numpy(x).cuda(), torch.from_numpy(idx).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize('name', ['s0', 's1'])
def test_encoder_and_descriptor_vs_fixture_and_oracle(name, sd, net):
    scene, seed, stride = (int(x) for x in G[name + '_case'])
    data = _pyramid(scene, seed, stride)
    feats = torch.ones(data['points'][0].shape[0], 1)
    with torch.no_grad():
        ref_list, ref_blocks = mo.kpencoder(sd, feats, data, return_all=True)
        ref_desc = mo.netvlad(sd, ref_list[-1]).numpy()
    d = _to_cuda(data)
    d['features'] = feats.cuda()
    with torch.no_grad():
        feats_list, blocks = net.encoder(d['features'], d, return_blocks=True)
        desc = net(d)['anc_global'].cpu().numpy()
    rows = int(G['rows'])
    for bn, t in blocks.items():
        got = t.cpu().numpy()
        ref = ref_blocks[bn.replace('encoder', '')].numpy()
        assert rel_err(got, ref) < REL, bn
        head = G['%s_%s_head' % (name, bn)]                       # the reference itself
        assert np.abs(got[:rows] - head).max() <= REL * max(1.0, np.abs(head).max()), bn
    assert desc.shape == (1, 256)
    assert rel_err(desc, ref_desc) < REL
    assert rel_err(desc, G[name + '_descriptor']) < REL          # vs the reference Python model
    assert abs(np.linalg.norm(desc) - 1.0) < 1e-5


def test_full_gpu_pipeline_matches_oracle(sd, net):
    """raw scan -> GPU pre-voxel + pyramid + tables -> encoder -> descriptor, vs the CPU oracle
    end to end (bit-exact tables, descriptor within 1e-4)."""
    from lcrnet_b200 import data as gdata
    raw = np.ascontiguousarray(synth.make_scan(2, 11)[::3])
    d = gdata.scans_collate_fn_stack_mode([raw], 4, 0.3, 1.275, [35, 35, 35, 35], pre_voxel=0.3, int32=False,
                                          upsampling=True)
    p0, l0 = on.grid_subsample(raw, np.array([len(raw)], dtype=np.int64), 0.3)
    ref = mo.precompute_pyramid(p0, l0, limits=[35, 35, 35, 35])
    for key in ('points', 'neighbors', 'subsampling', 'upsampling'):
        for a, b in zip(d[key], ref[key]):
            assert torch.equal(a.cpu(), b), key
    with torch.no_grad():
        want = mo.global_descriptor(sd, ref).numpy()
    d.pop('stack_size')
    got = net(d)['anc_global'].cpu().numpy()
    assert rel_err(got, want) < REL


def test_batched_stacks_equal_individual_forwards(net):
    """Many scans in one pass (stack_size=1) == the reference's one-scan-per-forward semantics."""
    from lcrnet_b200 import data as gdata
    scans = [np.ascontiguousarray(synth.make_scan(s, 40 + s)[::4]) for s in range(3)]
    lim = [30, 30, 30, 30]
    batch = gdata.scans_collate_fn_stack_mode(scans, 4, 0.3, 1.275, lim, pre_voxel=0.3)
    got = net(batch)['anc_global']
    assert got.shape == (3, 256)
    for i, s in enumerate(scans):
        one = gdata.scans_collate_fn_stack_mode([s], 4, 0.3, 1.275, lim, pre_voxel=0.3)
        one.pop('stack_size')
        single = net(one)['anc_global']
        assert rel_err(got[i].cpu().numpy(), single[0].cpu().numpy()) < 1e-5


@pytest.mark.parametrize('nq,ndb,k', [(50, 300, 5), (33, 1000, 25), (130, 64, 50)])
def test_l2_topk_vs_oracle(nq, ndb, k):
    from lcrnet_b200 import ops
    rng = np.random.default_rng(nq)
    db = rng.standard_normal((ndb, 256)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    q = db[rng.integers(0, ndb, nq)] + 0.05 * rng.standard_normal((nq, 256)).astype(np.float32)
    db[7] = db[3]                                             # exact duplicate rows: index tie-break
    ref_d, ref_i = mo.l2_topk(q, db, k)
    d2, idx = ops.l2_topk(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), k)
    d2, idx = d2.cpu().numpy(), idx.cpu().numpy()
    fin = np.isfinite(ref_d)
    assert np.array_equal(np.isfinite(d2), fin)
    assert np.allclose(d2[fin], ref_d[fin], rtol=1e-5, atol=1e-6)
    # indices: exact except where neighbouring distances are closer than fp32 summation noise
    gap_ok = np.ones_like(fin)
    gap = np.abs(np.diff(ref_d, axis=1)) > 1e-5
    gap_ok[:, 1:] &= gap
    gap_ok[:, :-1] &= gap
    sel = fin & gap_ok
    assert np.array_equal(idx[sel], ref_i[sel])
    assert (idx[~fin] == -1).all()
    # causal variant: query i only sees rows [0, valid[i])
    valid = rng.integers(0, ndb + 1, nq).astype(np.int32)
    ref_d, ref_i = mo.l2_topk(q, db, k, valid_counts=valid)
    d2, idx = ops.l2_topk(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), k,
                          valid_counts=torch.from_numpy(valid))
    idx = idx.cpu().numpy()
    assert ((idx < valid[:, None]) | (idx == -1)).all()
    assert np.array_equal(idx == -1, ref_i == -1)
    assert np.allclose(d2.cpu().numpy()[ref_i >= 0], ref_d[ref_i >= 0], rtol=1e-5, atol=1e-6)


def test_multistream_pipeline_matches_single_stream(net):
    """pipeline.DescriptorPipeline: the batch cut into chunks on 3 streams / host threads gives the
    descriptors of the single-stream pass (scans are independent units; bit-identical expected,
    bar 1e-6)."""
    from lcrnet_b200 import pipeline
    scans = [np.ascontiguousarray(synth.make_scan(i // 2, 7351 + i)[::8]) for i in range(7)]
    pts = torch.from_numpy(np.concatenate(scans, 0)).pin_memory()
    lens = [len(s) for s in scans]
    lim = [30, 30, 30, 30]
    one = pipeline.DescriptorPipeline(net, lim, n_streams=1)(pts, lens).cpu().numpy()
    pipe = pipeline.DescriptorPipeline(net, lim, n_streams=3)
    for src in (pts, pts.cuda()):
        many = pipe(src, lens)
        torch.cuda.current_stream().synchronize()
        assert np.abs(many.cpu().numpy() - one).max() <= 1e-6
    pipe.close()
    assert np.allclose(np.linalg.norm(one, axis=1), 1.0, atol=1e-5)


def test_multistream_pipeline_cold_caches():
    """ADVICE r1 (ops.py:40): derived weights (tf32 halves, transposes, packed BN) are shared between the
    pipeline's streams.  A FRESH network enters the 3-stream pipeline first (nothing primed by a single-stream
    pass), then its weights are replaced in place while the pipeline exists (every cache entry goes stale and is
    rebuilt from whichever side stream touches it first, protected by the entry's event); both results must equal
    the single-stream pass computed afterwards."""
    from lcrnet_b200 import model, pipeline
    scans = [np.ascontiguousarray(synth.make_scan(i // 2, 7351 + i)[::8]) for i in range(6)]
    pts = torch.from_numpy(np.concatenate(scans, 0)).pin_memory()
    lens = [len(s) for s in scans]
    lim = [30, 30, 30, 30]
    fresh = model.create_model(model.default_cfg()).eval()
    fresh.load_state_dict(checkpoint.random_state_dict('global_descriptor', 11), strict=True)
    fresh = fresh.cuda()
    pipe = pipeline.DescriptorPipeline(fresh, lim, n_streams=3)
    cold = pipe(pts, lens).cpu().numpy()                      # first use of this network at all
    fresh.load_state_dict({k: v.cuda() for k, v in checkpoint.random_state_dict('global_descriptor', 12).items()},
                          strict=True)                        # in-place copy: versions bump, caches stale
    stale = pipe(pts, lens).cpu().numpy()                     # rebuilt on the fly from the side streams
    pipe.close()
    one = pipeline.DescriptorPipeline(fresh, lim, n_streams=1)
    ref12 = one(pts, lens).cpu().numpy()
    fresh.load_state_dict({k: v.cuda() for k, v in checkpoint.random_state_dict('global_descriptor', 11).items()},
                          strict=True)
    ref11 = one(pts, lens).cpu().numpy()
    assert np.abs(cold - ref11).max() <= 1e-6
    assert np.abs(stale - ref12).max() <= 1e-6
    assert np.abs(ref11 - ref12).max() > 1e-3                 # the two weight sets really differ
