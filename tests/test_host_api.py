"""Host-side mirror of the reference's object API (re-targeted from seismic/test_seismic_utils.py
and the set-up code of the drivers). CPU only."""
import numpy as np
import pytest

import devito_fwi_b200 as b
from devito_fwi_b200 import configs
from devito_fwi_b200.sparse import resolve
from oracle import ref


@pytest.mark.parametrize('nbl', [0, 10, 40])
@pytest.mark.parametrize('shape', [(21, 21), (21, 21, 21)])
def test_damp(nbl, shape):
    """seismic/test_seismic_utils.py:12-36"""
    dims = len(shape)
    model = b.demo_model('constant-isotropic', nbl=nbl, shape=shape, spacing=tuple([10.] * dims))
    center = tuple(s // 2 + nbl for s in shape)
    if nbl == 0:
        assert model.damp == 0
        return
    assert model.grid.shape == tuple(s + 2 * nbl for s in shape)
    assert model.damp.data[center] == 0             # demo models are built with bcs="damp"
    with pytest.warns(UserWarning):
        model._initialize_bcs(bcs="mask")
    assert model.damp.data[center] == 1
    assert np.all(model.damp.data <= 1)
    with pytest.warns(UserWarning):
        model._initialize_bcs(bcs="damp")
    assert model.damp.data[center] == 0
    # against the oracle's restatement of initialize_damp
    want = ref.init_damp(model.grid.shape, nbl, [10.] * dims)
    assert np.array_equal(model.damp.data, want)


def test_default_geoms():
    """seismic/test_seismic_utils.py:39-97: dt override, shapes, resample, zero source."""
    model = b.demo_model('constant-isotropic', shape=(21, 21), spacing=(10., 10.), dt=1.0)
    assert model.critical_dt == 1.0
    geometry = b.setup_geometry(model, 250.0)
    nrec = 21
    assert geometry.grid == model.grid
    assert geometry.nrec == nrec and geometry.nsrc == 1
    assert geometry.src_type == "Ricker"
    assert geometry.rec.shape == (251, nrec)
    assert np.isclose(np.linalg.norm(geometry.rec.data), 0)
    assert geometry.src.shape == (251, 1)
    assert geometry.new_src(src_type=None).data.max() == 0 and geometry.new_src(src_type=None).shape == (251, 1)
    rec2 = geometry.rec.resample(num=501)
    assert rec2.shape == (501, nrec)
    assert np.isclose(rec2.time_range.step, 0.5)
    assert geometry.src.resample(dt=1.0) is not None
    with pytest.raises(ValueError):
        b.demo_model('constant-isotropic', shape=(21, 21), spacing=(10., 10.), dt=100.0).critical_dt


def test_model_update_and_cfl():
    g_true, g_init, g_const, mask = configs.marmousi()
    m = g_init.model
    assert m.grid.shape == (380, 186) and m.critical_dt == 2.95 and g_init.nt == 1357
    assert np.isclose(m._cfl_coeff, 0.5189321559)
    assert mask.shape == (300, 106) and not mask[:, :7].any()
    v = np.full(m.shape, 2.0, dtype=np.float32)
    v[0, 0] = 1.7
    m.update('vp', v)
    assert m.vp.data.shape == (380, 186)
    assert np.all(m.vp.data[:41, :41] == np.float32(1.7))       # edge replication into the corner
    with pytest.raises(ValueError):
        m.update('vp', np.zeros((3, 3), dtype=np.float32))
    ga, gb = configs.marmousi2()[:2]
    assert ga.model.grid.shape == (420, 220) and ga.nt == 1501 and ga.nrec == 340 and ga.nsrc == 31
    ca, cb = configs.circle()
    assert ca.model.grid.shape == (281, 281) and ca.nt == 1001 and ca.nsrc == 11 and ca.nrec == 201


def test_ricker_and_time_axis_match_oracle():
    g = configs.marmousi()[0]
    nt, stop, tv = ref.time_axis(0., 4000., 2.95)
    assert nt == g.nt
    assert np.allclose(g.src.data[:, 0], ref.ricker(0.007, tv).astype(np.float32))
    assert g.src.data.dtype == np.float32 and g.src.coordinates.data.dtype == np.float32


def test_sparse_resolution_weights():
    """Multilinear weights: partition of unity, exact at nodes, off-grid receivers."""
    g = configs.marmousi()[0]
    off, w = resolve(g.grid, g.rec_positions)
    assert off.shape == (300, 4) and (off >= 0).all()
    assert np.allclose(w.sum(axis=1), 1.0, atol=1e-6)
    # depth 60 m is a grid line: the two z1 corners carry zero weight
    assert np.allclose(w[:, 1], 0, atol=1e-7) and np.allclose(w[:, 3], 0, atol=1e-7)
    pitch = g.grid.pitch
    ix, iz = off[:, 0] // pitch, off[:, 0] % pitch
    assert np.array_equal(iz, np.full(300, 42)) and ix[0] == 41
    # outside the padded grid -> dropped corners
    off2, _ = resolve(g.grid, np.array([[-1300., 60.], [1e6, 60.]]))
    assert (off2 == -1).all()
    # 3-D
    grid3 = b.Grid(shape=(11, 12, 13), extent=(100., 110., 120.))
    off3, w3 = resolve(grid3, np.array([[33.3, 47.1, 58.9]]))
    assert off3.shape == (1, 8) and np.isclose(w3.sum(), 1.0, atol=1e-6)


def test_shot_partition():
    from devito_fwi_b200 import dist
    shots = [dist.local_shots(29, r, 8) for r in range(8)]
    assert sorted(sum(shots, [])) == list(range(29))
    assert max(map(len, shots)) - min(map(len, shots)) <= 1
    assert dist.local_shots(29) == list(range(29))


def test_launch_group_partition_logic():
    """resident.best_partition: surveys with more shots than resident clusters are cut into launch groups."""
    from devito_fwi_b200.resident import best_partition
    # Marmousi-like candidates: (cost of one wave, clusters resident at once, cluster size)
    cands = [(95 * 1.08 + 60, 33, 4), (76 * 1.10 + 60, 26, 5), (64 * 1.12 + 60, 22, 6), (48 * 1.16 + 60, 14, 8)]
    assert best_partition(cands, 29) == [(29, 4)]                    # one wave of the smallest cluster that holds them
    assert best_partition(cands, 5) == [(5, 8)]                      # few shots: the widest cluster
    assert best_partition(cands, 26) == [(26, 5)]
    # more than one wave: a full wave of 5-CTA clusters + the remainder on 8-CTA clusters (259) beats 33 x 4 + 7 x 8 (278)
    assert best_partition(cands, 40) == [(26, 5), (14, 8)]
    for n in (1, 13, 14, 15, 66, 67, 300):
        g = best_partition(cands, n)
        assert sum(k for k, _ in g) == n and all(k <= dict((c, s) for _, s, c in cands)[c] for k, c in g)
    assert best_partition([], 3) is None


def test_checkpoint_segment_plan():
    from devito_fwi_b200.checkpoint import plan_segments
    segs = plan_segments(1, 688)
    assert segs[0][0] == 1 and segs[-1][1] == 688
    assert all(a2 == b1 + 1 for (_, b1), (a2, _) in zip(segs, segs[1:]))          # contiguous, no overlap
    S = segs[0][1] - segs[0][0] + 1
    assert S == 38 and len(segs) == 19                                             # ~sqrt(2 * 688)
    assert plan_segments(5, 4) == [] and plan_segments(3, 3) == [(3, 3)]
    assert plan_segments(1, 10, segment=4) == [(1, 4), (5, 8), (9, 10)]


def test_checkpoint_keep_plan():
    """How many trailing steps keep their wavefield from pass 1 for a given amount of spare HBM (checkpoint.plan_keep)."""
    import math
    from devito_fwi_b200.checkpoint import plan_keep

    def need(steps, K, S, reserve=8):
        rest = steps - K
        return K + 2 + (S + 2 if rest > 0 else 0) + 2 * int(math.ceil(rest / float(S))) + reserve

    # everything fits: the whole history is kept, nothing is recomputed
    K, S = plan_keep(100, 1000)
    assert K == 100
    # the 592^3, nt=690 shot on a 180 GB B200: ~175 slices of 0.83 GB are spare
    K, S = plan_keep(688, 175)
    assert 80 <= K <= 120 and need(688, K, S) <= 175 and need(688, K + 1, S) > 175
    # tight memory: no room for a pass-1 history, the plan is the plain two-level scheme
    K, S = plan_keep(688, 60)
    assert K == 0 or need(688, K, S) <= 60
    # a fixed segment length is honoured
    K, S = plan_keep(50, 40, segment=5)
    assert S == 5 and need(50, K, 5) <= 40 and need(50, K + 1, 5) > 40
    assert plan_keep(0, 100)[0] == 0


def test_sparse_fuse_tables():
    """Bucketing of injection contributions by (plane, row) and of recorded points by (plane, 16-row tile) for the
    sweep kernels' service warps."""
    import devito_fwi_b200 as b
    from devito_fwi_b200 import sparse
    shape = (20, 37, 50)
    model = b.Model(origin=(0., 0., 0.), spacing=(10., 10., 10.), shape=shape, space_order=4,
                    vp=np.full(shape, 2.0, np.float32), nbl=3, bcs="damp")
    grid = model.grid
    rng = np.random.default_rng(0)
    pts = rng.uniform(5., 10. * np.array(shape) - 5., size=(60, 3))
    off, w = sparse.resolve(grid, pts)
    flat, pt = off.ravel(), np.repeat(np.arange(60), 8)
    order = np.lexsort((pt, flat))
    v_off = flat[order]
    cells, start = np.unique(v_off, return_index=True)
    cell_ptr = np.append(start, v_off.size).astype(np.int32)
    t, max_row_con = sparse.fuse_tables(grid, off, cells, cell_ptr)
    npl, nr = grid.shape[0], grid.shape[1]
    sr = grid.slice_shape[2]
    sp = grid.slice_shape[1] * sr
    nrt = (nr + 15) // 16
    assert np.array_equal(t['con_off'], v_off)
    assert t['con_rowptr'].shape == (npl * nr + 1,) and t['con_rowptr'][0] == 0 and t['con_rowptr'][-1] == v_off.size
    assert max_row_con == np.diff(t['con_rowptr']).max() and 0 < max_row_con <= sparse.MAX_ROW_CON
    for k in rng.integers(0, npl * nr, 80):
        p, r = divmod(int(k), nr)
        c = t['con_off'][t['con_rowptr'][k]:t['con_rowptr'][k + 1]]
        assert np.all(c // sp == p) and np.all((c % sp) // sr == r)
    # every point appears once, filed under its first in-grid corner, bucketed by (plane, 16-row tile)
    assert sorted(t['pt_order'].tolist()) == list(range(60))
    assert np.all(np.diff(t['pt_home']) >= 0) and t['pt_rowptr'][-1] == 60 and t['pt_rowptr'].shape == (npl * nrt + 1,)
    for k in rng.integers(0, npl * nrt, 60):
        p, rt = divmod(int(k), nrt)
        h = t['pt_home'][t['pt_rowptr'][k]:t['pt_rowptr'][k + 1]]
        assert np.all(h // sp == p) and np.all((h % sp) // sr // 16 == rt)
    for i in (0, 37, 59):
        assert t['pt_home'][i] == off[t['pt_order'][i]][0]
    # maps that the service warps do not take run as separate kernels: too dense, or a point outside the grid
    dense = np.stack(np.meshgrid(np.arange(0., 190., 2.5), np.arange(0., 360., 2.5), [25.], indexing='ij'), -1).reshape(-1, 3)
    off2, _ = sparse.resolve(grid, dense)
    f2 = np.sort(off2.ravel())
    c2, s2 = np.unique(f2, return_index=True)
    assert sparse.fuse_tables(grid, off2, c2, np.append(s2, f2.size).astype(np.int32)) == ({}, 0)
    off3, _ = sparse.resolve(grid, np.array([[-500., 10., 10.]]))
    assert (off3 < 0).all() and sparse.fuse_tables(grid, off3, np.zeros(0, np.int64), np.zeros(1, np.int32)) == ({}, 0)
