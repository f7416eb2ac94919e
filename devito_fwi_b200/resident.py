"""Host side of the SM-resident 2-D engine (csrc/resident2d.cu): decomposition plan, injection /
recording maps, device buffers of one survey, and the batched forward / gradient calls.

``ResidentSurvey`` is what the shot loops of fwi.py (fm_multi fwi.py:67-81, fwi_obj_multi
fwi.py:183-199) run on: all shots of the rank go through ONE kernel launch per sweep, one
thread-block cluster per shot.  It needs a 2-D model, space_order 4/6/8, the separable damping
profile of seismic/model.py:31-49 and a zero initial wavefield; otherwise callers fall back to
the per-shot streaming engine (``ResidentSurvey.supported`` tells).
"""
import ctypes

import numpy as np

from . import _lib
from .grid import Function, HALO
from .sparse import resolve
from .wavesolver import grid_struct, _ptr, _stream

__all__ = ['ResidentSurvey', 'plan_model']

MAX_CELLS = 1024      # RES2D_MAX_CELLS
MAX_CON = 1024        # RES2D_MAX_CON (contributions per CTA; point indices are staged as uint16)
MAX_CLUSTER = 16      # CTAs (SMs) per shot; 9..16 are non-portable cluster sizes
LAT_MAXV = 1024       # RES2D_LAT_MAXV / RES2D_LAT_MAXITP: staging limits of the 4-row-strip kernel (resident2d_lat.cu)
LAT_MAXITP = 1024


class Plan(ctypes.Structure):
    """struct b2fwi_res2d_plan"""
    _fields_ = [(n, ctypes.c_int32) for n in
                ('cluster', 'rows_per_thread', 'groups', 'threads', 'rows_cta', 'tile_rows', 'smem_bytes',
                 'wx0', 'wx1', 'wq0', 'wq1', 'tile_pitch')]


class Maps(ctypes.Structure):
    """struct b2fwi_res2d_maps (device pointers)"""
    _fields_ = [(n, ctypes.c_void_p) for n in
                ('inj_desc', 'inj_cptr', 'inj_pt', 'inj_w', 'thr_mask', 'thr_base',
                 'itp_desc', 'itp_pt', 'itp_off', 'itp_w')]


def plan_model(grid, space_order, nbl, min_cluster=1, min_rows=1):
    """b2fwi_res2d_plan_model (first fit from `min_cluster` upwards); a Plan, or None when the grid does not fit."""
    if grid.dim != 2 or space_order not in (4, 6, 8):
        return None
    g = grid_struct(grid, space_order)
    plan = Plan()
    rc = _lib.lib().b2fwi_res2d_plan_model(ctypes.byref(g), int(nbl), int(min_cluster), int(min_rows),
                                           ctypes.byref(plan))
    return plan if rc == 0 else None


def plan_exact(grid, space_order, nbl, cluster, rows):
    """b2fwi_res2d_plan_exact: the plan with `cluster` CTAs per shot and `rows` rows per thread, or None."""
    if grid.dim != 2 or space_order not in (4, 6, 8):
        return None
    g = grid_struct(grid, space_order)
    plan = Plan()
    rc = _lib.lib().b2fwi_res2d_plan_exact(ctypes.byref(g), int(nbl), int(cluster), int(rows), ctypes.byref(plan))
    return plan if rc == 0 else None


ROWS_PER_THREAD = (3, 4, 8, 12, 16)


def step_cost(plan):
    """Relative time of one time step of a cluster with this plan (fitted to B200 measurements, Marmousi / circle /
    Marmousi2, 1-29 shots, profiles/r02_res2d_plans.txt): a step is paced by the busiest of the 4 warp schedulers of an
    SM, i.e. by (warps per scheduler) x (rows per thread + window fill), plus a fixed part per step -- small for the
    short-strip kernels (everything per-point lives in registers / shared memory), large for the long-strip kernel
    (B and the history are re-read from L2 every row)."""
    wps = -(-(plan.threads // 32) // 4)
    P = plan.rows_per_thread
    if P <= 4:
        c = 0.56 + 0.227 * wps * (P + 2)
        return c * (1.1 if plan.threads > 384 else 1.0)       # 4 warps per scheduler: 128 registers, batches of 2 rows
    return 4.9 + 0.263 * max(wps, 3) * P * (1.35 if P >= 16 else 1.0)     # 16-row strips spill


def candidates(grid, space_order, nbl, min_rows=1):
    """Every feasible decomposition as (plan, clusters the device keeps resident)."""
    g = grid_struct(grid, space_order)
    out = []
    for C in range(1, MAX_CLUSTER + 1):
        for P in ROWS_PER_THREAD:
            if P < min_rows:
                continue
            plan = plan_exact(grid, space_order, nbl, C, P)
            if plan is None:
                continue
            n = ctypes.c_int32()
            rc = _lib.lib().b2fwi_res2d_max_active_clusters(ctypes.byref(g), ctypes.byref(plan), ctypes.byref(n))
            if rc == 0 and n.value > 0:
                out.append((plan, int(n.value)))
    return out


_CANDS = {}


def _cands(grid, space_order, nbl, min_rows):
    key = (grid._key(), space_order, nbl, min_rows)
    if key not in _CANDS:
        _CANDS[key] = candidates(grid, space_order, nbl, min_rows)
    return _CANDS[key]


def choose_plan(grid, space_order, nbl, nshots, min_rows=1):
    """Decomposition for `nshots` concurrent shots in ONE launch: minimises (waves of resident clusters) x (time per
    step) -- few shots get up to 16 SMs each and short strips, many shots the smallest cluster that fits."""
    best, best_cost = None, None
    for plan, slots in _cands(grid, space_order, nbl, min_rows):
        cost = -(-nshots // slots) * step_cost(plan)
        if best is None or cost < best_cost:
            best, best_cost = plan, cost
    return best


def partition_shots(grid, space_order, nbl, nshots, min_rows=1):
    """Split ``nshots`` concurrent shots into launch groups [(count, plan), ...].

    One launch keeps ``slots`` clusters resident; more shots than that run in waves, and the last wave is usually far
    from full. Cheaper: give each wave its own launch and its own decomposition - a full wave of the smallest cluster,
    then the remainder on wider clusters with shorter strips; best(n) = min over plans p of cost_p + best(n - slots_p)."""
    cands = [(step_cost(plan), slots, plan) for plan, slots in _cands(grid, space_order, nbl, min_rows)]
    return best_partition(cands, nshots)


def best_partition(cands, nshots):
    """``cands``: [(cost of one wave, resident clusters, cluster size)]; returns [(count, cluster), ...] covering
    ``nshots`` at minimum total cost: best(n) = min_p cost_p + best(n - min(n, slots_p))."""
    if not cands:
        return None
    memo = {}

    def best(n):
        if n <= 0:
            return 0.0, []
        if n not in memo:
            opts = []
            for cost, slots, cluster in cands:
                k = min(n, slots)
                rest_cost, rest = best(n - k)
                opts.append((cost + rest_cost, [(k, cluster)] + rest))
            memo[n] = min(opts, key=lambda o: o[0])
        return memo[n]

    return best(int(nshots))[1]


def build_maps(grid, plan, R, inj_coords, itp_coords=None):
    """numpy arrays of struct b2fwi_res2d_maps for a list of shots.

    inj_coords: per shot, [npoint, 2] positions injected into the field (source(s) on the forward
    sweep, receivers on the backward sweep).  itp_coords: per shot, receiver positions recorded by
    the forward sweep (or None).  Cells hit by several points gather their contributions in
    ascending point order, exactly like sparse.SparseMap."""
    C, P, T = plan.cluster, plan.rows_per_thread, plan.threads
    rows_cta = plan.rows_cta
    nzq = (grid.shape[1] + 3) // 4
    gpitch = grid.pitch
    spitch = plan.tile_pitch          # shared tile row pitch; one zero quad on each side of a row (resident2d.cu)
    nshots = len(inj_coords)
    inj_desc = np.zeros((nshots * C, 2), dtype=np.int32)
    thr_mask = np.zeros((nshots * C, T), dtype=np.uint64)
    thr_base = np.zeros((nshots * C, T), dtype=np.int32)
    cptr_all, pt_all, w_all = [], [], []
    ncontrib = 0
    cell_base = 0
    for s in range(nshots):
        off, w = resolve(grid, inj_coords[s])
        npoint, nc = off.shape
        flat_off, flat_w = off.ravel(), w.ravel()
        pt = np.repeat(np.arange(npoint, dtype=np.int32), nc)
        valid = flat_off >= 0
        v_off, v_w, v_pt = flat_off[valid], flat_w[valid], pt[valid]
        row, col = np.divmod(v_off - 0, gpitch)
        crank = row // rows_cta
        lr = row - crank * rows_cta
        tid = (lr // P) * nzq + col // 4
        bit = (lr % P) * 4 + col % 4
        order = np.lexsort((v_pt, bit, tid, crank))            # CTA, owner thread, bit, then point order
        crank, tid, bit, v_w, v_pt = crank[order], tid[order], bit[order], v_w[order], v_pt[order]
        key = (crank.astype(np.int64) * T + tid) * 64 + bit
        for c in range(C):
            sel = crank == c
            sc = s * C + c
            k = key[sel]
            cells, start = np.unique(k, return_index=True)
            ncell = cells.size
            if ncell > MAX_CELLS or k.size > MAX_CON or npoint > 65535:
                return None
            if P <= 4 and k.size and int(v_pt[sel].max()) - int(v_pt[sel].min()) + 1 > LAT_MAXV:
                return None         # the short-strip kernels stage the span of points one CTA uses
            inj_desc[sc] = (ncell, cell_base)
            cptr = np.concatenate([start, [k.size]]).astype(np.int32) + ncontrib
            cptr_all.append(cptr)
            pt_all.append(v_pt[sel])
            w_all.append(v_w[sel])
            ncontrib += int(k.size)
            cell_base += ncell + 1
            ctid = ((cells // 64) % T).astype(np.int64)
            cbit = (cells % 64).astype(np.uint64)
            np.bitwise_or.at(thr_mask[sc], ctid, np.uint64(1) << cbit)
            first = np.full(T, -1, dtype=np.int64)
            # cells are sorted by (tid, bit): the first occurrence of a tid is its base slot
            utid, ufirst = np.unique(ctid, return_index=True)
            first[utid] = ufirst
            thr_base[sc] = np.where(first >= 0, first, 0)
    maps = dict(inj_desc=inj_desc, inj_cptr=np.concatenate(cptr_all).astype(np.int32),
                inj_pt=np.concatenate(pt_all).astype(np.int32) if pt_all else np.zeros(1, np.int32),
                inj_w=np.concatenate(w_all).astype(np.float32) if w_all else np.zeros(1, np.float32),
                thr_mask=thr_mask, thr_base=thr_base)
    if maps['inj_pt'].size == 0:
        maps['inj_pt'] = np.zeros(1, np.int32)
        maps['inj_w'] = np.zeros(1, np.float32)

    itp_desc = np.zeros((nshots * C, 2), dtype=np.int32)
    ipt, ioff, iw = [], [], []
    base = 0
    if itp_coords is not None:
        nx = grid.shape[0]
        for s in range(nshots):
            off, w = resolve(grid, itp_coords[s])
            npoint = off.shape[0]
            row, col = np.divmod(np.where(off >= 0, off, 0), gpitch)
            # owner: the CTA holding the point's base row (corner 0); out-of-grid points go to an end CTA
            coords = np.ascontiguousarray(np.reshape(itp_coords[s], (-1, 2)), dtype=np.float32)
            ix = np.floor((coords[:, 0] - np.float32(grid.origin[0])) / np.float32(grid.spacing[0])).astype(np.int64)
            owner = np.clip(ix, 0, nx - 1) // rows_cta
            for c in range(C):
                sel = np.nonzero(owner == c)[0]
                if P <= 4 and sel.size > LAT_MAXITP:
                    return None
                sc = s * C + c
                itp_desc[sc] = (sel.size, base)
                trow = row[sel] - c * rows_cta + R
                ok = (off[sel] >= 0) & (trow >= 0) & (trow < plan.tile_rows)
                ioff.append(np.where(ok, trow * spitch + col[sel] + 4, -1).astype(np.int32))
                iw.append(w[sel].astype(np.float32))
                ipt.append(sel.astype(np.int32))
                base += sel.size
    maps.update(itp_desc=itp_desc,
                itp_pt=np.concatenate(ipt) if base else np.zeros(1, np.int32),
                itp_off=np.concatenate(ioff) if base else np.full((1, 4), -1, np.int32),
                itp_w=np.concatenate(iw) if base else np.zeros((1, 4), np.float32))
    return maps


class _DevMaps(object):
    def __init__(self, maps):
        import torch
        self.t = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int64) if v.dtype == np.uint64
                                      else np.ascontiguousarray(v)).cuda() for k, v in maps.items()}
        self.struct = Maps()
        for k, t in self.t.items():
            setattr(self.struct, k, t.data_ptr())

    def byref(self):
        return ctypes.byref(self.struct)


class ResidentSurvey(object):
    """Device-resident state of the shots ``shots`` of ``geometry`` (default: all of them)."""

    def __init__(self, geometry, shots=None, space_order=None, min_cluster=1, min_rows=1, plan=None):
        import torch
        self.geometry = geometry
        model = self.model = geometry.model
        self.grid = model.grid
        self.space_order = int(space_order or model.space_order)
        self.R = self.space_order // 2
        self.shots = list(range(geometry.nsrc)) if shots is None else list(shots)
        self.nshots = len(self.shots)
        if plan is not None:
            self.plan = plan
        else:
            self.plan = (choose_plan(self.grid, self.space_order, model.nbl, self.nshots, min_rows) if min_cluster == 1
                         else plan_model(self.grid, self.space_order, model.nbl, min_cluster, min_rows))
        if self.plan is None or not self.supported(geometry, self.space_order) or getattr(model, 'fs', False):
            # (free-surface models: the mirrored top rows exist in the streaming kernels only)
            raise ValueError("model does not fit the SM-resident engine")
        self.nt = geometry.nt
        self.dt = float(geometry.dt)
        self.nrec = geometry.nrec
        self.gs = grid_struct(self.grid, self.space_order)
        p = self.plan
        self.wshape = (p.wx1 - p.wx0, (p.wq1 - p.wq0) * 4)
        self.col0 = model.nbl - p.wq0 * 4
        # damping profiles (damp/dt), separable
        model._initialize_bcs(bcs="damp")
        px, pz = model.damp._profiles
        nzq4 = 4 * ((self.grid.shape[1] + 3) // 4)
        sz = np.zeros(nzq4, dtype=np.float32)
        sz[:pz.size] = pz / self.dt
        self._damp_ref = np.array(model.damp._buf.host_ro())
        model.damp._buf.dev()
        self.sx = torch.from_numpy((px / self.dt).astype(np.float32)).cuda()
        self.sz = torch.from_numpy(sz).cuda()
        # source wavelets [nshots][nt][1] and maps
        wav = np.asarray(geometry.src.data, dtype=np.float32)          # [nt, nsrc_total]
        self.src = torch.from_numpy(np.ascontiguousarray(wav[:, self.shots].T[:, :, None])).cuda()
        src_pos = [geometry.src_positions[i:i + 1] for i in self.shots]
        rec_pos = [geometry.rec_positions] * self.nshots
        mf = build_maps(self.grid, p, self.R, src_pos, rec_pos)
        mb = build_maps(self.grid, p, self.R, rec_pos, None)
        if (mf is None or mb is None) and p.rows_per_thread <= 4 and min_rows < 8:
            # beyond the staging limits of the short-strip kernels: plan again with longer strips
            self.__init__(geometry, shots, space_order, min_cluster, min_rows=8)
            return
        if mf is None or mb is None:
            raise ValueError("too many injection cells per CTA for the SM-resident engine")
        self.maps_fwd, self.maps_bwd = _DevMaps(mf), _DevMaps(mb)
        self.B = torch.zeros(self.grid.slice_shape, dtype=torch.float32, device='cuda')
        self.rec = torch.zeros((self.nshots, self.nt, self.nrec), dtype=torch.float32, device='cuda')
        self.hist = None
        self.illum = None
        self.grad = None

    @staticmethod
    def supported(geometry, space_order=None):
        model = geometry.model
        so = int(space_order or model.space_order)
        if model.dim != 2 or so not in (4, 6, 8) or np.dtype(model.dtype) != np.float32 or model.nbl == 0:
            return False
        if not isinstance(model.vp, Function) or HALO != 0:
            return False
        model._initialize_bcs(bcs="damp")
        if getattr(model.damp, '_profiles', None) is None:
            return False
        return plan_model(model.grid, so, model.nbl) is not None

    # ------------------------------------------------------------------
    @property
    def nbytes(self):
        """Device bytes held by this survey (history, records, accumulators)."""
        return sum(t.numel() * t.element_size() for t in (self.hist, self.illum, self.grad, self.rec, self.B)
                   if t is not None)

    def _check_model(self):
        """The survey froze dt, the damping profiles and the wavelet at construction (they are geometry-only in the
        reference's drivers): refuse to run on a model that has moved away from them."""
        model = self.model
        dt_now = float(model.critical_dt)         # raises ValueError when a user dt exceeds the CFL limit (model.py:366-369)
        if abs(dt_now - self.dt) > 1e-6 * abs(self.dt):
            raise ValueError("model.critical_dt changed from %g to %g since this survey was built" % (self.dt, dt_now))
        buf = model.damp._buf
        if buf._newer == 'host':                  # the host view of damp was handed out since the last check
            if not np.array_equal(buf._host, self._damp_ref):
                import warnings
                warnings.warn("model.damp.data was edited after the resident survey was built: the SM-resident engine "
                              "keeps using the separable profile of model.py:31-49; set fwi.ENGINE = 'stream' for a "
                              "custom damping field")
                self._damp_ref = np.array(buf._host)       # warn once per edit
            buf.dev()                             # device copy current again: the flag is cleared

    def set_model(self, vp_dev=None):
        """(Re)compute B = dt^2 vp^2 from the model's current velocity."""
        self._check_model()
        vp_dev = self.model.vp._buf.dev() if vp_dev is None else vp_dev
        _lib.check(_lib.lib().b2fwi_res2d_prepare(ctypes.byref(self.gs), _ptr(vp_dev), ctypes.c_float(self.dt),
                                                  _ptr(self.B), _stream()))

    def forward(self, save=False, illum=False):
        """All shots forward: returns the device tensor rec[nshots, nt, nrec]."""
        import torch
        steps = self.nt - 2
        if save and self.hist is None:
            self.hist = torch.empty((self.nshots, steps) + self.wshape, dtype=torch.float32, device='cuda')
        if illum and self.illum is None:
            self.illum = torch.empty((self.nshots,) + self.wshape, dtype=torch.float32, device='cuda')
        self.set_model()
        _lib.check(_lib.lib().b2fwi_res2d_forward(
            ctypes.byref(self.gs), ctypes.byref(self.plan), _ptr(self.B), _ptr(self.sx), _ptr(self.sz),
            ctypes.c_float(self.dt), self.nt, 1, self.nt - 2, self.nshots, _ptr(self.src), 1,
            self.maps_fwd.byref(), _ptr(self.rec), self.nrec, _ptr(self.hist) if save else None,
            _ptr(self.illum) if illum else None, _stream()))
        return self.rec

    def gradient(self, residual):
        """All shots backward + imaging from the saved history: returns grad[nshots, wx, wz4] (window layout)."""
        import torch
        assert self.hist is not None, "forward(save=True) first"
        assert tuple(residual.shape) == (self.nshots, self.nt, self.nrec) and residual.is_contiguous()
        if self.grad is None:
            self.grad = torch.empty((self.nshots,) + self.wshape, dtype=torch.float32, device='cuda')
        _lib.check(_lib.lib().b2fwi_res2d_gradient(
            ctypes.byref(self.gs), ctypes.byref(self.plan), _ptr(self.B), _ptr(self.sx), _ptr(self.sz),
            ctypes.c_float(self.dt), self.nt, 1, self.nt - 2, self.nshots, _ptr(residual), self.nrec,
            self.maps_bwd.byref(), _ptr(self.hist), _ptr(self.grad), _stream()))
        return self.grad

    def crop(self, window_field):
        """[.., wx, wz4] window layout -> [.., nx, nz] physical domain."""
        nz = self.model.shape[1]
        return window_field[..., self.col0:self.col0 + nz]

    def masks(self):
        """fp64 mute masks of all shots, [nshots, nx, nz] (b2fwi_geometry_mask; geometry-only, built once)."""
        import torch
        if getattr(self, '_masks', None) is None:
            nx, nz = self.model.shape
            m = torch.empty((self.nshots, nx, nz), dtype=torch.float64, device='cuda')
            for k, i in enumerate(self.shots):
                pts = np.ascontiguousarray(np.concatenate([self.geometry.src_positions[i:i + 1],
                                                           self.geometry.rec_positions]), dtype=np.float64)
                pts_dev = torch.from_numpy(pts).cuda()
                _lib.check(_lib.lib().b2fwi_geometry_mask(ctypes.byref(self.gs), self.model.nbl, _ptr(pts_dev),
                                                          pts.shape[0], _ptr(m[k]), _stream()))
            self._masks = m
        return self._masks

    def window_mask_accumulate_all(self, field, out, use_mask=True):
        """out[nx, nz] (fp64) += sum_shots crop(field[s]) * mask[s]."""
        nx, nz = self.model.shape
        _lib.check(_lib.lib().b2fwi_window_mask_accumulate_batch(
            self.nshots, nx, nz, _ptr(field), self.wshape[0] * self.wshape[1], self.wshape[1], self.col0,
            _ptr(self.masks()) if use_mask else None, _ptr(out), _stream()))

    def window_mask_accumulate(self, field_s, mask, out):
        """out[nx, nz] (fp64) += crop(field_s) * mask   (b2fwi_window_mask_accumulate)."""
        nx, nz = self.model.shape
        _lib.check(_lib.lib().b2fwi_window_mask_accumulate(nx, nz, _ptr(field_s), self.wshape[1], self.col0,
                                                           _ptr(mask), _ptr(out), _stream()))


def smoke_check(model, geom, d64, g64_full):
    """Used by __graft_entry__.smoke(): the resident engine on the smoke shot against the oracle's
    traces and (cropped) gradient."""
    import torch
    sv = ResidentSurvey(geom, [0])
    rec = sv.forward(save=True, illum=True)[0].cpu().numpy()
    res = torch.from_numpy(np.ascontiguousarray(d64, dtype=np.float32)[None]).cuda()
    grad = sv.crop(sv.gradient(res))[0].cpu().numpy()
    nbl = model.nbl
    g64 = g64_full[nbl:-nbl, nbl:-nbl]
    et = np.linalg.norm(rec - d64) / np.linalg.norm(d64)
    eg = np.linalg.norm(grad - g64) / np.linalg.norm(g64)
    print("smoke (resident engine, cluster=%d): traces rel-L2 %.2e, gradient rel-L2 %.2e" % (sv.plan.cluster, et, eg))
    assert et <= 1e-5 and eg <= 1e-4
