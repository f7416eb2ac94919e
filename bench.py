#!/usr/bin/env python
"""bench.py -- FWI-gradient throughput of the B200 path on the reference's Marmousi configuration.

Workload (BASELINE.json configs[2], marmousi_fwi.py): SMARMN 300x106 (+2*40 sponge = 380x186), h = 30 m,
space_order 8, dt = 2.95 ms, tn = 4000 ms (nt = 1357), 29 shots x 300 receivers, L2 misfit with
direct-wave subtraction, bathymetry mask and illumination preconditioning. One "step" = one evaluation of
the objective and its gradient over the whole survey, i.e. the reference's
    fwi_loss(x, geometry, obs, least_square, direct_wave, mask, precond=True)        (fwi.py:236-246)
Weak scaling: every rank (GPU) owns one full 29-shot survey (29*N shots in the job); the ranks'
[grad | illum | fval] are summed with ONE NCCL all-reduce per step, as fwi_obj_multi does.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-cpu] [--no-extra]
    torchrun --nproc-per-node N bench.py --gpus N ...

Prints ONE JSON line (rank 0). `value` = grid-point-steps per second of the whole job with observed data
already resident in HBM; `e2e` = the same through the public API with HOST buffers (H2D of the observed and
direct-wave records and of the model, D2H of (f, g)) inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHOTS_PER_RANK = 29
BYTES_FWD, BYTES_ADJ = 20, 32      # algorithmic bytes per grid-point-step (SURVEY.md section 8d / DESIGN.md)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons while the timed regions run: NVML every 10 ms (nvidia_ml_py), else nvidia-smi
    every 200 ms."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []              # (sm_mhz, [reason flags]) per sample
        self.sm_max = None
        self.source = None
        self.stop_flag = threading.Event()

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it holds plain indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.index
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            idx = int(vis.split(",")[self.index])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        masks = [0x8, 0x40, 0x20, 0x4]      # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap
        self.source = "nvml"
        while not self.stop_flag.is_set():
            r = int(get_reasons(h))
            self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), [bool(r & m) for m in masks]))
            self.stop_flag.wait(0.01)

    def _run_smi(self):
        self.source = "nvidia-smi"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    c = [x.strip() for x in out.split(",")]
                    self.sm_max = float(c[2])
                    self.rows.append((float(c[1]), [c[5 + k].lower().startswith("active") for k in range(4)]))
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for k, n in enumerate(self.NAMES) if any(r[1][k] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


def make_survey(world):
    """Marmousi geometries with SHOTS_PER_RANK * world shots (each rank gets the 29 positions of the
    reference's acquisition through the round-robin partition i % world == rank)."""
    from devito_fwi_b200 import configs
    from devito_fwi_b200.geometry import AcquisitionGeometry
    g_true, g_init, g_const, mask = configs.marmousi(nsrc=SHOTS_PER_RANK)
    if world > 1:
        src = np.repeat(g_true.src_positions, world, axis=0)
        mk = lambda g: AcquisitionGeometry(g.model, g.rec_positions, src, g.t0, g.tn, f0=g.f0,  # noqa: E731
                                           src_type=g.src_type)
        g_true, g_init, g_const = mk(g_true), mk(g_init), mk(g_const)
    return g_true, g_init, g_const, mask


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import devito_fwi_b200 as b
    from devito_fwi_b200 import fwi, _lib
    from devito_fwi_b200 import dist as bdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        bdist.init_from_env("nccl")
    lib = _lib.lib()

    g_true, g_init, g_const, mask = make_survey(world)
    model = g_init.model
    npts = int(np.prod(model.grid.shape))
    nt, nrec = g_init.nt, g_init.nrec
    steps_per_sweep = nt - 2
    nshots_job = g_init.nsrc
    my_shots = bdist.local_shots(nshots_job)

    # ---- set-up (untimed): observed and direct-wave data of this rank's shots, host + device copies
    from devito_fwi_b200.resident import ResidentSurvey
    obs, dw = [None] * nshots_job, [None] * nshots_job
    for geom, store in ((g_true, obs), (g_const, dw)):
        sv = ResidentSurvey(geom, my_shots)
        rec = sv.forward().clone()
        for k, i in enumerate(my_shots):
            r = b.Receiver(name='rec', grid=geom.grid, time_range=geom.time_axis, coordinates=geom.rec_positions)
            r.data[:] = rec[k].cpu().numpy()
            r._sdata.dev()
            store[i] = r
        del sv
    x0 = (1. / (model.vp.data[model.nbl:-model.nbl, model.nbl:-model.nbl].astype(np.float64) ** 2)).ravel()

    def step(host_buffers):
        if host_buffers:
            for i in my_shots:          # the caller hands HOST arrays: invalidate the device copies
                obs[i].data
                dw[i].data
        return fwi.fwi_loss(x0, g_init, obs, fwi.least_square, dw, mask, True, True)

    def timed(n, host_buffers):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.b2fwi_launch_count()
        e0.record()
        for _ in range(n):
            f, g, _ = step(host_buffers)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), lib.b2fwi_launch_count() - l0, (f, g)

    for _ in range(max(args.warmup, 3)):
        f_w, g_w, _ = step(False)
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches, (fval, grad) = timed(args.steps, host_buffers=False)
    for _ in range(2):
        step(True)
    ms_e2e, _, (f_e, g_e) = timed(args.steps, host_buffers=True)
    # same inputs every step: the engine is deterministic, so warm-up, device-resident and host-buffer steps must agree bit for bit
    repeatable = bool(f_w == fval and f_e == fval and np.array_equal(g_w, grad) and np.array_equal(g_e, grad))

    # ---- per-kernel roofline, measured live with CUDA events on the launching stream
    survey = fwi._resident_survey(g_init, my_shots)
    peak, peak_src = peaks()
    kern = {}
    if survey is not None:
        res = survey._res
        for name, fn, bpp in (("res2d_kernel<fwd>", lambda: survey.forward(save=True, illum=True), BYTES_FWD),
                              ("res2d_kernel<adj+img>", lambda: survey.gradient(res), BYTES_ADJ)):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            alg = float(bpp) * npts * steps_per_sweep * len(my_shots)
            kern[name] = {"ms_per_launch": round(ms, 4), "algorithmic_GB": round(alg / 1e9, 3),
                          "achieved": round(alg / ms / 1e6, 1), "gpts_per_s": round(alg / bpp / ms / 1e6, 1)}
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    work = 2.0 * npts * steps_per_sweep * nshots_job            # forward + adjoint grid-point-steps per step
    ms_step = ms_dev / args.steps
    ms_step_e2e = ms_e2e / args.steps
    out = {
        "metric": "FWI gradient throughput: fwd+adj stencil grid-point-steps per second, whole job",
        "value": round(work / ms_step / 1e6, 2), "unit": "Gpts/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "shots_per_s": round(nshots_job / ms_step * 1e3, 1),
        "bitwise_repeatable": repeatable, "fval": float(fval),
        "config": {"workload": "marmousi_fwi (BASELINE.json configs[2]): 380x186 padded, so=8, nt=1357, "
                               "%d shots/GPU x 300 rec, L2 + direct-wave + mask + illumination precond" % SHOTS_PER_RANK,
                   "shots_total": nshots_job, "engine": "resident2d" if survey is not None else "streaming",
                   "cluster_per_shot": int(survey.plan.cluster) if survey is not None else None,
                   "l2": "no explicit flush: each step streams a %.1f GB u.dt2 history through HBM (>> 126 MB L2)"
                         % (len(my_shots) * steps_per_sweep * 300 * 108 * 4 * 2 / 1e9),
                   "allreduce": "1 x NCCL sum of [grad|illum|fval] (%d doubles) per step" % (2 * 300 * 106 + 1)},
        "e2e": {"value": round(work / ms_step_e2e / 1e6, 2), "unit": "Gpts/s", "ms_per_step": round(ms_step_e2e, 3),
                "shots_per_s": round(nshots_job / ms_step_e2e * 1e3, 1),
                "h2d_bytes_per_step": int(len(my_shots) * 2 * nt * nrec * 4 + npts * 4),
                "d2h_bytes_per_step": int((2 * 300 * 106 + 1) * 8),
                "note": "fwi_loss() with host obs / direct-wave records (pinned) and host model, device copies "
                        "invalidated before every step; the H2D runs on a copy stream underneath the forward "
                        "sweep (queued first), so it is hidden when it takes less than the sweep; residual list "
                        "stays on the device until read (LazyResidual)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "objective": {"fval": float(fval), "grad_absmax": float(np.abs(grad).max())},
    }
    if kern:
        dom = max(kern, key=lambda k: kern[k]["ms_per_launch"])
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = int(json.load(f)[dom]["dram_bytes"])      # per launch, from the committed ncu capture
        except Exception:
            pass
        out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved"], "peak": peak,
                           "unit": "GB/s", "frac": round(kern[dom]["achieved"] / peak, 3), "traffic": traffic,
                           "peak_source": peak_src,
                           "note": "achieved = ALGORITHMIC bytes per launch (20 B fwd / 32 B adj+img per grid-point-step "
                                   "x points x steps x shots) / live CUDA-event duration. The 2-D wavefields are "
                                   "SM-resident (shared memory + registers): real DRAM traffic (`traffic`, ncu) is "
                                   "only the u.dt2 history, 17x below the algorithmic bytes, so HBM is not the binding "
                                   "roof of this kernel - instruction issue / barrier latency is (DESIGN.md 4.2)",
                           "kernels": kern}
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(sample_shots=8)
    if world == 1 and not args.no_extra:
        try:
            out["extra"] = extra_3d()
        except Exception as e:          # the headline line must survive a failure of the secondary workload
            out["extra"] = {"error": repr(e)[:200]}
        try:
            out["extra_2d"] = extra_2d()
        except Exception as e:
            out["extra_2d"] = {"error": repr(e)[:200]}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def cpu_gradient_shots(shot_ids, fast=True):
    """Oracle (CPU restatement, OpenMP) forward(save) + gradient of Marmousi shots; returns seconds."""
    from oracle import ref
    from devito_fwi_b200 import configs
    shape, spacing, nbl = (300, 106), (30., 30.), 40
    vp_true = configs.load_vp('SMARMN', 'vp.true', shape)
    vp_init = configs.load_vp('SMARMN', 'vp.smooth_20', shape)
    rm_true = ref.RefModel((0., 0.), spacing, shape, 8, vp_true, nbl=nbl, dt=2.95)
    rm_init = ref.RefModel((0., 0.), spacing, shape, 8, vp_init, nbl=nbl, dt=2.95)
    nt, _, tv = ref.time_axis(0., 4000., 2.95)
    wav = ref.ricker(0.007, tv).astype(np.float32)
    src = np.stack([np.linspace(0, 8970., SHOTS_PER_RANK), np.full(SHOTS_PER_RANK, 60.)], axis=1)
    rec = np.stack([np.linspace(30., 8940., 300), np.full(300, 60.)], axis=1)
    ref.lib(fast)
    obs = [ref.forward(rm_true, src[i], rec, wav, nt, 2.95, fast=fast)[0] for i in shot_ids]
    t0 = time.perf_counter()
    for k, i in enumerate(shot_ids):
        syn, u = ref.forward(rm_init, src[i], rec, wav, nt, 2.95, save=True, fast=fast)
        ref.gradient(rm_init, syn - obs[k], rec, u, nt, 2.95, fast=fast)
    return time.perf_counter() - t0, int(np.prod(rm_init.shape_pml)), nt


def _shot_worker(shot_id):
    sec, npts, nt = cpu_gradient_shots([shot_id])
    return sec, npts, nt


def _shot_worker_init():
    os.environ["OMP_NUM_THREADS"] = "1"


def cpu_shot_parallel(cores):
    """The same CPU code arranged the other way round: one single-threaded shot per core, `cores` shots at once
    (the reference runs its shots one after the other with OpenMP inside each - a 380x186 grid is too small for
    that to scale - so this is reported next to it, as the best the host can do with this code)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    pool = ctx.Pool(cores, initializer=_shot_worker_init)
    try:                                                            # bounded: a stuck worker must not stall the bench
        pool.map_async(_shot_worker, list(range(cores))).get(timeout=180)      # library load, first touch
        t0 = time.perf_counter()
        out = pool.map_async(_shot_worker, [i % SHOTS_PER_RANK for i in range(cores)]).get(timeout=180)
        sec = time.perf_counter() - t0
    finally:
        pool.terminate()
    npts, nt = out[0][1], out[0][2]
    return {"value": round(2.0 * npts * (nt - 2) * cores / sec / 1e9, 3), "unit": "Gpts/s",
            "shots_per_s": round(cores / sec, 3), "sample": "%d shots at once, one OpenMP thread each" % cores}


def cpu_baseline(sample_shots=2, shot_parallel=True):
    cores = os.cpu_count() or 1
    sec, npts, nt = cpu_gradient_shots(list(range(sample_shots)))     # includes first-touch warm-up
    sec, npts, nt = cpu_gradient_shots(list(range(sample_shots)))
    work = 2.0 * npts * (nt - 2) * sample_shots
    out = {"value": round(work / sec / 1e9, 3), "unit": "Gpts/s", "cores": cores, "kind": "port",
           "shots_per_s": round(sample_shots / sec, 3),
           "sample": "%d Marmousi shot-gradients (forward with saved history + adjoint/imaging), "
                     "oracle/fwi_oracle.c built -O3 -march=native -ffast-math -fopenmp (Devito's flag set), "
                     "%d OpenMP threads inside each shot, shots one after the other as the reference runs them; "
                     "CPU restatement, not Devito" % (sample_shots, cores)}
    if shot_parallel:
        try:
            out["shot_parallel"] = cpu_shot_parallel(cores)
        except Exception as e:
            out["shot_parallel"] = {"error": repr(e)[:160]}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. Devito itself is neither vendored
    nor installable offline, so this times the oracle port with all host threads (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    sample = 8
    cpu_gradient_shots([0])        # warm-up: build the library, touch memory
    for _ in range(max(args.warmup - 1, 0)):
        cpu_gradient_shots(list(range(sample)))
    t = 0.0
    for _ in range(args.steps):
        sec, npts, nt = cpu_gradient_shots(list(range(sample)))
        t += sec
    work = 2.0 * npts * (nt - 2) * sample * args.steps
    v = round(work / t / 1e9, 3)
    out = {"impl": "reference",
           "metric": "FWI gradient throughput: fwd+adj stencil grid-point-steps per second, whole job",
           "value": v, "unit": "Gpts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(t / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "shots_per_s": round(sample * args.steps / t, 3),
           "config": {"workload": "marmousi_fwi (BASELINE.json configs[2]): 380x186 padded, so=8, nt=1357, "
                                  "%d shots/GPU x 300 rec" % SHOTS_PER_RANK,
                      "sample_shots_per_step": sample},
           "cpu_baseline": {"value": v, "unit": "Gpts/s", "cores": cores, "kind": "port",
                            "sample": "each step = %d Marmousi shot-gradients of the 29-shot survey on %d OpenMP "
                                      "threads (oracle port, Devito flag set); Devito is not installable here"
                                      % (sample, cores)},
           "e2e": {"value": v, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:        # for transparency: the same code with one single-threaded shot per core (not how the reference runs)
        out["cpu_baseline"]["shot_parallel"] = cpu_shot_parallel(cores)
    except Exception as e:
        out["cpu_baseline"]["shot_parallel"] = {"error": repr(e)[:160]}
    print(json.dumps(out), flush=True)


def extra_2d():
    """The other named 2-D configurations (BASELINE.json configs[0], [1], [3]) through the public API, CUDA-event
    timed after a warm-up evaluation: circle (so=6, 11 shots) and Marmousi2 (31 shots) objective + gradient with
    the L2 misfit, marmousi_fm forward modelling of 21 shots. Secondary numbers; the headline is configs[2]."""
    import torch
    from devito_fwi_b200 import fwi, configs
    out = {}

    def ev_time(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def objective(name, g_true, g_init, dw_geom, mask):
        model = g_init.model
        obs = fwi.fm_multi(g_true)
        dw = fwi.fm_multi(dw_geom) if dw_geom is not None else None
        nbl = model.nbl
        x0 = (1. / (model.vp.data[nbl:-nbl, nbl:-nbl].astype(np.float64) ** 2)).ravel()
        ms = ev_time(lambda: fwi.fwi_loss(x0, g_init, obs, fwi.least_square, dw, mask, True, True))
        work = 2.0 * np.prod(model.grid.shape) * (g_init.nt - 2) * g_init.nsrc
        svs = fwi._resident_surveys(g_init, list(range(g_init.nsrc)))
        out[name] = {"grid": list(model.grid.shape), "so": model.space_order, "nt": g_init.nt, "shots": g_init.nsrc,
                     "ms_per_objective_gradient": round(ms, 3), "gpts_per_s": round(work / ms / 1e6, 1),
                     "shots_per_s": round(g_init.nsrc / ms * 1e3, 1),
                     "launch_groups_shots_x_cluster": [[sv.nshots, int(sv.plan.cluster)] for sv in svs] if svs else None}
        fwi._SURVEYS.clear()

    g_true, g_init = configs.circle()                          # circle_fwi.py:65 uses space_order 6 ...
    objective("circle_fwi", g_true, g_init, None, None)
    g_true, g_init = configs.circle(space_order=4)             # ... BASELINE.json's text says 4: both are reported
    objective("circle_fwi_so4", g_true, g_init, None, None)
    g_true, g_init, g_const, mask = configs.marmousi2()
    objective("marmousi2_fwi_L2", g_true, g_init, g_const, mask)
    g_true, g_init, g_const, _ = configs.marmousi(nsrc=21, tn=4500.)
    ms = ev_time(lambda: [fwi.fm_multi(g) for g in (g_true, g_init, g_const)])
    model = g_true.model
    work = 3.0 * np.prod(model.grid.shape) * (g_true.nt - 2) * g_true.nsrc
    out["marmousi_fm"] = {"grid": list(model.grid.shape), "so": model.space_order, "nt": g_true.nt, "shots": g_true.nsrc,
                          "models": 3, "ms_all_shots_3_models": round(ms, 3), "gpts_per_s": round(work / ms / 1e6, 1)}
    fwi._SURVEYS.clear()
    torch.cuda.empty_cache()
    return out


def extra_3d():
    """Secondary workload (BASELINE.json configs[4]): 3-D layered 512^3 (+2*40 = 592^3), so=8, one shot on the
    streaming engine (TMA-staged kernels): forward sweep that also lays down the on-device checkpoints, then the
    gradient pass (recompute with u.dt2 store + adjoint/imaging), against the HBM roofline. tn is shortened to
    300 ms (167 time levels) to keep the default run short; per-step cost does not depend on nt."""
    import torch
    import devito_fwi_b200 as b
    from devito_fwi_b200 import configs
    peak, _ = peaks()
    geom = configs.layered3d(n=512, space_order=8, tn=300., rec_decimate=4)
    model = geom.model
    npts = int(np.prod(model.grid.shape))
    solver = b.AcousticWaveSolver(model, geom, space_order=8)
    solver.forward(time_M=8)                                    # warm-up
    _, _, s_r = solver.forward()                                # plain forward modelling (ring buffer)
    res = b.Receiver(name='res', grid=model.grid, time_range=geom.time_axis, coordinates=geom.rec_positions)
    for rep in range(2):        # first pass warms the allocator (tens of GB of checkpoint / u.dt2 buffers), second is reported
        rec, cw, s_f = solver.forward(save='checkpoint')        # forward + checkpoints (pass 1)
        res._sdata.adopt_dev(rec._sdata.dev().clone())
        _, s_g = solver.gradient(rec=res, u=cw)                 # pass 2
        if rep == 0:
            del cw
    steps = geom.nt - 2
    t_shot = s_f.time + s_g.time
    out = {"workload": "layered3d 592^3 (512^3 + 2*40), so=8, nt=%d, %d receivers, 1 shot, streaming engine (TMA)" % (geom.nt, geom.nrec),
           "forward": {"ms_per_step": round(s_r.time / steps * 1e3, 4), "gpts_per_s": round(s_r.gpointss, 1),
                       "achieved_GBs_20B": round(s_r.gbytess, 1), "frac_of_measured_hbm": round(s_r.gbytess / peak, 3)},
           "shot_gradient": {"s": round(t_shot, 4),
                             "sweeps": "forward(+checkpoints) + recompute(+u.dt2 store) + adjoint/imaging",
                             "forward_s": round(s_f.time, 4), "recompute_adjoint_s": round(s_g.time, 4),
                             "achieved_GBs_52B_algorithmic": round(52.0 * npts * steps / t_shot / 1e9, 1),
                             "frac_of_measured_hbm_52B": round(52.0 * npts * steps / t_shot / 1e9 / peak, 3),
                             "GBs_with_recompute_76B": round(76.0 * npts * steps / t_shot / 1e9, 1)},
           "note": "fractions use ALGORITHMIC bytes (20 / 32 B per point-step); the kernels skip the c1 read inside "
                   "the undamped interior, so they move fewer bytes than that (profiles/r01_3d_tma_launches.txt: "
                   "18.7 / 22.7 / 31.0 B per point at 6.5 TB/s of DRAM traffic)",
           "hbm_peak_alloc_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
    del solver, rec, res, cw
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-only", action="store_true", help="print the cpu_baseline object alone (no GPU needed)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary 3-D workload (N=1 only)")
    args = ap.parse_args()
    import warnings
    warnings.filterwarnings("ignore")
    if args.cpu_only:
        print(json.dumps(cpu_baseline(sample_shots=8)), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
