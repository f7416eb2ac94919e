#!/usr/bin/env python
"""bench.py -- FWI-gradient throughput of the B200 path on the reference's Marmousi configuration.

Workload (BASELINE.json configs[2], marmousi_fwi.py): SMARMN 300x106 (+2*40 sponge = 380x186), h = 30 m,
space_order 8, dt = 2.95 ms, tn = 4000 ms (nt = 1357), 29 shots x 300 receivers, L2 misfit with
direct-wave subtraction, bathymetry mask and illumination preconditioning. One "step" = one evaluation of
the objective and its gradient over the whole survey, i.e. the reference's
    fwi_loss(x, geometry, obs, least_square, direct_wave, mask, precond=True)        (fwi.py:236-246)
Weak scaling (the headline `value`): every rank (GPU) owns one full 29-shot survey (29*N shots in the job); the
ranks' [grad | illum | fval] are summed with ONE NCCL all-reduce per step, as fwi_obj_multi does.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload marmousi|layered3d]
                    [--no-cpu] [--no-extra] [--no-strong] [--no-3d]
    torchrun --nproc-per-node N bench.py --gpus N ...

Prints ONE JSON line (rank 0). `value` = grid-point-steps per second of the whole job with observed data
already resident in HBM; `e2e` = the same through the public API with HOST buffers (H2D of the observed and
direct-wave records and of the model, D2H of (f, g)) inside the timed region. The same line carries, at every N:
  `strong`    the NAMED 29-shot Marmousi survey (and a 32-shot variant) sharded over the N ranks (i % N == rank);
  `layered3d` BASELINE.json configs[4]: 592^3 (512^3 + 2*40), so=8, the full nt=690 time axis, L2 gradient with
              on-device checkpointing -- one shot per GPU (weak) and 8 shots over N GPUs (strong), with its own
              roofline (52 algorithmic B per point-step), e2e (host records in, (f, g) out) and, at N=1, the CPU
              oracle timed on a few steps of the same grid.
`--workload layered3d` prints the 3-D workload as the top-level line instead.
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
MARMOUSI_WORKLOAD = ("marmousi_fwi (BASELINE.json configs[2]): 380x186 padded, so=8, nt=1357, %d shots/GPU x 300 rec, "
                     "L2 + direct-wave + mask + illumination precond")


def _events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _max_over_ranks(ms, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class MarmousiJob(object):
    """One Marmousi FWI job: geometry of `nshots_job` shots, observed / direct-wave records of this rank's shots
    (host + device copies), and the timed objective evaluation."""

    def __init__(self, g_true, g_init, g_const, mask):
        import devito_fwi_b200 as b
        from devito_fwi_b200 import dist as bdist
        from devito_fwi_b200.resident import ResidentSurvey
        self.g_init, self.mask = g_init, mask
        model = g_init.model
        self.npts = int(np.prod(model.grid.shape))
        self.nt, self.nrec = g_init.nt, g_init.nrec
        self.nshots_job = g_init.nsrc
        self.my_shots = bdist.local_shots(self.nshots_job)
        self.obs, self.dw = [None] * self.nshots_job, [None] * self.nshots_job
        for geom, store in ((g_true, self.obs), (g_const, self.dw)):
            if not self.my_shots:
                continue
            sv = ResidentSurvey(geom, self.my_shots)
            rec = sv.forward().clone()
            for k, i in enumerate(self.my_shots):
                r = b.Receiver(name='rec', grid=geom.grid, time_range=geom.time_axis, coordinates=geom.rec_positions)
                r.data[:] = rec[k].cpu().numpy()
                r._sdata.dev()
                store[i] = r
            del sv
        nbl = model.nbl
        self.x0 = (1. / (model.vp.data[nbl:-nbl, nbl:-nbl].astype(np.float64) ** 2)).ravel()

    def step(self, host_buffers):
        from devito_fwi_b200 import fwi
        if host_buffers:
            for i in self.my_shots:          # the caller hands HOST arrays: invalidate the device copies
                self.obs[i].data
                self.dw[i].data
        return fwi.fwi_loss(self.x0, self.g_init, self.obs, fwi.least_square, self.dw, self.mask, True, True)

    def timed(self, n, host_buffers, world, lib):
        import torch
        import torch.distributed as dist
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = _events()
        l0 = lib.b2fwi_launch_count()
        e0.record()
        for _ in range(n):
            f, g, _ = self.step(host_buffers)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return _max_over_ranks(e0.elapsed_time(e1), world), lib.b2fwi_launch_count() - l0, (f, g)

    @property
    def work(self):                      # forward + adjoint grid-point-steps per objective evaluation, whole job
        return 2.0 * self.npts * (self.nt - 2) * self.nshots_job

    def bytes_per_step(self):
        return (int(len(self.my_shots) * 2 * self.nt * self.nrec * 4 + self.npts * 4), int((2 * 300 * 106 + 1) * 8))


def strong_scaling(args, world, lib):
    """The named survey as ONE job sharded over the ranks: 29 shots (marmousi_fwi.py:94-117), and a 32-shot
    variant that divides evenly by 8."""
    from devito_fwi_b200 import configs, fwi
    from devito_fwi_b200 import dist as bdist
    out = {}
    for nshots in (29, 32):
        fwi._SURVEYS.clear()
        g_true, g_init, g_const, mask = configs.marmousi(nsrc=nshots)
        job = MarmousiJob(g_true, g_init, g_const, mask)
        for _ in range(3):
            job.step(False)
        ms, _, (fval, _) = job.timed(args.steps, False, world, lib)
        for _ in range(2):
            job.step(True)
        ms_e, _, _ = job.timed(args.steps, True, world, lib)
        svs = fwi._resident_surveys(g_init, job.my_shots) if job.my_shots else None
        ms, ms_e = ms / args.steps, ms_e / args.steps
        out["marmousi_%d_shots" % nshots] = {
            "shots_total": nshots, "shots_per_rank": [len(bdist.local_shots(nshots, r, world)) for r in range(world)],
            "ms_per_step": round(ms, 3), "shots_per_s": round(nshots / ms * 1e3, 1),
            "gpts_per_s": round(job.work / ms / 1e6, 1),
            "e2e_ms_per_step": round(ms_e, 3), "e2e_gpts_per_s": round(job.work / ms_e / 1e6, 1),
            "rank0_launch_groups_shots_x_cluster_x_rows": [[sv.nshots, int(sv.plan.cluster), int(sv.plan.rows_per_thread)]
                                                           for sv in svs] if svs else None,
            "fval": float(fval)}
        del job
    fwi._SURVEYS.clear()
    out["note"] = ("strong scaling: the job is fixed (29 / 32 shots), shots i % N == rank, one NCCL all-reduce of "
                   "[grad|illum|fval] per step; with few shots per GPU a shot runs on a cluster of up to 16 SMs "
                   "(resident2d_lat.cuh). The N=1 entry of the 29-shot survey is the headline workload itself.")
    return out


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from devito_fwi_b200 import fwi, _lib
    from devito_fwi_b200 import dist as bdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        bdist.init_from_env("nccl")
    lib = _lib.lib()
    if args.workload == "layered3d":
        out = layered3d(args, world, rank, local, lib, headline=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            print(json.dumps(out), flush=True)
        return

    job = MarmousiJob(*make_survey(world))
    g_init, my_shots, npts = job.g_init, job.my_shots, job.npts
    nt, steps_per_sweep, nshots_job = job.nt, job.nt - 2, job.nshots_job

    for _ in range(max(args.warmup, 3)):
        f_w, g_w, _ = job.step(False)
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches, (fval, grad) = job.timed(args.steps, False, world, lib)
    for _ in range(2):
        job.step(True)
    ms_e2e, _, (f_e, g_e) = job.timed(args.steps, True, world, lib)
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
            e0, e1 = _events()
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
    plan = survey.plan if survey is not None else None
    bytes_in, bytes_out = job.bytes_per_step()
    work = job.work
    del job, survey
    fwi._SURVEYS.clear()
    torch.cuda.empty_cache()

    strong = None
    if not args.no_strong:
        try:
            strong = strong_scaling(args, world, lib)
        except Exception as e:
            strong = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    l3d = None
    if not args.no_3d:
        try:
            l3d = layered3d(args, world, rank, local, lib, headline=False)
        except Exception as e:
            l3d = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    ms_step = ms_dev / args.steps
    ms_step_e2e = ms_e2e / args.steps
    out = {
        "metric": "FWI gradient throughput: fwd+adj stencil grid-point-steps per second, whole job",
        "value": round(work / ms_step / 1e6, 2), "unit": "Gpts/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "shots_per_s": round(nshots_job / ms_step * 1e3, 1),
        "bitwise_repeatable": repeatable, "fval": float(fval),
        "config": {"workload": MARMOUSI_WORKLOAD % SHOTS_PER_RANK,
                   "shots_total": nshots_job, "engine": "resident2d" if plan is not None else "streaming",
                   "cluster_per_shot": int(plan.cluster) if plan is not None else None,
                   "rows_per_thread": int(plan.rows_per_thread) if plan is not None else None,
                   "l2": "no explicit flush: each step streams a %.1f GB u.dt2 history through HBM (>> 126 MB L2)"
                         % (len(my_shots) * steps_per_sweep * 300 * 108 * 4 * 2 / 1e9),
                   "allreduce": "1 x NCCL sum of [grad|illum|fval] (%d doubles) per step" % (2 * 300 * 106 + 1)},
        "e2e": {"value": round(work / ms_step_e2e / 1e6, 2), "unit": "Gpts/s", "ms_per_step": round(ms_step_e2e, 3),
                "shots_per_s": round(nshots_job / ms_step_e2e * 1e3, 1),
                "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
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
        prof = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                prof = json.load(f)[dom]                        # per launch, from the committed ncu capture
        except Exception:
            pass
        out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved"], "peak": peak,
                           "unit": "GB/s", "frac": round(kern[dom]["achieved"] / peak, 3),
                           "traffic": int(prof["dram_bytes"]) if "dram_bytes" in prof else None,
                           "peak_source": peak_src,
                           "binding_roof": "issue",
                           "issue": {"bound": "issue", "unit": "share of the SMs' warp-issue slots used (ncu "
                                     "smsp__issue_active, committed capture)", "achieved": prof.get("issue_active_pct"),
                                     "peak": 100.0,
                                     "frac": round(prof["issue_active_pct"] / 100.0, 3) if "issue_active_pct" in prof else None},
                           "note": "the contract's HBM fraction: achieved = ALGORITHMIC bytes per launch (20 B fwd / 32 B "
                                   "adj+img per grid-point-step x points x steps x shots) / live CUDA-event duration. The "
                                   "2-D wavefields are SM-resident (shared memory + registers): real DRAM traffic (`traffic`, "
                                   "ncu) is only the u.dt2 history, 17x below the algorithmic bytes, so this kernel is NOT "
                                   "HBM-bound and the fraction can exceed 1; its binding roof is instruction issue / FP32 "
                                   "(`issue`, DESIGN.md 4.2). The HBM-bound kernels of this repository are the 3-D sweeps: "
                                   "see layered3d.roofline",
                           "kernels": kern}
    if strong is not None:
        out["strong"] = strong
    if l3d is not None:
        out["layered3d"] = l3d
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline()
    if world == 1 and not args.no_extra:
        try:
            out["extra_2d"] = extra_2d()
        except Exception as e:
            out["extra_2d"] = {"error": repr(e)[:200]}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def layered3d(args, world, rank, local, lib, headline):
    """BASELINE.json configs[4] (acoustic_example.py:26-63,96 + preset_models.py:109-126): 592^3 padded, so=8, full
    time axis (tn=1250 -> nt=690), 128x128 receivers (every 4th of the 512^2 grid of utils.py:12-47), L2 gradient
    with on-device checkpoints (fwi.StreamingSurvey). Weak: one shot per GPU; strong: 8 shots over the N GPUs."""
    import torch
    import torch.distributed as dist
    import devito_fwi_b200 as b
    from devito_fwi_b200 import configs, fwi
    from devito_fwi_b200 import dist as bdist
    peak, peak_src = peaks()
    geom0 = configs.layered3d(n=512, space_order=8, tn=args.tn3d, rec_decimate=4)
    model = geom0.model
    npts = int(np.prod(model.grid.shape))
    steps = geom0.nt - 2
    vp_true = b.Function(name='vp_true', grid=model.grid)        # "true" model: the layering + a fast block under the spread
    v = np.array(model.vp.data)
    nbl, n = model.nbl, 512
    lo, hi = nbl + 3 * n // 8, nbl + 5 * n // 8
    v[lo:hi, lo:hi, nbl + 10:nbl + 40] *= 1.1
    vp_true.data[...] = v
    del v

    def survey_of(nshots):
        centre = np.array(geom0.src_positions[0], dtype=np.float64)
        src = np.repeat(centre[None], nshots, axis=0)
        span = 0.5 * model.domain_size[0]
        src[:, 0] = centre[0] + (np.linspace(-0.5, 0.5, nshots) * span if nshots > 1 else 0.0)
        return b.AcquisitionGeometry(model, geom0.rec_positions, src, geom0.t0, geom0.tn, f0=geom0.f0, src_type='Ricker')

    def measure(nshots, reps, with_e2e, warm=True):
        geom = survey_of(nshots)
        sv = fwi.StreamingSurvey(geom)
        obs = sv.forward(vp=vp_true)                         # untimed set-up: observed data, device resident
        host_obs = {i: np.array(r.data) for i, r in obs.items()} if with_e2e else None
        if with_e2e:
            sv.host_buffer()                                 # pinned (f, g) mirror: allocated outside the timed region
        for r in obs.values():
            r._sdata.dev()
        if warm:
            sv.objective(obs, host=False)                    # warm-up (allocations of ~150 GB of checkpoints / kept wavefield)

        def timed(host):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = _events()
            l0 = lib.b2fwi_launch_count()
            e0.record()
            for _ in range(reps):
                if host:
                    for i, r in obs.items():                 # the caller hands HOST records: upload them again
                        r.data[:] = host_obs[i]
                f, g = sv.objective(obs, host=host)
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            return _max_over_ranks(e0.elapsed_time(e1), world) / reps / 1e3, lib.b2fwi_launch_count() - l0, f, g

        sec, launches, f, g = timed(False)
        res = {"shots_total": nshots, "shots_per_rank": [len(bdist.local_shots(nshots, r, world)) for r in range(world)],
               "s_per_step": round(sec, 4), "shots_per_s": round(nshots / sec, 4),
               "gpts_per_s_fwd_adj": round(2.0 * npts * steps * nshots / sec / 1e9, 2),
               "algorithmic_GBs_52B": round(52.0 * npts * steps * nshots / sec / 1e9, 1),
               "frac_of_measured_hbm_per_gpu": round(52.0 * npts * steps * max(len(sv.shots), 1) / sec / 1e9 / peak, 3),
               "gpu_launches": int(launches), "fval": float(f), "grad_absmax": float(sv.buf[:-1].abs().max())}
        if with_e2e:
            sec_e, _, f_e, g_e = timed(True)
            res["e2e"] = {"s_per_step": round(sec_e, 4), "shots_per_s": round(nshots / sec_e, 4),
                          "gpts_per_s_fwd_adj": round(2.0 * npts * steps * nshots / sec_e / 1e9, 2),
                          "h2d_bytes_per_step": int(len(sv.shots) * geom.nt * geom.nrec * 4),
                          "d2h_bytes_per_step": int(sv.buf.numel() * 4),
                          "fval_matches_device_run": bool(abs(f_e - float(f)) <= 1e-6 * abs(f_e))}
        hbm = round(torch.cuda.max_memory_allocated() / 1e9, 1)
        del sv, obs
        return res, hbm

    # per-sweep roofline, live (one shot, rank-local): plain forward, and the gradient call (recompute + adjoint/imaging)
    solver = b.AcousticWaveSolver(model, fwi._shot_geometry(survey_of(1), 0), space_order=8)
    solver.forward(time_M=8)
    _, _, s_fwd = solver.forward()
    del solver
    torch.cuda.empty_cache()
    weak, hbm_gb = measure(world, 1, True)
    strong, _ = measure(8, 1, False, warm=False)      # the allocator is warm: same buffer sizes as the weak run
    roof = {"bound": "hbm", "kernel": "step3d_tma_kernel (forward sweep, 20 B per point-step)",
            "achieved": round(s_fwd.gbytess, 1), "peak": peak, "unit": "GB/s", "frac": round(s_fwd.gbytess / peak, 3),
            "peak_source": peak_src, "traffic": traffic_3d(),
            "ms_per_launch": round(s_fwd.time / steps * 1e3, 4),
            "shot_gradient": {"algorithmic_B_per_point_step": 52, "achieved": weak["algorithmic_GBs_52B"] / world,
                              "frac": weak["frac_of_measured_hbm_per_gpu"],
                              "note": "forward (18.7 B per point-step moved: c1 is not read inside the undamped box) + "
                                      "recompute of the steps whose wavefield was not kept from pass 1 (18.7 B x ~85 %) + "
                                      "adjoint with imaging by parts from ONE stored wavefield value (31 B) = ~66 B per "
                                      "point-step actually streamed for 52 algorithmic; the recompute sweep is the price "
                                      "of checkpointing (checkpoint.py)"}}
    out = {"workload": "layered3d (BASELINE.json configs[4]): 592^3 = 512^3 + 2*40, so=8, nt=%d (tn=%g), %d receivers, "
                       "L2 gradient with on-device checkpointing, streaming engine (TMA)" % (geom0.nt, args.tn3d, geom0.nrec),
           "weak_one_shot_per_gpu": weak, "strong_8_shots": strong, "roofline": roof, "hbm_peak_alloc_GB": hbm_gb}
    if world == 1 and not args.no_cpu and rank == 0:
        try:
            out["cpu_baseline"] = cpu_baseline_3d()
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)[:200]}
    if not headline:
        return out
    sec = weak["s_per_step"]
    top = {"metric": "FWI gradient throughput: fwd+adj stencil grid-point-steps per second, whole job",
           "value": weak["gpts_per_s_fwd_adj"], "unit": "Gpts/s", "n_gpus": world, "steps": 1, "warmup": 1,
           "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "shots_per_s": weak["shots_per_s"],
           "config": {"workload": out["workload"], "shots_total": world,
                      "l2": "inputs larger than L2: every sweep streams 830 MB slices (>> 126 MB L2)"},
           "e2e": {"value": weak["e2e"]["gpts_per_s_fwd_adj"], "unit": "Gpts/s",
                   "h2d_bytes_per_step": weak["e2e"]["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": weak["e2e"]["d2h_bytes_per_step"]},
           "gpu_launches": weak["gpu_launches"], "roofline": roof, "layered3d": out}
    if "cpu_baseline" in out:
        top["cpu_baseline"] = out["cpu_baseline"]
    return top


def traffic_3d():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return int(json.load(f)["step3d_tma_kernel<fwd>"]["dram_bytes"])
    except Exception:
        return None


def cpu_baseline_3d(nsteps=3):
    """The CPU oracle (fast build, all host threads) on the 592^3 grid itself: `nsteps` forward steps with saved
    history and the adjoint + imaging steps over them; Gpts/s extrapolates to any nt (cost per step is constant)."""
    from oracle import ref
    threads = set_omp_threads()
    n, nbl = 512, 40
    vp = np.empty((n, n, n), dtype=np.float32)
    vp[..., :170] = 1.5
    vp[..., 170:340] = 2.5
    vp[..., 340:] = 3.5
    rm = ref.RefModel((0., 0., 0.), (15., 15., 15.), (n, n, n), 8, vp, nbl=nbl)
    del vp
    dt = float(rm.critical_dt)
    nt = nsteps + 2
    tv = np.arange(nt) * dt
    wav = ref.ricker(0.010, tv).astype(np.float32)
    src = np.array([[3832.5, 3832.5, 15.0]])
    g = np.linspace(0., 7665., 32)
    rx, ry = np.meshgrid(g, g, indexing='ij')
    rec = np.stack([rx.ravel(), ry.ravel(), np.full(rx.size, 30.0)], axis=1)
    ref.lib(True)
    u = np.zeros((nt,) + rm.shape_pml, dtype=np.float32)
    t0 = time.perf_counter()
    d, u = ref.forward(rm, src, rec, wav, nt, dt, save=True, u=u, fast=True)
    t1 = time.perf_counter()
    ref.gradient(rm, d, rec, u, nt, dt, fast=True)
    t2 = time.perf_counter()
    npts = int(np.prod(rm.shape_pml))
    return {"value": round(2.0 * npts * nsteps / (t2 - t0) / 1e9, 3), "unit": "Gpts/s", "cores": threads, "kind": "port",
            "forward_s_per_step": round((t1 - t0) / nsteps, 3), "adjoint_s_per_step": round((t2 - t1) / nsteps, 3),
            "shot_gradient_s_extrapolated_nt690": round((t2 - t0) / nsteps * 688, 1),
            "sample": "%d forward steps (saved history, includes the first touch of the arrays) + %d adjoint/imaging steps on "
                      "the 592^3 grid, oracle/fwi_oracle.c -O3 -march=native -ffast-math -fopenmp on %d OpenMP threads; "
                      "CPU restatement, not Devito" % (nsteps, nsteps, threads)}


# ------------------------------------------------------------------------------------------------
def set_omp_threads(n=None):
    """Make the oracle's OpenMP runtime use `n` threads (default: every host core) whatever the launcher put into
    OMP_NUM_THREADS (torchrun sets it to 1), and return the number the runtime then reports."""
    import ctypes
    n = int(n or os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    from oracle import ref
    L = ref.lib(True)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(n)
        gomp.omp_get_max_threads.restype = ctypes.c_int
        got = int(gomp.omp_get_max_threads())
    except Exception:
        got = n
    del L
    return got


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


class ShotPool(object):
    """The CPU port arranged for throughput: one single-threaded shot per core (the reference itself runs its shots
    one after the other with OpenMP inside each - a 380x186 grid is too small for that to scale)."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_shot_worker_init)
        self.pool.map_async(_shot_worker, list(range(cores))).get(timeout=300)      # library load, first touch

    def survey(self, nshots=SHOTS_PER_RANK):
        """All `nshots` shot-gradients of the survey; returns (seconds, npts, nt)."""
        t0 = time.perf_counter()
        out = self.pool.map_async(_shot_worker, list(range(nshots)), chunksize=1).get(timeout=600)
        return time.perf_counter() - t0, out[0][1], out[0][2]

    def close(self):
        self.pool.terminate()


def cpu_openmp_inside(sample_shots):
    """Reference-style arrangement: shots one after the other, every OpenMP thread inside each shot."""
    sec, npts, nt = cpu_gradient_shots(list(range(sample_shots)))     # first call includes first-touch warm-up
    sec, npts, nt = cpu_gradient_shots(list(range(sample_shots)))
    return {"value": round(2.0 * npts * (nt - 2) * sample_shots / sec / 1e9, 3), "unit": "Gpts/s",
            "shots_per_s": round(sample_shots / sec, 3),
            "sample": "%d shots one after the other, all OpenMP threads inside each (how the reference runs them)"
                      % sample_shots}


def cpu_baseline():
    """cpu_baseline leg: the oracle port (Devito's flag set) on the box's host cores, both arrangements; `value`
    is the better one, measured on the WHOLE 29-shot survey."""
    cores = os.cpu_count() or 1
    threads = set_omp_threads(cores)
    inside = cpu_openmp_inside(4)
    out = {"unit": "Gpts/s", "cores": cores, "omp_threads": threads, "kind": "port", "openmp_inside_shot": inside}
    try:
        pool = ShotPool(cores)
        try:
            sec, npts, nt = pool.survey()
        finally:
            pool.close()
        par = {"value": round(2.0 * npts * (nt - 2) * SHOTS_PER_RANK / sec / 1e9, 3), "unit": "Gpts/s",
               "shots_per_s": round(SHOTS_PER_RANK / sec, 3),
               "sample": "the whole %d-shot survey, one single-threaded shot per core, %d at a time" % (SHOTS_PER_RANK, cores)}
    except Exception as e:
        par = {"error": repr(e)[:160], "value": 0.0}
    out["shot_parallel"] = par
    best = par if par.get("value", 0.0) >= inside["value"] else inside
    out["value"] = best["value"]
    out["shots_per_s"] = best["shots_per_s"]
    out["sample"] = ("Marmousi shot-gradients (forward with saved history + adjoint/imaging) by oracle/fwi_oracle.c built "
                     "-O3 -march=native -ffast-math -fopenmp (Devito's flag set); value = the better of the two arrangements: "
                     + best["sample"] + "; CPU restatement, not Devito")
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. Devito itself is neither vendored
    nor installable offline, so this times the oracle port with all host threads (kind = "port"): every step is the
    WHOLE 29-shot survey (same config as the GPU arm's per-GPU workload), in the arrangement that is fastest on the
    host - one single-threaded shot per core; the reference-style arrangement (OpenMP inside each shot) is reported
    next to it. The launcher's OMP_NUM_THREADS (torchrun forces 1) is overridden: see set_omp_threads()."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    threads = set_omp_threads(cores)
    inside = cpu_openmp_inside(4)
    pool = ShotPool(cores)
    try:
        for _ in range(max(args.warmup - 1, 0)):
            pool.survey()
        t = 0.0
        for _ in range(args.steps):
            sec, npts, nt = pool.survey()
            t += sec
    finally:
        pool.close()
    work = 2.0 * npts * (nt - 2) * SHOTS_PER_RANK * args.steps
    v_par = work / t / 1e9
    v = round(max(v_par, inside["value"]), 3)
    out = {"impl": "reference",
           "metric": "FWI gradient throughput: fwd+adj stencil grid-point-steps per second, whole job",
           "value": v, "unit": "Gpts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(t / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "shots_per_s": round(SHOTS_PER_RANK * args.steps / t, 3),
           "config": {"workload": MARMOUSI_WORKLOAD % SHOTS_PER_RANK, "sample_shots_per_step": SHOTS_PER_RANK},
           "cpu_baseline": {"value": v, "unit": "Gpts/s", "cores": cores, "omp_threads": threads, "kind": "port",
                            "shot_parallel": {"value": round(v_par, 3), "unit": "Gpts/s"},
                            "openmp_inside_shot": inside,
                            "sample": "each step = all %d Marmousi shot-gradients of the survey, one single-threaded shot "
                                      "per core on %d cores (oracle port, Devito flag set: -O3 -march=native -ffast-math "
                                      "-fopenmp); the reference-style arrangement is under openmp_inside_shot (%d OpenMP "
                                      "threads reported by the runtime); Devito is not installable here"
                                      % (SHOTS_PER_RANK, cores, threads)},
           "e2e": {"value": v, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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

    def objective(name, g_true, g_init, dw_geom, mask, misfit=None):
        misfit = misfit or fwi.least_square
        model = g_init.model
        obs = fwi.fm_multi(g_true)
        dw = fwi.fm_multi(dw_geom) if dw_geom is not None else None
        nbl = model.nbl
        x0 = (1. / (model.vp.data[nbl:-nbl, nbl:-nbl].astype(np.float64) ** 2)).ravel()
        ms = ev_time(lambda: fwi.fwi_loss(x0, g_init, obs, misfit, dw, mask, True, True))
        work = 2.0 * np.prod(model.grid.shape) * (g_init.nt - 2) * g_init.nsrc
        svs = fwi._resident_surveys(g_init, list(range(g_init.nsrc)))
        out[name] = {"grid": list(model.grid.shape), "so": model.space_order, "nt": g_init.nt, "shots": g_init.nsrc,
                     "ms_per_objective_gradient": round(ms, 3), "gpts_per_s": round(work / ms / 1e6, 1),
                     "shots_per_s": round(g_init.nsrc / ms * 1e3, 1),
                     "misfit": getattr(misfit, '__name__', type(misfit).__name__) + (
                         "(method=%s)" % misfit.method if hasattr(misfit, 'method') else ""),
                     "launch_groups_shots_x_cluster": [[sv.nshots, int(sv.plan.cluster)] for sv in svs] if svs else None}
        fwi._SURVEYS.clear()

    g_true, g_init = configs.circle()                          # circle_fwi.py:65 uses space_order 6 ...
    objective("circle_fwi", g_true, g_init, None, None)
    g_true, g_init = configs.circle(space_order=4)             # ... BASELINE.json's text says 4: both are reported
    objective("circle_fwi_so4", g_true, g_init, None, None)
    g_true, g_init, g_const, mask = configs.marmousi2()
    objective("marmousi2_fwi_L2", g_true, g_init, g_const, mask)
    # BASELINE.json configs[3] as the reference runs it (marmousi2_fwi.py:131-132): back-and-forth W2 of whole records
    from devito_fwi_b200.misfit import qWasserstein
    objective("marmousi2_fwi_QW2D", g_true, g_init, g_const, mask,
              qWasserstein(method='2d', gamma=1.01, num_steps=15, step_scale=4.))
    g_true, g_init, g_const, _ = configs.marmousi(nsrc=21, tn=4500.)
    ms = ev_time(lambda: [fwi.fm_multi(g) for g in (g_true, g_init, g_const)])
    model = g_true.model
    work = 3.0 * np.prod(model.grid.shape) * (g_true.nt - 2) * g_true.nsrc
    out["marmousi_fm"] = {"grid": list(model.grid.shape), "so": model.space_order, "nt": g_true.nt, "shots": g_true.nsrc,
                          "models": 3, "ms_all_shots_3_models": round(ms, 3), "gpts_per_s": round(work / ms / 1e6, 1)}
    fwi._SURVEYS.clear()
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
    ap.add_argument("--workload", default="marmousi", choices=["marmousi", "layered3d"])
    ap.add_argument("--no-extra", action="store_true", help="skip the other named 2-D configurations (N=1 only)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling legs (29 / 32 shots over N ranks)")
    ap.add_argument("--no-3d", action="store_true", help="skip the 3-D 592^3 workload")
    ap.add_argument("--tn3d", type=float, default=1250., help="end time of the 3-D workload (1250 -> nt = 690)")
    args = ap.parse_args()
    import warnings
    warnings.filterwarnings("ignore")
    if args.cpu_only:
        print(json.dumps(cpu_baseline()), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
