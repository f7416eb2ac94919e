"""BASELINE.json configs[4]: 3-D layered model 512^3 (+2*40 sponge), so=8 - L2 FWI gradient of a small survey
with on-device checkpointing, shots partitioned across the GPUs of one box and ONE NCCL all-reduce of
[grad | fval].

    python examples/layered3d_gradient.py [--size 512] [--tn 1250] [--shots-per-gpu 1]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/layered3d_gradient.py

Observed data come from a perturbed model (a +10 % block 10-40 cells below the acquisition surface: with tn = 1.25 s
only the top layer is illuminated), as the survey's measurement table asks.
Per shot: forward(save='checkpoint') [pass 1, records the receivers] -> residual -> gradient(rec, u=<checkpoints>)
[pass 2: recompute + adjoint/imaging]; the reference's own sequence (acoustic_example.py:26-63 + checkpointing=True)
runs the forward sweep twice."""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")


def main():
    import torch
    import devito_fwi_b200 as b
    from devito_fwi_b200 import configs, dist
    from devito_fwi_b200.fwi import _shot_geometry

    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512, help="cells per dimension of the physical domain")
    ap.add_argument("--tn", type=float, default=1250.)
    ap.add_argument("--shots-per-gpu", type=int, default=1)
    ap.add_argument("--rec-decimate", type=int, default=4)
    args = ap.parse_args()

    dist.init_from_env()
    rank, world = dist.rank(), dist.world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))

    geom0 = configs.layered3d(n=args.size, space_order=8, tn=args.tn, rec_decimate=args.rec_decimate)
    model = geom0.model
    nshots = args.shots_per_gpu * world
    # sources on a line through the centre, depth as in the single-shot template
    centre = np.array(geom0.src_positions[0], dtype=np.float64)
    src = np.repeat(centre[None], nshots, axis=0)
    span = 0.5 * model.domain_size[0]
    src[:, 0] = centre[0] + (np.linspace(-0.5, 0.5, nshots) * span if nshots > 1 else 0.0)
    geom = b.AcquisitionGeometry(model, geom0.rec_positions, src, geom0.t0, geom0.tn, f0=geom0.f0, src_type='Ricker')

    # "true" model: same layering + a fast block under the spread
    vp_true = b.Function(name='vp_true', grid=model.grid)
    v = np.array(model.vp.data)
    nbl, n = model.nbl, args.size
    lo, hi = nbl + 3 * n // 8, nbl + 5 * n // 8
    v[lo:hi, lo:hi, nbl + 10:nbl + 40] *= 1.1
    vp_true.data[...] = v

    steps = geom.nt - 2
    npts = int(np.prod(model.grid.shape))
    buf = torch.zeros(model.grid.slice_elems + 1, dtype=torch.float32, device='cuda')    # [grad | fval]
    grad = b.Function(name='grad', grid=model.grid)
    grad._buf._dev = buf[:-1].view(model.grid.slice_shape)       # accumulate straight into the all-reduce buffer
    grad._buf._newer = 'dev'
    dist.all_reduce_sum(torch.zeros(1, device='cuda'))       # NCCL communicator set-up stays outside the timing
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.time()
    for i in dist.local_shots(nshots):
        g_i = _shot_geometry(geom, i)
        solver = b.AcousticWaveSolver(model, g_i, space_order=8, profile=False)
        obs = solver.forward(vp=vp_true)[0]._sdata.dev().clone()              # synthetic "observed" data
        rec, cw, _ = solver.forward(save='checkpoint')
        res = rec._sdata.dev() - obs
        buf[-1] += 0.5 * (res.double() ** 2).sum().float()
        r = b.Receiver(name='res', grid=model.grid, time_range=g_i.time_axis, coordinates=g_i.rec_positions)
        r._sdata.adopt_dev(res)
        solver.gradient(rec=r, u=cw, grad=grad)
        del cw
    dist.all_reduce_sum(buf)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    dt_all = time.time() - t0
    if rank == 0:
        g = buf[:-1]
        print(json.dumps({
            "workload": "layered3d %d^3 (+2*%d) so=8 nt=%d, %d receivers, %d shots on %d GPU(s)" % (
                args.size, model.nbl, geom.nt, geom.nrec, nshots, world),
            "seconds": round(dt_all, 3), "shots_per_s": round(nshots / dt_all, 3),
            "per_shot": "observed-data forward + forward(+checkpoints) + recompute + adjoint/imaging",
            "gpts_per_s_4_sweeps": round(4.0 * npts * steps * nshots / dt_all / 1e9, 1),
            "fval": float(buf[-1]), "grad_absmax": float(g.abs().max()),
            "allreduce_floats": int(buf.numel()),
            "hbm_peak_alloc_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
