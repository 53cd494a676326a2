"""View-parallel training WITH the f2 row (SURVEY §8(e) + §8(f2)), 2 GPUs, NCCL: every rank renders its own view with
the fused operator, the six raw-parameter gradients are summed over the ranks in ONE flat all-reduce, the one-launch
Adam applies the mean (grad_scale = 1 / world, DDP semantics, pipelines/base_pipeline.py:202-207), the densification
statistics are all-reduced (sum, sum, max) and the refinement runs on every rank with the same seed.
Checks: (i) the replicas stay bit-identical through refinements, (ii) the result equals a single-process run that
renders both views itself and averages the gradients (gradient accumulation), up to the order of the FP32 sums."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
H, W = 240, 320
GROUPS = ("means", "scales", "quats", "features_dc", "features_rest", "opacities")


def _setup():
    from oracle import oracle as orc
    from rasterizer.synthetic import look_at_viewmat, make_scene, projection_matrix

    teacher = make_scene(30_000, W, H, 0.02, 0.12, margin=1.0, seed=200)
    student = make_scene(20_000, W, H, 0.02, 0.12, margin=1.0, seed=201)
    fovx = math.radians(60.0)
    fovy = 2.0 * math.atan(0.5 * H / teacher["fy"])
    P = projection_matrix(0.001, 1000.0, fovx, fovy).astype(np.float64)
    cams = []
    for k in range(8):
        V = look_at_viewmat(yaw_deg=-14.0 + 4.0 * k, pitch_deg=5.0 * math.sin(k))
        cams.append((np.ascontiguousarray(V), np.ascontiguousarray((P @ V.astype(np.float64)).astype(np.float32))))

    def params_of(scene):
        raw = orc.raw_parameters(scene)
        return {"means": scene["means3d"], "scales": raw["scales_raw"], "quats": raw["quats_raw"], "features_dc": raw["features_dc"],
                "features_rest": raw["features_rest"], "opacities": raw["opacities_raw"]}

    return params_of(teacher), params_of(student), cams, (teacher["fx"], teacher["fy"], teacher["cx"], teacher["cy"])


def _train(rank, world, views_per_step, steps, dist=None):
    """world == 1 and views_per_step == 2: the single-process reference (both views, averaged gradients)."""
    from rasterizer.densify import DensifyConfig, DensifyStats, refinement_after
    from rasterizer.fused import RenderAux, render_gaussians
    from rasterizer.losses import l1_ssim_loss
    from rasterizer.optim import GaussianOptimizers

    tp, sp, cams, (fx, fy, cx, cy) = _setup()
    dev = torch.device("cuda", rank)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cams = [(cu(V), cu(PM)) for V, PM in cams]
    bg = torch.zeros(3, device=dev)
    teacher = {k: cu(v) for k, v in tp.items()}

    def render(p, cam, aux=None):
        rgb, _, _ = render_gaussians(p["means"], p["scales"], p["quats"], p["features_dc"], p["features_rest"], p["opacities"],
                                     cam[0], cam[1], fx, fy, cx, cy, H, W, 3, background=bg, render_depth=False, aux=aux)
        return torch.clamp(rgb, max=1.0)

    with torch.no_grad():
        gts = [render(teacher, c) for c in cams]
    p = {k: cu(v).requires_grad_(True) for k, v in sp.items()}
    opt = GaussianOptimizers(p)
    cfg = DensifyConfig(warmup_length=10, refine_every=10, reset_alpha_every=6, densify_grad_thresh=5e-5)
    stats = DensifyStats()
    total = world * views_per_step
    counts, snapshot = [], None
    for step in range(steps):
        grads = None
        for j in range(views_per_step):
            v = (step * total + rank * views_per_step + j) % len(cams)
            aux = RenderAux()
            for t in p.values():
                t.grad = None
            loss = l1_ssim_loss(render(p, cams[v], aux), gts[v], 0.2)
            loss.backward()
            flat = torch.cat([p[k].grad.reshape(-1) for k in GROUPS])
            grads = flat if grads is None else grads + flat
            stats.update(aux.xys_grad, aux.radii, (H, W))
        if dist is not None:
            dist.all_reduce(grads)                      # the one exchange step: 59 floats / Gaussian, summed
        off, gd = 0, {}
        for k in GROUPS:
            gd[k] = grads[off:off + p[k].numel()].view_as(p[k])
            off += p[k].numel()
        opt.optimizer_step_all(gd, grad_scale=1.0 / total)
        if step == cfg.refine_every:
            snapshot = {k: v.detach().cpu().clone() for k, v in p.items()}   # before the first refinement
        if step % cfg.refine_every == 0 and step > 0:
            stats.all_reduce()
            g = torch.Generator(device=dev).manual_seed(1000 + step)   # same draws on every rank
            info = refinement_after(p, opt, stats, cfg, step, 2, (H, W), generator=g)
            counts.append(info["n_after"])
    torch.cuda.synchronize()
    return {k: v.detach().cpu() for k, v in p.items()}, counts, snapshot


def _worker(rank, world, port, results):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    p, counts, snap = _train(rank, world, 1, 35, dist)
    results[rank] = ({k: v.numpy() for k, v in p.items()}, counts, {k: v.numpy() for k, v in snap.items()})
    dist.destroy_process_group()


def test_view_parallel_training_with_densification_2gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, 29700 + os.getpid() % 1000, results), nprocs=2, join=True)
    (p0, c0, s0), (p1, c1, _) = results[0], results[1]
    assert c0 == c1 and len(c0) == 3 and max(c0) > 20_000, (c0, c1)
    for k in GROUPS:
        assert np.array_equal(p0[k], p1[k]), f"replicas diverged in {k}"
    # single-process reference up to the first refinement (the all-reduced vis_counts start at 1 PER RANK, so later
    # split decisions are not comparable one to one): same parameters up to the order of the FP32 gradient sums
    _, cr, ref = _train(0, 1, 2, 11)
    print("[view-parallel f2] Gaussian counts after refinements (2 ranks):", c0)
    # (Adam turns the sign of a rounding-noise gradient into a full +-lr step, so a few elements legitimately differ:
    # the bound is on the fraction of such elements and on the norm)
    for k in GROUPS:
        a, b = s0[k], ref[k].numpy()
        err = float(np.linalg.norm(a - b) / np.linalg.norm(b))
        frac = float((np.abs(a - b) > 1e-3 * np.abs(b) + 1e-4).mean())
        print(f"[view-parallel f2] {k}: normwise {err:.2e}, fraction of differing elements {frac:.2e}")
        assert err < 5e-4 and frac < 5e-3, (k, err, frac)   # observed over runs: <= 2.3e-5, <= 2.5e-4
