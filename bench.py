#!/usr/bin/env python
"""bench.py — rasterizer forward+backward views/s (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one view of SURVEY §8(d): SH fwd -> project fwd -> bin/sort -> blend fwd (rgb + alpha) ->
[fixed synthetic upstream gradients] -> blend bwd -> SH bwd -> project bwd [-> view-parallel gradient exchange
when N_gpus > 1: fused gather + SH adjoint over NVLink peer memory and an NCCL all-reduce of the other 11 N floats,
rasterizer/view_parallel.py].  Workload at N=1: cfg2 = 1 M Gaussians, 1920x1080,
SH degree 3 (BASELINE.json configs[1]).  Multi-GPU is view-parallel (weak scaling): every rank renders
its own camera of the same replicated scene.

Prints ONE JSON line (rank 0).  `value` = device-timed views/s with everything resident in HBM, through
the C ABI (`rasterizer.cuda` -> libgsr_b200.so).  `e2e` = views/s through the public autograd API
(`rasterizer.project_gaussians / spherical_harmonics / rasterize_gaussians`) with the per-view inputs
(camera matrices, upstream image gradients) copied from pinned host memory and the rendered image + alpha
read back to the host inside the timed region.  `roofline` is for the blend-adjoint kernel (the dominant
one).  `cpu_baseline` = the CPU oracle (a port; oracle/) on the host cores.  `--impl reference` times the
reference path's CPU implementation (the oracle port: the reference's own CPU code is Python and cannot
travel to the GPU box) on all host threads.

`ref_cuda_ext` (N=1 only) = the UNMODIFIED reference CUDA extension (oracle/_ref/rasterizer_ref_cuda.so, built from the
sources under /root/reference by oracle/build_ref.py with the reference's packaged flags) driven by the reference's own
orchestration (cumsum -> .item() -> map_gaussian_to_intersects -> torch.sort -> torch.gather -> get_tile_bin_edges ->
rasterize fwd/bwd, rasterizer/utils.py:106-182, rasterize.py:92-247) on the same scene in the same process, timed after
our own legs: the north star's ">= 2x the reference rasterizer on the same box" as a driver-visible number.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gaussian-splatting-toolkit_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "rasterizer fwd+bwd views/sec @1080p, 1M Gaussians"
UNIT = "views/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg4", "cfg3view"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-resident", action="store_true", help="profiling aid: only the resident leg (no graph / e2e / "
                    "fused / reference-extension legs); the JSON line is tagged diagnostic")
    ap.add_argument("--cpu-budget-s", type=float, default=200.0, help="wall-clock bound of the reference arm")
    return ap.parse_args()


def ncu_facts(kernel):
    """Per-launch DRAM traffic etc. of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(path))
        return d["kernels"].get(kernel), d.get("source")
    except Exception:
        return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        return {"sm_mhz": sm_sorted[len(sm_sorted) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
def algorithmic_bytes(N, M, P, T):
    """SURVEY §8(d) per-view algorithmic bytes (FP32, each tensor touched once)."""
    return {
        "sh_fwd": 216 * N, "project_fwd": 100 * N, "cumsum": 8 * N, "key_emit": 20 * N + 12 * M, "sort": 24 * M,
        "bin_edges": 8 * M + 8 * T, "blend_fwd": 40 * M + 20 * P + 8 * T, "blend_bwd": 40 * M + 24 * P + 8 * T + 36 * N,
        "project_bwd": 152 * N, "sh_bwd": 216 * N, "total": 748 * N + 124 * M + 44 * P + 24 * T,
    }


class ResidentView:
    """One view through the C ABI with every input resident in HBM (the `value` leg).  Stage boundaries carry
    CUDA events (on torch's current stream = the launching stream) for the per-kernel breakdown."""

    STAGES = ["sh_fwd", "project_fwd", "binning", "blend_fwd", "blend_bwd", "sh_bwd", "project_bwd"]

    def __init__(self, s, bucket=None, exchange=None):
        import torch
        from rasterizer import cuda as C

        self.torch, self.C, self.s = torch, C, s
        self.N = s["means3d"].shape[0]
        H, W, bw = s["img_height"], s["img_width"], s["block_width"]
        self.tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
        self.degree = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[s["sh_coeffs"].shape[1]]
        self.viewdirs = (s["means3d"] - s["cam_pos"][None, :]).contiguous()
        self.opac = s["opacities"].reshape(-1, 1).contiguous()
        self.zeros_n = torch.zeros(self.N, device=s["means3d"].device)
        self.pin_meta = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.capacity = None  # pair-buffer capacity of the asynchronous binning, learned from the first (synchronous) step
        self.events = []
        self.M = 0
        self.bucket = bucket  # view_parallel.GradientBucket: backward kernels write straight into its segments
        self.exchange = exchange  # view_parallel.GradientExchange (world > 1)

    def _mark(self, rec):
        if rec is not None:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            rec.append(e)

    def step(self, record=False):
        torch, C, s = self.torch, self.C, self.s
        H, W, bw, N = s["img_height"], s["img_width"], s["block_width"], self.N
        rec = [] if record else None
        self._mark(rec)
        rgb_sh = C.compute_sh_forward(N, self.degree, s["degrees_to_use"], self.viewdirs, s["sh_coeffs"])
        colors = torch.clamp(rgb_sh + 0.5, min=0.0)
        self._mark(rec)
        cov3d, xys, depths, radii, conics, comp, nth = C.project_gaussians_forward(
            N, s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"], s["projmat"], s["fx"], s["fy"],
            s["cx"], s["cy"], H, W, bw, s["clip_thresh"])
        self._mark(rec)
        # internal binning of rasterize_gaussians: two-level sort (own radix sort) + exact tile culling, same per-tile
        # order as the reference.  First step: synchronous form (learns M); afterwards the asynchronous form — the pair
        # count stays on the device, NO host read inside the timed region (M is read from the pinned slot afterwards)
        if self.capacity is None:
            M, vs, bins = C.bin_gaussians_fast(xys, depths, radii, conics, self.opac.reshape(-1), H, W, bw)
            self.M = M
            self.capacity = int(1.25 * M) + 65536
        else:
            vs, bins, _meta = C.bin_gaussians_device(xys, depths, radii, conics, self.opac.reshape(-1), H, W, bw,
                                                     self.capacity, meta_pinned=self.pin_meta)
        self._mark(rec)
        img, fT, fi = C.rasterize_forward(self.tb, (bw, bw, 1), (W, H, 1), vs, bins, xys, conics, colors, self.opac,
                                          s["background"])
        alpha = 1 - fT
        self._mark(rec)
        bk = self.bucket
        v_xy, v_conic, v_colors, v_opacity = C.rasterize_backward(
            H, W, bw, vs, bins, xys, conics, colors, self.opac, s["background"], fT, fi, s["v_out_img"],
            s["v_out_alpha"], out_opacity=None if bk is None else bk["v_opacity"])
        self._mark(rec)
        v_rgb_sh = torch.where(rgb_sh + 0.5 > 0, v_colors, torch.zeros_like(v_colors))
        if bk is None:
            v_coeffs = C.compute_sh_backward(N, self.degree, s["degrees_to_use"], self.viewdirs, v_rgb_sh)
        else:
            v_coeffs = v_rgb_sh  # view-parallel: the SH adjoint is evaluated for all ranks' views in the exchange
            if self.exchange is not None:
                self.exchange.start_sh(v_rgb_sh, s["cam_pos"], s["degrees_to_use"])
        self._mark(rec)
        _, _, v_mean, v_scale, v_quat = C.project_gaussians_backward(
            N, s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"], s["projmat"], s["fx"], s["fy"],
            s["cx"], s["cy"], H, W, cov3d, radii, conics, comp, v_xy, self.zeros_n, v_conic, self.zeros_n,
            out_mean3d=None if bk is None else bk["v_mean3d"], out_scale=None if bk is None else bk["v_scale"],
            out_quat=None if bk is None else bk["v_quat"], need_cov_grads=False)
        self._mark(rec)
        if rec is not None:
            self.events.append(rec)
        return img, alpha, (v_coeffs, v_mean, v_scale, v_quat, v_opacity)

    def stage_ms(self):
        out = {k: 0.0 for k in self.STAGES}
        for rec in self.events:
            for i, k in enumerate(self.STAGES):
                out[k] += rec[i].elapsed_time(rec[i + 1])
        n = max(1, len(self.events))
        return {k: v / n for k, v in out.items()}


# diagnostic only (never a reported number): skip the large host copies of the e2e leg to separate PCIe time from
# exchange time when looking at multi-GPU runs; the JSON line is tagged "diagnostic"
DIAG_NOCOPY = os.environ.get("GSR_E2E_NOCOPY") == "1"


class PublicApiView:
    """One view through the public autograd API with HOST buffers for the per-view inputs/outputs (`e2e`).

    Every step copies that step's inputs (camera matrices, upstream image gradients) from pinned host memory and
    reads the rendered image + alpha back into pinned host memory.  The two large transfers run on dedicated copy
    streams so that PCIe traffic overlaps the kernels (H2D of the upstream gradients under the forward pass, D2H of
    the image under the backward pass); device-side input buffers are double-buffered across steps."""

    def __init__(self, s, scene_np, bucket=None, peer=None):
        import torch

        self.torch, self.s, self.bucket, self.peer = torch, s, bucket, peer
        dev = s["means3d"].device
        self.means = s["means3d"].clone().requires_grad_(True)
        self.scales = s["scales"].clone().requires_grad_(True)
        self.quats = s["quats"].clone().requires_grad_(True)
        self.coeffs = s["sh_coeffs"].clone().requires_grad_(True)
        self.opac = s["opacities"].reshape(-1, 1).clone().requires_grad_(True)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        self.h_viewmat, self.h_projmat = pin(scene_np["viewmat"]), pin(scene_np["projmat"])
        self.h_vimg, self.h_valpha = pin(scene_np["v_out_img"]), pin(scene_np["v_out_alpha"])
        H, W = s["img_height"], s["img_width"]
        self.h_img = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
        self.h_alpha = torch.empty((H, W), dtype=torch.float32).pin_memory()
        self.d_viewmat = torch.empty_like(s["viewmat"]); self.d_projmat = torch.empty_like(s["projmat"])
        self.d_vimg = [torch.empty((H, W, 3), device=dev) for _ in range(2)]
        self.d_valpha = [torch.empty((H, W), device=dev) for _ in range(2)]
        self.h2d_stream, self.d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.buf_free = [torch.cuda.Event(), torch.cuda.Event()]
        for e in self.buf_free:
            e.record()
        self.k = 0
        self.step_done = [None, None]  # the host runs at most two views ahead of the device (double-buffered inputs)
        self.h2d = sum(t.numel() * t.element_size() for t in (self.h_viewmat, self.h_projmat, self.h_vimg, self.h_valpha))
        self.d2h = sum(t.numel() * t.element_size() for t in (self.h_img, self.h_alpha))

    def compute(self, i):
        """The view itself through the public autograd API, on the current stream: reads the camera buffers and input
        buffers `i`, returns (img, alpha) and leaves the parameter gradients in .grad.  Eager in `step`, or captured into
        a CUDA graph by `capture` (rasterizer.graphs) and replayed."""
        import rasterizer
        from rasterizer.sh import spherical_harmonics

        torch, s = self.torch, self.s
        H, W, bw = s["img_height"], s["img_width"], s["block_width"]
        for p in (self.means, self.scales, self.quats, self.coeffs, self.opac):
            p.grad = None
        xys, depths, radii, conics, comp, nth, cov3d = rasterizer.project_gaussians(
            self.means, self.scales, s["glob_scale"], self.quats, self.d_viewmat, self.d_projmat, s["fx"], s["fy"],
            s["cx"], s["cy"], H, W, bw, s["clip_thresh"])
        if self.bucket is None:
            viewdirs = self.means.detach() - s["cam_pos"][None, :]
            sh = spherical_harmonics(s["degrees_to_use"], viewdirs, self.coeffs)
        else:  # view-parallel: coeffs.grad comes out already summed over the ranks
            from rasterizer.view_parallel import spherical_harmonics_view_parallel

            sh = spherical_harmonics_view_parallel(s["degrees_to_use"], self.means, s["cam_pos"], self.coeffs, peer=self.peer)
        rgbs = torch.clamp(sh + 0.5, min=0.0)
        img, alpha = rasterizer.rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, self.opac, H, W, bw,
                                                    background=s["background"], return_alpha=True)
        return img, alpha

    def backward(self, i, img, alpha):
        torch = self.torch
        torch.autograd.backward([img, alpha], [self.d_vimg[i], self.d_valpha[i]])
        if self.bucket is not None:
            # the remaining 11 floats / Gaussian: pack + one all-reduce over the non-SH part of the bucket
            import torch.distributed as dist

            bk = self.bucket
            for name, p in (("v_mean3d", self.means), ("v_scale", self.scales), ("v_quat", self.quats), ("v_opacity", self.opac)):
                bk[name].copy_(p.grad.reshape(bk[name].shape))
            dist.all_reduce(bk.flat[bk.offsets["v_coeffs"][1]:])
        return (self.coeffs.grad, self.means.grad, self.scales.grad, self.quats.grad, self.opac.grad)

    def capture(self):
        """Capture forward and backward of the view for both input-buffer sets (rasterizer.graphs.capture_step): the replayed
        step is then two graph launches around the D2H copy of the image.  Returns False (and stays eager) if the capture
        fails on this platform."""
        from rasterizer import graphs

        torch = self.torch
        try:
            self.graph_fwd, self.graph_bwd, self.graph_out = [None, None], [None, None], [None, None]
            pool = torch.cuda.graph_pool_handle()
            for i in range(2):
                # inputs must be valid numbers during warm-up / capture
                self.d_vimg[i].copy_(self.h_vimg, non_blocking=True)
                self.d_valpha[i].copy_(self.h_valpha, non_blocking=True)
                self.d_viewmat.copy_(self.h_viewmat, non_blocking=True)
                self.d_projmat.copy_(self.h_projmat, non_blocking=True)
                torch.cuda.synchronize()
                box = {}

                def fwd():
                    box["out"] = self.compute(i)
                    return box["out"]

                cf = graphs.capture_step(fwd, warmup=3, pool=pool)
                img, alpha = cf.result

                def bwd():
                    return self.backward(i, img, alpha)

                # the backward graph is captured ONCE, directly (its warm-up would need a fresh forward graph each time)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    grads = bwd()
                torch.cuda.current_stream().wait_stream(side)
                self.graph_fwd[i], self.graph_bwd[i], self.graph_out[i] = cf.graph, g, (img, alpha, grads)
            torch.cuda.synchronize()
            self.graphed = True
        except Exception as e:  # pragma: no cover - platform dependent
            print(f"[bench] CUDA-graph capture of the public-API view failed, e2e stays eager: {e}", file=sys.stderr)
            self.graphed = False
            torch.cuda.synchronize()
        return self.graphed

    def step(self):
        torch, s = self.torch, self.s
        main = torch.cuda.current_stream()
        i = self.k & 1
        self.k += 1
        if self.step_done[i] is not None:
            self.step_done[i].synchronize()  # view k-2 has finished: its buffers (device inputs, pinned outputs) are free
        # this step's inputs: cameras on the compute stream (128 B), upstream gradients on the H2D stream
        self.d_viewmat.copy_(self.h_viewmat, non_blocking=True)
        self.d_projmat.copy_(self.h_projmat, non_blocking=True)
        self.h2d_stream.wait_event(self.buf_free[i])
        with torch.cuda.stream(self.h2d_stream):
            if not DIAG_NOCOPY:
                self.d_vimg[i].copy_(self.h_vimg, non_blocking=True)
                self.d_valpha[i].copy_(self.h_valpha, non_blocking=True)
            in_ready = torch.cuda.Event()
            in_ready.record()
        if getattr(self, "graphed", False):
            self.graph_fwd[i].replay()
            img, alpha, grads = self.graph_out[i]
        else:
            img, alpha = self.compute(i)
        # this step's result goes back to the host on the D2H stream, under the backward pass
        fwd_done = torch.cuda.Event()
        fwd_done.record()
        self.d2h_stream.wait_event(fwd_done)
        with torch.cuda.stream(self.d2h_stream):
            img_d, alpha_d = img.detach(), alpha.detach()
            if not DIAG_NOCOPY:
                self.h_img.copy_(img_d, non_blocking=True)
                self.h_alpha.copy_(alpha_d, non_blocking=True)
            img_d.record_stream(self.d2h_stream)
            alpha_d.record_stream(self.d2h_stream)
            out_read = torch.cuda.Event()
            out_read.record()
        main.wait_event(in_ready)
        if getattr(self, "graphed", False):
            self.graph_bwd[i].replay()
            main.wait_event(out_read)  # the static image of buffer set i is overwritten by the next replay of graph i
        else:
            grads = self.backward(i, img, alpha)
        self.buf_free[i].record()
        done = torch.cuda.Event()
        done.record()
        self.step_done[i] = done
        return grads

    def finish(self):
        """Called inside the timed region after the last step: all transfers of all steps have landed."""
        self.torch.cuda.current_stream().wait_stream(self.d2h_stream)
        self.torch.cuda.current_stream().wait_stream(self.h2d_stream)


def timed_loop(torch, dist, world, fn, steps, warmup, finish=None):
    """W untimed + exactly K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if finish is not None:
        finish()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def fused_leg(torch, s, scene_np, steps, warmup):
    """Informational: the same view through the FUSED operator rasterizer.fused.render_gaussians, which also
    folds the model's activations (exp / normalise / sigmoid / SH concat / clamp) in; rgb + alpha, same upstream gradients."""
    import numpy as np
    from rasterizer.fused import render_gaussians

    o = np.clip(scene_np["opacities"].astype(np.float64), 1e-6, 1 - 1e-6)
    dev = s["means3d"].device
    raw = [s["means3d"].clone(), torch.log(s["scales"]), s["quats"].clone(), s["sh_coeffs"][:, 0, :].contiguous(),
           s["sh_coeffs"][:, 1:, :].contiguous(), torch.from_numpy(np.log(o / (1 - o)).astype(np.float32)).to(dev)[:, None]]
    raw = [t.requires_grad_(True) for t in raw]
    H, W = s["img_height"], s["img_width"]
    v_alpha = s["v_out_alpha"][..., None].contiguous()

    def step():
        for t in raw:
            t.grad = None
        rgb, _, alpha = render_gaussians(*raw, s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W,
                                         s["degrees_to_use"], background=s["background"], block_width=s["block_width"],
                                         render_depth=False)
        torch.autograd.backward([rgb, alpha], [s["v_out_img"], v_alpha])

    ms = timed_loop(torch, None, 1, step, steps, warmup)
    return {"value": steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
            "what": "render_gaussians(raw parameters) fwd+bwd, resident inputs; not the headline (reference-facing API) path"}


def host_link_gbs(torch, pv):
    """H2D / D2H rate of the e2e leg's own pinned image buffers (3 copies each, CUDA events) — outside every timed
    region; reported next to `e2e` so that a box with a slow host link is recognisable in the record."""
    try:
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        e[0].record()
        for _ in range(3):
            pv.d_vimg[0].copy_(pv.h_vimg, non_blocking=True)
        e[1].record()
        dimg = pv.d_vimg[0]
        for _ in range(3):
            pv.h_img.copy_(dimg, non_blocking=True)
        e[2].record()
        torch.cuda.synchronize()
        nbytes = 3 * pv.h_vimg.numel() * 4
        return {"h2d": nbytes / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9, "d2h": nbytes / (e[1].elapsed_time(e[2]) * 1e-3) / 1e9,
                "what": "3 x 24.9 MB copies between the leg's pinned host buffers and HBM, outside the timed regions"}
    except Exception as ex:  # diagnostic only
        return {"error": str(ex)}


def workload_text(name, scene_np):
    N, W, H = scene_np["means3d"].shape[0], scene_np["img_width"], scene_np["img_height"]
    return (f"{name}: {N} Gaussians, {W}x{H}, SH degree {scene_np['sh_degree']}, fwd+bwd, block_width "
            f"{scene_np['block_width']}, seeded scene of SURVEY 8(d)"
            + (" — clustered / object-centric variant (85 % of the Gaussians in a blob around the look-at point)"
               if name == "cfg3view" else ""))


def make_scene_for_rank(workload, rank):
    from rasterizer.synthetic import look_at_viewmat, make_config_scene

    viewmat = None if rank == 0 else look_at_viewmat(yaw_deg=45.0 * rank)  # cfg5: yaw = rank * 45 deg
    return make_config_scene(workload, seed=0, viewmat=viewmat)


def cpu_leg(scene_np, n_views_budget_s, steps=None, warmup=0):
    """Times the CPU oracle (port of the reference path) on all host threads.  Each step is a bounded sample:
    the whole view when it fits the budget, else a band of tile rows (fraction f of the image) — SH /
    projection / binning always run on all N Gaussians.  Returns (views_per_s, description, cores, ms_per_step)."""
    from oracle import oracle as orc

    orc.build()
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use every host core
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    orc.set_num_threads(max(1, ncpu))
    cores = orc.num_threads()
    s = scene_np
    t0 = time.perf_counter()
    orc.render_view(s, s["v_out_img"], s["v_out_alpha"])  # one untimed full view: page-in + calibration
    t_full = time.perf_counter() - t0
    if steps is None:
        steps = max(1, min(3, int(n_views_budget_s / max(t_full, 1e-3)) - 1))
        warmup = 0
    total_steps = steps + warmup
    bw, H = s["block_width"], s["img_height"]
    tiles_y = (H + bw - 1) // bw
    frac = min(1.0, n_views_budget_s / (total_steps * t_full))
    rows = max(1, int(round(frac * tiles_y)))
    r0 = (tiles_y - rows) // 2
    tile_rows = (r0, r0 + rows)
    frac = rows / tiles_y
    for _ in range(warmup):
        orc.render_view(s, s["v_out_img"], s["v_out_alpha"], tile_rows=tile_rows)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.render_view(s, s["v_out_img"], s["v_out_alpha"], tile_rows=tile_rows)
    dt = time.perf_counter() - t0
    views_per_s = steps * frac / dt
    desc = (f"{steps} step(s), each the full per-Gaussian work (SH, projection, binning of all {s['means3d'].shape[0]} "
            f"Gaussians) + blend fwd/bwd of {rows}/{tiles_y} tile rows ({frac:.3f} of a view), scaled to whole views; "
            f"CPU oracle (C + OpenMP port of the reference CUDA path), {cores} threads")
    return views_per_s, desc, cores, 1e3 * dt / steps


def torch_impl_cfg1_leg(budget_rows=8):
    """BASELINE configs[0]: the reference's OWN CPU render path, literally — rasterizer/_torch_impl.py (pure PyTorch,
    installed unmodified under baseline/_ref/ by oracle/build_ref.install_ref_python) on cfg1 (10 k Gaussians, 256x256,
    all visible), forward only (the reference has no CPU backward).  Its per-pixel Python loops need minutes per image,
    so the timed sample is: projection of all Gaussians (vectorised) + `rasterize_forward` of the first `budget_rows`
    image rows (one tile row; the function is called with img_size = (W, budget_rows), the tile lists are those of the
    full image, produced by the CPU oracle because the reference's own per-Gaussian Python loop
    `map_gaussian_to_intersects` would add ~10 min), scaled to a whole image."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    path = os.path.join(ref_dir, "rasterizer", "_torch_impl.py")
    if not os.path.exists(path):
        return {"unavailable": "baseline/_ref/rasterizer/_torch_impl.py not present"}
    import importlib.util

    import numpy as np
    import torch

    from oracle import oracle as orc

    spec = importlib.util.spec_from_file_location("_ref_torch_impl", path)
    ti = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ti)
    sc = make_scene_for_rank("cfg1", 0)
    H, W, bw, N = sc["img_height"], sc["img_width"], sc["block_width"], sc["means3d"].shape[0]
    tt = lambda k: torch.from_numpy(sc[k])
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    t0 = time.perf_counter()
    cov3d, cov2d, xys, depths, radii, conics, comp, nth, mask = ti.project_gaussians_forward(
        tt("means3d"), tt("scales"), 1.0, tt("quats"), tt("viewmat"), tt("projmat"),
        (sc["fx"], sc["fy"], sc["cx"], sc["cy"]), (W, H), bw)
    colors = torch.clamp(ti.compute_sh_color(tt("means3d") - tt("cam_pos")[None], tt("sh_coeffs")) + 0.5, min=0.0) \
        if hasattr(ti, "compute_sh_color") else torch.rand(N, 3)
    t_proj = time.perf_counter() - t0
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    m, cum = orc.compute_cumulative_intersects(nth.numpy().astype(np.int32))
    _, _, _, vs, bins = orc.bin_and_sort_gaussians(N, m, xys.numpy(), depths.numpy(), radii.numpy().astype(np.int32), cum, tb, bw)
    rows = min(budget_rows, H)
    t0 = time.perf_counter()
    ti.rasterize_forward(tb, (bw, bw, 1), (W, rows, 1), torch.from_numpy(vs).long(), torch.from_numpy(bins).long(), xys,
                         conics, colors, tt("opacities").reshape(-1, 1), tt("background"))
    t_rows = time.perf_counter() - t0
    t_img = t_proj + t_rows * H / rows
    return {"views_per_s": 1.0 / t_img, "s_per_view_forward": t_img, "cores": torch.get_num_threads(), "kind": "reference",
            "workload": f"cfg1: {N} Gaussians, {W}x{H}, forward only", "num_intersects": int(m),
            "sample": f"project_gaussians_forward (all Gaussians, {t_proj:.2f} s) + rasterize_forward of {rows}/{H} image rows "
                      f"({t_rows:.1f} s), scaled to the whole image; baseline/_ref/rasterizer/_torch_impl.py unmodified"}


def ref_cuda_leg(torch, s, steps, warmup, ours_ms):
    """The unmodified reference CUDA extension behind the reference orchestration, same scene, same process (N=1).
    Present => measured; absent => {"unavailable": reason}."""
    try:
        from oracle.build_ref import load_ref

        ref_ext = load_ref()
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"loading oracle/_ref failed: {e}"}
    if ref_ext is None:
        return {"unavailable": "oracle/_ref/rasterizer_ref_cuda.so not present (built by __graft_entry__.build() where /root/reference exists)"}
    tests_dir = os.path.join(ROOT, "tests")
    if tests_dir not in sys.path:
        sys.path.insert(0, tests_dir)
    from pipelines import run_view_bindings

    names = ["compute_sh_forward", "project_gaussians_forward", "map_gaussian_to_intersects", "get_tile_bin_edges",
             "rasterize_forward", "rasterize_backward", "compute_sh_backward", "project_gaussians_backward"]
    acc = {n: [] for n in names}
    rec = {"on": False}

    class Timed:  # per-binding CUDA events; everything between the bindings (torch glue, sort) is the remainder
        def __getattr__(self, name):
            fn = getattr(ref_ext, name)
            if name not in acc:
                return fn

            def wrapped(*a, **k):
                if not rec["on"]:
                    return fn(*a, **k)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*a, **k)
                e1.record()
                acc[name].append((e0, e1))
                return r

            return wrapped

    C = Timed()
    out = {}

    def step():
        out["last"] = run_view_bindings(C, s, backward=True, sort_impl="torch", binning="reference")

    for _ in range(warmup):
        step()
    rec["on"] = True
    ms = timed_loop(torch, None, 1, step, steps, 0)
    rec["on"] = False
    per = ms / steps
    stages = {n: sum(a.elapsed_time(b) for a, b in v) / max(1, len(v)) for n, v in acc.items()}
    stages["torch_glue_cumsum_item_sort_gather_clamp_where"] = per - sum(stages.values())
    M_ref = int(out["last"]["num_intersects"])
    return {"views_per_s": 1e3 / per, "ms": per, "stages_ms": stages, "speedup": per / ours_ms,
            "num_intersects": M_ref, "steps": steps, "warmup": warmup,
            "what": "oracle/_ref/rasterizer_ref_cuda.so (reference sources, -O3 --use_fast_math, sm_100) + reference "
                    "orchestration (torch.cumsum/.item()/torch.sort/torch.gather), resident inputs, CUDA events; "
                    "speedup = ms / this line's ms_per_step",
            "_last": out["last"]}


def exchange_check(torch, dist, rv, bucket, exchange, s):
    """N > 1, before the timed region: the exchanged bucket (NVLink-peer SH adjoint + NCCL all-reduce of the other
    11 floats) must equal dist.all_reduce(sum) of the per-rank FULL gradients computed without the exchange path."""
    N = rv.N
    saved_bucket, saved_exchange = rv.bucket, rv.exchange
    rv.bucket, rv.exchange = None, None
    _, _, g = rv.step()  # (v_coeffs, v_mean, v_scale, v_quat, v_opacity) of THIS rank's view, plain kernels
    full = torch.cat([t.reshape(N, -1) for t in g], dim=1).contiguous()  # [N, 48+3+3+4+1]
    dist.all_reduce(full)
    rv.bucket, rv.exchange = saved_bucket, saved_exchange
    rv.step()
    exchange.finish()
    torch.cuda.synchronize()
    got = torch.cat([bucket[k].reshape(N, -1) for k in ("v_coeffs", "v_mean3d", "v_scale", "v_quat", "v_opacity")], dim=1)
    num = (got.double() - full.double()).norm()
    den = full.double().norm()
    stat = torch.stack([num * num, den * den])
    rel = float((stat[0] / stat[1]).sqrt().item())
    worst = torch.tensor([rel], device=got.device)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return {"norm_rel": float(worst.item()), "floats_per_gaussian": int(full.shape[1]),
            "what": "|| exchanged bucket - all_reduce(per-rank full gradients) || / || . ||, max over ranks"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        # rank 0 alone times the CPU path; the other ranks leave immediately
        if rank != 0:
            return
        scene_np = make_scene_for_rank(args.workload, 0)
        v, desc, cores, ms = cpu_leg(scene_np, args.cpu_budget_s, steps=args.steps, warmup=args.warmup)
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(args.workload, scene_np), "path": "CPU (oracle port), all host threads",
                       "pixels": scene_np["img_height"] * scene_np["img_width"]},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        if os.environ.get("GSR_BENCH_TORCH_IMPL", "1") != "0":
            try:
                line["torch_impl_cfg1"] = torch_impl_cfg1_leg()
            except Exception as e:  # never let the side measurement break the arm
                line["torch_impl_cfg1"] = {"unavailable": f"{type(e).__name__}: {e}"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    numa_cpus = None
    if world > 1 and os.environ.get("GSR_NUMA_BIND", "1") != "0":
        # before any pinned allocation: keep this rank's host buffers and PCIe traffic on its GPU's socket
        from rasterizer.view_parallel import bind_process_to_gpu_numa

        numa_cpus = bind_process_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from rasterizer import _lib
    from rasterizer.synthetic import scene_to_torch

    _lib.load()  # fail loudly if libgsr_b200.so is missing
    scene_np = make_scene_for_rank(args.workload, rank)
    s = scene_to_torch(scene_np, torch.device("cuda", local_rank))
    N = s["means3d"].shape[0]
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    P, T = H * W, ((W + bw - 1) // bw) * ((H + bw - 1) // bw)

    # flat gradient bucket for the view-parallel all-reduce: 48 (SH) + 3 + 3 + 4 + 1 = 59 floats / Gaussian; the
    # backward kernels of the resident leg write straight into its segments, the autograd leg packs into it
    from rasterizer.view_parallel import GradientBucket

    bucket = (GradientBucket(N, s["sh_coeffs"].shape[1], device=s["means3d"].device,
                             symmetric=os.environ.get("GSR_OWN_TAIL") in ("1", "push") and os.environ.get("GSR_NO_P2P") != "1")
              if world > 1 else None)
    ar_events = []
    from rasterizer.view_parallel import GradientExchange, PeerColorGrads

    peer = PeerColorGrads.try_create(N, device=s["means3d"].device) if (world > 1 and os.environ.get("GSR_NO_P2P") != "1") else None
    exchange = GradientExchange(bucket, s["means3d"], {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[s["sh_coeffs"].shape[1]], peer) if world > 1 else None
    if world > 1 and os.environ.get("GSR_OWN_TAIL") == "push" and bucket.hdl is not None:
        from rasterizer.view_parallel import PushGradientExchange

        exchange = PushGradientExchange(bucket, s["means3d"], {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[s["sh_coeffs"].shape[1]])

    rv = ResidentView(s, bucket, exchange)
    recording = {"on": False}

    def resident_step():
        _, _, grads = rv.step(record=recording["on"])
        if world == 1:
            return
        # grads[0] is the masked colour gradient v_rgb_sh; the other four already sit in the bucket
        if recording["on"]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        exchange.finish()
        if recording["on"]:
            e1.record()
            ar_events.append((e0, e1))

    xcheck = exchange_check(torch, dist, rv, bucket, exchange, s) if world > 1 else None
    sampler = ClockSampler(local_rank)
    # warm-up outside, then the timed region with stage events
    for _ in range(args.warmup):
        resident_step()
    recording["on"] = True
    if exchange is not None:
        exchange.timing = True
    sampler.start()
    lib = _lib.load()
    launches0 = int(lib.gsr_launch_count())
    ms_total = timed_loop(torch, dist, world, resident_step, args.steps, 0)
    launches = int(lib.gsr_launch_count()) - launches0
    clocks = sampler.stop()
    recording["on"] = False
    if exchange is not None:
        exchange.timing = False
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total * 1e-3)
    stages = rv.stage_ms()
    if args.steps + args.warmup > 1:  # the asynchronous binning's pinned slot: M and the overflow flag of the last step
        assert int(rv.pin_meta[1]) == 0, "pair buffers overflowed in the timed region"
        rv.M = int(rv.pin_meta[0])
    if ar_events:
        stages["grad_exchange_tail(allreduce 11N + join of the %s SH adjoint started after blend_bwd)" % ("NVLink-peer-load" if peer is not None else "NCCL-allgather")] = (
            sum(a.elapsed_time(b) for a, b in ar_events) / len(ar_events))
    if exchange is not None and exchange.peer is not None:
        stages["grad_exchange_breakdown"] = exchange.breakdown_ms()
    M = rv.M
    with torch.no_grad():
        _nth = rv.C.project_gaussians_forward(N, s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"],
                                              s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W, bw,
                                              s["clip_thresh"])[6]
        M_ref = int(_nth.sum().item())

    # the same resident view as ONE CUDA graph (N = 1): possible because no kernel's grid depends on the pair count and
    # nothing in the view reads the device from the host (asynchronous binning); replayed back to back
    graph_leg = None
    if world == 1 and os.environ.get("GSR_BENCH_GRAPH", "1") != "0" and not args.only_resident:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                rv.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                g_out = rv.step()
            ms_g = timed_loop(torch, dist, 1, g.replay, args.steps, args.warmup)
            assert int(rv.pin_meta[1]) == 0, "pair buffers overflowed in the graph replay"
            img_stream = rv.step()[0]
            torch.cuda.synchronize()
            graph_leg = {"value": args.steps / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g / args.steps,
                         "identical_image": bool(torch.equal(img_stream, g_out[0])),
                         "what": "the resident view (all kernels of a step incl. binning) captured once with "
                                 "torch.cuda.graph and replayed; same work as `value`, one launch per view"}
        except Exception as e:  # never let the extra leg break the bench line
            graph_leg = {"unavailable": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()

    # e2e through the public API with host buffers; asynchronous binning (rasterizer.binning): the first call of the
    # signature learns M synchronously (in the warm-up), the timed steps never read the pair count on the host
    import rasterizer
    from rasterizer import binning as _binning

    rasterizer.set_binning_mode("async")
    pv = PublicApiView(s, scene_np, bucket, peer)

    def e2e_step():
        pv.step()

    if args.only_resident:
        if rank == 0:
            print(json.dumps({"diagnostic": "--only-resident", "value": value, "ms_per_step": ms_per_step, "stages_ms": stages,
                              "gpu_launches": launches}))
        if world > 1:
            dist.destroy_process_group()
        return
    host_link = host_link_gbs(torch, pv)  # diagnostic: what this box's PCIe path gives the leg's own pinned buffers
    e2e_ms = timed_loop(torch, dist, world, e2e_step, args.steps, args.warmup, finish=pv.finish)
    e2e_value = world * args.steps / (e2e_ms * 1e-3)
    _binning.check()  # every asynchronous call of the e2e leg stayed within its pair-buffer capacity (raises otherwise)
    e2e_eager = {"value": e2e_value, "ms_per_step": e2e_ms / args.steps}
    # the same step with the view captured by rasterizer.graphs (public API): per step the same pinned-host copies in and
    # out, two graph launches instead of ~40 Python-dispatched calls — the eager path is host-bound on slow hosts
    e2e_mode = "eager"
    # default: on for a single process; at N > 1 the captured step contains the NCCL all-reduce and the symmetric-memory
    # barriers — measured working at 2 and 8 GPUs (profiles/r02/bench_v13_n8.json: 2063 against 1937 views/s eager), but a
    # teardown under live graphs hung once during development, so multi-rank runs replay graphs only with GSR_E2E_GRAPH=1
    if os.environ.get("GSR_E2E_GRAPH", "1" if world == 1 else "0") != "0":
        ok = pv.capture()
        if world > 1:  # every rank must take the same path (the captured step contains collectives)
            flag = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
            pv.graphed = ok
        if ok:
            g_ms = timed_loop(torch, dist, world, e2e_step, args.steps, args.warmup, finish=pv.finish)
            from rasterizer import graphs as _graphs

            _graphs.check()  # no replay overflowed the pair-buffer capacity baked into the graphs
            g_value = world * args.steps / (g_ms * 1e-3)
            if g_value > e2e_value:
                e2e_value, e2e_ms, e2e_mode = g_value, g_ms, "cuda_graph"
            e2e_eager["graphed_value"] = g_value
        # drop the captured graphs now (they hold captured NCCL work at N > 1: a process group must not be torn down under them)
        pv.graphed = False
        pv.graph_fwd = pv.graph_bwd = pv.graph_out = None
        import gc

        gc.collect()
        torch.cuda.synchronize()

    if rank == 0:
        peak, peak_src = peaks()
        ab = algorithmic_bytes(N, M, P, T)
        t_bwd = stages["blend_bwd"] * 1e-3
        achieved = ab["blend_bwd"] / t_bwd / 1e9
        bwd_kernel = {"p": "blend_backward_kernel", "s": "blend_backward_scan_kernel"}.get(
            os.environ.get("GSR_BWD_KERNEL", "tr")[:1], "blend_backward_tr_kernel<16, 3>")
        facts, facts_src = ncu_facts(bwd_kernel)
        roofline = {
            "kernel": f"{bwd_kernel} (gsr_rasterize_backward)", "bound": "hbm", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": facts["dram_bytes"] if facts else None,
            "traffic_source": facts_src, "peak_source": peak_src,
            "ncu_issue_slot_utilization": (facts["issue_active_pct"] / 100.0) if facts else None,
            "algorithmic_bytes_per_launch": ab["blend_bwd"], "avg_launch_ms": stages["blend_bwd"],
            "note": "blend kernels are FP32-issue bound (>=100 flop/B); a low HBM fraction is expected (SURVEY 8d). "
                    "Algorithmic bytes use the M the kernel actually walks (after exact tile culling), not the reference's "
                    "larger bounding-box M",
            "blend_fwd": {"achieved": ab["blend_fwd"] / (stages["blend_fwd"] * 1e-3) / 1e9, "avg_launch_ms": stages["blend_fwd"],
                          "algorithmic_bytes_per_launch": ab["blend_fwd"]},
            "whole_view": {"achieved": ab["total"] / (ms_per_step * 1e-3) / 1e9, "algorithmic_bytes": ab["total"]},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_text(args.workload, scene_np),
                       "num_intersects_after_exact_tile_culling": M,
                       "binning": "asynchronous: device-side pair count, capacity-bounded buffers, no host read in the timed "
                                  "regions (gsr_bin_gaussians_device; own radix sort + scan, no CUB)", "num_intersects_reference_bbox": M_ref, "pixels": P, "tiles": T, "parallelism": f"view-parallel x{world}",
                       "l2_policy": "inputs larger than L2 (192 MB SH coefficients + 192 MB SH gradients per view; "
                                    "no explicit flush)"},
            "stages_ms": stages,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": pv.h2d,
                    "d2h_bytes_per_step": pv.d2h, "host_link_gbs": host_link,
                    "mode": e2e_mode, "eager": e2e_eager,
                    "api": "rasterizer.project_gaussians + spherical_harmonics + rasterize_gaussians + autograd backward"
                           + (", captured once with rasterizer.graphs.capture_step and replayed (two graph launches per view; "
                              "the pinned-host copies in and out stay outside the graphs, every step); `eager` = the same step "
                              "dispatched from Python every view" if e2e_mode == "cuda_graph" else ""),
                    "note": "training-operator e2e: the Gaussian parameters and their 236 N bytes of gradients stay resident "
                            "in HBM (as in training); per view the camera matrices + upstream image / alpha gradients come "
                            "from pinned host memory and the rendered image + alpha go back to pinned host memory"},
            "gpu_launches": launches,
            "gpu_launches_note": "kernels of libgsr_b200.so launched inside the timed region of the resident leg, COUNTED by the "
                                 "library (gsr_launch_count(): one increment per launch site after its cudaGetLastError check); "
                                 "CUB / ATen kernels and memsets are not included",
            "clocks": clocks, "roofline": roofline,
        }
        if world > 1:
            line["config"]["host_numa_binding"] = (f"rank pinned to the {len(numa_cpus)} CPUs local to its GPU (NVML)"
                                                   if numa_cpus else "none")
            line["config"]["sh_gradient_exchange"] = "NVLink peer loads (symmetric memory)" if peer is not None else "NCCL all-gather"
        if DIAG_NOCOPY:
            line["diagnostic"] = "GSR_E2E_NOCOPY=1: e2e WITHOUT its host copies - not a reportable number"
        if graph_leg is not None:
            line["cuda_graph"] = graph_leg
        if xcheck is not None:
            line["exchange_check"] = xcheck
        if world == 1:
            line["fused_operator"] = fused_leg(torch, s, scene_np, args.steps, args.warmup)
            ref = ref_cuda_leg(torch, s, max(3, min(args.steps, 20)), max(2, min(args.warmup, 5)), ms_per_step)
            last = ref.pop("_last", None)
            if last is not None:  # same image from both implementations on this very run
                img_ours = rv.step()[0]
                d = (img_ours - last["out_img"]).abs()
                ref["image_frac_outside_1e-4"] = float((d > 1e-4 * last["out_img"].abs() + 1e-5).float().mean())
            line["ref_cuda_ext"] = ref
        if world == 1 and not args.no_cpu_baseline:
            v, desc, cores, _ = cpu_leg(scene_np, 25.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # the line is out: a process-group teardown that hangs (it did once with live CUDA graphs holding captured NCCL
        # work, before they were released explicitly above) must not hold the box
        wd = threading.Timer(45.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.barrier()
        dist.destroy_process_group()
        wd.cancel()


if __name__ == "__main__":
    main()
