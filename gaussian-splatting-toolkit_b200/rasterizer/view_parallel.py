"""View-parallel multi-GPU plumbing: every rank holds the full, replicated Gaussian set and
renders its own training view; the only exchange step is ONE all-reduce (sum) of the parameter gradients —
59 floats per Gaussian (SH 48 + means 3 + scales 3 + quats 4 + opacity 1 at SH degree 3) — after backward.

Reference precedent: torch DDP's bucketed all-reduce of the `gauss_params` gradients
(gs_toolkit/pipelines/base_pipeline.py:202-207, implicit in loss.backward(), engine/trainer.py:498); views
are drawn per rank because `random` is seeded with seed + global_rank (scripts/train.py:54).  DDP averages;
`GradientBucket.all_reduce(average=True)` does the same.

Here the gradients live in ONE flat FP32 buffer whose segments are handed to the backward kernels as their
output pointers (the C ABI takes raw output pointers), so there is no per-tensor bucketing or copy: one
`ncclAllReduce` over the whole buffer, or two (SH segment first, on a side stream, overlapping the projection
backward) with `overlap=True`.  Works on any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

# segment order: the SH gradient (81 % of the bytes) first so that its all-reduce can start first
SEGMENTS = ("v_coeffs", "v_mean3d", "v_scale", "v_quat", "v_opacity")


def segment_shapes(num_points: int, sh_bases: int) -> Dict[str, Tuple[int, ...]]:
    return {"v_coeffs": (num_points, sh_bases, 3), "v_mean3d": (num_points, 3), "v_scale": (num_points, 3),
            "v_quat": (num_points, 4), "v_opacity": (num_points, 1)}


def floats_per_gaussian(sh_bases: int) -> int:
    return 3 * sh_bases + 3 + 3 + 4 + 1


class GradientBucket:
    """Flat gradient buffer + typed views into it."""

    def __init__(self, num_points: int, sh_bases: int = 16, device="cuda", symmetric: bool = False, group=None):
        """`symmetric=True` (NCCL job on NVLink-connected GPUs): the buffer is allocated in symmetric memory and every rank
        maps its peers' buckets (`peer_tails`), so that the non-SH segment can be all-reduced by this package's own
        reduce-scatter / all-gather kernels over peer memory (GradientExchange) instead of NCCL.  Falls back to an ordinary
        allocation when symmetric memory is unavailable.  Opt-in: measured on 2 / 4 / 8 B200s the load-based kernels take
        0.17 / 0.28 / 0.40 ms for the 44 MB tail of 1 M Gaussians against 0.13 / 0.20 / 0.33 ms for NCCL's all-reduce running
        beside the peer-load SH adjoint (profiles/r02/exchange_tail_own_vs_nccl.txt), so NCCL stays the default."""
        self.num_points, self.sh_bases = num_points, sh_bases
        self.hdl, self.peer_tails = None, None
        shapes = segment_shapes(num_points, sh_bases)
        # every segment starts on a 16-byte boundary for ANY N (the projection adjoint requires a 16-byte aligned
        # v_quat, project.cu): offsets are rounded up to a multiple of 4 floats, the <= 3 pad floats stay zero and
        # simply ride along in the all-reduce
        sizes, starts, off = {}, {}, 0
        for name, shape in shapes.items():
            n = 1
            for d in shape:
                n *= d
            starts[name], sizes[name] = off, n
            off = (off + n + 3) // 4 * 4
        self.flat = None
        if symmetric and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                dev = device if isinstance(device, torch.device) else torch.device(device)
                if dev.index is None:
                    dev = torch.device("cuda", torch.cuda.current_device())
                flat = symm_mem.empty(off, dtype=torch.float32, device=dev)
                flat.zero_()
                self.hdl = symm_mem.rendezvous(flat, group if group is not None else dist.group.WORLD)
                self.flat = flat
            except Exception:  # pragma: no cover - depends on the platform
                self.hdl, self.flat = None, None
        if self.flat is None:
            self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        self.offsets: Dict[str, Tuple[int, int]] = {}
        for name, shape in shapes.items():
            lo, n = starts[name], sizes[name]
            self.views[name] = self.flat[lo:lo + n].view(shape)
            self.offsets[name] = (lo, lo + n)
        if self.hdl is not None:
            hi = self.offsets["v_coeffs"][1]
            hi = (hi + 3) // 4 * 4  # = start of v_mean3d
            self.tail_start, self.tail_len = hi, off - hi
            self.peer_tails = [self.hdl.get_buffer(r, (self.tail_len,), torch.float32, hi) for r in range(self.hdl.world_size)]

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]

    def pack(self, grads: Dict[str, torch.Tensor]) -> None:
        """Copy gradients produced elsewhere into the bucket (no-op for tensors that already alias it)."""
        for name in SEGMENTS:
            g = grads[name]
            v = self.views[name]
            if g.data_ptr() != v.data_ptr():
                v.copy_(g.reshape(v.shape))

    def all_reduce(self, group=None, average: bool = False, overlap_stream: Optional["torch.cuda.Stream"] = None):
        """Sum (or average) the bucket over the ranks.  With `overlap_stream` the SH segment is reduced on that
        stream first (call after the SH backward has been enqueued, before the projection backward) and the
        rest on the current stream; returns when both are enqueued."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        world = dist.get_world_size(group)
        if overlap_stream is None:
            dist.all_reduce(self.flat, group=group)
        else:
            lo, hi = self.offsets["v_coeffs"]
            cur = torch.cuda.current_stream()
            overlap_stream.wait_stream(cur)
            with torch.cuda.stream(overlap_stream):
                dist.all_reduce(self.flat[lo:hi], group=group)
            dist.all_reduce(self.flat[hi:], group=group)
            cur.wait_stream(overlap_stream)
        if average:
            self.flat.div_(world)


class PeerColorGrads:
    """NVLink peer-memory mailbox for the colour gradients: every rank owns one symmetric-memory buffer
    [N*3 colour-gradient floats | 3 camera-centre floats | pad] that all peers map into their address space
    (torch.distributed._symmetric_memory: CUDA VMM handles exchanged once at rendezvous).  The multi-view SH adjoint
    kernel then LOADS the peers' colour gradients straight over NVLink while it writes the summed SH gradient —
    all-gather and compute are one kernel, there is no staging buffer and no NCCL call for the SH segment.
    Two device-side barriers per exchange (signal pads, stream-ordered, no host sync): "all published" before the
    kernel, "all consumed" after it."""

    def __init__(self, num_points: int, group=None, device=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.num_points = num_points
        group = group if group is not None else dist.group.WORLD
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(num_points * 3 + 4, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.world, self.rank = self.hdl.world_size, self.hdl.rank
        self.local_rgb = self.buf[: 3 * num_points].view(num_points, 3)
        self.local_cam = self.buf[3 * num_points: 3 * num_points + 3]
        self.peer_rgb = [self.hdl.get_buffer(r, (num_points, 3), torch.float32, 0) for r in range(self.world)]
        self.peer_cam = [self.hdl.get_buffer(r, (3,), torch.float32, 3 * num_points) for r in range(self.world)]

    @staticmethod
    def try_create(num_points: int, group=None, device=None):
        """None when symmetric memory is unavailable (no P2P mapping): callers fall back to the NCCL all-gather."""
        try:
            return PeerColorGrads(num_points, group, device)
        except Exception:  # pragma: no cover - depends on the platform
            return None


def exchange_gradients(bucket: GradientBucket, v_rgb_sh: torch.Tensor, means3d: torch.Tensor, cam_pos: torch.Tensor,
                       degree: int, degrees_to_use: int, group=None, average: bool = False,
                       peer: "PeerColorGrads" = None) -> None:
    """The view-parallel gradient exchange with the SH segment computed instead of communicated.

    v_coeffs of rank r is the outer product Y(means - cam_r) (x) v_rgb_r, so the SUM over ranks can be evaluated
    locally from every rank's 3-float colour gradient: one kernel (gsr_compute_sh_backward_multiview) reads all ranks'
    v_rgb_sh (12 B / Gaussian / rank) and camera centres and writes the summed SH gradient into the bucket; only the
    remaining 11 floats / Gaussian (means, scales, quats, opacity) go through an all-reduce: 12 W + 44 bytes per
    Gaussian on the wire instead of 236 (W = world size).  With `peer` (PeerColorGrads) the kernel loads the peers'
    colour gradients directly over NVLink from symmetric memory; without it they are all-gathered with NCCL first.
    The caller must have put v_mean3d, v_scale, v_quat, v_opacity into `bucket` (the SH segment is overwritten)."""
    from . import cuda as _C

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        _C.compute_sh_backward_multiview(degree, degrees_to_use, means3d, cam_pos.reshape(1, 3).contiguous(),
                                         [v_rgb_sh.contiguous()], out=bucket["v_coeffs"])
        return
    world = dist.get_world_size(group)
    lo, hi = bucket.offsets["v_coeffs"]
    if peer is not None:
        # NVLink peer-memory path: publish, barrier, [NCCL all-reduce of the 11 N rest] + fused gather/adjoint kernel
        if v_rgb_sh.data_ptr() != peer.local_rgb.data_ptr():
            peer.local_rgb.copy_(v_rgb_sh)
        peer.local_cam.copy_(cam_pos.reshape(3))
        peer.hdl.barrier(channel=0)
        dist.all_reduce(bucket.flat[hi:], group=group)
        _C.compute_sh_backward_multiview(degree, degrees_to_use, means3d, peer.peer_cam, peer.peer_rgb,
                                         out=bucket["v_coeffs"])
        peer.hdl.barrier(channel=1)
    else:
        v_all = torch.empty((world,) + tuple(v_rgb_sh.shape), dtype=torch.float32, device=v_rgb_sh.device)
        cams = torch.empty((world, 3), dtype=torch.float32, device=v_rgb_sh.device)
        dist.all_gather_into_tensor(v_all, v_rgb_sh.contiguous(), group=group)
        dist.all_gather_into_tensor(cams, cam_pos.reshape(3).contiguous(), group=group)
        dist.all_reduce(bucket.flat[hi:], group=group)
        _C.compute_sh_backward_multiview(degree, degrees_to_use, means3d, cams, v_all, out=bucket["v_coeffs"])
    if average:
        bucket.flat.div_(world)


class _SphericalHarmonicsViewParallel(torch.autograd.Function):
    """spherical_harmonics whose backward returns the coefficient gradient already SUMMED over the ranks of
    `group` (all-gather of the 3-float colour gradients + multi-view adjoint kernel), see exchange_gradients."""

    @staticmethod
    def forward(ctx, degrees_to_use, means3d, cam_pos, coeffs, group, average, peer=None):
        from . import cuda as _C
        from .sh import deg_from_sh

        ctx.degrees_to_use, ctx.degree, ctx.group, ctx.average = degrees_to_use, deg_from_sh(coeffs.shape[-2]), group, average
        ctx.peer = peer
        viewdirs = (means3d - cam_pos.reshape(1, 3)).contiguous()
        ctx.save_for_backward(means3d, cam_pos)
        return _C.compute_sh_forward(coeffs.shape[0], ctx.degree, degrees_to_use, viewdirs, coeffs)

    @staticmethod
    def backward(ctx, v_colors):
        from . import cuda as _C

        means3d, cam_pos = ctx.saved_tensors
        v_colors = v_colors.contiguous()
        world = dist.get_world_size(ctx.group) if (dist.is_available() and dist.is_initialized()) else 1
        peer = ctx.peer
        if world > 1 and peer is not None:
            # NVLink peer-memory path (PeerColorGrads): publish, barrier, the kernel loads the peers' colour gradients
            peer.local_rgb.copy_(v_colors)
            peer.local_cam.copy_(cam_pos.reshape(3))
            peer.hdl.barrier(channel=0)
            v_coeffs = _C.compute_sh_backward_multiview(ctx.degree, ctx.degrees_to_use, means3d, peer.peer_cam, peer.peer_rgb)
            peer.hdl.barrier(channel=1)
            if ctx.average:
                v_coeffs.div_(world)
            return None, None, None, v_coeffs, None, None, None
        if world == 1:
            v_all, cams = [v_colors], cam_pos.reshape(1, 3).contiguous()
        else:
            v_all = torch.empty((world,) + tuple(v_colors.shape), dtype=torch.float32, device=v_colors.device)
            cams = torch.empty((world, 3), dtype=torch.float32, device=v_colors.device)
            dist.all_gather_into_tensor(v_all, v_colors, group=ctx.group)
            dist.all_gather_into_tensor(cams, cam_pos.reshape(3).contiguous(), group=ctx.group)
        v_coeffs = _C.compute_sh_backward_multiview(ctx.degree, ctx.degrees_to_use, means3d, cams, v_all)
        if ctx.average and world > 1:
            v_coeffs.div_(world)
        return None, None, None, v_coeffs, None, None, None


def spherical_harmonics_view_parallel(degrees_to_use: int, means3d: torch.Tensor, cam_pos: torch.Tensor,
                                      coeffs: torch.Tensor, group=None, average: bool = False,
                                      peer: "PeerColorGrads" = None) -> torch.Tensor:
    """Drop-in for `spherical_harmonics(degrees_to_use, means3d - cam_pos, coeffs)` in a view-parallel job:
    same colours; `coeffs.grad` comes out already reduced over `group`, so the SH coefficients must be left out
    of the gradient all-reduce (only 11 of the 59 floats per Gaussian remain to be reduced).  With `peer` (a
    PeerColorGrads mailbox) the backward loads the peers' colour gradients over NVLink instead of all-gathering them."""
    return _SphericalHarmonicsViewParallel.apply(degrees_to_use, means3d.detach().contiguous(), cam_pos.contiguous(),
                                                 coeffs.contiguous(), group, average, peer)


class GradientExchange:
    """The exchange of `exchange_gradients` split in two so that it overlaps the tail of the backward pass:

        ex.start_sh(v_rgb_sh, cam_pos)     # right after the blend adjoint: publish colour gradients, fused
                                           # gather + SH adjoint on a SIDE stream (NVLink peer loads)
        ... projection adjoint writes v_mean3d / v_scale / v_quat into the bucket on the main stream ...
        ex.finish()                        # NCCL all-reduce of the 11 N remaining floats on the main stream,
                                           # concurrently with the side-stream kernel; then join

    Without a PeerColorGrads mailbox (no symmetric memory) start_sh() only remembers its arguments and finish()
    runs the NCCL all-gather variant."""

    def __init__(self, bucket: GradientBucket, means3d: torch.Tensor, degree: int, peer: "PeerColorGrads" = None,
                 group=None):
        self.bucket, self.means3d, self.degree, self.peer, self.group = bucket, means3d, degree, peer, group
        self.side = torch.cuda.Stream(means3d.device) if peer is not None else None
        self._pending = None
        # optional per-component device timing (bench.py): lists of CUDA-event tuples, see breakdown_ms()
        self.timing = False
        self._ev_side, self._ev_main = [], []

    def _ev(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def breakdown_ms(self) -> Dict[str, float]:
        """Average device times of the exchange's components over the timed calls (call after a synchronize):
        side stream = publish colour gradients + 'all published' barrier, then the peer-load multi-view SH adjoint;
        main stream = NCCL all-reduce of the 11 N other floats, then the join with the side stream + 'all consumed' barrier."""
        out = {}
        if self._ev_side:
            n = len(self._ev_side)
            out["side.publish_and_barrier"] = sum(a.elapsed_time(b) for a, b, _ in self._ev_side) / n
            out["side.sh_adjoint_multiview_peer_loads"] = sum(b.elapsed_time(c) for _, b, c in self._ev_side) / n
        if self._ev_main:
            n = len(self._ev_main)
            key = "main.peer_reduce_scatter_allgather_11N" if self.bucket.hdl is not None else "main.nccl_allreduce_11N"
            out[key] = sum(a.elapsed_time(b) for a, b, _ in self._ev_main) / n
            out["main.join_side_and_barrier"] = sum(b.elapsed_time(c) for _, b, c in self._ev_main) / n
        return out

    def start_sh(self, v_rgb_sh: torch.Tensor, cam_pos: torch.Tensor, degrees_to_use: int) -> None:
        from . import cuda as _C

        if self.peer is None:
            self._pending = (v_rgb_sh, cam_pos, degrees_to_use)
            return
        peer, cur = self.peer, torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            e0 = self._ev() if self.timing else None
            if v_rgb_sh.data_ptr() != peer.local_rgb.data_ptr():
                peer.local_rgb.copy_(v_rgb_sh)
                v_rgb_sh.record_stream(self.side)
            peer.local_cam.copy_(cam_pos.reshape(3))
            peer.hdl.barrier(channel=0)
            e1 = self._ev() if self.timing else None
            _C.compute_sh_backward_multiview(self.degree, degrees_to_use, self.means3d, peer.peer_cam, peer.peer_rgb,
                                             out=self.bucket["v_coeffs"])
            if self.timing:
                self._ev_side.append((e0, e1, self._ev()))

    def finish(self, average: bool = False) -> None:
        bucket = self.bucket
        if self.peer is None:
            v_rgb_sh, cam_pos, degrees_to_use = self._pending
            self._pending = None
            exchange_gradients(bucket, v_rgb_sh, self.means3d, cam_pos, self.degree, degrees_to_use, self.group, average)
            return
        hi = bucket.offsets["v_coeffs"][1]
        e0 = self._ev() if self.timing else None
        if bucket.hdl is not None:
            # own all-reduce of the 11 N non-SH floats over NVLink peer memory: reduce-scatter + all-gather kernels between
            # device-side barriers ("every tail written" / "every slice reduced"; the final "all consumed" barrier below also
            # covers the gathers).  A GPU receives 2 x (W-1)/W x 44 N bytes; no NCCL call on the path.
            from . import cuda as _C

            rank = bucket.hdl.rank
            bucket.hdl.barrier(channel=0)
            _C.peer_all_reduce_phase("reduce_scatter", rank, bucket.peer_tails, bucket.tail_len)
            bucket.hdl.barrier(channel=1)
            _C.peer_all_reduce_phase("all_gather", rank, bucket.peer_tails, bucket.tail_len)
        else:
            dist.all_reduce(bucket.flat[hi:], group=self.group)
        e1 = self._ev() if self.timing else None
        torch.cuda.current_stream().wait_stream(self.side)
        self.peer.hdl.barrier(channel=1)
        if self.timing:
            self._ev_main.append((e0, e1, self._ev()))
        if average:
            bucket.flat.div_(dist.get_world_size(self.group))


class PushGradientExchange:
    """GradientExchange with PUSHED data (opt-in, NVLink-connected GPUs): everything that crosses NVLink is a remote STORE
    into a peer's symmetric staging buffer — posted, no round trip per access — and every reduction reads local memory.

        start_sh   side stream: push this rank's colour gradients + camera centre into slot `rank` of every rank's staging
                   buffer; barrier; multi-view SH adjoint over the W LOCAL slots -> bucket["v_coeffs"]
        finish     main stream: push slice w of this rank's 11 N non-SH floats into slot `rank` of rank w's staging buffer;
                   barrier; sum the W local slots of this rank's slice (fixed order: replicas bit-identical) and store the
                   total into slice `rank` of EVERY rank's bucket; join the side stream; barrier ("all gathered" + "all
                   consumed")

    Needs a symmetric bucket (GradientBucket(symmetric=True)).  Same interface as GradientExchange."""

    def __init__(self, bucket: GradientBucket, means3d: torch.Tensor, degree: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        if bucket.hdl is None:
            raise RuntimeError("PushGradientExchange needs GradientBucket(symmetric=True)")
        self.bucket, self.means3d, self.degree, self.group = bucket, means3d, degree, group
        W, r, N = bucket.hdl.world_size, bucket.hdl.rank, bucket.num_points
        self.world, self.rank = W, r
        self.rgb_floats = 3 * N
        self.cam_off = (3 * N + 3) // 4 * 4            # camera centre behind the colour gradients, 16-byte aligned
        S = self.cam_off + 4                           # floats per colour slot
        L = ((bucket.tail_len + W - 1) // W + 3) // 4 * 4   # floats per slice of the non-SH tail
        self.S, self.L = S, L
        self.n_mine = max(0, min(L, bucket.tail_len - r * L))
        dev = means3d.device
        self.stage = symm_mem.empty(W * S + W * L, dtype=torch.float32, device=dev)
        self.stage.zero_()
        self.hdl = symm_mem.rendezvous(self.stage, group if group is not None else dist.group.WORLD)
        # destinations on every rank w: MY colour slot, MY camera slot, MY tail slot; and slice `rank` of w's bucket tail
        self.dst_rgb = [self.hdl.get_buffer(w, (S,), torch.float32, r * S) for w in range(W)]
        self.dst_cam = [self.hdl.get_buffer(w, (4,), torch.float32, r * S + self.cam_off) for w in range(W)]
        self.dst_tail = [self.hdl.get_buffer(w, (L,), torch.float32, W * S + r * L) for w in range(W)]
        self.dst_slice = [bucket.peer_tails[w][r * L:] for w in range(W)] if self.n_mine > 0 else None
        # local views of the W slots
        self.loc_rgb = [self.stage[w * S: w * S + 3 * N].view(N, 3) for w in range(W)]
        self.loc_cam = [self.stage[w * S + self.cam_off: w * S + self.cam_off + 3] for w in range(W)]
        self.loc_tail = self.stage[W * S:]
        self.cam4 = torch.zeros(4, dtype=torch.float32, device=dev)
        self.side = torch.cuda.Stream(dev)
        self.peer = self  # bench.py looks at `.peer is not None` for the breakdown
        self.timing = False
        self._ev_side, self._ev_main = [], []

    def _ev(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def breakdown_ms(self) -> Dict[str, float]:
        out = {}
        if self._ev_side:
            n = len(self._ev_side)
            out["side.push_rgb_and_barrier"] = sum(a.elapsed_time(b) for a, b, _ in self._ev_side) / n
            out["side.sh_adjoint_multiview_local_slots"] = sum(b.elapsed_time(c) for _, b, c in self._ev_side) / n
        if self._ev_main:
            n = len(self._ev_main)
            out["main.push_slices_barrier_reduce_broadcast_11N"] = sum(a.elapsed_time(b) for a, b, _ in self._ev_main) / n
            out["main.join_side_and_barrier"] = sum(b.elapsed_time(c) for _, b, c in self._ev_main) / n
        return out

    def start_sh(self, v_rgb_sh: torch.Tensor, cam_pos: torch.Tensor, degrees_to_use: int) -> None:
        from . import cuda as _C

        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            e0 = self._ev() if self.timing else None
            v = v_rgb_sh.contiguous()
            _C.peer_push(self.dst_rgb, v.view(-1), 0, self.rgb_floats)
            v.record_stream(self.side)
            self.cam4[:3].copy_(cam_pos.reshape(3))
            _C.peer_push(self.dst_cam, self.cam4, 0, 4)
            self.hdl.barrier(channel=0)
            e1 = self._ev() if self.timing else None
            _C.compute_sh_backward_multiview(self.degree, degrees_to_use, self.means3d, self.loc_cam, self.loc_rgb,
                                             out=self.bucket["v_coeffs"])
            if self.timing:
                self._ev_side.append((e0, e1, self._ev()))

    def finish(self, average: bool = False) -> None:
        from . import cuda as _C

        bucket = self.bucket
        e0 = self._ev() if self.timing else None
        tail = bucket.flat[bucket.tail_start:]
        _C.peer_push(self.dst_tail, tail, self.L, self.L, bucket.tail_len)
        self.hdl.barrier(channel=1)
        if self.n_mine > 0:
            _C.peer_reduce_broadcast(self.dst_slice, self.loc_tail, self.L, self.n_mine)
        e1 = self._ev() if self.timing else None
        torch.cuda.current_stream().wait_stream(self.side)
        self.hdl.barrier(channel=2)
        if self.timing:
            self._ev_main.append((e0, e1, self._ev()))
        if average:
            bucket.flat.div_(self.world)


def bind_process_to_gpu_numa(device_index: int) -> Optional[list]:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index` (its NUMA node), so that the
    pinned host buffers it allocates afterwards (first touch) and its H2D / D2H traffic stay on the GPU's own socket —
    with 8 ranks streaming images over PCIe, cross-socket traffic is the first thing to saturate.  Returns the CPU
    list, or None when NVML / the affinity call is unavailable (nothing is changed then)."""
    try:
        import os

        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # pragma: no cover - depends on the platform
        return None


def view_for_rank(step: int, rank: int, world_size: int, num_views: int) -> int:
    """Round-robin view assignment: rank r renders view (step * world + r) mod num_views."""
    return (step * world_size + rank) % num_views


def all_reduce_densification_stats(xys_grad_norm: torch.Tensor, vis_counts: torch.Tensor,
                                   max_2dsize: torch.Tensor, group=None) -> None:
    """Replicas must take identical split / cull decisions: sum, sum, max over ranks of the statistics the
    reference model keeps per Gaussian (gs_toolkit/models/vanilla_gs.py:344-372)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(xys_grad_norm, group=group)
    dist.all_reduce(vis_counts, group=group)
    dist.all_reduce(max_2dsize, op=dist.ReduceOp.MAX, group=group)
