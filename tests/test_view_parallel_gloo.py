"""CPU, world_size = 2, gloo: the host-side logic of the view-parallel path (flat gradient bucket, one
all-reduce, round-robin view assignment, densification statistics)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gaussian-splatting-toolkit_b200")


def _worker(rank, world, port, results):
    sys.path.insert(0, PKG)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rasterizer.view_parallel import (GradientBucket, all_reduce_densification_stats, floats_per_gaussian,
                                          view_for_rank)

    N, K = 1000, 16
    bucket = GradientBucket(N, K, device="cpu")
    assert bucket.flat.numel() == 59 * N == floats_per_gaussian(K) * N
    g = torch.Generator().manual_seed(100 + rank)
    grads = {"v_coeffs": torch.randn(N, K, 3, generator=g), "v_mean3d": torch.randn(N, 3, generator=g),
             "v_scale": torch.randn(N, 3, generator=g), "v_quat": torch.randn(N, 4, generator=g),
             "v_opacity": torch.randn(N, 1, generator=g)}
    # "kernel writes straight into the bucket" for one tensor, pack() for the others
    bucket["v_quat"].copy_(grads["v_quat"])
    grads["v_quat"] = bucket["v_quat"]
    bucket.pack(grads)
    bucket.all_reduce(average=False)
    # expected: sum over both ranks, recomputed locally from the seeds
    exp = {}
    for r in range(world):
        gr = torch.Generator().manual_seed(100 + r)
        for name, shape in (("v_coeffs", (N, K, 3)), ("v_mean3d", (N, 3)), ("v_scale", (N, 3)), ("v_quat", (N, 4)),
                            ("v_opacity", (N, 1))):
            t = torch.randn(*shape, generator=gr)
            exp[name] = exp.get(name, 0) + t
    ok = all(torch.allclose(bucket[name], exp[name], atol=1e-6) for name in exp)
    # averaging variant (DDP semantics)
    b2 = GradientBucket(N, K, device="cpu")
    b2.flat.fill_(float(rank + 1))
    b2.all_reduce(average=True)
    ok = ok and torch.allclose(b2.flat, torch.full_like(b2.flat, 1.5))
    # statistics: sum, sum, max
    a, b, c = torch.full((N,), float(rank)), torch.ones(N), torch.full((N,), float(10 * rank))
    all_reduce_densification_stats(a, b, c)
    ok = ok and float(a[0]) == 1.0 and float(b[0]) == 2.0 and float(c[0]) == 10.0
    # the same through the densification-statistics holder of rasterizer.densify (f2)
    from rasterizer.densify import DensifyStats

    st = DensifyStats()
    st.all_reduce()   # unset statistics: no-op
    st.xys_grad_norm, st.vis_counts, st.max_2Dsize = torch.full((N,), 0.5 + rank), torch.full((N,), 2.0), torch.full((N,), 0.1 * rank)
    st.all_reduce()
    ok = ok and float(st.xys_grad_norm[3]) == 2.0 and float(st.vis_counts[3]) == 4.0 and abs(float(st.max_2Dsize[3]) - 0.1) < 1e-7
    views = [view_for_rank(step, rank, world, 7) for step in range(7)]
    gathered = [None] * world
    dist.all_gather_object(gathered, views)
    if rank == 0:
        flat = [v for per_step in zip(*gathered) for v in per_step]
        ok = ok and sorted(flat) == sorted(list(range(7)) * 2) and flat[:7] == list(range(7))
    results[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def test_bucket_layout_is_contiguous_and_aligned():
    sys.path.insert(0, PKG)
    from rasterizer.view_parallel import GradientBucket, SEGMENTS

    b = GradientBucket(4096, 16, device="cpu")
    end = 0
    for name in SEGMENTS:
        lo, hi = b.offsets[name]
        assert lo == end and (lo * 4) % 16 == 0
        assert b[name].data_ptr() == b.flat.data_ptr() + 4 * lo and b[name].is_contiguous()
        end = hi
    assert end == b.flat.numel()


def test_gradient_bucket_segments_are_16_byte_aligned_for_any_n():
    """after densification N is arbitrary (odd); the projection adjoint needs a 16-byte aligned v_quat
    segment, so every segment offset is padded to a multiple of 4 floats."""
    import torch

    from rasterizer.view_parallel import SEGMENTS, GradientBucket, segment_shapes

    for n in (1, 7, 1001, 33_149):
        b = GradientBucket(n, sh_bases=16, device="cpu")
        shapes = segment_shapes(n, 16)
        for name in SEGMENTS:
            lo, hi = b.offsets[name]
            assert lo % 4 == 0 and (b[name].data_ptr() - b.flat.data_ptr()) == 4 * lo
            assert tuple(b[name].shape) == shapes[name] and hi - lo == b[name].numel()
        b["v_quat"].fill_(1.0)
        b["v_opacity"].fill_(2.0)
        assert float(b.flat.sum()) == 4.0 * n + 2.0 * n  # segments do not overlap, padding stays zero
