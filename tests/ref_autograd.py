"""TEST INFRASTRUCTURE: autograd wrappers over a native bindings module with the reference's 11 names (the compiled
reference extension from oracle/_ref), so that a model-style step (torch glue + operators + autograd, the way
gs_toolkit/models/vanilla_gs.py:759-855 runs) can be timed and compared with this package.  The orchestration inside
each Function follows the reference wrappers (rasterizer/sh.py:62-97, project_gaussians.py:83-232,
rasterize.py:92-247, utils.py:106-182) including torch.cumsum + .item(), torch.sort and torch.gather."""
import torch


def make_ops(C):
    class SH(torch.autograd.Function):
        @staticmethod
        def forward(ctx, degrees_to_use, viewdirs, coeffs):
            ctx.meta = (degrees_to_use, {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[coeffs.shape[-2]])
            ctx.save_for_backward(viewdirs)
            return C.compute_sh_forward(coeffs.shape[0], ctx.meta[1], degrees_to_use, viewdirs, coeffs)

        @staticmethod
        def backward(ctx, v_colors):
            (viewdirs,) = ctx.saved_tensors
            return None, None, C.compute_sh_backward(v_colors.shape[0], ctx.meta[1], ctx.meta[0], viewdirs, v_colors.contiguous())

    class Project(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, H, W, bw, clip):
            n = means3d.shape[0]
            cov3d, xys, depths, radii, conics, comp, nth = C.project_gaussians_forward(
                n, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, H, W, bw, clip)
            ctx.meta = (n, glob_scale, fx, fy, cx, cy, H, W)
            ctx.save_for_backward(means3d, scales, quats, viewmat, projmat, cov3d, radii, conics, comp)
            return xys, depths, radii, conics, comp, nth, cov3d

        @staticmethod
        def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_comp, v_nth, v_cov3d):
            means3d, scales, quats, viewmat, projmat, cov3d, radii, conics, comp = ctx.saved_tensors
            n, glob_scale, fx, fy, cx, cy, H, W = ctx.meta
            _, _, v_mean, v_scale, v_quat = C.project_gaussians_backward(
                n, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, H, W, cov3d, radii, conics, comp,
                v_xys.contiguous(), v_depths.contiguous(), v_conics.contiguous(), v_comp.contiguous())
            return (v_mean, v_scale, None, v_quat) + (None,) * 10

    class Rasterize(torch.autograd.Function):
        @staticmethod
        def forward(ctx, xys, depths, radii, conics, nth, colors, opacity, H, W, bw, background):
            n = xys.size(0)
            tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
            cum = torch.cumsum(nth, dim=0, dtype=torch.int32)
            M = cum[-1].item()
            isect, gids = C.map_gaussian_to_intersects(n, M, xys, depths, radii, cum, tb, bw)
            ks, order = torch.sort(isect)
            vs = torch.gather(gids, 0, order)
            bins = C.get_tile_bin_edges(M, ks, tb)
            img, fT, fi = C.rasterize_forward(tb, (bw, bw, 1), (W, H, 1), vs, bins, xys, conics, colors, opacity, background)
            ctx.meta = (H, W, bw)
            ctx.save_for_backward(vs, bins, xys, conics, colors, opacity, background, fT, fi)
            return img, 1 - fT

        @staticmethod
        def backward(ctx, v_img, v_alpha):
            vs, bins, xys, conics, colors, opacity, background, fT, fi = ctx.saved_tensors
            H, W, bw = ctx.meta
            v_xy, v_conic, v_colors, v_opacity = C.rasterize_backward(H, W, bw, vs, bins, xys, conics, colors, opacity,
                                                                      background, fT, fi, v_img.contiguous(), v_alpha.contiguous())
            return v_xy, None, None, v_conic, None, v_colors, v_opacity, None, None, None, None

    return SH.apply, Project.apply, Rasterize.apply
