// ext.cpp — `rasterizer.csrc`-compatible pybind module on top of the C ABI of libgsr_b200.so.
//
// The reference's native module (gs_toolkit/gs_components/rasterizer/cuda/csrc/ext.cpp:6-17) exports eleven functions
// with torch::Tensor signatures (bindings.h:19-115); its Python loader does `from rasterizer import csrc as _C`
// (rasterizer/cuda/_backend.py:61-63).  This file exports the same eleven names with the same positional
// arguments and the same returned tuples, each a thin shim: check the inputs the way the reference's CHECK_INPUT
// does, allocate the outputs from torch's caching allocator, take torch's CURRENT stream and call the C entry point
// (include/gsr_b200.h).  Dropped into the reference package as `rasterizer/csrc*.so` it lets the UNMODIFIED reference
// wrappers (rasterizer/{project_gaussians,rasterize,sh,utils}.py) and models run on the B200 kernels
// (tests/test_gpu_dropin_reference_callers.py).  No kernel lives here; this is host glue only.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <tuple>

#include "../../../include/gsr_b200.h"

namespace {

#define GSR_CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define GSR_CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define GSR_CHECK_INPUT(x) \
  GSR_CHECK_CUDA(x);       \
  GSR_CHECK_CONTIGUOUS(x)
#define GSR_DEVICE_GUARD(t) const at::cuda::OptionalCUDAGuard device_guard(device_of(t))

inline void *cur_stream() { return (void *)at::cuda::getCurrentCUDAStream().stream(); }
inline void ok(int rc, const char *what) {
  TORCH_CHECK(rc == GSR_OK, "libgsr_b200 ", what, " failed (status ", rc, "): ", gsr_last_error());
}
inline const float *F(const torch::Tensor &t) { return t.data_ptr<float>(); }
inline float *Fm(torch::Tensor &t) { return t.data_ptr<float>(); }
inline const int32_t *I(const torch::Tensor &t) { return t.data_ptr<int32_t>(); }
inline torch::Tensor f32(const torch::Tensor &like, at::IntArrayRef shape) {
  return torch::empty(shape, like.options().dtype(torch::kFloat32));
}
inline torch::Tensor i32(const torch::Tensor &like, at::IntArrayRef shape) {
  return torch::empty(shape, like.options().dtype(torch::kInt32));
}
using T3 = std::tuple<int, int, int>;

// bindings.cu:39-56
std::tuple<torch::Tensor, torch::Tensor> compute_cov2d_bounds_tensor(const int num_pts, torch::Tensor &covs2d) {
  GSR_DEVICE_GUARD(covs2d);
  GSR_CHECK_INPUT(covs2d);
  torch::Tensor conics = f32(covs2d, {num_pts, covs2d.size(1)});
  torch::Tensor radii = f32(covs2d, {num_pts, 1});
  ok(gsr_compute_cov2d_bounds(num_pts, F(covs2d), Fm(conics), Fm(radii), cur_stream()), "compute_cov2d_bounds");
  return std::make_tuple(conics, radii);
}

// bindings.cu:58-79
torch::Tensor compute_sh_forward_tensor(unsigned num_points, unsigned degree, unsigned degrees_to_use,
                                        torch::Tensor &viewdirs, torch::Tensor &coeffs) {
  GSR_DEVICE_GUARD(viewdirs);
  unsigned num_bases = (degree + 1) * (degree + 1);
  TORCH_CHECK(coeffs.ndimension() == 3 && coeffs.size(0) == num_points && coeffs.size(1) == num_bases &&
                  coeffs.size(2) == 3,
              "coeffs must have dimensions (N, D, 3)");
  GSR_CHECK_INPUT(viewdirs);
  GSR_CHECK_INPUT(coeffs);
  torch::Tensor colors = f32(coeffs, {num_points, 3});
  ok(gsr_compute_sh_forward((int)num_points, (int)degree, (int)degrees_to_use, F(viewdirs), F(coeffs), Fm(colors),
                            cur_stream()),
     "compute_sh_forward");
  return colors;
}

// bindings.cu:81-103
torch::Tensor compute_sh_backward_tensor(unsigned num_points, unsigned degree, unsigned degrees_to_use,
                                         torch::Tensor &viewdirs, torch::Tensor &v_colors) {
  GSR_DEVICE_GUARD(viewdirs);
  TORCH_CHECK(viewdirs.ndimension() == 2 && viewdirs.size(0) == num_points && viewdirs.size(1) == 3,
              "viewdirs must have dimensions (N, 3)");
  TORCH_CHECK(v_colors.ndimension() == 2 && v_colors.size(0) == num_points && v_colors.size(1) == 3,
              "v_colors must have dimensions (N, 3)");
  GSR_CHECK_INPUT(viewdirs);
  GSR_CHECK_INPUT(v_colors);
  unsigned num_bases = (degree + 1) * (degree + 1);
  torch::Tensor v_coeffs = f32(v_colors, {num_points, num_bases, 3});
  ok(gsr_compute_sh_backward((int)num_points, (int)degree, (int)degrees_to_use, F(viewdirs), F(v_colors),
                             Fm(v_coeffs), cur_stream()),
     "compute_sh_backward");
  return v_coeffs;
}

// bindings.cu:105-159
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
project_gaussians_forward_tensor(const int num_points, torch::Tensor &means3d, torch::Tensor &scales,
                                 const float glob_scale, torch::Tensor &quats, torch::Tensor &viewmat,
                                 torch::Tensor &projmat, const float fx, const float fy, const float cx,
                                 const float cy, const unsigned img_height, const unsigned img_width,
                                 const unsigned block_width, const float clip_thresh) {
  GSR_DEVICE_GUARD(means3d);
  GSR_CHECK_INPUT(means3d);
  GSR_CHECK_INPUT(scales);
  GSR_CHECK_INPUT(quats);
  GSR_CHECK_INPUT(viewmat);
  GSR_CHECK_INPUT(projmat);
  torch::Tensor cov3d = f32(means3d, {num_points, 6});
  torch::Tensor xys = f32(means3d, {num_points, 2});
  torch::Tensor depths = f32(means3d, {num_points});
  torch::Tensor radii = i32(means3d, {num_points});
  torch::Tensor conics = f32(means3d, {num_points, 3});
  torch::Tensor compensation = f32(means3d, {num_points});
  torch::Tensor num_tiles_hit = i32(means3d, {num_points});
  ok(gsr_project_gaussians_forward(num_points, F(means3d), F(scales), glob_scale, F(quats), F(viewmat), F(projmat), fx,
                                   fy, cx, cy, img_height, img_width, block_width, clip_thresh, Fm(cov3d), Fm(xys),
                                   Fm(depths), radii.data_ptr<int32_t>(), Fm(conics), Fm(compensation),
                                   num_tiles_hit.data_ptr<int32_t>(), cur_stream()),
     "project_gaussians_forward");
  return std::make_tuple(cov3d, xys, depths, radii, conics, compensation, num_tiles_hit);
}

// bindings.cu:161-216
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
project_gaussians_backward_tensor(const int num_points, torch::Tensor &means3d, torch::Tensor &scales,
                                  const float glob_scale, torch::Tensor &quats, torch::Tensor &viewmat,
                                  torch::Tensor &projmat, const float fx, const float fy, const float cx,
                                  const float cy, const unsigned img_height, const unsigned img_width,
                                  torch::Tensor &cov3d, torch::Tensor &radii, torch::Tensor &conics,
                                  torch::Tensor &compensation, torch::Tensor &v_xy, torch::Tensor &v_depth,
                                  torch::Tensor &v_conic, torch::Tensor &v_compensation) {
  GSR_DEVICE_GUARD(means3d);
  GSR_CHECK_INPUT(means3d);
  GSR_CHECK_INPUT(scales);
  GSR_CHECK_INPUT(quats);
  GSR_CHECK_INPUT(viewmat);
  GSR_CHECK_INPUT(projmat);
  GSR_CHECK_INPUT(cov3d);
  GSR_CHECK_INPUT(radii);
  GSR_CHECK_INPUT(conics);
  GSR_CHECK_INPUT(compensation);
  GSR_CHECK_INPUT(v_xy);
  GSR_CHECK_INPUT(v_depth);
  GSR_CHECK_INPUT(v_conic);
  GSR_CHECK_INPUT(v_compensation);
  torch::Tensor v_cov2d = f32(means3d, {num_points, 3});
  torch::Tensor v_cov3d = f32(means3d, {num_points, 6});
  torch::Tensor v_mean3d = f32(means3d, {num_points, 3});
  torch::Tensor v_scale = f32(means3d, {num_points, 3});
  torch::Tensor v_quat = f32(means3d, {num_points, 4});
  ok(gsr_project_gaussians_backward(num_points, F(means3d), F(scales), glob_scale, F(quats), F(viewmat), F(projmat),
                                    fx, fy, cx, cy, img_height, img_width, F(cov3d), I(radii), F(conics),
                                    F(compensation), F(v_xy), F(v_depth), F(v_conic), F(v_compensation), Fm(v_cov2d),
                                    Fm(v_cov3d), Fm(v_mean3d), Fm(v_scale), Fm(v_quat), cur_stream()),
     "project_gaussians_backward");
  return std::make_tuple(v_cov2d, v_cov3d, v_mean3d, v_scale, v_quat);
}

// bindings.cu:218-251
std::tuple<torch::Tensor, torch::Tensor> map_gaussian_to_intersects_tensor(
    const int num_points, const int num_intersects, const torch::Tensor &xys, const torch::Tensor &depths,
    const torch::Tensor &radii, const torch::Tensor &cum_tiles_hit, const T3 tile_bounds,
    const unsigned block_width) {
  GSR_DEVICE_GUARD(xys);
  GSR_CHECK_INPUT(xys);
  GSR_CHECK_INPUT(depths);
  GSR_CHECK_INPUT(radii);
  GSR_CHECK_INPUT(cum_tiles_hit);
  torch::Tensor gaussian_ids = i32(xys, {num_intersects});
  torch::Tensor isect_ids = torch::empty({num_intersects}, xys.options().dtype(torch::kInt64));
  ok(gsr_map_gaussian_to_intersects(num_points, num_intersects, F(xys), F(depths), I(radii), I(cum_tiles_hit),
                                    (unsigned)std::get<0>(tile_bounds), (unsigned)std::get<1>(tile_bounds),
                                    block_width, isect_ids.data_ptr<int64_t>(), gaussian_ids.data_ptr<int32_t>(),
                                    cur_stream()),
     "map_gaussian_to_intersects");
  return std::make_tuple(isect_ids, gaussian_ids);
}

// bindings.cu:253-267
torch::Tensor get_tile_bin_edges_tensor(int num_intersects, const torch::Tensor &isect_ids_sorted,
                                        const T3 tile_bounds) {
  GSR_DEVICE_GUARD(isect_ids_sorted);
  GSR_CHECK_INPUT(isect_ids_sorted);
  const int num_tiles = std::get<0>(tile_bounds) * std::get<1>(tile_bounds);
  torch::Tensor tile_bins = i32(isect_ids_sorted, {num_tiles, 2});
  ok(gsr_get_tile_bin_edges(num_intersects, isect_ids_sorted.data_ptr<int64_t>(), num_tiles,
                            tile_bins.data_ptr<int32_t>(), cur_stream()),
     "get_tile_bin_edges");
  return tile_bins;
}

template <bool ND>
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> rasterize_forward_impl(
    const T3 tile_bounds, const T3 block, const T3 img_size, const torch::Tensor &gaussian_ids_sorted,
    const torch::Tensor &tile_bins, const torch::Tensor &xys, const torch::Tensor &conics,
    const torch::Tensor &colors, const torch::Tensor &opacities, const torch::Tensor &background) {
  GSR_DEVICE_GUARD(xys);
  GSR_CHECK_INPUT(gaussian_ids_sorted);
  GSR_CHECK_INPUT(tile_bins);
  GSR_CHECK_INPUT(xys);
  GSR_CHECK_INPUT(conics);
  GSR_CHECK_INPUT(colors);
  GSR_CHECK_INPUT(opacities);
  GSR_CHECK_INPUT(background);
  (void)tile_bounds;
  const int channels = colors.size(1);
  const unsigned img_width = std::get<0>(img_size), img_height = std::get<1>(img_size);
  const unsigned block_width = std::get<0>(block);
  torch::Tensor out_img = f32(xys, {(int64_t)img_height, (int64_t)img_width, channels});
  torch::Tensor final_Ts = f32(xys, {(int64_t)img_height, (int64_t)img_width});
  torch::Tensor final_idx = i32(xys, {(int64_t)img_height, (int64_t)img_width});
  if (ND) {
    ok(gsr_nd_rasterize_forward(img_height, img_width, block_width, (unsigned)channels, (int)xys.size(0),
                                I(gaussian_ids_sorted), I(tile_bins), F(xys), F(conics), F(colors), F(opacities),
                                F(background), Fm(out_img), Fm(final_Ts), final_idx.data_ptr<int32_t>(), cur_stream()),
       "nd_rasterize_forward");
  } else {
    TORCH_CHECK(channels == 3, "rasterize_forward: colors must have 3 channels (use nd_rasterize_forward)");
    ok(gsr_rasterize_forward(img_height, img_width, block_width, (int)xys.size(0), I(gaussian_ids_sorted),
                             I(tile_bins), F(xys), F(conics), F(colors), F(opacities), F(background), Fm(out_img),
                             Fm(final_Ts), final_idx.data_ptr<int32_t>(), cur_stream()),
       "rasterize_forward");
  }
  return std::make_tuple(out_img, final_Ts, final_idx);
}

template <bool ND>
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> rasterize_backward_impl(
    const unsigned img_height, const unsigned img_width, const unsigned block_width,
    const torch::Tensor &gaussians_ids_sorted, const torch::Tensor &tile_bins, const torch::Tensor &xys,
    const torch::Tensor &conics, const torch::Tensor &colors, const torch::Tensor &opacities,
    const torch::Tensor &background, const torch::Tensor &final_Ts, const torch::Tensor &final_idx,
    const torch::Tensor &v_output, const torch::Tensor &v_output_alpha) {
  GSR_DEVICE_GUARD(xys);
  GSR_CHECK_INPUT(xys);
  GSR_CHECK_INPUT(colors);
  TORCH_CHECK(xys.ndimension() == 2 && xys.size(1) == 2, "xys must have dimensions (num_points, 2)");
  TORCH_CHECK(colors.ndimension() == 2, "colors must have 2 dimensions");
  const int num_points = xys.size(0);
  const int channels = colors.size(1);
  // the reference reads these through .contiguous() (bindings.cu:430-469, 500-528)
  const torch::Tensor ids = gaussians_ids_sorted.contiguous(), bins = tile_bins.contiguous(), con = conics.contiguous(),
                      op = opacities.contiguous(), bg = background.contiguous(), fT = final_Ts.contiguous(),
                      fi = final_idx.contiguous(), vo = v_output.contiguous(), va = v_output_alpha.contiguous();
  torch::Tensor v_xy = f32(xys, {num_points, 2});
  torch::Tensor v_conic = f32(xys, {num_points, 3});
  torch::Tensor v_colors = f32(xys, {num_points, channels});
  torch::Tensor v_opacity = f32(xys, {num_points, 1});
  if (ND) {
    ok(gsr_nd_rasterize_backward(img_height, img_width, block_width, (unsigned)channels, num_points, I(ids), I(bins),
                                 F(xys), F(con), F(colors), F(op), F(bg), F(fT), I(fi), F(vo), F(va), Fm(v_xy),
                                 Fm(v_conic), Fm(v_colors), Fm(v_opacity), cur_stream()),
       "nd_rasterize_backward");
  } else {
    TORCH_CHECK(channels == 3, "rasterize_backward: colors must have 3 channels (use nd_rasterize_backward)");
    ok(gsr_rasterize_backward(img_height, img_width, block_width, num_points, I(ids), I(bins), F(xys), F(con),
                              F(colors), F(op), F(bg), F(fT), I(fi), F(vo), F(va), Fm(v_xy), Fm(v_conic), Fm(v_colors),
                              Fm(v_opacity), cur_stream()),
       "rasterize_backward");
  }
  return std::make_tuple(v_xy, v_conic, v_colors, v_opacity);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "rasterizer.csrc-compatible bindings over libgsr_b200.so (B200-native kernels)";
  // the same eleven names as the reference's ext.cpp:6-17
  m.def("nd_rasterize_forward", &rasterize_forward_impl<true>);
  m.def("nd_rasterize_backward", &rasterize_backward_impl<true>);
  m.def("rasterize_forward", &rasterize_forward_impl<false>);
  m.def("rasterize_backward", &rasterize_backward_impl<false>);
  m.def("project_gaussians_forward", &project_gaussians_forward_tensor);
  m.def("project_gaussians_backward", &project_gaussians_backward_tensor);
  m.def("compute_sh_forward", &compute_sh_forward_tensor);
  m.def("compute_sh_backward", &compute_sh_backward_tensor);
  m.def("compute_cov2d_bounds", &compute_cov2d_bounds_tensor);
  m.def("map_gaussian_to_intersects", &map_gaussian_to_intersects_tensor);
  m.def("get_tile_bin_edges", &get_tile_bin_edges_tensor);
  m.def("gsr_version", []() { return std::string(gsr_version()); });
}
