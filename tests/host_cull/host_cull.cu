// host_cull.cu — TEST INFRASTRUCTURE: runs the product's own culling helpers (csrc/tile_cull.cuh, csrc/blend_common.cuh,
// __host__ __device__) on the CPU so that tests/test_cull_host.py can compare them with a brute-force evaluation of the
// reference's per-pixel alpha test (forward.cu:360-363).  Built by the test with nvcc; nothing in the product uses it.
#include "blend_common.cuh"
#include "tile_cull.cuh"

using namespace gsr;

extern "C" {

// the record planes r0 / r1 exactly as gather_record (blend_common.cuh) stages them, then the staging thread's mask
unsigned host_block_mask_16(float x, float y, float a, float b, float c, float opac, float tile_x0, float tile_y0) {
  float ex, ey;
  alpha_extents(a, b, c, opac, ex, ey);
  const float4 r0 = make_float4(x, y, ex, ey);
  const float4 r1 = make_float4(-0.5f * kLog2e * a, -kLog2e * b, -0.5f * kLog2e * c, opac);
  return block_mask_16(r0, r1, tile_x0, tile_y0);
}

// the tiles the binning keeps for one Gaussian — the count pass's own function (cull_tiles, csrc/tile_cull.cuh): returns the
// count, writes at most `cap` tile ids in visiting order and the cached 64-bit mask (0 for boxes of more than 64 tiles)
int host_kept_tiles(float x, float y, int radius, float a, float b, float c, float opac, int tiles_x, int tiles_y,
                    int block_width, int *out, int cap, unsigned long long *mask_out) {
  int x0, y0, x1, y1;
  tile_bbox(x, y, (float)radius, tiles_x, tiles_y, block_width, x0, y0, x1, y1);
  if (mask_out) *mask_out = 0ull;
  if ((x1 - x0) * (y1 - y0) <= 0) return 0;
  int n = 0;
  u64 mask = 0ull;
  const int count = cull_tiles(make_float2(x, y), radius, a, b, c, opac, x0, y0, x1, y1, tiles_x, block_width, mask,
                               [&](int tile) {
                                 if (n < cap) out[n] = tile;
                                 ++n;
                               });
  if (mask_out) *mask_out = mask;
  return count == n ? n : -1;
}

// the fill pass's walk of a cached mask (walk_tile_mask, csrc/tile_cull.cuh)
int host_walk_mask(unsigned long long mask, int x0, int y0, int bw_tiles, int tiles_x, int *out, int cap) {
  int n = 0;
  walk_tile_mask((u64)mask, x0, y0, bw_tiles, tiles_x, [&](int tile) {
    if (n < cap) out[n] = tile;
    ++n;
  });
  return n;
}

// the reference's bounding box of tiles (helpers.cuh:11-34 as restated in common.cuh)
void host_tile_bbox(float x, float y, int radius, int tiles_x, int tiles_y, int block_width, int *box) {
  tile_bbox(x, y, (float)radius, tiles_x, tiles_y, block_width, box[0], box[1], box[2], box[3]);
}
}
