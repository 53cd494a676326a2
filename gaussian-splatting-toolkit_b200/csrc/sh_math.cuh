// sh_math.cuh — real spherical-harmonics basis (degree <= 4) as a device function, for kernels that fuse the SH
// evaluation with other per-Gaussian work (fused.cu).  Same constants and polynomial forms as sh.cu / the reference
// (csrc/sh.cuh:6-98).
#pragma once
#include "common.cuh"

namespace gsr {

// Y[0..(deg+1)^2) for direction (dx,dy,dz) (normalised here, as sh.cuh:44-48 does)
GSR_HD void sh_basis_all(int deg, float dx, float dy, float dz, float *Y) {
  Y[0] = 0.28209479177387814f;
  if (deg < 1) return;
  const float C1 = 0.4886025119029199f;
  float norm = sqrtf(dx * dx + dy * dy + dz * dz);
  float x = dx / norm, y = dy / norm, z = dz / norm;
  Y[1] = -C1 * y;
  Y[2] = C1 * z;
  Y[3] = -C1 * x;
  if (deg < 2) return;
  float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
  Y[4] = 1.0925484305920792f * xy;
  Y[5] = -1.0925484305920792f * yz;
  Y[6] = 0.31539156525252005f * (2.f * zz - xx - yy);
  Y[7] = -1.0925484305920792f * xz;
  Y[8] = 0.5462742152960396f * (xx - yy);
  if (deg < 3) return;
  Y[9] = -0.5900435899266435f * y * (3.f * xx - yy);
  Y[10] = 2.890611442640554f * xy * z;
  Y[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
  Y[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
  Y[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy);
  Y[14] = 1.445305721320277f * z * (xx - yy);
  Y[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
  if (deg < 4) return;
  Y[16] = 2.5033429417967046f * xy * (xx - yy);
  Y[17] = -1.7701307697799304f * yz * (3.f * xx - yy);
  Y[18] = 0.9461746957575601f * xy * (7.f * zz - 1.f);
  Y[19] = -0.6690465435572892f * yz * (7.f * zz - 3.f);
  Y[20] = 0.10578554691520431f * (zz * (35.f * zz - 30.f) + 3.f);
  Y[21] = -0.6690465435572892f * xz * (7.f * zz - 3.f);
  Y[22] = 0.47308734787878004f * (xx - yy) * (7.f * zz - 1.f);
  Y[23] = -1.7701307697799304f * xz * (xx - 3.f * yy);
  Y[24] = 0.6258357354491761f * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
}

}  // namespace gsr
