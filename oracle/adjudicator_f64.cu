/*
 * oracle/adjudicator_f64.cu — float64 adjudicator for the gradient parity tests.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/ may build, load or call this file.
 *
 * The reference's backward (RAST/cuda_rasterizer/backward.cu) sums ~10^2..10^4 float terms per Gaussian with
 * global float atomics in a run-dependent order, and recovers the transmittance by repeated float division, so the
 * reference itself is only an approximation of the function it implements.  To decide whether a deviation between
 * this repository's CUDA path and the reference CUDA build is OUR error or THEIRS, this file evaluates the same
 * function in double precision:
 *
 *   adj_render_backward   RAST/cuda_rasterizer/backward.cu:509-754 (renderCUDA), one thread per pixel, the
 *                         reference's per-pixel recursion statement by statement, every float replaced by double,
 *                         double atomics.
 *   adj_gauss_backward    backward.cu:144-310 (computeCov2DCUDA), :433-506 (preprocessCUDA), :20-139 (SH),
 *                         :314-377 (cov3D), :380-428 (normal), one thread per Gaussian, in double.
 *
 * Inputs are the float32 forward intermediates the reference's backward itself reads (means2D, conic_opacity, rgb,
 * depths, normals, cov3D, clamp flags, accum_alpha, n_contrib, ranges, point_list) — the tests take them from the
 * reference CUDA build's own geometry/binning/image buffers.
 *
 * Control flow is the float32 one: whether a (pixel, Gaussian) pair contributes is decided exactly as the reference
 * decides it — `power > 0`, `alpha < 1/255` evaluated in FLOAT with the same expression and the same expf (this
 * file is compiled by the same nvcc without fast-math, so `exp(float)` is the same instruction sequence) — because
 * a pair that flips at the alpha threshold changes a gradient by ~4e-3*|dL/dpix|, which is not rounding error.
 * Only the VALUES are double.  Like the reference, the gradient ignores the min(0.99, .) clamp.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#define TILE 16

namespace {

__device__ __forceinline__ void atomic_add_d(double* p, double v) { atomicAdd(p, v); }

__global__ void adj_render_backward_kernel(
    int W, int H, const uint32_t* __restrict__ ranges, const uint32_t* __restrict__ point_list,
    const float* __restrict__ bg, const float* __restrict__ means2D, const float* __restrict__ conic_opacity,
    const float* __restrict__ colors, const float* __restrict__ depths, const float* __restrict__ norms,
    const float* __restrict__ accum_alphas, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_depth,
    const float* __restrict__ dL_dpix_norm, const float* __restrict__ dL_dpix_alpha,
    double* __restrict__ dL_dmean2D /*[P,3]*/, double* __restrict__ dL_dconic /*[P,4]*/,
    double* __restrict__ dL_dopacity, double* __restrict__ dL_dcolors, double* __restrict__ dL_ddepths,
    double* __restrict__ dL_dnorm3D) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= W * H) return;
  const int gx = (W + TILE - 1) / TILE;
  const size_t HW = (size_t)H * W;
  const int px = pix % W, py = pix / W;
  const int tile = (py / TILE) * gx + (px / TILE);
  const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
  const float pfx = (float)px, pfy = (float)py;
  const double ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H;
  const double T_final = 1.0 - (double)accum_alphas[pix];
  double T = T_final;
  const uint32_t last_contributor = n_contrib[pix];
  double accum_rec[3] = {0, 0, 0}, accum_red = 0, accum_ren[3] = {0, 0, 0}, accum_rea = 0;
  double last_color[3] = {0, 0, 0}, last_depth = 0, last_norm[3] = {0, 0, 0}, last_alpha = 0;
  double dLp[3], dLn[3];
  for (int i = 0; i < 3; i++) { dLp[i] = dL_dpix[i * HW + pix]; dLn[i] = dL_dpix_norm[i * HW + pix]; }
  const double dLd = dL_dpix_depth[pix], dLa = dL_dpix_alpha[pix];
  double bg_dot = 0;
  for (int i = 0; i < 3; i++) bg_dot += (double)bg[i] * dLp[i];
  uint32_t contributor = r1 - r0;
  for (uint32_t k = r1; k-- > r0;) {
    contributor--;
    if (contributor >= last_contributor) continue;
    const uint32_t id = point_list[k];
    const float mxf = means2D[2 * id], myf = means2D[2 * id + 1];
    const float* co = conic_opacity + 4 * (size_t)id;
    // ---- float32 control flow, the reference's expressions (backward.cu:617-628) ----
    {
      const float dxf = mxf - pfx, dyf = myf - pfy;
      const float power_f = -0.5f * (co[0] * dxf * dxf + co[2] * dyf * dyf) - co[1] * dxf * dyf;
      if (power_f > 0.0f) continue;
      const float G_f = exp(power_f);
      const float alpha_f = min(0.99f, co[3] * G_f);
      if (alpha_f < 1.0f / 255.0f) continue;
    }
    // ---- float64 values ----
    const double dx = (double)mxf - (double)pfx, dy = (double)myf - (double)pfy;
    const double c0 = co[0], c1 = co[1], c2 = co[2], op = co[3];
    const double power = -0.5 * (c0 * dx * dx + c2 * dy * dy) - c1 * dx * dy;
    const double G = exp(power);
    const double alpha = fmin((double)0.99f, op * G);
    T = T / (1.0 - alpha);
    const double weight = alpha * T;
    double dL_dalpha = 0.0;
    for (int ch = 0; ch < 3; ch++) {
      const double c = colors[3 * id + ch];
      accum_rec[ch] = last_alpha * last_color[ch] + (1.0 - last_alpha) * accum_rec[ch];
      last_color[ch] = c;
      dL_dalpha += (c - accum_rec[ch]) * dLp[ch];
      atomic_add_d(&dL_dcolors[3 * id + ch], weight * dLp[ch]);
    }
    const double dep = depths[id];
    accum_red = last_alpha * last_depth + (1.0 - last_alpha) * accum_red;
    last_depth = dep;
    dL_dalpha += (dep - accum_red) * dLd;
    atomic_add_d(&dL_ddepths[id], weight * dLd);
    for (int ch = 0; ch < 3; ch++) {
      const double n = norms[3 * id + ch];
      accum_ren[ch] = last_alpha * last_norm[ch] + (1.0 - last_alpha) * accum_ren[ch];
      last_norm[ch] = n;
      dL_dalpha += (n - accum_ren[ch]) * dLn[ch];
      atomic_add_d(&dL_dnorm3D[3 * id + ch], weight * dLn[ch]);
    }
    accum_rea = last_alpha + (1.0 - last_alpha) * accum_rea;
    dL_dalpha += (1.0 - accum_rea) * dLa;
    dL_dalpha *= T;
    last_alpha = alpha;
    dL_dalpha += (-T_final / (1.0 - alpha)) * bg_dot;
    const double dL_dG = op * dL_dalpha;
    const double gdx = G * dx, gdy = G * dy;
    const double dG_ddelx = -gdx * c0 - gdy * c1;
    const double dG_ddely = -gdy * c2 - gdx * c1;
    atomic_add_d(&dL_dmean2D[3 * id], dL_dG * dG_ddelx * ddelx_dx);
    atomic_add_d(&dL_dmean2D[3 * id + 1], dL_dG * dG_ddely * ddely_dy);
    atomic_add_d(&dL_dmean2D[3 * id + 2], fabs(dL_dG * dG_ddelx * ddelx_dx) + fabs(dL_dG * dG_ddely * ddely_dy));
    atomic_add_d(&dL_dconic[4 * id], -0.5 * gdx * dx * dL_dG);
    atomic_add_d(&dL_dconic[4 * id + 1], -0.5 * gdx * dy * dL_dG);
    atomic_add_d(&dL_dconic[4 * id + 3], -0.5 * gdy * dy * dL_dG);
    atomic_add_d(&dL_dopacity[id], G * dL_dalpha);
  }
}

__device__ void dnormvdv3(const double v[3], const double dv[3], double o[3]) {
  /* auxiliary.h:108-119 */
  const double sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double invsum32 = 1.0 / sqrt(sum2 * sum2 * sum2);
  o[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * invsum32;
  o[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * invsum32;
  o[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * invsum32;
}

__constant__ double D_SH_C1 = 0.4886025119029199;
__constant__ double D_SH_C2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
                                  0.5462742152960396};
__constant__ double D_SH_C3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                                  -0.4570457994644658, 1.445305721320277, -0.5900435899266435};

// All dL_* arrays hold the blend-adjoint sums on entry and the final gradients on return (zeros for culled ones).
__global__ void adj_gauss_backward_kernel(
    int P, int D, int M, const float* __restrict__ means3D, const int* __restrict__ radii,
    const float* __restrict__ shs, const unsigned char* __restrict__ clamped /*[P,3]*/,
    const float* __restrict__ scales, const float* __restrict__ rotations, float scale_modifier_f,
    const float* __restrict__ cov3Ds, const float* __restrict__ norm3Ds, const float* __restrict__ viewf,
    const float* __restrict__ projf, int W, int H, float tan_fovx_f, float tan_fovy_f, float kernel_size_f,
    const float* __restrict__ camposf, const float* __restrict__ conic_opacity, const double* __restrict__ dL_dmean2D,
    const double* __restrict__ dL_dconic, double* __restrict__ dL_dopacity, const double* __restrict__ dL_dcolor,
    const double* __restrict__ dL_ddepth, double* __restrict__ dL_dmean3D, double* __restrict__ dL_dcov3D,
    const double* __restrict__ dL_dnorm3D, double* __restrict__ dL_dsh, double* __restrict__ dL_dscale,
    double* __restrict__ dL_drot) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P || !(radii[idx] > 0)) return;
  // the float quantities the reference forms on the host / as kernel parameters stay float (they ARE the inputs)
  const float h_y_f = H / (2.0f * tan_fovy_f), h_x_f = W / (2.0f * tan_fovx_f);
  const double h_x = h_x_f, h_y = h_y_f, tan_fovx = tan_fovx_f, tan_fovy = tan_fovy_f, kernel_size = kernel_size_f;
  const double scale_modifier = scale_modifier_f;
  double view[16], proj[16];
  for (int i = 0; i < 16; i++) { view[i] = viewf[i]; proj[i] = projf[i]; }
  const double mx = means3D[3 * idx], my = means3D[3 * idx + 1], mz = means3D[3 * idx + 2];
  const float* cov3D = cov3Ds + 6 * (size_t)idx;
  double o_mean[3];
  /* ---- computeCov2DCUDA ---- */
  {
    const double dLcx = dL_dconic[4 * idx], dLcy = dL_dconic[4 * idx + 1], dLcz = dL_dconic[4 * idx + 3];
    const double combined_opacity = conic_opacity[4 * idx + 3];
    double tx = view[0] * mx + view[4] * my + view[8] * mz + view[12];
    double ty = view[1] * mx + view[5] * my + view[9] * mz + view[13];
    const double tz = view[2] * mx + view[6] * my + view[10] * mz + view[14];
    const double limx = (double)(1.3f * tan_fovx_f), limy = (double)(1.3f * tan_fovy_f);
    const double txtz = tx / tz, tytz = ty / tz;
    tx = fmin(limx, fmax(-limx, txtz)) * tz;
    ty = fmin(limy, fmax(-limy, tytz)) * tz;
    const double x_grad_mul = (txtz < -limx || txtz > limx) ? 0 : 1;
    const double y_grad_mul = (tytz < -limy || tytz > limy) ? 0 : 1;
    const double a0 = h_x / tz, a2 = -(h_x * tx) / (tz * tz), b1 = h_y / tz, b2 = -(h_y * ty) / (tz * tz);
    const double Wm[3][3] = {{view[0], view[4], view[8]}, {view[1], view[5], view[9]}, {view[2], view[6], view[10]}};
    double T[3][3];
    for (int r = 0; r < 3; r++) { T[0][r] = Wm[0][r] * a0 + Wm[2][r] * a2; T[1][r] = Wm[1][r] * b1 + Wm[2][r] * b2; T[2][r] = 0; }
    const double Vrk[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
    double X[3][2];
    for (int c = 0; c < 3; c++) {
      X[c][0] = T[0][0] * Vrk[0][c] + T[0][1] * Vrk[1][c] + T[0][2] * Vrk[2][c];
      X[c][1] = T[1][0] * Vrk[0][c] + T[1][1] * Vrk[1][c] + T[1][2] * Vrk[2][c];
    }
    const double c00 = X[0][0] * T[0][0] + X[1][0] * T[0][1] + X[2][0] * T[0][2];
    const double c01 = X[0][1] * T[0][0] + X[1][1] * T[0][1] + X[2][1] * T[0][2];
    const double c11 = X[0][1] * T[1][0] + X[1][1] * T[1][1] + X[2][1] * T[1][2];
    const double det_0 = fmax(1e-6, c00 * c11 - c01 * c01);
    const double det_1 = fmax(1e-6, (c00 + kernel_size) * (c11 + kernel_size) - c01 * c01);
    const double coef = sqrt(det_0 / (det_1 + 1e-6) + 1e-6);
    const double opacity = combined_opacity / (coef + 1e-6);
    const double dL_dcoef = dL_dopacity[idx] * opacity;
    const double dL_dsqrtcoef = dL_dcoef * 0.5 * 1. / (coef + 1e-6);
    const double dL_ddet0 = dL_dsqrtcoef / (det_1 + 1e-6);
    const double dL_ddet1 = dL_dsqrtcoef * det_0 * (-1.0 / (det_1 * det_1 + 1e-6));
    const double dcoef_da = dL_ddet0 * c11 + dL_ddet1 * (c11 + kernel_size);
    const double dcoef_db = dL_ddet0 * (-2. * c01) + dL_ddet1 * (-2. * c01);
    const double dcoef_dc = dL_ddet0 * c00 + dL_ddet1 * (c00 + kernel_size);
    const double a = c00 + kernel_size, b = c01, c = c11 + kernel_size;
    const double denom = a * c - b * b;
    double dL_da = 0, dL_db = 0, dL_dc = 0;
    const double denom2inv = 1.0 / ((denom * denom) + (double)0.0000001f);
    double* oc = dL_dcov3D + 6 * (size_t)idx;
    if (denom2inv != 0) {
      dL_da = denom2inv * (-c * c * dLcx + 2 * b * c * dLcy + (denom - a * c) * dLcz);
      dL_dc = denom2inv * (-a * a * dLcz + 2 * a * b * dLcy + (denom - a * c) * dLcx);
      dL_db = denom2inv * 2 * (b * c * dLcx - (denom + 2 * b * b) * dLcy + a * b * dLcz);
      if (det_0 <= 1e-6 || det_1 <= 1e-6) dL_dopacity[idx] = 0;
      else { dL_da += dcoef_da; dL_dc += dcoef_dc; dL_db += dcoef_db; dL_dopacity[idx] = dL_dopacity[idx] * coef; }
      oc[0] = (T[0][0] * T[0][0] * dL_da + T[0][0] * T[1][0] * dL_db + T[1][0] * T[1][0] * dL_dc);
      oc[3] = (T[0][1] * T[0][1] * dL_da + T[0][1] * T[1][1] * dL_db + T[1][1] * T[1][1] * dL_dc);
      oc[5] = (T[0][2] * T[0][2] * dL_da + T[0][2] * T[1][2] * dL_db + T[1][2] * T[1][2] * dL_dc);
      oc[1] = 2 * T[0][0] * T[0][1] * dL_da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][1] * dL_dc;
      oc[2] = 2 * T[0][0] * T[0][2] * dL_da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][2] * dL_dc;
      oc[4] = 2 * T[0][2] * T[0][1] * dL_da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_db + 2 * T[1][1] * T[1][2] * dL_dc;
    } else {
      for (int i = 0; i < 6; i++) oc[i] = 0;
    }
    double dT[2][3];
    for (int k = 0; k < 3; k++) {
      dT[0][k] = 2 * (T[0][0] * Vrk[k][0] + T[0][1] * Vrk[k][1] + T[0][2] * Vrk[k][2]) * dL_da +
                 (T[1][0] * Vrk[k][0] + T[1][1] * Vrk[k][1] + T[1][2] * Vrk[k][2]) * dL_db;
      dT[1][k] = 2 * (T[1][0] * Vrk[k][0] + T[1][1] * Vrk[k][1] + T[1][2] * Vrk[k][2]) * dL_dc +
                 (T[0][0] * Vrk[k][0] + T[0][1] * Vrk[k][1] + T[0][2] * Vrk[k][2]) * dL_db;
    }
    const double dL_dJ00 = Wm[0][0] * dT[0][0] + Wm[0][1] * dT[0][1] + Wm[0][2] * dT[0][2];
    const double dL_dJ02 = Wm[2][0] * dT[0][0] + Wm[2][1] * dT[0][1] + Wm[2][2] * dT[0][2];
    const double dL_dJ11 = Wm[1][0] * dT[1][0] + Wm[1][1] * dT[1][1] + Wm[1][2] * dT[1][2];
    const double dL_dJ12 = Wm[2][0] * dT[1][0] + Wm[2][1] * dT[1][1] + Wm[2][2] * dT[1][2];
    const double tzi = 1.0 / tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
    const double dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
    const double dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
    const double dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * tx) * tz3 * dL_dJ02 + (2 * h_y * ty) * tz3 * dL_dJ12;
    o_mean[0] = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
    o_mean[1] = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
    o_mean[2] = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;
  }
  /* ---- backward preprocessCUDA ---- */
  {
    const double hw = proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15];
    const double m_w = 1.0 / (hw + (double)0.0000001f);
    const double mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
    const double mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
    const double gx_ = dL_dmean2D[3 * idx], gy_ = dL_dmean2D[3 * idx + 1];
    o_mean[0] += (proj[0] * m_w - proj[3] * mul1) * gx_ + (proj[1] * m_w - proj[3] * mul2) * gy_;
    o_mean[1] += (proj[4] * m_w - proj[7] * mul1) * gx_ + (proj[5] * m_w - proj[7] * mul2) * gy_;
    o_mean[2] += (proj[8] * m_w - proj[11] * mul1) * gx_ + (proj[9] * m_w - proj[11] * mul2) * gy_;
    const double mul3 = view[2] * mx + view[6] * my + view[10] * mz + view[14];
    o_mean[0] += (view[2] - view[3] * mul3) * dL_ddepth[idx];
    o_mean[1] += (view[6] - view[7] * mul3) * dL_ddepth[idx];
    o_mean[2] += (view[10] - view[11] * mul3) * dL_ddepth[idx];
  }
  if (shs) {
    const double o[3] = {mx - (double)camposf[0], my - (double)camposf[1], mz - (double)camposf[2]};
    const double len = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    const double x = o[0] / len, y = o[1] / len, z = o[2] / len;
    const float* sh = shs + (size_t)idx * M * 3;
    double* dsh = dL_dsh + (size_t)idx * M * 3;
    double dRGB[3];
    for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[3 * idx + ch] * (clamped[3 * idx + ch] ? 0 : 1);
    double dx_[3] = {0, 0, 0}, dy_[3] = {0, 0, 0}, dz_[3] = {0, 0, 0};
    double w[16];
    for (int k = 0; k < 16; k++) w[k] = 0;
    w[0] = 0.28209479177387814;
    if (D > 0) {
      w[1] = -D_SH_C1 * y; w[2] = D_SH_C1 * z; w[3] = -D_SH_C1 * x;
      for (int ch = 0; ch < 3; ch++) { dx_[ch] = -D_SH_C1 * sh[9 + ch]; dy_[ch] = -D_SH_C1 * sh[3 + ch]; dz_[ch] = D_SH_C1 * sh[6 + ch]; }
      if (D > 1) {
        const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        w[4] = D_SH_C2[0] * xy; w[5] = D_SH_C2[1] * yz; w[6] = D_SH_C2[2] * (2. * zz - xx - yy); w[7] = D_SH_C2[3] * xz; w[8] = D_SH_C2[4] * (xx - yy);
        for (int ch = 0; ch < 3; ch++) {
          const double s4 = sh[12 + ch], s5 = sh[15 + ch], s6 = sh[18 + ch], s7 = sh[21 + ch], s8 = sh[24 + ch];
          dx_[ch] += D_SH_C2[0] * y * s4 + D_SH_C2[2] * 2. * -x * s6 + D_SH_C2[3] * z * s7 + D_SH_C2[4] * 2. * x * s8;
          dy_[ch] += D_SH_C2[0] * x * s4 + D_SH_C2[1] * z * s5 + D_SH_C2[2] * 2. * -y * s6 + D_SH_C2[4] * 2. * -y * s8;
          dz_[ch] += D_SH_C2[1] * y * s5 + D_SH_C2[2] * 2. * 2. * z * s6 + D_SH_C2[3] * x * s7;
        }
        if (D > 2) {
          w[9] = D_SH_C3[0] * y * (3. * xx - yy); w[10] = D_SH_C3[1] * xy * z; w[11] = D_SH_C3[2] * y * (4. * zz - xx - yy);
          w[12] = D_SH_C3[3] * z * (2. * zz - 3. * xx - 3. * yy); w[13] = D_SH_C3[4] * x * (4. * zz - xx - yy);
          w[14] = D_SH_C3[5] * z * (xx - yy); w[15] = D_SH_C3[6] * x * (xx - 3. * yy);
          for (int ch = 0; ch < 3; ch++) {
            const double s9 = sh[27 + ch], s10 = sh[30 + ch], s11 = sh[33 + ch], s12 = sh[36 + ch], s13 = sh[39 + ch], s14 = sh[42 + ch], s15 = sh[45 + ch];
            dx_[ch] += (D_SH_C3[0] * s9 * 3. * 2. * xy + D_SH_C3[1] * s10 * yz + D_SH_C3[2] * s11 * -2. * xy + D_SH_C3[3] * s12 * -3. * 2. * xz +
                        D_SH_C3[4] * s13 * (-3. * xx + 4. * zz - yy) + D_SH_C3[5] * s14 * 2. * xz + D_SH_C3[6] * s15 * 3. * (xx - yy));
            dy_[ch] += (D_SH_C3[0] * s9 * 3. * (xx - yy) + D_SH_C3[1] * s10 * xz + D_SH_C3[2] * s11 * (-3. * yy + 4. * zz - xx) +
                        D_SH_C3[3] * s12 * -3. * 2. * yz + D_SH_C3[4] * s13 * -2. * xy + D_SH_C3[5] * s14 * -2. * yz + D_SH_C3[6] * s15 * -3. * 2. * xy);
            dz_[ch] += (D_SH_C3[1] * s10 * xy + D_SH_C3[2] * s11 * 4. * 2. * yz + D_SH_C3[3] * s12 * 3. * (2. * zz - xx - yy) +
                        D_SH_C3[4] * s13 * 4. * 2. * xz + D_SH_C3[5] * s14 * (xx - yy));
          }
        }
      }
    }
    const int ncoef = (D + 1) * (D + 1);
    for (int k = 0; k < ncoef && k < M; k++) for (int ch = 0; ch < 3; ch++) dsh[3 * k + ch] = w[k] * dRGB[ch];
    const double ddir[3] = {dx_[0] * dRGB[0] + dx_[1] * dRGB[1] + dx_[2] * dRGB[2],
                            dy_[0] * dRGB[0] + dy_[1] * dRGB[1] + dy_[2] * dRGB[2],
                            dz_[0] * dRGB[0] + dz_[1] * dRGB[1] + dz_[2] * dRGB[2]};
    double dm[3];
    dnormvdv3(o, ddir, dm);
    o_mean[0] += dm[0]; o_mean[1] += dm[1]; o_mean[2] += dm[2];
  }
  dL_dmean3D[3 * idx] = o_mean[0]; dL_dmean3D[3 * idx + 1] = o_mean[1]; dL_dmean3D[3 * idx + 2] = o_mean[2];
  if (scales) {
    const float* q = rotations + 4 * (size_t)idx;
    const double r = q[0], x = q[1], y = q[2], z = q[3];
    double R[3][3];
    R[0][0] = 1. - 2. * (y * y + z * z); R[0][1] = 2. * (x * y - r * z); R[0][2] = 2. * (x * z + r * y);
    R[1][0] = 2. * (x * y + r * z); R[1][1] = 1. - 2. * (x * x + z * z); R[1][2] = 2. * (y * z - r * x);
    R[2][0] = 2. * (x * z - r * y); R[2][1] = 2. * (y * z + r * x); R[2][2] = 1. - 2. * (x * x + y * y);
    const float scf[3] = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
    const double s[3] = {scale_modifier * scf[0], scale_modifier * scf[1], scale_modifier * scf[2]};
    double Mm[3][3];
    for (int c = 0; c < 3; c++) { Mm[c][0] = s[0] * R[c][0]; Mm[c][1] = s[1] * R[c][1]; Mm[c][2] = s[2] * R[c][2]; }
    const double* g = dL_dcov3D + 6 * (size_t)idx;
    const double Sg[3][3] = {{g[0], 0.5 * g[1], 0.5 * g[2]}, {0.5 * g[1], g[3], 0.5 * g[4]}, {0.5 * g[2], 0.5 * g[4], g[5]}};
    double dM[3][3], dMt[3][3];
    for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++)
      dM[c][rr] = (2.0 * Mm[0][rr]) * Sg[c][0] + (2.0 * Mm[1][rr]) * Sg[c][1] + (2.0 * Mm[2][rr]) * Sg[c][2];
    for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) dMt[c][rr] = dM[rr][c];
    dL_dscale[3 * idx] = R[0][0] * dMt[0][0] + R[1][0] * dMt[0][1] + R[2][0] * dMt[0][2];
    dL_dscale[3 * idx + 1] = R[0][1] * dMt[1][0] + R[1][1] * dMt[1][1] + R[2][1] * dMt[1][2];
    dL_dscale[3 * idx + 2] = R[0][2] * dMt[2][0] + R[1][2] * dMt[2][1] + R[2][2] * dMt[2][2];
    for (int k = 0; k < 3; k++) { dMt[0][k] *= s[0]; dMt[1][k] *= s[1]; dMt[2][k] *= s[2]; }
    double dq[4];
    dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
    dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
    dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
    dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
    {
      /* normal -> rotation (backward.cu:380-428); the axis and its sign are float32 control flow */
      double ax[3];
      if (scf[0] > scf[2] && scf[1] > scf[2]) { ax[0] = 0; ax[1] = 0; ax[2] = 1; }
      else if (scf[0] > scf[1] && scf[2] > scf[1]) { ax[0] = 0; ax[1] = 1; ax[2] = 0; }
      else { ax[0] = 1; ax[1] = 0; ax[2] = 0; }
      const float* n3 = norm3Ds + 3 * (size_t)idx;
      const double rn0 = R[0][0] * ax[0] + R[1][0] * ax[1] + R[2][0] * ax[2];
      const double rn1 = R[0][1] * ax[0] + R[1][1] * ax[1] + R[2][1] * ax[2];
      const double rn2 = R[0][2] * ax[0] + R[1][2] * ax[1] + R[2][2] * ax[2];
      if (rn0 * n3[0] + rn1 * n3[1] + rn2 * n3[2] < 0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; }
      const double* dn = dL_dnorm3D + 3 * (size_t)idx;
      double dRt[3][3];
      for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) dRt[c][rr] = dn[rr] * ax[c];
      dq[0] += 2 * z * (dRt[0][1] - dRt[1][0]) + 2 * y * (dRt[2][0] - dRt[0][2]) + 2 * x * (dRt[1][2] - dRt[2][1]);
      dq[1] += 2 * y * (dRt[1][0] + dRt[0][1]) + 2 * z * (dRt[2][0] + dRt[0][2]) + 2 * r * (dRt[1][2] - dRt[2][1]) - 4 * x * (dRt[2][2] + dRt[1][1]);
      dq[2] += 2 * x * (dRt[1][0] + dRt[0][1]) + 2 * r * (dRt[2][0] - dRt[0][2]) + 2 * z * (dRt[1][2] + dRt[2][1]) - 4 * y * (dRt[2][2] + dRt[0][0]);
      dq[3] += 2 * r * (dRt[0][1] - dRt[1][0]) + 2 * x * (dRt[2][0] + dRt[0][2]) + 2 * y * (dRt[1][2] + dRt[2][1]) - 4 * z * (dRt[1][1] + dRt[0][0]);
    }
    for (int k = 0; k < 4; k++) dL_drot[4 * idx + k] = dq[k];
  }
}

}  // namespace

extern "C" {

// Returns 0 on success, otherwise the cudaError_t of the failing call.  All pointers are device pointers; the
// double outputs must be zero-filled by the caller.  Runs on the given stream and synchronises it before returning.
int adj_backward_f64(int P, int D, int M, int W, int H,
                     const uint32_t* ranges, const uint32_t* point_list, const float* bg, const float* means2D,
                     const float* conic_opacity, const float* colors, const float* depths, const float* norms,
                     const float* accum_alphas, const uint32_t* n_contrib, const float* dL_dpix,
                     const float* dL_dpix_depth, const float* dL_dpix_norm, const float* dL_dpix_alpha,
                     const float* means3D, const int* radii, const float* shs, const unsigned char* clamped,
                     const float* scales, const float* rotations, float scale_modifier, const float* cov3Ds,
                     const float* view, const float* proj, float tan_fovx, float tan_fovy, float kernel_size,
                     const float* campos,
                     double* dL_dmean2D, double* dL_dconic, double* dL_dopacity, double* dL_dcolors,
                     double* dL_ddepths, double* dL_dnorm3D, double* dL_dmean3D, double* dL_dcov3D, double* dL_dsh,
                     double* dL_dscale, double* dL_drot, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int N = W * H;
  if (N > 0 && P > 0) {
    adj_render_backward_kernel<<<(N + 127) / 128, 128, 0, st>>>(
        W, H, ranges, point_list, bg, means2D, conic_opacity, colors, depths, norms, accum_alphas, n_contrib, dL_dpix,
        dL_dpix_depth, dL_dpix_norm, dL_dpix_alpha, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_ddepths,
        dL_dnorm3D);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    adj_gauss_backward_kernel<<<(P + 127) / 128, 128, 0, st>>>(
        P, D, M, means3D, radii, shs, clamped, scales, rotations, scale_modifier, cov3Ds, norms, view, proj, W, H,
        tan_fovx, tan_fovy, kernel_size, campos, conic_opacity, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors,
        dL_ddepths, dL_dmean3D, dL_dcov3D, dL_dnorm3D, dL_dsh, dL_dscale, dL_drot);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return (int)cudaStreamSynchronize(st);
}

}  // extern "C"
