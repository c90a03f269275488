// Device helpers shared by the fused fine-stage kernels: ray/point reconstruction and the
// reference's grid-query arithmetic (model/nerf.py:597-672, 1202-1209, model/grid.py:49-59),
// spelled with explicit roundings so discrete decisions downstream stay bit-identical.
#pragma once
#include "../../include/fgs_b200_fine.h"
#include "ray_math.cuh"

namespace fine {

struct Ray {
  float sx, sy, sz;  // start = o + d * t_min      (render_utils_kernel.cu:72-74)
  float ux, uy, uz;  // dir   = d / ||d||          (render_utils_kernel.cu:75-77)
};

__device__ __forceinline__ Ray make_ray(const float* __restrict__ o, const float* __restrict__ d,
                                        float t_min) {
  const float dx = d[0], dy = d[1], dz = d[2];
  const float rn = ray_norm(dx, dy, dz);
  Ray r;
  r.sx = __fmaf_rn(dx, t_min, o[0]);
  r.sy = __fmaf_rn(dy, t_min, o[1]);
  r.sz = __fmaf_rn(dz, t_min, o[2]);
  r.ux = __fdiv_rn(dx, rn);
  r.uy = __fdiv_rn(dy, rn);
  r.uz = __fdiv_rn(dz, rn);
  return r;
}

// render_utils_kernel.cu:184-187
__device__ __forceinline__ void ray_point(const Ray& r, float stepdist, int step, float& px,
                                          float& py, float& pz) {
  const float dist = __fmul_rn(stepdist, (float)step);
  px = __fmaf_rn(dist, r.ux, r.sx);
  py = __fmaf_rn(dist, r.uy, r.sy);
  pz = __fmaf_rn(dist, r.uz, r.sz);
}

// Grid geometry in grid_sample's (W,H,D) = (Z,Y,X) order.
struct GridGeo {
  float mn[3], ext[3];  // xyz_min and (xyz_max - xyz_min), world x,y,z order
  int X, Y, Z;
};

__device__ __forceinline__ GridGeo make_geo(const float* mn, const float* mx, int X, int Y, int Z) {
  GridGeo g;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    g.mn[c] = mn[c];
    g.ext[c] = __fsub_rn(mx[c], mn[c]);
  }
  g.X = X; g.Y = Y; g.Z = Z;
  return g;
}

// continuous index of world point p in (W,H,D) order: ind[0] <- z, ind[1] <- y, ind[2] <- x
// = ((ind_norm + 1) / 2) * (size - 1) with ind_norm = ((p - min)/(max - min)).flip(-1)*2 - 1
// (model/nerf.py:604-607; identical to what F.grid_sample computes internally for the direct
// sample of nerf.py:654-657 / grid.py:55-56 / nerf.py:1206-1207).
__device__ __forceinline__ void cont_index(const GridGeo& g, float px, float py, float pz,
                                           float& iw, float& ih, float& id) {
  iw = aten_unnormalize(ref_normalize(pz, g.mn[2], g.ext[2]), g.Z);
  ih = aten_unnormalize(ref_normalize(py, g.mn[1], g.ext[1]), g.Y);
  id = aten_unnormalize(ref_normalize(px, g.mn[0], g.ext[0]), g.X);
}

// One displaced tap of sample_sdfs (model/nerf.py:611-619): u = clamp(ind + off, 0, S-1) per
// component, renormalised (u/(S-1))*2-1 and un-normalised again inside grid_sample.
__device__ __forceinline__ float tap_clamp(float ind, float off, int S) {
  return fminf(fmaxf(__fadd_rn(ind, off), 0.f), (float)(S - 1));
}
__device__ __forceinline__ float tap_resample_index(float u, int S) {
  const float cn = __fsub_rn(__fmul_rn(__fdiv_rn(u, (float)(S - 1)), 2.f), 1.f);
  return aten_unnormalize(cn, S);
}

// Tap geometry for axis a (0 = W/z, 1 = H/y, 2 = D/x), displacement +-disp.
struct TapPair {
  Tri lo, hi;
  float diff;  // u_hi - u_lo along the axis (model/nerf.py:621-622)
};

__device__ __forceinline__ TapPair make_taps(const GridGeo& g, float iw, float ih, float id, int axis,
                                             float disp) {
  // components that are not displaced still go through ind + 0 -> clamp -> renormalise
  const float uw = tap_clamp(iw, 0.f, g.Z), uh = tap_clamp(ih, 0.f, g.Y), ud = tap_clamp(id, 0.f, g.X);
  const float rw = tap_resample_index(uw, g.Z), rh = tap_resample_index(uh, g.Y),
              rd = tap_resample_index(ud, g.X);
  TapPair t;
  float ulo, uhi;
  if (axis == 0) {
    ulo = tap_clamp(iw, -disp, g.Z); uhi = tap_clamp(iw, disp, g.Z);
    t.lo = tri_setup(tap_resample_index(ulo, g.Z), rh, rd);
    t.hi = tri_setup(tap_resample_index(uhi, g.Z), rh, rd);
  } else if (axis == 1) {
    ulo = tap_clamp(ih, -disp, g.Y); uhi = tap_clamp(ih, disp, g.Y);
    t.lo = tri_setup(rw, tap_resample_index(ulo, g.Y), rd);
    t.hi = tri_setup(rw, tap_resample_index(uhi, g.Y), rd);
  } else {
    ulo = tap_clamp(id, -disp, g.X); uhi = tap_clamp(id, disp, g.X);
    t.lo = tri_setup(rw, rh, tap_resample_index(ulo, g.X));
    t.hi = tri_setup(rw, rh, tap_resample_index(uhi, g.X));
  }
  t.diff = __fsub_rn(uhi, ulo);
  return t;
}

// grad = (f_hi - f_lo) / diff / voxel_size   (model/nerf.py:624-626)
__device__ __forceinline__ float tap_grad(float flo, float fhi, float diff, float voxel_size) {
  return __fdiv_rn(__fdiv_rn(__fsub_rn(fhi, flo), diff), voxel_size);
}

// x = hi + mid + O(2^-17 |x|), hi = bf16(x), mid = bf16(x - hi): the operand split of the 3xBF16 GEMMs (gemm_tcw.cu),
// two values packed per word, element 0 in the low half.
__device__ __forceinline__ void bf16_split2(float x0, float x1, uint32_t& hi, uint32_t& mid) {
  const uint32_t h0 = (__float_as_uint(x0) + 0x8000u) & 0xFFFF0000u, h1 = (__float_as_uint(x1) + 0x8000u) & 0xFFFF0000u;
  const uint32_t m0 = __float_as_uint(x0 - __uint_as_float(h0)) + 0x8000u;
  const uint32_t m1 = __float_as_uint(x1 - __uint_as_float(h1)) + 0x8000u;
  hi = __byte_perm(h0, h1, 0x7632);
  mid = __byte_perm(m0, m1, 0x7632);
}

// nerf.neus_alpha_from_sdf_scatter, model/nerf.py:510-544 (cos_anneal_ratio = 1)
__device__ __forceinline__ float neus_alpha(float sdf, float gx, float gy, float gz, float vx, float vy,
                                            float vz, float dist, float inv_s) {
  const float tc = sum3_inner(__fmul_rn(vx, gx), __fmul_rn(vy, gy), __fmul_rn(vz, gz));
  // -(relu(-tc*0.5+0.5)*0 + relu(-tc)*1)
  const float r1 = __fmul_rn(fmaxf(__fadd_rn(__fmul_rn(-tc, 0.5f), 0.5f), 0.f), 0.f);
  const float r2 = __fmul_rn(fmaxf(-tc, 0.f), 1.f);
  const float ic = -__fadd_rn(r1, r2);
  const float h = __fmul_rn(__fmul_rn(ic, dist), 0.5f);
  const float nx = __fadd_rn(sdf, h), pv = __fsub_rn(sdf, h);
  const float pc = torch_sigmoid(__fmul_rn(pv, inv_s));
  const float nc = torch_sigmoid(__fmul_rn(nx, inv_s));
  const float p = __fsub_rn(pc, nc);
  const float a = __fdiv_rn(__fadd_rn(p, 1e-5f), __fadd_rn(pc, 1e-5f));
  return fminf(fmaxf(a, 0.f), 1.f);
}


// normal = l2_normalize(g / (||g|| + 1e-7))   (model/nerf.py:835, :480-483)
struct NormalCtx {
  float gn, den, ux, uy, uz, ss, L, nx, ny, nz;
};
__device__ __forceinline__ NormalCtx make_normal(float gx, float gy, float gz) {
  NormalCtx c;
  c.gn = __fsqrt_rn(sum3_inner(__fmul_rn(gx, gx), __fmul_rn(gy, gy), __fmul_rn(gz, gz)));
  c.den = __fadd_rn(c.gn, 1e-7f);
  c.ux = __fdiv_rn(gx, c.den); c.uy = __fdiv_rn(gy, c.den); c.uz = __fdiv_rn(gz, c.den);
  c.ss = sum3_inner(__fmul_rn(c.ux, c.ux), __fmul_rn(c.uy, c.uy), __fmul_rn(c.uz, c.uz));
  c.L = __fsqrt_rn(fmaxf(c.ss, 1.1920928955078125e-07f));
  c.nx = __fdiv_rn(c.ux, c.L); c.ny = __fdiv_rn(c.uy, c.L); c.nz = __fdiv_rn(c.uz, c.L);
  return c;
}


__device__ __forceinline__ int lanemask_lt() {
  int m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace fine
