// Bit-exact scalar building blocks shared by the sampling kernels and the fused fine-stage
// kernels.  Every function spells out the reference's operation/rounding order with __f*_rn
// intrinsics so the compiler cannot re-contract or re-associate (ids and masks downstream
// must be bit-identical to the reference).
#pragma once
#include "common.cuh"

struct RayIsect {
  float t_min, t_max;
};

// model/cuda/render_utils_kernel.cu:23-33
__device__ __forceinline__ RayIsect ray_aabb(float ox, float oy, float oz, float dx, float dy,
                                             float dz, float mn0, float mn1, float mn2, float mx0,
                                             float mx1, float mx2, float near_, float far_) {
  const float vx = (dx == 0.f) ? (float)1e-6 : dx;
  const float vy = (dy == 0.f) ? (float)1e-6 : dy;
  const float vz = (dz == 0.f) ? (float)1e-6 : dz;
  const float ax = __fdiv_rn(__fsub_rn(mx0, ox), vx);
  const float ay = __fdiv_rn(__fsub_rn(mx1, oy), vy);
  const float az = __fdiv_rn(__fsub_rn(mx2, oz), vz);
  const float bx = __fdiv_rn(__fsub_rn(mn0, ox), vx);
  const float by = __fdiv_rn(__fsub_rn(mn1, oy), vy);
  const float bz = __fdiv_rn(__fsub_rn(mn2, oz), vz);
  RayIsect r;
  r.t_min = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far_), near_);
  r.t_max = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far_), near_);
  return r;
}

// kernel.cu:48-51 / 68-71: SASS evaluates dx*dx+dy*dy+dz*dz as fma(dz,dz, fma(dx,dx, dy*dy)).
__device__ __forceinline__ float ray_norm(float dx, float dy, float dz) {
  return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
}

// kernel.cu:53
__device__ __forceinline__ long long ray_nsamples(float t_min, float t_max, float rnorm,
                                                  float stepdist) {
  const float c = ceilf(__fdiv_rn(__fmul_rn(__fsub_rn(t_max, t_min), rnorm), stepdist));
  return (long long)fmax((double)c, 1.);
}

// ---- ATen grid_sampler_3d (bilinear, zeros padding, align_corners=True) -------------------
// torch/include/ATen/native/cuda/GridSampler.cuh:23-27 (unnormalize) and the forward kernel of
// ATen/native/cuda/GridSampler.cu: floor, eight 3-factor weights in tnw,tne,tsw,tse,bnw,bne,
// bsw,bse order, then out += value * weight sequentially (FMA) over the in-bounds corners.

// normalised [-1,1] coordinate -> continuous index:  ((c + 1) / 2) * (size - 1)
__device__ __forceinline__ float aten_unnormalize(float c, int size) {
  return __fmul_rn(__fmul_rn(__fadd_rn(c, 1.f), 0.5f), (float)(size - 1));
}

// world coordinate -> normalised coordinate as the reference builds it
// (model/nerf.py:604,654,1206, model/grid.py:55):  ((p - min) / (max - min)) * 2 - 1
__device__ __forceinline__ float ref_normalize(float p, float mn, float extent) {
  return __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p, mn), extent), 2.f), 1.f);
}

struct Tri {
  int x0, y0, z0;  // floor corner in (W,H,D) = (grid Z, grid Y, grid X) index space
  float w[8];      // tnw,tne,tsw,tse,bnw,bne,bsw,bse
};

__device__ __forceinline__ Tri tri_setup(float ix, float iy, float iz) {
  Tri t;
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  t.x0 = (int)fx; t.y0 = (int)fy; t.z0 = (int)fz;
  const float x1 = (float)(t.x0 + 1), y1 = (float)(t.y0 + 1), z1 = (float)(t.z0 + 1);
  const float x0 = (float)t.x0, y0 = (float)t.y0, z0 = (float)t.z0;
  const float ax = __fsub_rn(x1, ix), bx = __fsub_rn(ix, x0);
  const float ay = __fsub_rn(y1, iy), by = __fsub_rn(iy, y0);
  const float az = __fsub_rn(z1, iz), bz = __fsub_rn(iz, z0);
  t.w[0] = __fmul_rn(__fmul_rn(ax, ay), az);
  t.w[1] = __fmul_rn(__fmul_rn(bx, ay), az);
  t.w[2] = __fmul_rn(__fmul_rn(ax, by), az);
  t.w[3] = __fmul_rn(__fmul_rn(bx, by), az);
  t.w[4] = __fmul_rn(__fmul_rn(ax, ay), bz);
  t.w[5] = __fmul_rn(__fmul_rn(bx, ay), bz);
  t.w[6] = __fmul_rn(__fmul_rn(ax, by), bz);
  t.w[7] = __fmul_rn(__fmul_rn(bx, by), bz);
  return t;
}

// corner c -> offsets (dx,dy,dz) in ATen order
#define TRI_DX(c) ((c) & 1)
#define TRI_DY(c) (((c) >> 1) & 1)
#define TRI_DZ(c) (((c) >> 2) & 1)

// single-channel sample of a [D][H][W] grid
__device__ __forceinline__ float tri_sample1(const float* __restrict__ g, const Tri& t, int D, int H,
                                             int W) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int x = t.x0 + TRI_DX(c), y = t.y0 + TRI_DY(c), z = t.z0 + TRI_DZ(c);
    if (x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D)
      acc = __fmaf_rn(__ldg(g + ((size_t)z * H + y) * W + x), t.w[c], acc);
  }
  return acc;
}

// scatter  grad[corner] += w[corner] * go   (ATen backward: safe_add_3d(..., tnw * gOut))
__device__ __forceinline__ void tri_scatter1(float* __restrict__ g, const Tri& t, int D, int H, int W,
                                             float go) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int x = t.x0 + TRI_DX(c), y = t.y0 + TRI_DY(c), z = t.z0 + TRI_DZ(c);
    if (x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D)
      atomicAdd(g + ((size_t)z * H + y) * W + x, __fmul_rn(t.w[c], go));
  }
}

// the same scatter, also marking the 4^3 bricks it touches (flags may be NULL; plain byte stores of 1)
__device__ __forceinline__ void tri_scatter1_flag(float* __restrict__ g, unsigned char* __restrict__ flags, const Tri& t, int D,
                                                  int H, int W, float go) {
  const int nbY = (H + 3) >> 2, nbZ = (W + 3) >> 2;
  int last = -1;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int x = t.x0 + TRI_DX(c), y = t.y0 + TRI_DY(c), z = t.z0 + TRI_DZ(c);
    if (x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D) {
      atomicAdd(g + ((size_t)z * H + y) * W + x, __fmul_rn(t.w[c], go));
      if (flags) {
        const int b = ((z >> 2) * nbY + (y >> 2)) * nbZ + (x >> 2);
        if (b != last && flags[b] == 0) flags[b] = 1;      // (the L1-cached test spares most of the stores)
        last = b;
      }
    }
  }
}

// 3-element sums in the association torch's CUDA reduction kernels use (probed bit-for-bit on
// a B200 with tools/probe_torch_arith.py, 2^20 random triples each):
//   reduce over a contiguous innermost dim of 3 ((a*b).sum(-1), x.norm(dim=-1),
//   torch.sum(x**2,-1)):  (e0 + e2) + e1        -- two lanes share a row, then combine
//   reduce over a strided dim of 3 (grad.norm(dim=1) of [M,3,nd]):  (e0 + e1) + e2
__device__ __forceinline__ float sum3_inner(float e0, float e1, float e2) {
  return __fadd_rn(__fadd_rn(e0, e2), e1);
}
__device__ __forceinline__ float sum3_outer(float e0, float e1, float e2) {
  return __fadd_rn(__fadd_rn(e0, e1), e2);
}
// torch.sigmoid (CUDA): 1 / (1 + exp(-x)), probed bit-exact
__device__ __forceinline__ float torch_sigmoid(float x) {
  return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
}
