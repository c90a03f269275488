// Coarse stage of the per-ray path (sm_100a): nerf.forward_coarse (model/nerf.py:943-1075), its backward and the
// C-ABI that sequences them (include/fgs_b200_fine.h, "coarse stage").
//
// Shared with the fine stage: march_kernel / expand1_kernel (sampling, MaskCache, level-1 ids), compact2_kernel,
// composite_kernel (+ loss), weights1_bwd_kernel and the tensor-core MLP kernels of gemm_tcw.cu (the refnet
// 90 -> 192 -> 192 -> 3 runs zero-padded on the 256-wide chain / weight-gradient kernels).  Kernels of this file:
//   coarse_sample1_kernel     thread per level-1 sample: smoothed SDF (8 taps), gradient GRID (3 x 8 taps), NeuS alpha
//   coarse_weights1_kernel    warp per ray: BOTH transmittance recurrences of :978-990 in one in-order sweep
//   coarse_feat_kernel        thread per level-2 sample: refnet input row [k0 | xyz_emb | reflect_emb | normal | view_emb]
//   coarse_feat_bwd_kernel    k0 scatter (vector reductions, touched-brick flags) and the reflect / normal chain
//   coarse_sample1_bwd_kernel NeuS alpha backward and the 8 + 24 trilinear scatters into g_smooth / g_grad_grid
#include "fine_common.cuh"
#include "tc_api.cuh"

namespace fine {
int launch_march_expand(const FgsFineParams& P, cudaStream_t st);
int launch_compact2(const FgsFineParams& P, cudaStream_t st);
int launch_composite(const FgsFineParams& P, const FgsFineLoss* L, cudaStream_t st);
int launch_weights1_bwd(const FgsFineParams& P, const float* g_last, cudaStream_t st);
}  // namespace fine

namespace {

using namespace fine;
constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr int W = FGS_MLP_WIDTH;
constexpr int kLast = FGS_MLP_LAYERS - 1;      // slot of the refnet's output layer (composite_kernel adds its bias gradient there)

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float4 fma4(float4 v, float w, float4 acc) {
  acc.x = __fmaf_rn(v.x, w, acc.x);
  acc.y = __fmaf_rn(v.y, w, acc.y);
  acc.z = __fmaf_rn(v.z, w, acc.z);
  acc.w = __fmaf_rn(v.w, w, acc.w);
  return acc;
}

// column map of the refnet input row (concat order of model/nerf.py:1005)
struct RowMap {
  int k0, xyz, ref, nrm, view, dim;
};
__host__ __device__ inline RowMap row_map(const FgsCoarse& C) {
  RowMap m;
  m.k0 = 0;
  m.xyz = (C.k0_dim > 0 && C.k0_dim < FGS_K0_DIM) ? C.k0_dim : FGS_K0_DIM;
  m.ref = m.xyz + 3 + 6 * C.pos_pe;
  m.nrm = m.ref + 3 + 6 * C.ref_pe;
  m.view = m.nrm + 3;
  m.dim = m.view + 3 + 6 * C.view_pe;
  return m;
}

// ---- level 1: values (model/nerf.py:969-976) ------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
coarse_sample1_kernel(const __grid_constant__ FgsFineParams P, const __grid_constant__ FgsCoarse C) {
  const GridGeo g = make_geo(P.xyz_min, P.xyz_max, P.X, P.Y, P.Z);
  const size_t V = (size_t)P.X * P.Y * P.Z;
  const int M1 = min(P.counters[FGS_CNT_M1], P.cap_m1);
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < M1; i += gridDim.x * kBlock) {
    const int ray = P.ray_id1[i];
    if (ray < 0) continue;
    const int step = P.step_id1[i];
    const float vx = P.viewdirs[3 * (size_t)ray], vy = P.viewdirs[3 * (size_t)ray + 1], vz = P.viewdirs[3 * (size_t)ray + 2];
    const Ray r = make_ray(P.rays_o + 3 * (size_t)ray, P.rays_d + 3 * (size_t)ray, P.t_min[ray]);
    float px, py, pz, iw, ih, id;
    ray_point(r, P.stepdist, step, px, py, pz);
    cont_index(g, px, py, pz, iw, ih, id);
    const Tri t = tri_setup(iw, ih, id);
    // grid_sampler(ray_pts, smooth_conv(sdf)) and grid_sampler(ray_pts, neus_sdf_gradient()): one F.grid_sample each,
    // the 3-channel one accumulates every channel over the corners in the same order
    const float sdf = tri_sample1(C.sdf_smooth, t, g.X, g.Y, g.Z);
    const float gx = tri_sample1(C.grad_grid, t, g.X, g.Y, g.Z);
    const float gy = tri_sample1(C.grad_grid + V, t, g.X, g.Y, g.Z);
    const float gz = tri_sample1(C.grad_grid + 2 * V, t, g.X, g.Y, g.Z);
    P.sdf1[i] = sdf;
    P.grad1[i] = gx;
    P.grad1[(size_t)P.cap_m1 + i] = gy;
    P.grad1[2 * (size_t)P.cap_m1 + i] = gz;
    P.alpha1[i] = neus_alpha(sdf, gx, gy, gz, vx, vy, vz, P.dist, P.inv_s);
  }
}

// ---- level 1: the two alpha2weight passes (model/nerf.py:978-990; render_utils_kernel.cu:577-605) --------------------
// Pass 1 runs over ALL level-1 samples of the ray; a sample is kept iff its pass-1 weight > fast_color_thres; pass 2 is
// the same recurrence over the kept samples only.  The decision for sample j only needs pass 1 up to j, so both
// recurrences advance together, in order.  Pass 1's stop ends the sweep (later weights are 0: never kept).
__global__ void __launch_bounds__(kBlock) coarse_weights1_kernel(const __grid_constant__ FgsFineParams P) {
  const int lane = lane_id();
  const float thres = P.fast_color_thres;
  int tot_a = 0;
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < P.n_rays; ray += gridDim.x * kWarps) {
    const int base = P.off1[ray];
    const int c1 = P.cnt1[ray];
    if (c1 == 0 || base + c1 > P.cap_m1) {
      if (lane == 0) { P.cnt2[ray] = 0; P.alphainv_last[ray] = 1.f; }
      continue;
    }
    if (lane == 0) tot_a += c1;
    float T1c = 1.f, T2c = 1.f;
    bool stop1 = false, stop2 = false;
    int c2 = 0;
    for (int b = 0; b < c1; b += 32) {
      const bool mine = b + lane < c1;
      const int idx = base + b + lane;
      const float alpha = mine ? P.alpha1[idx] : 0.f;
      const int cnt = min(32, c1 - b);
      float myW = 0.f, myT = 1.f;
      bool myL2 = false, myVis = false;
      if (!stop1) {
        for (int j = 0; j < cnt; ++j) {
          const float aj = __shfl_sync(0xffffffffu, alpha, j);
          const float w1j = __fmul_rn(T1c, aj);
          T1c = (float)__dmul_rn(__dsub_rn(1., (double)aj), (double)T1c);
          if (w1j > thres) {                                  // model/nerf.py:981
            bool vis = false;
            float Tj = 1.f, w2j = 0.f;
            if (!stop2) {
              vis = true;
              Tj = T2c;
              w2j = __fmul_rn(T2c, aj);
              T2c = (float)__dmul_rn(__dsub_rn(1., (double)aj), (double)T2c);
              if ((double)T2c < 1e-3) stop2 = true;
            }
            if (lane == j) { myL2 = true; myVis = vis; myW = w2j; myT = Tj; }
          }
          if ((double)T1c < 1e-3) { stop1 = true; break; }
        }
      }
      if (mine) {
        P.w1[idx] = myW;
        P.T1[idx] = myT;
        P.flag1[idx] = FGS_F_ALPHA | (myVis ? FGS_F_VISIT : 0) | (myL2 ? FGS_F_L2 : 0);
      }
      c2 += __popc(__ballot_sync(0xffffffffu, myL2));
    }
    if (lane == 0) {
      P.cnt2[ray] = c2;
      P.alphainv_last[ray] = T2c;
    }
  }
  __shared__ int s_a[kWarps];
  if (lane == 0) s_a[threadIdx.x >> 5] = tot_a;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) t += s_a[i];
    if (t) atomicAdd(P.counters + FGS_CNT_MA, t);
  }
}

struct Sample {
  int idx1, ray, step;
  float px, py, pz;
};
__device__ __forceinline__ Sample load_sample(const FgsFineParams& P, int s) {
  Sample q;
  q.idx1 = P.idx1_of2[s];
  q.ray = P.ray_id2[s];
  q.step = P.step_id1[q.idx1];
  const Ray r = make_ray(P.rays_o + 3 * (size_t)q.ray, P.rays_d + 3 * (size_t)q.ray, P.t_min[q.ray]);
  ray_point(r, P.stepdist, q.step, q.px, q.py, q.pz);
  return q;
}

// ---- level 2: the refnet input row (model/nerf.py:992-1005) ---------------------------------------------------------
constexpr int kTaskK0 = 0, kTaskPE = 1, kTaskRefl = 2, kNumTasks = 3;

__device__ __forceinline__ void encode3(float* __restrict__ dst, const float v[3], int pe) {
  // cat([v, sin(v * 2^k), cos(v * 2^k)]) with (v.unsqueeze(-1) * freq).flatten(-2): component-major, then k
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dst[c] = v[c];
    float f = 1.f;
    for (int k = 0; k < pe; ++k) {
      float sn, cs;
      sincosf(__fmul_rn(v[c], f), &sn, &cs);
      dst[3 + c * pe + k] = sn;
      dst[3 + 3 * pe + c * pe + k] = cs;
      f *= 2.f;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
coarse_feat_kernel(const __grid_constant__ FgsFineParams P, const __grid_constant__ FgsCoarse C) {
  const GridGeo g = make_geo(P.xyz_min, P.xyz_max, P.X, P.Y, P.Z);
  const RowMap rm = row_map(C);
  const int M2 = min(P.counters[FGS_CNT_M2], P.cap_m2);
  const int task = blockIdx.y;
  for (int s = blockIdx.x * kBlock + threadIdx.x; s < M2; s += gridDim.x * kBlock) {
    const Sample q = load_sample(P, s);
    float* __restrict__ row = P.X0 + (size_t)s * C.in_ld;
    if (task == kTaskK0) {
      // DenseGrid.forward (model/grid.py:49-59), channels-last storage
      float iw, ih, id;
      cont_index(g, q.px, q.py, q.pz, iw, ih, id);
      const Tri t = tri_setup(iw, ih, id);
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int x = t.x0 + TRI_DX(c), y = t.y0 + TRI_DY(c), z = t.z0 + TRI_DZ(c);
        if (x >= 0 && x < g.Z && y >= 0 && y < g.Y && z >= 0 && z < g.X) {
          const float* v = P.k0 + (((size_t)z * g.Y + y) * g.Z + x) * FGS_K0_DIM;
          a0 = fma4(__ldg(reinterpret_cast<const float4*>(v)), t.w[c], a0);
          a1 = fma4(__ldg(reinterpret_cast<const float4*>(v) + 1), t.w[c], a1);
          a2 = fma4(__ldg(reinterpret_cast<const float4*>(v) + 2), t.w[c], a2);
        }
      }
      if (rm.xyz == FGS_K0_DIM) {
        st4(row + rm.k0, a0);
        st4(row + rm.k0 + 4, a1);
        st4(row + rm.k0 + 8, a2);
      } else {                                              // fewer feature channels: only they enter the row
        const float a[FGS_K0_DIM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
        for (int ch = 0; ch < FGS_K0_DIM; ++ch)
          if (ch < rm.xyz) row[rm.k0 + ch] = a[ch];
      }
    } else if (task == kTaskPE) {
      const float rx[3] = {__fdiv_rn(__fsub_rn(q.px, g.mn[0]), g.ext[0]), __fdiv_rn(__fsub_rn(q.py, g.mn[1]), g.ext[1]),
                           __fdiv_rn(__fsub_rn(q.pz, g.mn[2]), g.ext[2])};
      encode3(row + rm.xyz, rx, C.pos_pe);
      const float v[3] = {P.viewdirs[3 * (size_t)q.ray], P.viewdirs[3 * (size_t)q.ray + 1], P.viewdirs[3 * (size_t)q.ray + 2]};
      encode3(row + rm.view, v, C.view_pe);
      for (int c = rm.dim; c < C.in_ld; ++c) row[c] = 0.f;
    } else {
      // normal, reflection direction and its encoding (model/nerf.py:991, 996-998)
      const float gx = P.grad1[q.idx1], gy = P.grad1[(size_t)P.cap_m1 + q.idx1], gz = P.grad1[2 * (size_t)P.cap_m1 + q.idx1];
      const NormalCtx nc = make_normal(gx, gy, gz);
      const float v[3] = {P.viewdirs[3 * (size_t)q.ray], P.viewdirs[3 * (size_t)q.ray + 1], P.viewdirs[3 * (size_t)q.ray + 2]};
      const float dot = sum3_inner(__fmul_rn(v[0], nc.nx), __fmul_rn(v[1], nc.ny), __fmul_rn(v[2], nc.nz));
      const float n3[3] = {nc.nx, nc.ny, nc.nz};
      float rr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) rr[c] = __fsub_rn(v[c], __fmul_rn(__fmul_rn(2.f, dot), n3[c]));
      encode3(row + rm.ref, rr, C.ref_pe);
      row[rm.nrm] = nc.nx; row[rm.nrm + 1] = nc.ny; row[rm.nrm + 2] = nc.nz;
      st4(P.normal2 + 4 * (size_t)s, make_float4(nc.nx, nc.ny, nc.nz, dot));
    }
  }
}

__global__ void __launch_bounds__(kBlock)
coarse_feat_bwd_kernel(const __grid_constant__ FgsFineParams P, const __grid_constant__ FgsCoarse C, float weight_orientation) {
  const GridGeo g = make_geo(P.xyz_min, P.xyz_max, P.X, P.Y, P.Z);
  const RowMap rm = row_map(C);
  const int M2 = min(P.counters[FGS_CNT_M2], P.cap_m2);
  const int task = blockIdx.y;     // 0: k0 scatter, 1: reflect / normal chain
  for (int s = blockIdx.x * kBlock + threadIdx.x; s < M2; s += gridDim.x * kBlock) {
    const Sample q = load_sample(P, s);
    const float* __restrict__ row = P.X0 + (size_t)s * C.in_ld;
    const float* __restrict__ drow = P.dX0 + (size_t)s * C.in_ld;
    if (task == 0) {
      float iw, ih, id;
      cont_index(g, q.px, q.py, q.pz, iw, ih, id);
      const Tri t = tri_setup(iw, ih, id);
      float4 d0, d1, d2;
      if (rm.xyz == FGS_K0_DIM) {
        d0 = ld4(drow + rm.k0); d1 = ld4(drow + rm.k0 + 4); d2 = ld4(drow + rm.k0 + 8);
      } else {                                              // the grid's pad channels receive exactly 0
        float d[FGS_K0_DIM];
#pragma unroll
        for (int ch = 0; ch < FGS_K0_DIM; ++ch) d[ch] = ch < rm.xyz ? drow[rm.k0 + ch] : 0.f;
        d0 = make_float4(d[0], d[1], d[2], d[3]); d1 = make_float4(d[4], d[5], d[6], d[7]); d2 = make_float4(d[8], d[9], d[10], d[11]);
      }
      const int nbY = (g.Y + 3) >> 2, nbZ = (g.Z + 3) >> 2;
      int last_brick = -1;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int x = t.x0 + TRI_DX(c), y = t.y0 + TRI_DY(c), z = t.z0 + TRI_DZ(c);
        if (x >= 0 && x < g.Z && y >= 0 && y < g.Y && z >= 0 && z < g.X) {
          if (P.k0_brick_flags) {
            const int brick = ((z >> 2) * nbY + (y >> 2)) * nbZ + (x >> 2);
            if (brick != last_brick && P.k0_brick_flags[brick] == 0) P.k0_brick_flags[brick] = 1;
            last_brick = brick;
          }
          const size_t vox = ((size_t)z * g.Y + y) * g.Z + x;
          float* v = P.g_k0 + vox * FGS_K0_DIM;
          const float w = t.w[c];
          red_add_v4(v, w * d0.x, w * d0.y, w * d0.z, w * d0.w);
          red_add_v4(v + 4, w * d1.x, w * d1.y, w * d1.z, w * d1.w);
          red_add_v4(v + 8, w * d2.x, w * d2.y, w * d2.z, w * d2.w);
        }
      }
      continue;
    }
    // ---- reflection encoding -> reflect dir -> normal <- normal feature, orientation loss; normal -> gradient ----
    const float gx = P.grad1[q.idx1], gy = P.grad1[(size_t)P.cap_m1 + q.idx1], gz = P.grad1[2 * (size_t)P.cap_m1 + q.idx1];
    const NormalCtx nc = make_normal(gx, gy, gz);
    const float v[3] = {P.viewdirs[3 * (size_t)q.ray], P.viewdirs[3 * (size_t)q.ray + 1], P.viewdirs[3 * (size_t)q.ray + 2]};
    const float n3[3] = {nc.nx, nc.ny, nc.nz};
    float dr[3];
    const int pe = C.ref_pe;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float acc = drow[rm.ref + c];
      float f = 1.f;
      for (int k = 0; k < pe; ++k) {
        const int is = rm.ref + 3 + c * pe + k, ic = is + 3 * pe;
        acc += f * (row[ic] * drow[is] - row[is] * drow[ic]);
        f *= 2.f;
      }
      dr[c] = acc;
    }
    const float vn = v[0] * n3[0] + v[1] * n3[1] + v[2] * n3[2];
    const float drn = dr[0] * n3[0] + dr[1] * n3[1] + dr[2] * n3[2];
    float dn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) dn[c] = -2.f * (drn * v[c] + vn * dr[c]) + drow[rm.nrm + c];
    // orientation loss (model/nerf.py:469-478): sum_s w.detach * min(0, n.(-v))^2
    if (weight_orientation > 0.f && -vn < 0.f) {
      const float w = P.w1[q.idx1];
      const float k = weight_orientation * w * 2.f * (-vn);
#pragma unroll
      for (int c = 0; c < 3; ++c) dn[c] += k * (-v[c]);
    }
    // normal = u / L, L = sqrt(max(|u|^2, eps))
    float du[3];
    {
      const float u3[3] = {nc.ux, nc.uy, nc.uz};
      const float invL = 1.f / nc.L;
      if (nc.ss > 1.1920928955078125e-07f) {
        const float dnu = dn[0] * u3[0] + dn[1] * u3[1] + dn[2] * u3[2];
        const float k = dnu * invL * invL * invL;
#pragma unroll
        for (int c = 0; c < 3; ++c) du[c] = dn[c] * invL - k * u3[c];
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) du[c] = dn[c] * invL;
      }
    }
    // u = g / (|g| + 1e-7)
    {
      const float g3[3] = {gx, gy, gz};
      const float inv = 1.f / nc.den;
      const float dug = du[0] * g3[0] + du[1] * g3[1] + du[2] * g3[2];
      const float k = (nc.gn > 0.f) ? dug * inv * inv / nc.gn : 0.f;
      P.d_grad1[q.idx1] = du[0] * inv - k * g3[0];
      P.d_grad1[(size_t)P.cap_m1 + q.idx1] = du[1] * inv - k * g3[1];
      P.d_grad1[2 * (size_t)P.cap_m1 + q.idx1] = du[2] * inv - k * g3[2];
    }
  }
}

// ---- level 1 backward: NeuS alpha (model/nerf.py:510-544) and the trilinear scatters into the two grids -------------
__global__ void __launch_bounds__(kBlock)
coarse_sample1_bwd_kernel(const __grid_constant__ FgsFineParams P, const __grid_constant__ FgsCoarse C) {
  const GridGeo g = make_geo(P.xyz_min, P.xyz_max, P.X, P.Y, P.Z);
  const size_t V = (size_t)P.X * P.Y * P.Z;
  const int M1 = min(P.counters[FGS_CNT_M1], P.cap_m1);
  for (int idx = blockIdx.x * kBlock + threadIdx.x; idx < M1; idx += gridDim.x * kBlock) {
    const int ray = P.ray_id1[idx];
    if (ray < 0) continue;
    const int fl = P.flag1[idx];
    if (!(fl & FGS_F_VISIT)) continue;      // (kept samples behind the second pass's stop carry weight 0: no gradient)
    const float vx = P.viewdirs[3 * (size_t)ray], vy = P.viewdirs[3 * (size_t)ray + 1], vz = P.viewdirs[3 * (size_t)ray + 2];
    const float d_alpha = P.d_w1[idx];
    const float sdf = P.sdf1[idx];
    const float gx = P.grad1[idx], gy = P.grad1[(size_t)P.cap_m1 + idx], gz = P.grad1[2 * (size_t)P.cap_m1 + idx];
    const float tc = vx * gx + vy * gy + vz * gz;
    const float ic = -fmaxf(-tc, 0.f);
    const float h = ic * P.dist * 0.5f;
    const float pc = sigm((sdf - h) * P.inv_s), nc = sigm((sdf + h) * P.inv_s);
    const float u = pc - nc + 1e-5f, q = pc + 1e-5f;
    const float a_raw = u / q;
    float d_sdf = 0.f, d_g[3] = {0.f, 0.f, 0.f};
    if (a_raw >= 0.f && a_raw <= 1.f) {
      const float d_pc = d_alpha * (q - u) / (q * q);
      const float d_nc = -d_alpha / q;
      const float d_prev = d_pc * pc * (1.f - pc) * P.inv_s;
      const float d_next = d_nc * nc * (1.f - nc) * P.inv_s;
      d_sdf = d_prev + d_next;
      const float d_ic = (d_next - d_prev) * P.dist * 0.5f;
      const float d_tc = (tc < 0.f) ? d_ic : 0.f;
      d_g[0] = d_tc * vx; d_g[1] = d_tc * vy; d_g[2] = d_tc * vz;
    }
    // every visited sample is in the pruned set: add what came back through the normal
    d_g[0] += P.d_grad1[idx];
    d_g[1] += P.d_grad1[(size_t)P.cap_m1 + idx];
    d_g[2] += P.d_grad1[2 * (size_t)P.cap_m1 + idx];
    const Ray r = make_ray(P.rays_o + 3 * (size_t)ray, P.rays_d + 3 * (size_t)ray, P.t_min[ray]);
    float px, py, pz, iw, ih, id;
    ray_point(r, P.stepdist, P.step_id1[idx], px, py, pz);
    cont_index(g, px, py, pz, iw, ih, id);
    const Tri t = tri_setup(iw, ih, id);
    if (d_sdf != 0.f) tri_scatter1(C.g_smooth, t, g.X, g.Y, g.Z, d_sdf);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (d_g[c] != 0.f) tri_scatter1(C.g_grad_grid + c * V, t, g.X, g.Y, g.Z, d_g[c]);
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
// w_split of the coarse stage: per refnet layer the straight TF32 hi / lo planes [round_up(out, 8)][in_ld] (forward) and
// the transposed bf16 hi / mid planes [in_ld][ldt] (backward-data), as fine::wsplit_layout
struct CoarseLayout {
  size_t hi[3], lo[3], thi[3], tlo[3];
  int in_ld[3], out[3], ldt[3], slot[3];
  size_t total;
};
CoarseLayout coarse_layout(int in_ld) {
  CoarseLayout L;
  const int ild[3] = {in_ld, W, W}, out[3] = {W, W, 3}, slot[3] = {0, 1, kLast};
  size_t o = 0;
  for (int l = 0; l < 3; ++l) {
    L.in_ld[l] = ild[l]; L.out[l] = out[l]; L.slot[l] = slot[l];
    L.ldt[l] = out[l] == 3 ? 8 : W;
    const size_t s = (size_t)((out[l] + 7) & ~7) * ild[l], t = (size_t)ild[l] * L.ldt[l] / 2;
    L.hi[l] = o; o += s;
    L.lo[l] = o; o += s;
    L.thi[l] = o; o += t;
    L.tlo[l] = o; o += t;
  }
  L.total = o;
  return L;
}

int check_coarse(const FgsFineParams* p, const FgsCoarse* c) {
  FGS_REQUIRE(p != nullptr && c != nullptr, "params is NULL");
  FGS_REQUIRE(p->n_rays > 0 && p->n_rays <= (1 << 20), "n_rays out of range (1..2^20)");
  FGS_REQUIRE(p->X >= 3 && p->Y >= 3 && p->Z >= 3, "grid too small");
  FGS_REQUIRE(p->cap_words > 0 && p->cap_m1 > 0 && p->cap_m2 > 0, "capacities must be > 0");
  FGS_REQUIRE(p->stepdist > 0.f, "stepdist must be > 0");
  FGS_REQUIRE(c->sdf_smooth && c->grad_grid, "coarse grids are NULL");
  FGS_REQUIRE(c->pos_pe >= 0 && c->pos_pe <= 10 && c->view_pe >= 0 && c->view_pe <= 10 && c->ref_pe >= 0 && c->ref_pe <= 10,
              "encoding frequencies out of range (0..10)");
  FGS_REQUIRE(c->k0_dim >= 0 && c->k0_dim <= FGS_K0_DIM, "k0_dim out of range (0 = 12)");
  FGS_REQUIRE(c->in_dim == row_map(*c).dim, "in_dim does not match k0_dim + the three encodings + normal");
  FGS_REQUIRE(c->in_ld >= c->in_dim && (c->in_ld & 31) == 0 && c->in_ld <= 256, "in_ld must be a multiple of 32 in [in_dim, 256]");
  FGS_REQUIRE(p->w_split && p->hbits && p->X0 && p->H && p->logit, "MLP workspaces are NULL");
  FGS_REQUIRE(c->width >= 0 && c->width <= W, "refnet width out of range (0..256)");
  return 0;
}

// N / K extent of the hidden layers in the GEMMs: the width rounded up to whole 64-k slabs of the bf16 kernels
int gemm_width(const FgsCoarse& C) { return C.width > 0 ? ((C.width + 63) & ~63) : W; }

float* hidden(const FgsFineParams& P, int i) { return P.H + (size_t)i * P.cap_m2 * W; }
uint32_t* hbits(const FgsFineParams& P, int i) { return P.hbits + (size_t)i * P.cap_m2 * 8; }

}  // namespace

extern "C" {

int64_t fgs_coarse_wsplit_floats(int in_ld) { return (int64_t)coarse_layout(in_ld).total; }

int fgs_coarse_forward(const FgsFineParams* p, const FgsCoarse* c, fgs_stream_t stream) {
  int rc = check_coarse(p, c);
  if (rc) return rc;
  const FgsFineParams& P = *p;
  const FgsCoarse& C = *c;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = launch_march_expand(P, st))) return rc;
  coarse_sample1_kernel<<<fgs::stream_grid(P.cap_m1, kBlock, 8), kBlock, 0, st>>>(P, C);
  if ((rc = fgs::check_launch("coarse level1", st))) return rc;
  coarse_weights1_kernel<<<fgs::stream_grid(P.n_rays, kWarps, 8), kBlock, 0, st>>>(P);
  if ((rc = fgs::check_launch("coarse weights1", st))) return rc;
  if ((rc = launch_compact2(P, st))) return rc;
  coarse_feat_kernel<<<dim3(fgs::stream_grid(P.cap_m2, kBlock, 4), kNumTasks), kBlock, 0, st>>>(P, C);
  if ((rc = fgs::check_launch("coarse feat", st))) return rc;
  // operand planes of the three weight matrices, then the refnet as one persistent launch (model/nerf.py:1006)
  const CoarseLayout L = coarse_layout(C.in_ld);
  SplitJobHost jobs[6];
  for (int l = 0; l < 3; ++l) {
    const float* Wl = P.W[L.slot[l]];
    FGS_REQUIRE(Wl != nullptr && P.B[L.slot[l]] != nullptr, "refnet weights (slots 0, 1, 7) are NULL");
    jobs[2 * l] = {Wl, P.w_split + L.hi[l], P.w_split + L.lo[l], L.out[l], L.in_ld[l], L.in_ld[l], L.in_ld[l], 0, 0};
    jobs[2 * l + 1] = {Wl, P.w_split + L.thi[l], P.w_split + L.tlo[l], L.out[l], L.in_ld[l], L.in_ld[l], L.ldt[l], 1, 1};
  }
  if ((rc = weight_split_jobs(jobs, 6, st))) return rc;
  const int Wp = gemm_width(C);
  auto fwd = [&](int l, const float* A, int lda, int K, float* Cout, int ldc, int N, int relu, int hid) {
    ChainStepHost s{};
    s.A = A; s.lda = lda; s.K = K;
    s.Bh = P.w_split + L.hi[l]; s.Bl = P.w_split + L.lo[l]; s.ldb = L.in_ld[l];
    s.C = Cout; s.ldc = ldc; s.N = N; s.n0 = 0;
    s.bias = P.B[L.slot[l]]; s.relu = relu;
    s.bits_in = nullptr; s.bits_out = relu ? hbits(P, hid) : nullptr;
    return s;
  };
  const ChainStepHost steps[3] = {fwd(0, P.X0, C.in_ld, C.in_ld, hidden(P, 0), W, Wp, 1, 0),
                                  fwd(1, hidden(P, 0), W, Wp, hidden(P, 1), W, Wp, 1, 1),
                                  fwd(2, hidden(P, 1), W, Wp, P.logit, 4, 3, 0, -1)};
  if ((rc = mlp_chain(steps, 3, P.cap_m2, P.counters + FGS_CNT_M2, st))) return rc;
  if (P.skip_composite) return 0;
  return launch_composite(P, nullptr, st);
}

int fgs_coarse_backward(const FgsFineParams* p, const FgsFineLoss* l, const FgsCoarse* c, fgs_stream_t stream) {
  int rc = check_coarse(p, c);
  if (rc) return rc;
  FGS_REQUIRE(l && l->g_alphainv_last, "loss descriptor incomplete");
  FGS_REQUIRE(c->g_smooth && c->g_grad_grid, "coarse gradient grids are NULL");
  FGS_REQUIRE(p->dZ && p->dlogit_p && p->dX0 && p->dlogit, "backward workspaces are NULL");
  const FgsFineParams& P = *p;
  const FgsCoarse& C = *c;
  cudaStream_t st = (cudaStream_t)stream;
  const int* m2 = P.counters + FGS_CNT_M2;
  const int cap = P.cap_m2;
  const CoarseLayout L = coarse_layout(C.in_ld);
  const size_t plane = (size_t)P.cap_m2 * W;
  auto zhi = [&](int i) { return reinterpret_cast<uint16_t*>(P.dZ + (size_t)i * P.cap_m2 * W); };
  auto zmid = [&](int i) { return zhi(i) + plane; };
  uint16_t* lg_hi = P.dlogit_p;
  uint16_t* lg_mid = P.dlogit_p + (size_t)P.cap_m2 * 8;
  auto step = [&](int layer, const uint16_t* ahi, const uint16_t* amid, int lda, int K, uint16_t* chi, uint16_t* cmid,
                  float* c32, int ldc, int N, int hid, float* colsum) {
    BwStepHost s{};
    s.Ahi = ahi; s.Amid = amid; s.lda = lda; s.K = K;
    s.Bh = P.w_split + L.thi[layer]; s.Bl = P.w_split + L.tlo[layer]; s.ldb = L.ldt[layer];
    s.C = c32; s.ldc = ldc; s.Chi = chi; s.Cmid = cmid; s.out_planes = chi != nullptr;
    s.N = N; s.n0 = 0; s.bits_in = hid >= 0 ? hbits(P, hid) : nullptr; s.colsum = colsum;
    return s;
  };
  // backward-data chain (3xBF16): dlogit -> dZ1 (masked by relu'(H1)) -> dZ0 (relu'(H0)) -> dX0; bias gradients of
  // layers 1 and 0 are the column sums of dZ1 / dZ0 (the last layer's came from the compositing kernel)
  const int Wp = gemm_width(C);
  const BwStepHost steps[3] = {
      step(2, lg_hi, lg_mid, 8, 3, zhi(0), zmid(0), nullptr, 0, Wp, 1, P.gB[1]),
      step(1, zhi(0), zmid(0), W, Wp, zhi(1), zmid(1), nullptr, 0, Wp, 0, P.gB[0]),
      step(0, zhi(1), zmid(1), W, Wp, nullptr, nullptr, P.dX0, C.in_ld, C.in_ld, -1, nullptr)};
  if ((rc = mlp_bwd_planes(steps, 3, cap, m2, st))) return rc;
  coarse_feat_bwd_kernel<<<dim3(fgs::stream_grid(P.cap_m2, kBlock, 4), 2), kBlock, 0, st>>>(P, C, l->weight_orientation);
  if ((rc = fgs::check_launch("coarse feat_bwd", st))) return rc;
  if ((rc = launch_weights1_bwd(P, l->g_alphainv_last, st))) return rc;
  coarse_sample1_bwd_kernel<<<fgs::stream_grid(P.cap_m1, kBlock, 8), kBlock, 0, st>>>(P, C);
  if ((rc = fgs::check_launch("coarse level1_bwd", st))) return rc;
  // weight gradients: dW_l = dZ_l^T X_l
  DwProblemHost dw[2];
  dw[0] = DwProblemHost{nullptr, P.X0, P.gW[0], nullptr, Wp, C.in_ld, 0, C.in_ld, C.in_ld};
  dw[0].Ahi = zhi(1); dw[0].Amid = zmid(1); dw[0].ldp = W;
  dw[1] = DwProblemHost{nullptr, hidden(P, 0), P.gW[1], nullptr, Wp, Wp, 0, W, W};
  dw[1].Ahi = zhi(0); dw[1].Amid = zmid(0); dw[1].ldp = W;
  if ((rc = mlp_dw(dw, 2, cap, m2, st, true))) return rc;
  return dw_small(3, Wp, cap, P.dlogit, 4, hidden(P, 1), W, P.gW[kLast], W, m2, st);
}

}  // extern "C"
