// Level-2 feature kernels of the fused fine stage (sm_100a).
//
// feat_kernel: one thread per level-2 sample builds the whole rgbnet input row
//   [k0 12 | xyz_emb 33 | viewdirs_emb 21 | sdf 1 | all_feat 24 | all_grad 12 | gradient 3]
// (model/nerf.py:835-876) plus the normal, the reflection direction and its encoding
// (:835, :879-883) in ONE pass: the reference runs grid_sample (k0), sample_sdfs (24 taps ->
// 1 grid_sample + ~25 index kernels), 3 encodings, 4 concats and ~15 elementwise kernels.
// k0 is channels-last so a corner's 12 channels are three 128-bit loads.
// feat_bwd_kernel: the matching backward: k0 scatter with 128-bit vector reductions
// (red.global.add.v4.f32), 24-tap SDF scatter, and the gradient/normal/reflection chain back to
// the level-1 quantities (d_sdf1, d_grad1) that level1_bwd_kernel turns into SDF-grid gradients.
#include "fine_common.cuh"

namespace {

using namespace fine;
constexpr int kBlock = 128;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

__device__ __forceinline__ float4 fma4(float4 v, float w, float4 acc) {
  acc.x = __fmaf_rn(v.x, w, acc.x);
  acc.y = __fmaf_rn(v.y, w, acc.y);
  acc.z = __fmaf_rn(v.z, w, acc.z);
  acc.w = __fmaf_rn(v.w, w, acc.w);
  return acc;
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

// displace_list = sorted(set(grad_feat + (1.0,))) = 0.5, 1.0, 1.5, 2.0 (model/nerf.py:843-850)
__device__ __forceinline__ float disp_of(int d) { return 0.5f * (float)(d + 1); }

// independent parts of a level-2 row (blockIdx.y of feat_kernel / feat_bwd_kernel)
// forward: k0 | xyz / view encodings | reflection chain | 4 displacement levels
// backward: k0 scatter | - | reflection chain | (tap axis, displacement pair) x 6
constexpr int kTaskK0 = 0, kTaskPE = 1, kTaskRefl = 2, kTaskTap0 = 3, kNumTasks = 7, kNumTasksBwd = 9;

// (launch bound: 6 CTAs per SM = 80 registers instead of the 96 ptxas picks on its own: 46.6 -> 45.4 us on the headline
// workload; 64 / 48 registers spill into the encoding tasks and gain nothing more, 101 registers lose 4 us --
// profiles/r02_variant_sweep.jsonl.  The same sweep moved feat_bwd_kernel by less than its noise at every bound.)
__global__ void __launch_bounds__(kBlock, 6) feat_kernel(const __grid_constant__ FgsFineParams P, int only_task) {
  const GridGeo g = make_geo(P.xyz_min, P.xyz_max, P.X, P.Y, P.Z);
  const int M2 = min(P.counters[FGS_CNT_M2], P.cap_m2);
  // blockIdx.y = which independent part of the row this CTA builds (uniform per CTA: no divergence).  With one
  // thread building a whole row only ~1300 warps existed for ~40 K samples (sm__warps_active 14 %, 79 us).
  const int task = blockIdx.y;
  if (only_task >= 0 && task != only_task) return;    // timing experiments (fgs_debug_feat_task)
  for (int s = blockIdx.x * kBlock + threadIdx.x; s < M2; s += gridDim.x * kBlock) {
    const Sample q = load_sample(P, s);
    float* __restrict__ row = P.X0 + (size_t)s * FGS_X0_LD;
    float iw, ih, id;
    cont_index(g, q.px, q.py, q.pz, iw, ih, id);
    // ---- k0: DenseGrid.forward (model/grid.py:49-59), channels-last storage ----
    if (task == kTaskK0) {
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
      st4(row + FGS_X0_K0, a0);
      st4(row + FGS_X0_K0 + 4, a1);
      st4(row + FGS_X0_K0 + 8, a2);
    }
    // ---- positional encodings (model/nerf.py:836-839, 870-873) ----
    // (sincosf shares the argument reduction of the sinf / cosf pair it replaces; the 54 values + sdf are assembled
    // in registers and leave as 16-byte stores: columns 12..63 are 13 float4, 64..66 scalars)
    if (task == kTaskPE) {
      float e54[56];
      const float rx[3] = {__fdiv_rn(__fsub_rn(q.px, g.mn[0]), g.ext[0]),
                           __fdiv_rn(__fsub_rn(q.py, g.mn[1]), g.ext[1]),
                           __fdiv_rn(__fsub_rn(q.pz, g.mn[2]), g.ext[2])};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        e54[c] = rx[c];
        float f = 1.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          float sn, cs;
          sincosf(__fmul_rn(rx[c], f), &sn, &cs);
          e54[3 + c * 5 + k] = sn;
          e54[18 + c * 5 + k] = cs;
          f *= 2.f;
        }
      }
      const float v[3] = {P.viewdirs[3 * (size_t)q.ray], P.viewdirs[3 * (size_t)q.ray + 1],
                          P.viewdirs[3 * (size_t)q.ray + 2]};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        e54[33 + c] = v[c];
        float f = 1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float sn, cs;
          sincosf(__fmul_rn(v[c], f), &sn, &cs);
          e54[33 + 3 + c * 3 + k] = sn;
          e54[33 + 12 + c * 3 + k] = cs;
          f *= 2.f;
        }
      }
      e54[54] = P.sdf1[q.idx1];
      static_assert(FGS_X0_XYZ % 4 == 0 && FGS_X0_VIEW == FGS_X0_XYZ + 33 && FGS_X0_SDF == FGS_X0_XYZ + 54, "row layout");
#pragma unroll
      for (int j = 0; j < 13; ++j)
        st4(row + FGS_X0_XYZ + 4 * j, make_float4(e54[4 * j], e54[4 * j + 1], e54[4 * j + 2], e54[4 * j + 3]));
      row[FGS_X0_XYZ + 52] = e54[52];
      row[FGS_X0_XYZ + 53] = e54[53];
      row[FGS_X0_SDF] = e54[54];
    }
    const float v[3] = {P.viewdirs[3 * (size_t)q.ray], P.viewdirs[3 * (size_t)q.ray + 1],
                        P.viewdirs[3 * (size_t)q.ray + 2]};
    // ---- hierarchical SDF taps (sample_sdfs, model/nerf.py:597-637, use_grad_norm=True) ----
    if (task >= kTaskTap0 && task < kTaskTap0 + 4) {
      const int d = task - kTaskTap0;
      float gr[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const TapPair tp = make_taps(g, iw, ih, id, a, disp_of(d));
        const float flo = tri_sample1(P.sdf, tp.lo, g.X, g.Y, g.Z);
        const float fhi = tri_sample1(P.sdf, tp.hi, g.X, g.Y, g.Z);
        row[FGS_X0_FEAT + (2 * a) * 4 + d] = flo;
        row[FGS_X0_FEAT + (2 * a + 1) * 4 + d] = fhi;
        gr[a] = tap_grad(flo, fhi, tp.diff, P.voxel_size);
      }
      const float nrm = __fsqrt_rn(sum3_outer(__fmul_rn(gr[0], gr[0]), __fmul_rn(gr[1], gr[1]),
                                              __fmul_rn(gr[2], gr[2])));
      const float den = __fadd_rn(nrm, 1e-5f);
#pragma unroll
      for (int a = 0; a < 3; ++a) row[FGS_X0_AGRAD + a * 4 + d] = __fdiv_rn(gr[a], den);
    }
    if (task != kTaskRefl) continue;
    const float gx = P.grad1[q.idx1], gy = P.grad1[(size_t)P.cap_m1 + q.idx1],
                gz = P.grad1[2 * (size_t)P.cap_m1 + q.idx1];
    row[FGS_X0_GRAD] = gx; row[FGS_X0_GRAD + 1] = gy; row[FGS_X0_GRAD + 2] = gz;
#pragma unroll
    for (int c = FGS_X0_DIM; c < FGS_X0_LD; ++c) row[c] = 0.f;
    // ---- normal, reflection direction and its encoding (model/nerf.py:835, 879-882) ----
    // (one task: splitting the 24 sincosf over three per-component tasks made the kernel slower, 47 -> 56 us -- every
    // extra task repeats the sample's address chain, and the kernel is bound by that, not by the longest task)
    const NormalCtx nc = make_normal(gx, gy, gz);
    const float dot = sum3_inner(__fmul_rn(v[0], nc.nx), __fmul_rn(v[1], nc.ny), __fmul_rn(v[2], nc.nz));
    const float n3[3] = {nc.nx, nc.ny, nc.nz};
    float* __restrict__ xr = P.XR + (size_t)s * FGS_XR_LD + FGS_MLP_WIDTH;
    {
      float e64[FGS_XR_LD - FGS_MLP_WIDTH];             // 51 values + zero pad, assembled in registers, 16 float4 stores
#pragma unroll
      for (int c = FGS_XR_DIM - FGS_MLP_WIDTH; c < FGS_XR_LD - FGS_MLP_WIDTH; ++c) e64[c] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float r = __fsub_rn(v[c], __fmul_rn(__fmul_rn(2.f, dot), n3[c]));
        e64[c] = r;
        float f = 1.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float sn, cs;
          sincosf(__fmul_rn(r, f), &sn, &cs);
          e64[3 + c * 8 + k] = sn;
          e64[27 + c * 8 + k] = cs;
          f *= 2.f;
        }
      }
#pragma unroll
      for (int j = 0; j < (FGS_XR_LD - FGS_MLP_WIDTH) / 4; ++j)
        st4(xr + 4 * j, make_float4(e64[4 * j], e64[4 * j + 1], e64[4 * j + 2], e64[4 * j + 3]));
    }
    st4(P.normal2 + 4 * (size_t)s, make_float4(nc.nx, nc.ny, nc.nz, dot));
  }
}

__global__ void __launch_bounds__(kBlock)
feat_bwd_kernel(const __grid_constant__ FgsFineParams P, float weight_orientation, int only_task) {
  const GridGeo g = make_geo(P.xyz_min, P.xyz_max, P.X, P.Y, P.Z);
  const int M2 = min(P.counters[FGS_CNT_M2], P.cap_m2);
  __shared__ float s_acc[32 * kBlock];
  const int task = blockIdx.y;     // k0 scatter | (unused) | reflection chain | the three tap axes | (unused)
  if (task == kTaskPE) return;     // (the encodings of xyz / viewdirs have no trainable inputs)
  if (only_task >= 0 && task != only_task) return;
  for (int s = blockIdx.x * kBlock + threadIdx.x; s < M2; s += gridDim.x * kBlock) {
    const Sample q = load_sample(P, s);
    const float* __restrict__ row = P.X0 + (size_t)s * FGS_X0_LD;
    const float* __restrict__ drow = P.dX0 + (size_t)s * FGS_X0_LD;
    float iw, ih, id;
    cont_index(g, q.px, q.py, q.pz, iw, ih, id);
    // ---- k0 scatter (grid_sample backward wrt input), 3 vector reductions per corner ----
    if (task == kTaskK0) {
      const Tri t = tri_setup(iw, ih, id);
      const float4 d0 = ld4(drow + FGS_X0_K0), d1 = ld4(drow + FGS_X0_K0 + 4), d2 = ld4(drow + FGS_X0_K0 + 8);
      const int nbY = (g.Y + 3) >> 2, nbZ = (g.Z + 3) >> 2;
      int last_brick = -1;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int x = t.x0 + TRI_DX(c), y = t.y0 + TRI_DY(c), z = t.z0 + TRI_DZ(c);
        if (x >= 0 && x < g.Z && y >= 0 && y < g.Y && z >= 0 && z < g.X) {
          if (P.k0_brick_flags) {   // touched-brick flag for the sparse masked Adam / gradient exchange
            const int brick = ((z >> 2) * nbY + (y >> 2)) * nbZ + (x >> 2);
            if (brick != last_brick && P.k0_brick_flags[brick] == 0) P.k0_brick_flags[brick] = 1;
            last_brick = brick;
          }
          const size_t vox = ((size_t)z * g.Y + y) * g.Z + x;
          if (P.k0_vox_flags) P.k0_vox_flags[vox] = 1;
          float* v = P.g_k0 + vox * FGS_K0_DIM;
          const float w = t.w[c];
          red_add_v4(v, w * d0.x, w * d0.y, w * d0.z, w * d0.w);
          red_add_v4(v + 4, w * d1.x, w * d1.y, w * d1.z, w * d1.w);
          red_add_v4(v + 8, w * d2.x, w * d2.y, w * d2.z, w * d2.w);
        }
      }
    }
    // ---- 24 hierarchical taps: direct (all_feat) + through the normalised gradients.  One task per AXIS: the eight
    // taps of an axis (4 displacements x lo / hi) lie on one line -- they share the 2 x 2 footprint on the other two axes
    // and cover at most 8 cells along the axis -- so their 64 corner contributions are summed per voxel in shared memory
    // first (32 accumulators per thread) before the reductions (shared-memory pre-aggregation of the scatter).  Two tasks
    // per axis (displacements {0.5, 1.0} and {1.5, 2.0}: 32 corner contributions each): three axis-wide tasks left the
    // kernel with 4 K long-running warps, 22-26 us per task; 192 -> ~100 atomics per sample over the six tasks. ----
    if (task >= kTaskTap0 && task < kTaskTap0 + 6) {
      const int a = (task - kTaskTap0) >> 1;                // 0 = W / grid z, 1 = H / grid y, 2 = D / grid x
      const int d_begin = ((task - kTaskTap0) & 1) * 2;     // displacements {0.5, 1.0} or {1.5, 2.0}
      float* acc = s_acc + threadIdx.x;                     // acc[q * kBlock], q = slot * 4 + corner on the other axes
#pragma unroll
      for (int q = 0; q < 32; ++q) acc[q * kBlock] = 0.f;
      const float ia[3] = {iw, ih, id};
      const int S[3] = {g.Z, g.Y, g.X};
      const int b0 = (int)floorf(ia[a]) - 3;               // cells b0 .. b0 + 7 hold every corner of the eight taps
      int o0 = 0, o1 = 0;                                   // floor corner on the two other axes (the same for all taps)
#pragma unroll 1
      for (int d = d_begin; d < d_begin + 2; ++d) {
        const float disp = disp_of(d);
        float gr[3], dG[3], diff[3];
#pragma unroll
        for (int a2 = 0; a2 < 3; ++a2) {
          diff[a2] = __fsub_rn(tap_clamp(ia[a2], disp, S[a2]), tap_clamp(ia[a2], -disp, S[a2]));
          gr[a2] = tap_grad(row[FGS_X0_FEAT + (2 * a2) * 4 + d], row[FGS_X0_FEAT + (2 * a2 + 1) * 4 + d], diff[a2], P.voxel_size);
          dG[a2] = drow[FGS_X0_AGRAD + a2 * 4 + d];
        }
        const float nrm = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
        const float inv = 1.f / (nrm + 1e-5f);
        const float dotg = dG[0] * gr[0] + dG[1] * gr[1] + dG[2] * gr[2];
        const float k = (nrm > 0.f) ? dotg * inv * inv / nrm : 0.f;
        const float dgr = dG[a] * inv - k * gr[a];
        const float df = dgr / diff[a] / P.voxel_size;
        const float go[2] = {drow[FGS_X0_FEAT + (2 * a) * 4 + d] - df, drow[FGS_X0_FEAT + (2 * a + 1) * 4 + d] + df};
        const TapPair tp = make_taps(g, iw, ih, id, a, disp);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const Tri& t = side ? tp.hi : tp.lo;
          const int ai = (a == 0 ? t.x0 : a == 1 ? t.y0 : t.z0) - b0;
          o0 = (a == 0) ? t.y0 : t.x0;
          o1 = (a == 2) ? t.y0 : t.z0;
          if (go[side] == 0.f) continue;
          if (ai < 0 || ai > 6) {                           // cannot happen for points inside the bbox; stay correct anyway
            tri_scatter1_flag(P.g_sdf, P.sdf_brick_flags, t, g.X, g.Y, g.Z, go[side]);
            continue;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int bit = (c >> a) & 1;
            const int oc = (a == 0) ? (c >> 1) : (a == 1) ? ((c & 1) | ((c >> 2) << 1)) : (c & 3);
            acc[((ai + bit) * 4 + oc) * kBlock] += __fmul_rn(t.w[c], go[side]);
          }
        }
      }
      int last_brick = -1;
#pragma unroll 4
      for (int q = 0; q < 32; ++q) {
        const float v = acc[q * kBlock];
        if (v == 0.f) continue;
        const int line = b0 + (q >> 2), c0 = o0 + (q & 1), c1 = o1 + ((q >> 1) & 1);
        const int x = (a == 0) ? line : c0;                 // W / grid z
        const int y = (a == 1) ? line : (a == 0 ? c0 : c1); // H / grid y
        const int z = (a == 2) ? line : c1;                 // D / grid x
        if (x >= 0 && x < g.Z && y >= 0 && y < g.Y && z >= 0 && z < g.X) {
          atomicAdd(P.g_sdf + ((size_t)z * g.Y + y) * g.Z + x, v);
          if (P.sdf_brick_flags) {
            const int b = ((z >> 2) * ((g.Y + 3) >> 2) + (y >> 2)) * ((g.Z + 3) >> 2) + (x >> 2);
            if (b != last_brick && P.sdf_brick_flags[b] == 0) P.sdf_brick_flags[b] = 1;
            last_brick = b;
          }
        }
      }
    }
    if (task != kTaskRefl) continue;
    // ---- reflection encoding -> reflect dir -> normal -> gradient ----
    const float gx = P.grad1[q.idx1], gy = P.grad1[(size_t)P.cap_m1 + q.idx1],
                gz = P.grad1[2 * (size_t)P.cap_m1 + q.idx1];
    const NormalCtx nc = make_normal(gx, gy, gz);
    const float v[3] = {P.viewdirs[3 * (size_t)q.ray], P.viewdirs[3 * (size_t)q.ray + 1],
                        P.viewdirs[3 * (size_t)q.ray + 2]};
    const float n3[3] = {nc.nx, nc.ny, nc.nz};
    const float* __restrict__ xr = P.XR + (size_t)s * FGS_XR_LD + FGS_MLP_WIDTH;
    const float* __restrict__ dxr = P.dXR + (size_t)s * FGS_XR_LD + FGS_MLP_WIDTH;
    float dr[3];
    {
      float xl[52], dl[52];                              // the 51 encoding values / their gradients: 13 float4 loads each
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const float4 a4 = ld4(xr + 4 * j), b4 = ld4(dxr + 4 * j);
        xl[4 * j] = a4.x; xl[4 * j + 1] = a4.y; xl[4 * j + 2] = a4.z; xl[4 * j + 3] = a4.w;
        dl[4 * j] = b4.x; dl[4 * j + 1] = b4.y; dl[4 * j + 2] = b4.z; dl[4 * j + 3] = b4.w;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float acc = dl[c];
        float f = 1.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc += f * (xl[27 + c * 8 + k] * dl[3 + c * 8 + k] - xl[3 + c * 8 + k] * dl[27 + c * 8 + k]);
          f *= 2.f;
        }
        dr[c] = acc;
      }
    }
    const float vn = v[0] * n3[0] + v[1] * n3[1] + v[2] * n3[2];
    const float drn = dr[0] * n3[0] + dr[1] * n3[1] + dr[2] * n3[2];
    float dn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) dn[c] = -2.f * (drn * v[c] + vn * dr[c]);
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
    float dg[3];
    {
      const float g3[3] = {gx, gy, gz};
      const float inv = 1.f / nc.den;
      const float dug = du[0] * g3[0] + du[1] * g3[1] + du[2] * g3[2];
      const float k = (nc.gn > 0.f) ? dug * inv * inv / nc.gn : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) dg[c] = du[c] * inv - k * g3[c];
    }
    P.d_sdf1[q.idx1] = drow[FGS_X0_SDF];
    P.d_grad1[q.idx1] = dg[0] + drow[FGS_X0_GRAD];
    P.d_grad1[(size_t)P.cap_m1 + q.idx1] = dg[1] + drow[FGS_X0_GRAD + 1];
    P.d_grad1[2 * (size_t)P.cap_m1 + q.idx1] = dg[2] + drow[FGS_X0_GRAD + 2];
  }
}

}  // namespace

namespace fine {

int g_only_task = -1;

int launch_feat(const FgsFineParams& P, cudaStream_t st) {
  const int grid = fgs::stream_grid(P.cap_m2, kBlock, 4);
  feat_kernel<<<dim3(grid, kNumTasks), kBlock, 0, st>>>(P, g_only_task);
  return fgs::check_launch("fine feat", st);
}

int launch_feat_bwd(const FgsFineParams& P, float weight_orientation, cudaStream_t st) {
  const int grid = fgs::stream_grid(P.cap_m2, kBlock, 4);
  feat_bwd_kernel<<<dim3(grid, kNumTasksBwd), kBlock, 0, st>>>(P, weight_orientation, g_only_task);
  return fgs::check_launch("fine feat_bwd", st);
}

}  // namespace fine

// Timing aid: run only one of the independent row parts of feat_kernel / feat_bwd_kernel (-1 = all).
extern "C" int fgs_debug_feat_task(int task) {
  fine::g_only_task = task;
  return 0;
}
