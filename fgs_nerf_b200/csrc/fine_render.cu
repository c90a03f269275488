// Compositing, loss and the per-ray backward of the fused fine stage (sm_100a).
//
// composite_kernel (warp per ray): the segment sums of model/nerf.py:888-920 (the reference
// calls torch_scatter.segment_coo 3-5 times) over the ray's contiguous level-2 run, the
// background/clamp epilogue, and -- when a loss descriptor is given -- the loss block of
// model/nerf_training.py:307-327 together with the first backward hop: d rgb_marched and
// d sigmoid_rgb are pushed straight back through the sums into dlogit and d_w1, so neither the
// per-ray gradients nor the per-sample rgb ever round-trip through HBM.
// weights1_bwd_kernel (warp per ray) + sample1_bwd_kernel (thread per level-1 sample): alpha2weight
// backward (render_utils_kernel.cu:654-677) as a reverse warp-shuffle recurrence, then the NeuS alpha
// backward (model/nerf.py:510-544) and the 1 + 6 trilinear scatters into the SDF gradient grid.
#include "fine_common.cuh"

namespace {

using namespace fine;
constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

template <bool LOSS>
__global__ void __launch_bounds__(kBlock)
composite_kernel(const __grid_constant__ FgsFineParams P, const __grid_constant__ FgsFineLoss L) {
  const int lane = lane_id();
  double l_mse = 0., l_srgb = 0., l_ori = 0., l_ent = 0., l_rgbper = 0.;
  float gb7[3] = {0.f, 0.f, 0.f};          // bias gradient of the last MLP layer = column sums of dlogit
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < P.n_rays; ray += gridDim.x * kWarps) {
    const int b = P.off2[ray];
    int c = P.cnt2[ray];
    if (b + c > P.cap_m2) c = 0;
    float a_rgb[3] = {0.f, 0.f, 0.f}, a_s[3] = {0.f, 0.f, 0.f}, a_n[3] = {0.f, 0.f, 0.f};
    float a_w = 0.f, a_d = 0.f, a_o = 0.f;
    for (int i = lane; i < c; i += 32) {
      const int s = b + i;
      const int idx1 = P.idx1_of2[s];
      const float w = P.w1[idx1];
      const float4 z = ld4(P.logit + 4 * (size_t)s);
      const float r0 = sigm(z.x), r1 = sigm(z.y), r2 = sigm(z.z);
      a_rgb[0] += w * r0; a_rgb[1] += w * r1; a_rgb[2] += w * r2;
      a_s[0] += w * sigm(r0); a_s[1] += w * sigm(r1); a_s[2] += w * sigm(r2);
      a_w += w;
      const float4 n = ld4(P.normal2 + 4 * (size_t)s);
      if (P.render_normal) { a_n[0] += w * n.x; a_n[1] += w * n.y; a_n[2] += w * n.z; }
      if (P.render_depth) a_d += w * (float)P.step_id1[idx1] * P.dist;
      if (LOSS) {
        const float ndv = fminf(0.f, -n.w);
        a_o += w * ndv * ndv;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a_rgb[k] = warp_sum(a_rgb[k]);
      a_s[k] = warp_sum(a_s[k]);
      if (P.render_normal) a_n[k] = warp_sum(a_n[k]);
    }
    a_w = warp_sum(a_w);
    if (P.render_depth) a_d = warp_sum(a_d);
    float pre_r[3], pre_s[3], out_r[3], out_s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pre_r[k] = a_rgb[k] + (1.f - a_w) * P.bg;
      pre_s[k] = a_s[k] + (1.f - a_w) * P.bg;
      out_r[k] = fminf(fmaxf(pre_r[k], 0.f), 1.f);
      out_s[k] = fminf(fmaxf(pre_s[k], 0.f), 1.f);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        P.rgb_marched[3 * (size_t)ray + k] = out_r[k];
        P.sigmoid_rgb[3 * (size_t)ray + k] = out_s[k];
        if (P.render_normal) P.normal_marched[3 * (size_t)ray + k] = a_n[k];
      }
      P.cum_weights[ray] = a_w;
      if (P.render_depth) P.depth[ray] = a_d;
    }
    if (!LOSS) continue;
    // ---- loss block (model/nerf_training.py:307-327) ----
    const float inv_cnt = 1.f / (3.f * (float)L.n_rays_global);
    float g_r[3], g_s[3], gsum = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float t = L.target[3 * (size_t)ray + k];
      const float er = out_r[k] - t, es = out_s[k] - t;
      if (lane == 0) { l_mse += (double)er * er; l_srgb += (double)es * es; }
      g_r[k] = (pre_r[k] >= 0.f && pre_r[k] <= 1.f) ? L.weight_main * 2.f * er * inv_cnt : 0.f;
      g_s[k] = (pre_s[k] >= 0.f && pre_s[k] <= 1.f) ? L.sigmoid_rgb_loss * 2.f * es * inv_cnt : 0.f;
      gsum += g_r[k] + g_s[k];
    }
    {
      const float o_all = warp_sum(a_o);
      if (lane == 0) l_ori += (double)o_all;
    }
    for (int i = lane; i < c; i += 32) {
      const int s = b + i;
      const int idx1 = P.idx1_of2[s];
      const float w = P.w1[idx1];
      const float4 z = ld4(P.logit + 4 * (size_t)s);
      const float r[3] = {sigm(z.x), sigm(z.y), sigm(z.z)};
      float dz[3], dw = -P.bg * gsum;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float sr = sigm(r[k]);
        const float d_rgb = w * g_r[k] + w * g_s[k] * sr * (1.f - sr);
        float d_rgb_k = d_rgb;
        if (L.weight_rgbper > 0.f) {
          // weight_rgbper * sum_s w.detach() * sum_c (rgb - target[ray])^2 / N   (model/nerf_training.py:311-314)
          const float e = r[k] - L.target[3 * (size_t)ray + k];
          d_rgb_k += L.weight_rgbper * w * 2.f * e / (float)L.n_rays_global;
          l_rgbper += (double)w * (double)e * (double)e;
        }
        dz[k] = d_rgb_k * r[k] * (1.f - r[k]);
        dw += r[k] * g_r[k] + sr * g_s[k];
      }
      st4(P.dlogit + 4 * (size_t)s, make_float4(dz[0], dz[1], dz[2], 0.f));
      if (P.dlogit_p) {
        // the same row as bf16 hi / mid planes [cap_m2][8]: the operand format of the backward GEMMs (3xBF16)
        uint32_t h01, m01, h2, m2;
        bf16_split2(dz[0], dz[1], h01, m01);
        bf16_split2(dz[2], 0.f, h2, m2);
        reinterpret_cast<uint4*>(P.dlogit_p)[s] = make_uint4(h01, h2, 0u, 0u);
        reinterpret_cast<uint4*>(P.dlogit_p + (size_t)P.cap_m2 * 8)[s] = make_uint4(m01, m2, 0u, 0u);
      }
      gb7[0] += dz[0]; gb7[1] += dz[1]; gb7[2] += dz[2];
      P.d_w1[idx1] = dw;
    }
    if (lane == 0) {
      // entropy_last_loss on alphainv_cum[..., -1] = the LAST ray only (nerf_training.py:316-319)
      float g_last = 0.f;
      if (L.has_last_ray && ray == P.n_rays - 1 && L.weight_entropy_last > 0.f) {
        const float a = P.alphainv_last[ray];
        const float p = fminf(fmaxf(a, 1e-6f), 1.f - 1e-6f);
        l_ent += -((double)p * log((double)p) + (1. - (double)p) * log(1. - (double)p));
        if (a >= 1e-6f && a <= 1.f - 1e-6f) g_last = L.weight_entropy_last * (logf(1.f - p) - logf(p));
      }
      L.g_alphainv_last[ray] = g_last;
    }
  }
  // loss terms and the last layer's bias gradient: one atomic per CTA and term (8192 warps adding into the same
  // handful of addresses serialise in L2: ~50 K same-address atomics per launch were most of this kernel's time)
  if (LOSS) {
    __shared__ double s_l[kWarps][5];
    __shared__ float s_g[kWarps][3];
    const int w = threadIdx.x >> 5;
    const double t_per = (L.weight_rgbper > 0.f) ? warp_sum(l_rgbper) : 0.;     // accumulated by every lane
#pragma unroll
    for (int k = 0; k < 3; ++k) gb7[k] = warp_sum(gb7[k]);
    if (lane == 0) {
      s_l[w][0] = l_mse; s_l[w][1] = l_ent; s_l[w][2] = l_ori; s_l[w][3] = l_srgb; s_l[w][4] = t_per;
      s_g[w][0] = gb7[0]; s_g[w][1] = gb7[1]; s_g[w][2] = gb7[2];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      const int k = threadIdx.x;
      if (k < 5) {
        double t = 0.;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) t += s_l[i][k];
        const double inv_cnt = 1. / (3. * (double)L.n_rays_global);
        if (t != 0.) {
          if (k == 0) atomicAdd(L.loss_out + 0, (double)L.weight_main * t * inv_cnt);
          else if (k == 1) atomicAdd(L.loss_out + 1, (double)L.weight_entropy_last * t);
          else if (k == 2) atomicAdd(L.loss_out + 2, (double)L.weight_orientation * t);
          else if (k == 3) atomicAdd(L.loss_out + 3, (double)L.sigmoid_rgb_loss * t * inv_cnt);
          else atomicAdd(L.loss_out + max(L.rgbper_slot, 0), (double)L.weight_rgbper * t / (double)L.n_rays_global);
        }
      } else if (P.gB[FGS_MLP_LAYERS - 1]) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) t += s_g[i][k - 5];
        if (t != 0.f) atomicAdd(P.gB[FGS_MLP_LAYERS - 1] + (k - 5), t);
      }
    }
  }
}

// Backward of level 1 in two kernels (as the forward: the recurrence is per ray, the expensive part per sample):
//   weights1_bwd_kernel  warp per ray: alpha2weight backward (render_utils_kernel.cu:654-677) as a reverse
//                        warp-shuffle recurrence; leaves d_alpha of every visited sample in d_w1
//   sample1_bwd_kernel   thread per level-1 sample: NeuS alpha backward (model/nerf.py:510-544) and the 1 + 6
//                        trilinear scatters into the SDF gradient grid
__global__ void __launch_bounds__(kBlock)
weights1_bwd_kernel(const __grid_constant__ FgsFineParams P, const float* __restrict__ g_last) {
  const int lane = lane_id();
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < P.n_rays; ray += gridDim.x * kWarps) {
    const int base = P.off1[ray];
    const int c1 = P.cnt1[ray];
    if (c1 == 0 || base + c1 > P.cap_m1) continue;
    float back = __fmul_rn(g_last ? g_last[ray] : 0.f, P.alphainv_last[ray]);
    for (int top = c1; top > 0; top -= 32) {
      const int i = top - 1 - lane;
      const bool valid = i >= 0;
      const int idx = base + (valid ? i : 0);
      const int fl = valid ? P.flag1[idx] : 0;
      const bool part = fl & FGS_F_VISIT;
      const bool l2 = fl & FGS_F_L2;
      const float gw = l2 ? P.d_w1[idx] : 0.f;
      const float w = part ? P.w1[idx] : 0.f;
      const int cnt = min(32, top);
      float myBack = 0.f;
      if (__any_sync(0xffffffffu, part)) {
        for (int j = 0; j < cnt; ++j) {
          if (j == lane) myBack = back;
          const float gwj = __shfl_sync(0xffffffffu, gw, j);
          const float wj = __shfl_sync(0xffffffffu, w, j);
          back = __fmaf_rn(gwj, wj, back);
        }
      }
      if (!part) continue;
      const float a = P.alpha1[idx];
      const double den = __dadd_rn((double)__fsub_rn(1.f, a), 1e-10);
      P.d_w1[idx] = (float)__dsub_rn((double)__fmul_rn(gw, P.T1[idx]), __ddiv_rn((double)myBack, den));   // d_alpha
    }
  }
}

__global__ void __launch_bounds__(kBlock) sample1_bwd_kernel(const __grid_constant__ FgsFineParams P) {
  const GridGeo g = make_geo(P.xyz_min, P.xyz_max, P.X, P.Y, P.Z);
  const int M1 = min(P.counters[FGS_CNT_M1], P.cap_m1);
  for (int idx = blockIdx.x * kBlock + threadIdx.x; idx < M1; idx += gridDim.x * kBlock) {
    const int ray = P.ray_id1[idx];
    if (ray < 0) continue;
    const int fl = P.flag1[idx];
    if (!(fl & FGS_F_VISIT)) continue;
    const bool l2 = fl & FGS_F_L2;
    const float vx = P.viewdirs[3 * (size_t)ray], vy = P.viewdirs[3 * (size_t)ray + 1],
                vz = P.viewdirs[3 * (size_t)ray + 2];
    const float d_alpha = P.d_w1[idx];
    // ---- NeuS alpha backward (model/nerf.py:510-544) ----
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
    if (l2) {
      d_sdf += P.d_sdf1[idx];
      d_g[0] += P.d_grad1[idx];
      d_g[1] += P.d_grad1[(size_t)P.cap_m1 + idx];
      d_g[2] += P.d_grad1[2 * (size_t)P.cap_m1 + idx];
    }
    // ---- scatter: direct sample + 6 finite-difference taps (model/nerf.py:639-672) ----
    const Ray r = make_ray(P.rays_o + 3 * (size_t)ray, P.rays_d + 3 * (size_t)ray, P.t_min[ray]);
    float px, py, pz, iw, ih, id;
    ray_point(r, P.stepdist, P.step_id1[idx], px, py, pz);
    cont_index(g, px, py, pz, iw, ih, id);
    if (d_sdf != 0.f) tri_scatter1_flag(P.g_sdf, P.sdf_brick_flags, tri_setup(iw, ih, id), g.X, g.Y, g.Z, d_sdf);
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      const float dgr = d_g[2 - ax];  // axis 0 = z, 1 = y, 2 = x
      if (dgr == 0.f) continue;
      const TapPair tp = make_taps(g, iw, ih, id, ax, 1.0f);
      const float df = dgr / tp.diff / P.voxel_size;
      tri_scatter1_flag(P.g_sdf, P.sdf_brick_flags, tp.lo, g.X, g.Y, g.Z, -df);
      tri_scatter1_flag(P.g_sdf, P.sdf_brick_flags, tp.hi, g.X, g.Y, g.Z, df);
    }
  }
}

}  // namespace

namespace fine {

int launch_composite(const FgsFineParams& P, const FgsFineLoss* L, cudaStream_t st) {
  const int grid = fgs::stream_grid(P.n_rays, kWarps, 8);
  if (L) {
    composite_kernel<true><<<grid, kBlock, 0, st>>>(P, *L);
  } else {
    FgsFineLoss z = {};
    composite_kernel<false><<<grid, kBlock, 0, st>>>(P, z);
  }
  return fgs::check_launch("fine composite", st);
}

// alpha2weight backward alone (shared with the coarse stage)
int launch_weights1_bwd(const FgsFineParams& P, const float* g_last, cudaStream_t st) {
  const int grid = fgs::stream_grid(P.n_rays, kWarps, 8);
  weights1_bwd_kernel<<<grid, kBlock, 0, st>>>(P, g_last);
  return fgs::check_launch("fine weights1_bwd", st);
}

int launch_level1_bwd(const FgsFineParams& P, const float* g_last, cudaStream_t st) {
  const int grid = fgs::stream_grid(P.n_rays, kWarps, 8);
  weights1_bwd_kernel<<<grid, kBlock, 0, st>>>(P, g_last);
  int rc = fgs::check_launch("fine weights1_bwd", st);
  if (rc) return rc;
  sample1_bwd_kernel<<<fgs::stream_grid(P.cap_m1, kBlock, 8), kBlock, 0, st>>>(P);
  return fgs::check_launch("fine level1_bwd", st);
}

}  // namespace fine
