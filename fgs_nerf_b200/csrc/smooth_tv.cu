// smooth_grad_tv regulariser of the fine stage, differentiated analytically (sm_100a).
//
// Reference (every tv_every-th step): nerf.density_total_variation(sdf_tv=0, smooth_grad_tv=w)
// (model/nerf.py:430-447) builds the full central-difference gradient grid G = neus_sdf_gradient()
// (:485-494, a fresh [1,3,X,Y,Z] tensor *every forward*, :856), runs a 3x3x3 binomial Conv3d with
// replicate padding over it (:226-258, :439), takes mean((conv(G).detach() - G)^2 [nonempty]) and
// lets autograd walk back through the slicing ops into sdf.grid.grad.
// Here: three streaming stencil kernels, no autograd graph, computed only on TV steps:
//   G  = central differences (interior only, faces 0)                        [3][V]
//   E  = -2 w / n_sel * (conv3(G) - G) * nonempty   (= dLoss/dG), loss sum   [3][V]
//   g_sdf[i] += (E_x[i-1] - E_x[i+1]) / (2 voxel)  (+ y, z)  with E taken only at interior rows
#include "common.cuh"
#include "../../include/fgs_b200_fine.h"

namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock)
tv_grad_grid_kernel(const float* __restrict__ sdf, float* __restrict__ G, int X, int Y, int Z,
                    float voxel_size) {
  const long long V = (long long)X * Y * Z;
  const long long sj = Z, si = (long long)Z * Y;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < V;
       idx += (long long)gridDim.x * kBlock) {
    long long r = idx;
    const int k = (int)(r % Z); r /= Z;
    const int j = (int)(r % Y); r /= Y;
    const int i = (int)r;
    // (a - b) / 2 / voxel_size   (model/nerf.py:492-494)
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (i >= 1 && i <= X - 2) gx = __fdiv_rn(__fmul_rn(__fsub_rn(sdf[idx + si], sdf[idx - si]), 0.5f), voxel_size);
    if (j >= 1 && j <= Y - 2) gy = __fdiv_rn(__fmul_rn(__fsub_rn(sdf[idx + sj], sdf[idx - sj]), 0.5f), voxel_size);
    if (k >= 1 && k <= Z - 2) gz = __fdiv_rn(__fmul_rn(__fsub_rn(sdf[idx + 1], sdf[idx - 1]), 0.5f), voxel_size);
    G[idx] = gx;
    G[V + idx] = gy;
    G[2 * V + idx] = gz;
  }
}

__global__ void __launch_bounds__(kBlock)
tv_err_kernel(const float* __restrict__ G, const unsigned char* __restrict__ nonempty,
              float* __restrict__ E, int X, int Y, int Z, float scale, double* __restrict__ loss_out,
              double loss_scale) {
  const long long V = (long long)X * Y * Z;
  double lsum = 0.;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < 3 * V;
       idx += (long long)gridDim.x * kBlock) {
    const long long vox = idx % V;
    const float* __restrict__ Gc = G + (idx - vox);
    float e = 0.f;
    if (!nonempty || nonempty[vox]) {
      long long r = vox;
      const int k = (int)(r % Z); r /= Z;
      const int j = (int)(r % Y); r /= Y;
      const int i = (int)r;
      // binomial 3x3x3 kernel (1,2,1)^3 / 64 with replicate padding
      float conv = 0.f;
#pragma unroll
      for (int di = -1; di <= 1; ++di) {
        const int ii = min(max(i + di, 0), X - 1);
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
          const int jj = min(max(j + dj, 0), Y - 1);
#pragma unroll
          for (int dk = -1; dk <= 1; ++dk) {
            const int kk = min(max(k + dk, 0), Z - 1);
            const float w = (float)((2 - (di != 0)) * (2 - (dj != 0)) * (2 - (dk != 0))) * (1.f / 64.f);
            conv = fmaf(w, Gc[((long long)ii * Y + jj) * Z + kk], conv);
          }
        }
      }
      const float err = conv - Gc[vox];
      lsum += (double)err * err;
      e = scale * err;  // dLoss/dG = -2 w / n_sel * err
    }
    E[idx] = e;
  }
  if (loss_out) {
    // block reduction -> one atomic per CTA
    __shared__ double s[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane_id() == 0) s[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.;
      for (int w = 0; w < kBlock / 32; ++w) t += s[w];
      if (t != 0.) atomicAdd(loss_out, t * loss_scale);
    }
  }
}

__global__ void __launch_bounds__(kBlock)
tv_scatter_kernel(const float* __restrict__ E, float* __restrict__ g_sdf, int X, int Y, int Z,
                  float voxel_size) {
  const long long V = (long long)X * Y * Z;
  const long long sj = Z, si = (long long)Z * Y;
  const float inv = 0.5f / voxel_size;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < V;
       idx += (long long)gridDim.x * kBlock) {
    long long r = idx;
    const int k = (int)(r % Z); r /= Z;
    const int j = (int)(r % Y); r /= Y;
    const int i = (int)r;
    float acc = 0.f;
    // G_x[i'] = (s[i'+1] - s[i'-1]) / 2 / vs for 1 <= i' <= X-2: s[i] appears in G_x[i-1] (+) and G_x[i+1] (-)
    if (i - 1 >= 1 && i - 1 <= X - 2) acc += E[idx - si];
    if (i + 1 >= 1 && i + 1 <= X - 2) acc -= E[idx + si];
    if (j - 1 >= 1 && j - 1 <= Y - 2) acc += E[V + idx - sj];
    if (j + 1 >= 1 && j + 1 <= Y - 2) acc -= E[V + idx + sj];
    if (k - 1 >= 1 && k - 1 <= Z - 2) acc += E[2 * V + idx - 1];
    if (k + 1 >= 1 && k + 1 <= Z - 2) acc -= E[2 * V + idx + 1];
    if (acc != 0.f) g_sdf[idx] += acc * inv;
  }
}

// ---- 128-bit versions for Z % 4 == 0 (the scalar kernels above issue 6 / 27 / 6 dependent 32-bit loads per
// voxel and ran at 16-43 % of the HBM roofline; here a thread owns 4 consecutive z voxels: the stencil rows are
// float4 loads plus two halo scalars) ----
__device__ __forceinline__ float4 ld4g(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4g(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float cdiff(float a, float b, float voxel_size) {
  return __fdiv_rn(__fmul_rn(__fsub_rn(a, b), 0.5f), voxel_size);   // (a - b) / 2 / voxel_size (model/nerf.py:492-494)
}

__global__ void __launch_bounds__(kBlock)
tv_grad_grid_v4_kernel(const float* __restrict__ sdf, float* __restrict__ G, int X, int Y, int Z, float voxel_size) {
  const long long V = (long long)X * Y * Z, V4 = V >> 2;
  const long long sj = Z, si = (long long)Z * Y;
  const int Z4 = Z >> 2;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < V4; q += (long long)gridDim.x * kBlock) {
    long long r = q;
    const int k0 = (int)(r % Z4) * 4; r /= Z4;
    const int j = (int)(r % Y); r /= Y;
    const int i = (int)r;
    const long long idx = q << 2;
    float4 gx = z4, gy = z4, gz;
    if (i >= 1 && i <= X - 2) {
      const float4 a = ld4g(sdf + idx + si), b = ld4g(sdf + idx - si);
      gx = make_float4(cdiff(a.x, b.x, voxel_size), cdiff(a.y, b.y, voxel_size), cdiff(a.z, b.z, voxel_size), cdiff(a.w, b.w, voxel_size));
    }
    if (j >= 1 && j <= Y - 2) {
      const float4 a = ld4g(sdf + idx + sj), b = ld4g(sdf + idx - sj);
      gy = make_float4(cdiff(a.x, b.x, voxel_size), cdiff(a.y, b.y, voxel_size), cdiff(a.z, b.z, voxel_size), cdiff(a.w, b.w, voxel_size));
    }
    {
      const float4 c = ld4g(sdf + idx);
      const float left = k0 > 0 ? sdf[idx - 1] : 0.f, right = k0 + 4 < Z ? sdf[idx + 4] : 0.f;
      gz.x = k0 >= 1 ? cdiff(c.y, left, voxel_size) : 0.f;              // k = k0 (interior iff k0 >= 1; Z >= 4)
      gz.y = cdiff(c.z, c.x, voxel_size);                               // k0 + 1 is always interior
      gz.z = cdiff(c.w, c.y, voxel_size);                               // k0 + 2 too
      gz.w = k0 + 4 < Z ? cdiff(right, c.z, voxel_size) : 0.f;          // k0 + 3 interior iff not the last voxel
    }
    st4g(G + idx, gx);
    st4g(G + V + idx, gy);
    st4g(G + 2 * V + idx, gz);
  }
}

__global__ void __launch_bounds__(kBlock)
tv_err_v4_kernel(const float* __restrict__ G, const unsigned char* __restrict__ nonempty, float* __restrict__ E,
                 int X, int Y, int Z, float scale, double* __restrict__ loss_out, double loss_scale) {
  const long long V = (long long)X * Y * Z, V4 = V >> 2;
  const int Z4 = Z >> 2;
  double lsum = 0.;
  for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < 3 * V4; q += (long long)gridDim.x * kBlock) {
    const long long vq = q % V4;
    const long long vox = vq << 2;
    const float* __restrict__ Gc = G + ((q - vq) << 2);
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    uchar4 ne = make_uchar4(1, 1, 1, 1);
    if (nonempty) ne = *reinterpret_cast<const uchar4*>(nonempty + vox);
    if (ne.x | ne.y | ne.z | ne.w) {
      long long r = vq;
      const int k0 = (int)(r % Z4) * 4; r /= Z4;
      const int j = (int)(r % Y); r /= Y;
      const int i = (int)r;
      // binomial (1,2,1)^3 / 64 with replicate padding: 9 rows, each convolved along z from one float4 + 2 halos
      float4 conv = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 centre = conv;
#pragma unroll
      for (int di = -1; di <= 1; ++di) {
        const int ii = min(max(i + di, 0), X - 1);
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
          const int jj = min(max(j + dj, 0), Y - 1);
          const float* __restrict__ row = Gc + ((long long)ii * Y + jj) * Z + k0;
          const float4 v = ld4g(row);
          const float left = k0 > 0 ? row[-1] : v.x, right = k0 + 4 < Z ? row[4] : v.w;
          const float w = (float)((2 - (di != 0)) * (2 - (dj != 0))) * (1.f / 64.f);
          conv.x = fmaf(w, (left + v.y) + 2.f * v.x, conv.x);
          conv.y = fmaf(w, (v.x + v.z) + 2.f * v.y, conv.y);
          conv.z = fmaf(w, (v.y + v.w) + 2.f * v.z, conv.z);
          conv.w = fmaf(w, (v.z + right) + 2.f * v.w, conv.w);
          if (di == 0 && dj == 0) centre = v;
        }
      }
      const float ex = ne.x ? conv.x - centre.x : 0.f, ey = ne.y ? conv.y - centre.y : 0.f;
      const float ez = ne.z ? conv.z - centre.z : 0.f, ew = ne.w ? conv.w - centre.w : 0.f;
      lsum += ((double)ex * ex + (double)ey * ey) + ((double)ez * ez + (double)ew * ew);
      e = make_float4(scale * ex, scale * ey, scale * ez, scale * ew);   // dLoss/dG = -2 w / n_sel * err
    }
    st4g(E + (q << 2), e);
  }
  if (loss_out) {
    __shared__ double s[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane_id() == 0) s[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.;
      for (int w = 0; w < kBlock / 32; ++w) t += s[w];
      if (t != 0.) atomicAdd(loss_out, t * loss_scale);
    }
  }
}

__global__ void __launch_bounds__(kBlock)
tv_scatter_v4_kernel(const float* __restrict__ E, float* __restrict__ g_sdf, int X, int Y, int Z, float voxel_size) {
  const long long V = (long long)X * Y * Z, V4 = V >> 2;
  const long long sj = Z, si = (long long)Z * Y;
  const int Z4 = Z >> 2;
  const float inv = 0.5f / voxel_size;
  for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < V4; q += (long long)gridDim.x * kBlock) {
    long long r = q;
    const int k0 = (int)(r % Z4) * 4; r /= Z4;
    const int j = (int)(r % Y); r /= Y;
    const int i = (int)r;
    const long long idx = q << 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // G_x[i'] = (s[i'+1] - s[i'-1]) / 2 / vs for 1 <= i' <= X-2: s[i] appears in G_x[i-1] (+) and G_x[i+1] (-)
    if (i - 1 >= 1 && i - 1 <= X - 2) { const float4 v = ld4g(E + idx - si); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    if (i + 1 >= 1 && i + 1 <= X - 2) { const float4 v = ld4g(E + idx + si); acc.x -= v.x; acc.y -= v.y; acc.z -= v.z; acc.w -= v.w; }
    if (j - 1 >= 1 && j - 1 <= Y - 2) { const float4 v = ld4g(E + V + idx - sj); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    if (j + 1 >= 1 && j + 1 <= Y - 2) { const float4 v = ld4g(E + V + idx + sj); acc.x -= v.x; acc.y -= v.y; acc.z -= v.z; acc.w -= v.w; }
    {
      const float* __restrict__ Ez = E + 2 * V + idx;
      const float4 v = ld4g(Ez);
      const float left = k0 > 0 ? Ez[-1] : 0.f, right = k0 + 4 < Z ? Ez[4] : 0.f;
      // voxel k takes +E_z[k-1] if 1 <= k-1 <= Z-2 and -E_z[k+1] if 1 <= k+1 <= Z-2
      const int k1 = k0 + 1, k2 = k0 + 2, k3 = k0 + 3;
      if (k0 - 1 >= 1) acc.x += left;
      if (k0 + 1 <= Z - 2) acc.x -= v.y;
      if (k1 - 1 >= 1) acc.y += v.x;
      if (k1 + 1 <= Z - 2) acc.y -= v.z;
      if (k2 - 1 >= 1 && k2 - 1 <= Z - 2) acc.z += v.y;
      if (k2 + 1 <= Z - 2) acc.z -= v.w;
      if (k3 - 1 <= Z - 2) acc.w += v.z;
      if (k3 + 1 <= Z - 2) acc.w -= right;
    }
    if (acc.x != 0.f || acc.y != 0.f || acc.z != 0.f || acc.w != 0.f) {
      float4 g = ld4g(g_sdf + idx);
      if (acc.x != 0.f) g.x += acc.x * inv;
      if (acc.y != 0.f) g.y += acc.y * inv;
      if (acc.z != 0.f) g.z += acc.z * inv;
      if (acc.w != 0.f) g.w += acc.w * inv;
      st4g(g_sdf + idx, g);
    }
  }
}

// nerf.density_total_variation(sdf_tv) of the coarse stage (ori_tv: through autograd in the reference; here its
// closed form): per voxel the signs of the up-to-6 masked differences it takes part in.  total_variation,
// model/nerf.py:1212-1221: |v.diff(axis)| over pairs with both voxels in the mask.
__global__ void __launch_bounds__(kBlock)
l1_tv_kernel(const float* __restrict__ sdf, const unsigned char* __restrict__ mask, float* __restrict__ g_sdf, int X, int Y,
             int Z, float scale, double* __restrict__ loss_out, double loss_scale) {
  const long long V = (long long)X * Y * Z;
  const long long sj = Z, si = (long long)Z * Y;
  double lsum = 0.;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < V; idx += (long long)gridDim.x * kBlock) {
    if (!mask[idx]) continue;
    long long r = idx;
    const int k = (int)(r % Z); r /= Z;
    const int j = (int)(r % Y); r /= Y;
    const int i = (int)r;
    const float c = sdf[idx];
    float acc = 0.f;
    auto fwd = [&](long long o) {          // pair (idx, idx + o): d = sdf[idx + o] - c, this voxel gets -sign(d)
      if (mask[idx + o]) {
        const float d = sdf[idx + o] - c;
        lsum += (double)fabsf(d);
        acc -= (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
      }
    };
    auto bwd = [&](long long o) {          // pair (idx - o, idx): d = c - sdf[idx - o], this voxel gets +sign(d)
      if (mask[idx - o]) {
        const float d = c - sdf[idx - o];
        acc += (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
      }
    };
    if (i + 1 < X) fwd(si);
    if (i > 0) bwd(si);
    if (j + 1 < Y) fwd(sj);
    if (j > 0) bwd(sj);
    if (k + 1 < Z) fwd(1);
    if (k > 0) bwd(1);
    if (acc != 0.f) g_sdf[idx] += scale * acc;
  }
  if (loss_out) {
    __shared__ double s[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane_id() == 0) s[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.;
      for (int w = 0; w < kBlock / 32; ++w) t += s[w];
      if (t != 0.) atomicAdd(loss_out, t * loss_scale);
    }
  }
}

// The unmasked branch of total_variation (model/nerf.py:1212-1221: `... / 3 / v.sum()`): pass 1 = S and sum(v)
__global__ void __launch_bounds__(kBlock)
l1_tv_sums_kernel(const float* __restrict__ sdf, int X, int Y, int Z, double* __restrict__ out2) {
  const long long V = (long long)X * Y * Z;
  const long long sj = Z, si = (long long)Z * Y;
  double s_abs = 0., s_val = 0.;
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < V; idx += (long long)gridDim.x * kBlock) {
    long long r = idx;
    const int k = (int)(r % Z); r /= Z;
    const int j = (int)(r % Y); r /= Y;
    const int i = (int)r;
    const float c = sdf[idx];
    s_val += (double)c;
    if (i + 1 < X) s_abs += (double)fabsf(sdf[idx + si] - c);
    if (j + 1 < Y) s_abs += (double)fabsf(sdf[idx + sj] - c);
    if (k + 1 < Z) s_abs += (double)fabsf(sdf[idx + 1] - c);
  }
  __shared__ double s[2][kBlock / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_abs += __shfl_xor_sync(0xffffffffu, s_abs, o);
    s_val += __shfl_xor_sync(0xffffffffu, s_val, o);
  }
  if (lane_id() == 0) { s[0][threadIdx.x >> 5] = s_abs; s[1][threadIdx.x >> 5] = s_val; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.;
    for (int w = 0; w < kBlock / 32; ++w) t += s[threadIdx.x][w];
    atomicAdd(out2 + threadIdx.x, t);
  }
}

// pass 2: d/dv [ w S / (3 sum) ] = w / (3 sum) * dS/dv - w S / (3 sum^2)
__global__ void __launch_bounds__(kBlock)
l1_tv_unmasked_grad_kernel(const float* __restrict__ sdf, float* __restrict__ g_sdf, int X, int Y, int Z, double weight,
                           const double* __restrict__ sums, double* __restrict__ loss_out) {
  const long long V = (long long)X * Y * Z;
  const long long sj = Z, si = (long long)Z * Y;
  const double S = sums[0], T = sums[1];
  const float scale = (float)(weight / (3.0 * T));
  const float shift = (float)(-weight * S / (3.0 * T * T));
  if (loss_out && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(loss_out, weight * S / (3.0 * T));
  for (long long idx = blockIdx.x * (long long)kBlock + threadIdx.x; idx < V; idx += (long long)gridDim.x * kBlock) {
    long long r = idx;
    const int k = (int)(r % Z); r /= Z;
    const int j = (int)(r % Y); r /= Y;
    const int i = (int)r;
    const float c = sdf[idx];
    float acc = 0.f;
    auto sgn = [](float d) { return (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f); };
    if (i + 1 < X) acc -= sgn(sdf[idx + si] - c);
    if (i > 0) acc += sgn(c - sdf[idx - si]);
    if (j + 1 < Y) acc -= sgn(sdf[idx + sj] - c);
    if (j > 0) acc += sgn(c - sdf[idx - sj]);
    if (k + 1 < Z) acc -= sgn(sdf[idx + 1] - c);
    if (k > 0) acc += sgn(c - sdf[idx - 1]);
    g_sdf[idx] += scale * acc + shift;
  }
}

}  // namespace

extern "C" int fgs_smooth_grad_tv_add_grad(const float* sdf, float* g_sdf, const uint8_t* nonempty,
                                           float* scratch, int X, int Y, int Z, float voxel_size,
                                           float weight, double n_sel, double* loss_out,
                                           fgs_stream_t stream) {
  FGS_REQUIRE(X >= 3 && Y >= 3 && Z >= 3, "grid must be at least 3^3");
  FGS_REQUIRE(n_sel > 0, "n_sel must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  const long long V = (long long)X * Y * Z;
  float* G = scratch;
  float* E = scratch + 3 * V;
  if ((Z & 3) == 0 && Z >= 4 && ((((uintptr_t)sdf | (uintptr_t)g_sdf | (uintptr_t)scratch | (uintptr_t)nonempty) & 15) == 0)) {
    const int grid4 = fgs::stream_grid(V >> 2, kBlock, 8);
    FGS_PREFER_MAX_SHARED(tv_grad_grid_v4_kernel);
    FGS_PREFER_MAX_SHARED(tv_err_v4_kernel);
    FGS_PREFER_MAX_SHARED(tv_scatter_v4_kernel);
    tv_grad_grid_v4_kernel<<<grid4, kBlock, 0, st>>>(sdf, G, X, Y, Z, voxel_size);
    int rc4 = fgs::check_launch("smooth_tv grad grid", st);
    if (rc4) return rc4;
    tv_err_v4_kernel<<<fgs::stream_grid(3 * (V >> 2), kBlock, 8), kBlock, 0, st>>>(
        G, nonempty, E, X, Y, Z, (float)(-2.0 * (double)weight / n_sel), loss_out, (double)weight / n_sel);
    if ((rc4 = fgs::check_launch("smooth_tv err", st))) return rc4;
    tv_scatter_v4_kernel<<<grid4, kBlock, 0, st>>>(E, g_sdf, X, Y, Z, voxel_size);
    return fgs::check_launch("smooth_tv scatter", st);
  }
  const int grid = fgs::stream_grid(V, kBlock, 8);
  FGS_PREFER_MAX_SHARED(tv_grad_grid_kernel);
  FGS_PREFER_MAX_SHARED(tv_err_kernel);
  FGS_PREFER_MAX_SHARED(tv_scatter_kernel);
  tv_grad_grid_kernel<<<grid, kBlock, 0, st>>>(sdf, G, X, Y, Z, voxel_size);
  int rc = fgs::check_launch("smooth_tv grad grid", st);
  if (rc) return rc;
  tv_err_kernel<<<fgs::stream_grid(3 * V, kBlock, 8), kBlock, 0, st>>>(
      G, nonempty, E, X, Y, Z, (float)(-2.0 * (double)weight / n_sel), loss_out, (double)weight / n_sel);
  if ((rc = fgs::check_launch("smooth_tv err", st))) return rc;
  tv_scatter_kernel<<<grid, kBlock, 0, st>>>(E, g_sdf, X, Y, Z, voxel_size);
  return fgs::check_launch("smooth_tv scatter", st);
}

extern "C" int fgs_sdf_gradient_grid(const float* sdf, float* G, int X, int Y, int Z, float voxel_size, fgs_stream_t stream) {
  FGS_REQUIRE(sdf && G && X >= 3 && Y >= 3 && Z >= 3, "sdf_gradient_grid: grid must be at least 3^3");
  cudaStream_t st = (cudaStream_t)stream;
  const long long V = (long long)X * Y * Z;
  if ((Z & 3) == 0 && ((((uintptr_t)sdf | (uintptr_t)G) & 15) == 0) && (V & 3) == 0)
    tv_grad_grid_v4_kernel<<<fgs::stream_grid(V >> 2, kBlock, 8), kBlock, 0, st>>>(sdf, G, X, Y, Z, voxel_size);
  else
    tv_grad_grid_kernel<<<fgs::stream_grid(V, kBlock, 8), kBlock, 0, st>>>(sdf, G, X, Y, Z, voxel_size);
  return fgs::check_launch("sdf_gradient_grid", st);
}

extern "C" int fgs_sdf_gradient_grid_bwd(const float* g_G, float* g_sdf, int X, int Y, int Z, float voxel_size,
                                         fgs_stream_t stream) {
  FGS_REQUIRE(g_G && g_sdf && X >= 3 && Y >= 3 && Z >= 3, "sdf_gradient_grid_bwd: grid must be at least 3^3");
  cudaStream_t st = (cudaStream_t)stream;
  const long long V = (long long)X * Y * Z;
  if ((Z & 3) == 0 && ((((uintptr_t)g_sdf | (uintptr_t)g_G) & 15) == 0) && (V & 3) == 0)
    tv_scatter_v4_kernel<<<fgs::stream_grid(V >> 2, kBlock, 8), kBlock, 0, st>>>(g_G, g_sdf, X, Y, Z, voxel_size);
  else
    tv_scatter_kernel<<<fgs::stream_grid(V, kBlock, 8), kBlock, 0, st>>>(g_G, g_sdf, X, Y, Z, voxel_size);
  return fgs::check_launch("sdf_gradient_grid_bwd", st);
}

extern "C" int fgs_l1_tv_add_grad(const float* sdf, const uint8_t* mask, float* g_sdf, int X, int Y, int Z, double weight,
                                  double mask_count, double* loss_out, fgs_stream_t stream) {
  FGS_REQUIRE(sdf && g_sdf && X >= 2 && Y >= 2 && Z >= 2, "l1_tv_add_grad: bad grid");
  FGS_REQUIRE(mask != nullptr && mask_count > 0, "l1_tv_add_grad: the non-empty mask (and its count) is required");
  cudaStream_t st = (cudaStream_t)stream;
  const long long V = (long long)X * Y * Z;
  const double s = weight / 3.0 / mask_count;
  l1_tv_kernel<<<fgs::stream_grid(V, kBlock, 8), kBlock, 0, st>>>(sdf, mask, g_sdf, X, Y, Z, (float)s, loss_out, s);
  return fgs::check_launch("l1_tv_add_grad", st);
}

extern "C" int fgs_l1_tv_unmasked_add_grad(const float* sdf, float* g_sdf, int X, int Y, int Z, double weight, double* scratch2,
                                           double* loss_out, fgs_stream_t stream) {
  FGS_REQUIRE(sdf && g_sdf && scratch2 && X >= 2 && Y >= 2 && Z >= 2, "l1_tv_unmasked_add_grad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long V = (long long)X * Y * Z;
  int rc = fgs::check_cuda(cudaMemsetAsync(scratch2, 0, 2 * sizeof(double), st), "l1_tv scratch memset");
  if (rc) return rc;
  const int grid = fgs::stream_grid(V, kBlock, 8);
  l1_tv_sums_kernel<<<grid, kBlock, 0, st>>>(sdf, X, Y, Z, scratch2);
  if ((rc = fgs::check_launch("l1_tv_add_grad", st))) return rc;
  l1_tv_unmasked_grad_kernel<<<grid, kBlock, 0, st>>>(sdf, g_sdf, X, Y, Z, weight, scratch2, loss_out);
  return fgs::check_launch("l1_tv_add_grad", st);
}
