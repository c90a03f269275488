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
