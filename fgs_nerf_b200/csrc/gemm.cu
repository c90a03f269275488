// fp32 GEMMs of the colour MLPs (rgbnet / refnet, model/nerf.py:125-151, 877-884) with fused
// bias / ReLU / ReLU-mask epilogues (sm_100a).
//
// The reference runs each nn.Linear as a cuBLAS SGEMM (TF32 off) + a separate bias/ReLU kernel,
// and autograd adds two more SGEMMs and a mask kernel per layer in backward.  This file is the
// true-fp32 SIMT path: a 128x128x8 register-tiled kernel (8x8 outputs per thread, register
// prefetch of the next k-slab) with the epilogue fused.  It is the parity reference for -- and
// the fallback-free small-shape path beside -- the tcgen05 tensor-core kernels in mlp_tc.cu.
//   mode NT: C[M,N]  = A[M,K] B[N,K]^T   forward      (activations x nn.Linear weight)
//   mode NN: C[M,N]  = A[M,K] B[K,N]     backward-data (dZ x W)
//   mode TN: C[M,N] += A[K,M]^T B[K,N]   backward-weight (dZ^T x X), split over K with atomics
// The sample-dimension extent (M for NT/NN, K for TN) is read on the device (rows_dev) so the
// launch needs no host synchronisation on the data-dependent sample count.
#include "common.cuh"
#include "../../include/fgs_b200_fine.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, kThreads = 256;

struct GemmArgs {
  const float* A; const float* B; float* C;
  const float* bias; const float* relu_mask; const int* rows_dev;
  int M, N, K, lda, ldb, ldc, ld_mask, relu;
};

// A_KMAJ: A is [M][K] (k contiguous) else [K][M].  B_KMAJ: B is [N][K] else [K][N].
template <bool A_KMAJ, bool B_KMAJ, bool SPLITK>
__global__ void __launch_bounds__(kThreads) gemm_kernel(const GemmArgs a) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  int M = a.M, K = a.K;
  if (a.rows_dev) {
    const int r = *a.rows_dev;
    if (SPLITK) K = min(K, r); else M = min(M, r);
  }
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M) return;
  int kbeg = 0, kend = K;
  if (SPLITK) {
    int chunk = (K + gridDim.z - 1) / gridDim.z;
    chunk = (chunk + BK - 1) / BK * BK;
    kbeg = blockIdx.z * chunk;
    kend = min(K, kbeg + chunk);
    if (kbeg >= kend) return;
  }
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tiles = [&](int k0) {
    if (A_KMAJ) {
      const int row = tid >> 1, kk = (tid & 1) * 4;
      const int m = m0 + row;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk + i;
        ra[i] = (m < M && k < kend) ? __ldg(a.A + (size_t)m * a.lda + k) : 0.f;
      }
    } else {
      const int kk = tid >> 5, mm = (tid & 31) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + mm + i;
        ra[i] = (m < M && k < kend) ? __ldg(a.A + (size_t)k * a.lda + m) : 0.f;
      }
    }
    if (B_KMAJ) {
      const int row = tid >> 1, kk = (tid & 1) * 4;
      const int n = n0 + row;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk + i;
        rb[i] = (n < a.N && k < kend) ? __ldg(a.B + (size_t)n * a.ldb + k) : 0.f;
      }
    } else {
      const int kk = tid >> 5, nn = (tid & 31) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + nn + i;
        rb[i] = (n < a.N && k < kend) ? __ldg(a.B + (size_t)k * a.ldb + n) : 0.f;
      }
    }
  };
  auto store_tiles = [&]() {
    if (A_KMAJ) {
      const int row = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) As[kk + i][row] = ra[i];
    } else {
      const int kk = tid >> 5, mm = (tid & 31) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) As[kk][mm + i] = ra[i];
    }
    if (B_KMAJ) {
      const int row = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[kk + i][row] = rb[i];
    } else {
      const int kk = tid >> 5, nn = (tid & 31) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[kk][nn + i] = rb[i];
    }
  };

  load_tiles(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    __syncthreads();
    store_tiles();
    __syncthreads();
    if (k0 + BK < kend) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (SPLITK) {
        atomicAdd(a.C + (size_t)m * a.ldc + n, v);
      } else {
        if (a.bias) v += a.bias[n];
        if (a.relu) v = fmaxf(v, 0.f);
        if (a.relu_mask && !(a.relu_mask[(size_t)m * a.ld_mask + n] > 0.f)) v = 0.f;
        a.C[(size_t)m * a.ldc + n] = v;
      }
    }
  }
}

// out[n] += sum over rows of X[r][n]  (bias gradients); rows read on the device.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int ld, int rows, int N, const int* __restrict__ rows_dev,
              float* __restrict__ out) {
  if (rows_dev) rows = min(rows, *rows_dev);
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rlane = threadIdx.x >> 5;  // 8 row lanes
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc = 0.f;
  if (n < N)
    for (int r = r0 + rlane; r < r1; r += 8) acc += X[(size_t)r * ld + n];
  __shared__ float s[8][33];
  s[rlane][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rlane == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][threadIdx.x & 31];
    if (t != 0.f) atomicAdd(out + n, t);
  }
}

}  // namespace

namespace fine {

int gemm(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
         int ldc, const float* bias, int relu, const float* relu_mask, int ld_mask,
         const int32_t* rows_dev, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  GemmArgs a{A, B, C, bias, relu_mask, rows_dev, M, N, K, lda, ldb, ldc, ld_mask, relu};
  dim3 grid(fgs::cdiv(N, BN), fgs::cdiv(M, BM), 1);
  switch (mode) {
    case 0: gemm_kernel<true, true, false><<<grid, kThreads, 0, st>>>(a); break;
    case 1: gemm_kernel<true, false, false><<<grid, kThreads, 0, st>>>(a); break;
    case 2: {
      const int tiles = grid.x * grid.y;
      int splits = (2 * FGS_NUM_SMS + tiles - 1) / tiles;
      const int max_splits = fgs::cdiv(K, 4 * BK);
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
      grid.z = splits;
      gemm_kernel<false, false, true><<<grid, kThreads, 0, st>>>(a);
      break;
    }
    default: return fgs::fail_arg("gemm mode must be 0,1,2");
  }
  return fgs::check_launch("gemm", st);
}

int colsum(const float* X, int ld, int rows, int N, const int32_t* rows_dev, float* out, cudaStream_t st) {
  if (rows <= 0 || N <= 0) return 0;
  dim3 grid(fgs::cdiv(N, 32), 64);
  colsum_kernel<<<grid, 256, 0, st>>>(X, ld, rows, N, rows_dev, out);
  return fgs::check_launch("colsum", st);
}

}  // namespace fine

extern "C" int fgs_gemm(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                        float* C, int ldc, const float* bias, int relu, const float* relu_mask,
                        int ld_mask, const int32_t* rows_dev, fgs_stream_t stream) {
  FGS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
  return fine::gemm(mode, M, N, K, A, lda, B, ldb, C, ldc, bias, relu, relu_mask, ld_mask, rows_dev,
                    (cudaStream_t)stream);
}
