// Internal interface of the tensor-core GEMM translation unit (gemm_tcw.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fgs_b200_fine.h"

namespace fine {

struct SplitJobHost {
  const float* W; float* hi; float* lo;
  int rows, cols, ld, ld_out, transpose;  // transpose: out[c][r] (ld_out >= rows) else out[r][c]
  int bf16;                               // 1: bf16 hi / mid planes (ld_out in 2-byte elements), 0: TF32 hi / lo
};
// hi / lo TF32 planes of up to 16 weight matrices in one launch (gemm_tcw.cu)
int weight_split_jobs(const SplitJobHost* jobs, int n, cudaStream_t st);

// C[M,N] = A[M,K] * (Bh + Bl)[N,K]^T (+bias, ReLU -> bits_out, masked by bits_in)   (gemm_tcw.cu)
int gemm_w(int M, int N, int K, const float* A, int lda, const float* Bh, const float* Bl, int ldb, float* C,
           int ldc, const float* bias, int relu, const uint32_t* bits_in, uint32_t* bits_out,
           const int32_t* rows_dev, cudaStream_t st);

// C[M,N] += A[K,M]^T B[K,N], split over K; colsum[m] += sum_k A[k][m]   (gemm_tcw.cu)
int gemm_tn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            float* colsum, const int32_t* rows_dev, cudaStream_t st);

// One step of the fused MLP chain (gemm_tcw.cu: mlp_chain_kernel): C = act(A W^T + b) on output columns
// [n0, n0 + 256) with pre-split weight planes Bh / Bl.
struct ChainStepHost {
  const float* A; int lda; int K;
  const float* Bh; const float* Bl; int ldb;
  float* C; int ldc; int N; int n0;
  const float* bias; int relu;
  const uint32_t* bits_in; uint32_t* bits_out;
};
int mlp_chain(const ChainStepHost* steps, int n_steps, int M, const int32_t* rows_dev, cudaStream_t st);

// One weight-gradient problem of mlp_dw (gemm_tcw.cu: mlp_dw_kernel): C[M,N] += A[K,M]^T B[K,N], colsum[m] += sum_k A[k][m]
struct DwProblemHost {
  const float* A; const float* B; float* C; float* colsum;
  int M, N, lda, ldb, ldc;
  // bf16 variant: A is given as bf16 hi / mid planes [K][ldp] (ldp elements, % 8 == 0) instead of fp32, and colsum
  // must be NULL (the bias gradient is a by-product of whoever wrote the planes)
  const uint16_t* Ahi = nullptr; const uint16_t* Amid = nullptr; int ldp = 0;
};
// bf16: 3xBF16 operands (kind::f16, 64-sample slabs) instead of 3xTF32
int mlp_dw(const DwProblemHost* probs, int n, int K, const int32_t* rows_dev, cudaStream_t st, bool bf16 = false);
// C[M <= 4][N] += A[K][M]^T B[K][N], fp32 SIMT: layers with a handful of out-features (the last refnet layer)
int dw_small(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             const int32_t* rows_dev, cudaStream_t st);
// leave n SMs free while mlp_dw runs (a data-parallel step overlaps the gradient exchange with it)
void set_dw_sm_reserve(int n_sms);

// One step of the backward-data chain on bf16 planes (gemm_tcw.cu: mlp_bwd_planes_kernel):
//   out[:, n0 : n0 + 256) = (A W^T-planes) masked by bits_in, A = Ahi + Amid [rows][lda] bf16 planes (lda elements),
//   written as bf16 planes Chi / Cmid [rows][256] (out_planes) or as fp32 C [rows][ldc]; colsum += column sums.
struct BwStepHost {
  const uint16_t* Ahi; const uint16_t* Amid; int lda; int K;
  const float* Bh; const float* Bl; int ldb;           // bf16 hi / mid weight planes [N][ldb] (as mlp_chain, bf16 = true)
  float* C; int ldc; uint16_t* Chi; uint16_t* Cmid; int out_planes;
  int N; int n0;
  const uint32_t* bits_in; float* colsum;
};
int mlp_bwd_planes(const BwStepHost* steps, int n_steps, int M, const int32_t* rows_dev, cudaStream_t st);
// fp32 [rows][ld] -> bf16 hi / mid planes [rows][ldp] (+ optional column sums over the first min(rows, *rows_dev) rows)
int split_planes(const float* X, int ld, int rows, int cols, uint16_t* hi, uint16_t* mid, int ldp, float* colsum,
                 const int32_t* rows_dev, cudaStream_t st);

// Layout of FgsFineParams::w_split (in floats): per layer the straight TF32 hi / lo planes [out][in_ld] (forward, 3xTF32)
// and the transposed bf16 hi / mid planes [in_ld][ldt] (backward-data, 3xBF16; ldt in 2-byte elements).
struct WSplitLayout {
  size_t hi[FGS_MLP_LAYERS], lo[FGS_MLP_LAYERS], thi[FGS_MLP_LAYERS], tlo[FGS_MLP_LAYERS];
  int in_ld[FGS_MLP_LAYERS], out[FGS_MLP_LAYERS], ldt[FGS_MLP_LAYERS];
  size_t total;
};
inline WSplitLayout wsplit_layout() {
  WSplitLayout L;
  const int in_ld[FGS_MLP_LAYERS] = {FGS_X0_LD, 256, 256, 256, FGS_XR_LD, 256, 256, 256};
  size_t o = 0;
  for (int l = 0; l < FGS_MLP_LAYERS; ++l) {
    L.in_ld[l] = in_ld[l];
    L.out[l] = (l == FGS_MLP_LAYERS - 1) ? 3 : FGS_MLP_WIDTH;
    L.ldt[l] = (l == FGS_MLP_LAYERS - 1) ? 8 : FGS_MLP_WIDTH;     // bf16 rows are 16-byte multiples
    // straight planes are allocated with the row count rounded up to 8 (TMA row groups)
    const size_t s = (size_t)((L.out[l] + 7) & ~7) * in_ld[l], t = (size_t)in_ld[l] * L.ldt[l] / 2;
    L.hi[l] = o; o += s;
    L.lo[l] = o; o += s;
    L.thi[l] = o; o += t;
    L.tlo[l] = o; o += t;
  }
  L.total = o;
  return L;
}

}  // namespace fine
