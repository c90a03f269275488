// Tensor-core GEMMs of the colour MLPs: tcgen05.mma (kind::tf32) with fp32-grade accuracy --
// 3xTF32 operand split + K-chunked accumulation -- accumulators in TMEM, fused bias / ReLU /
// ReLU-mask / split-K epilogues (sm_100a).
//
// Why not plain TF32: the reference runs its nn.Linear layers as true-fp32 SGEMMs (TF32 off,
// model/nerf.py:125-151) and the parity bar is 1e-4 relative on colours AND gradients.
//  (1) 3xTF32: each fp32 operand x is split on the fly into
//        hi = rn_tf32(x)          (cvt.rna: 10-bit mantissa, round to nearest),
//        lo = rn_tf32(x - hi)     (the residue, again rounded to TF32: split error <= 2^-23 |x|),
//      and D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi is issued as three MMAs per k-step (the dropped
//      A_lo*B_lo term is ~2^-22 relative).
//  (2) K-chunked accumulation: the TMEM accumulator rounds TOWARD ZERO at every MMA (measured:
//      error grows linearly with the number of accumulations, ~3e-6 relative after the 96 MMAs of
//      a K=256 layer -- enough to switch the ReLU branch of a few of the 2e7 hidden units of a
//      batch, and one switched unit moves its sample's gradients by O(1)).  So the tensor core only
//      accumulates 32 k (12 MMAs) into a fresh TMEM region; two regions alternate, and the CUDA
//      cores add each finished chunk into fp32 REGISTER accumulators with round-to-nearest adds
//      while the tensor core works on the next chunk.  The result is at fp32-SGEMM noise level.
//
// Shape of the kernel (one CTA = one 128 x BN output tile, BN <= 256, one CTA per SM):
//   * 256 threads load a 16-wide k-slab of A and B from global, split it in registers and store
//     hi/lo tiles to shared memory in the canonical no-swizzle K-major UMMA layout (8x16-byte
//     core matrices).  Operands whose reduction index is the SLOW global dimension (W in dZ*W, dZ
//     and X in dZ^T*X) are transposed by the loader itself -- coalesced 32-bit loads along the
//     contiguous dimension, 128-bit conflict-free smem stores -- because kind::tf32 reads MN-major
//     tiles only in the 128B_BASE32B swizzle (tools/umma_probe.cu: a no-swizzle MN-major
//     descriptor multiplies by zeros).  No transposed copies in HBM;
//   * four smem stages (4 x 48 KB) so the loaders run up to four slabs ahead of the tensor core,
//     register prefetch two slabs ahead of the smem stores; one elected thread issues the 6 MMAs
//     of a slab and commits them to the stage's mbarrier (and the chunk's region mbarrier);
//   * all 512 TMEM columns: two 256-column chunk regions;
//   * epilogue from the register accumulators: bias / ReLU / mask -> 128-bit stores, or
//     red.global.add for the split-K weight-gradient products.
// The sample-dimension extent is read on the device (rows_dev), like the SIMT path in gemm.cu.
#include "common.cuh"
#include "../../include/fgs_b200_fine.h"

namespace {

constexpr int BM = 128, BK = 16, kThreads = 256;
constexpr int kTmemCols = 512;
constexpr uint32_t kATile = BM * BK * 4;        // 8 KB
constexpr uint32_t kBTile = 256 * BK * 4;       // 16 KB
constexpr uint32_t kStage = 2 * kATile + 2 * kBTile;  // 48 KB: A_hi, A_lo, B_hi, B_lo
constexpr int kStages = 4;
constexpr int kChunkSlabs = 2;                  // 32 k per TMEM accumulation chunk
constexpr uint32_t kSmemBytes = kStages * kStage + 128;

struct TcArgs {
  const float* A; const float* B; float* C;
  const float* bias; const float* relu_mask; const int* rows_dev;
  int M, N, K, lda, ldb, ldc, ld_mask, relu;
  float* colsum;  // SPLITK only (may be NULL): colsum[m] += sum_k A[k][m]  (bias gradients ride along)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  // bounded spin: a descriptor / phase bug must surface as an error, never as a hung GPU
  for (uint32_t it = 0; it < (1u << 26) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  if (!done) __trap();
}

// K-major no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  // arrives on the mbarrier once every MMA issued so far by this thread has finished
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 split_lo(float4 x, float4& hi) {
  hi.x = rn_tf32(x.x); hi.y = rn_tf32(x.y); hi.z = rn_tf32(x.z); hi.w = rn_tf32(x.w);
  return make_float4(rn_tf32(x.x - hi.x), rn_tf32(x.y - hi.y), rn_tf32(x.z - hi.z), rn_tf32(x.w - hi.w));
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct Slab {
  float4 a[2], b[4];
};

// A_MN: A is [K][M] (m contiguous) else [M][K].  B_MN: B is [K][N] (n contiguous) else [N][K].
// Either way the smem tiles are K-major: element (row, k) at (row/8)*512 + (k/4)*128 + (row%8)*16 + (k%4)*4.
template <bool A_MN, bool B_MN, bool SPLITK>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  const uint32_t stage_bar0 = smem + kStages * kStage;       // kStages mbarriers
  const uint32_t region_bar0 = stage_bar0 + 8 * kStages;     // 2 mbarriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + kStages * kStage + 8 * kStages + 16 + 16);

  int M = a.M, K = a.K;
  if (a.rows_dev) {
    const int r = *a.rows_dev;
    if (SPLITK) K = min(K, r); else M = min(M, r);
  }
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * 256;
  if (m0 >= M || n0 >= a.N) return;
  int kbeg = 0, kend = K;
  if (SPLITK) {
    int chunk = (K + gridDim.z - 1) / gridDim.z;
    chunk = (chunk + BK * kChunkSlabs - 1) / (BK * kChunkSlabs) * (BK * kChunkSlabs);
    kbeg = blockIdx.z * chunk;
    kend = min(K, kbeg + chunk);
    if (kbeg >= kend) return;
  }
  const int BN = min(256, (a.N - n0 + 15) & ~15);  // UMMA N: multiple of 16
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: TMEM allocation (warp 0), mbarriers (thread 0) ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < kStages + 2; ++i) mbar_init(stage_bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32 (K-major), N>>3, M>>4
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                         ((uint32_t)(BM >> 4) << 24);

  float csum = 0.f;
  // slot idx -> (row = tile row, c = 16-byte k-chunk): K-major operands take one 128-bit load per
  // slot; transposed operands take four coalesced 32-bit loads (k .. k+3 of one row)
  auto load_slab = [&](Slab& s, int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + kThreads * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!A_MN) {
        const int r = idx >> 2, c = idx & 3;
        const int m = m0 + r, k = k0 + 4 * c;
        if (m < M && k < kend) v = ldg4(a.A + (size_t)m * a.lda + k);
      } else {
        const int r = idx & 127, c = idx >> 7;
        const int m = m0 + r, k = k0 + 4 * c;
        if (m < M) {
          const float* p = a.A + (size_t)k * a.lda + m;
          if (k < kend) v.x = __ldg(p);
          if (k + 1 < kend) v.y = __ldg(p + a.lda);
          if (k + 2 < kend) v.z = __ldg(p + 2 * (size_t)a.lda);
          if (k + 3 < kend) v.w = __ldg(p + 3 * (size_t)a.lda);
        }
        if (SPLITK) csum += (v.x + v.y) + (v.z + v.w);   // both slots of a thread belong to row idx & 127
      }
      s.a[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + kThreads * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!B_MN) {
        const int r = idx >> 2, c = idx & 3;
        const int n = n0 + r, k = k0 + 4 * c;
        if (r < BN && n < a.N && k < kend) v = ldg4(a.B + (size_t)n * a.ldb + k);
      } else {
        const int r = idx & 255, c = idx >> 8;
        const int n = n0 + r, k = k0 + 4 * c;
        if (r < BN && n < a.N) {
          const float* p = a.B + (size_t)k * a.ldb + n;
          if (k < kend) v.x = __ldg(p);
          if (k + 1 < kend) v.y = __ldg(p + a.ldb);
          if (k + 2 < kend) v.z = __ldg(p + 2 * (size_t)a.ldb);
          if (k + 3 < kend) v.w = __ldg(p + 3 * (size_t)a.ldb);
        }
      }
      s.b[i] = v;
    }
  };
  auto store_slab = [&](const Slab& s, uint32_t stage_base) {
    const uint32_t a_hi = stage_base, a_lo = stage_base + kATile;
    const uint32_t b_hi = stage_base + 2 * kATile, b_lo = b_hi + kBTile;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + kThreads * i;
      const int r = A_MN ? (idx & 127) : (idx >> 2), c = A_MN ? (idx >> 7) : (idx & 3);
      const uint32_t off = (uint32_t)((r >> 3) * 512 + c * 128 + (r & 7) * 16);
      float4 hi;
      const float4 lo = split_lo(s.a[i], hi);
      st_shared_v4(a_hi + off, hi);
      st_shared_v4(a_lo + off, lo);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + kThreads * i;
      const int r = B_MN ? (idx & 255) : (idx >> 2), c = B_MN ? (idx >> 8) : (idx & 3);
      if (r >= BN) continue;
      const uint32_t off = (uint32_t)((r >> 3) * 512 + c * 128 + (r & 7) * 16);
      float4 hi;
      const float4 lo = split_lo(s.b[i], hi);
      st_shared_v4(b_hi + off, hi);
      st_shared_v4(b_lo + off, lo);
    }
  };

  // register accumulators: thread (warp w, lane l) owns tile row 32*(w%4)+l, columns 128*(w/4) .. +127
  float acc[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) acc[i] = 0.f;
  const int col_half = (warp >> 2) * 128;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;

  auto drain = [&](int c) {
    mbar_wait(region_bar0 + 8 * (c & 1), (uint32_t)((c >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (col_half + 32 * q < BN) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + lane_base + (uint32_t)((c & 1) * 256 + col_half + 32 * q);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 * q + j] = __fadd_rn(acc[32 * q + j], __uint_as_float(r[j]));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  const int n_slabs = (kend - kbeg + BK - 1) / BK;
  const int n_chunks = (n_slabs + kChunkSlabs - 1) / kChunkSlabs;
  Slab s0, s1;  // slabs kt (even / odd alternate): prefetch distance 2
  load_slab(s0, kbeg);
  if (n_slabs > 1) load_slab(s1, kbeg + BK);
  for (int kt = 0; kt < n_slabs; ++kt) {
    const int st = kt % kStages;
    const uint32_t stage_base = smem + st * kStage;
    const int c = kt / kChunkSlabs;
    const bool first_in_chunk = (kt % kChunkSlabs) == 0;
    const bool last_in_chunk = (kt % kChunkSlabs) == kChunkSlabs - 1 || kt == n_slabs - 1;
    if (kt >= kStages) mbar_wait(stage_bar0 + 8 * st, (uint32_t)(((kt / kStages) - 1) & 1));
    if (kt & 1) {
      store_slab(s1, stage_base);
      if (kt + 2 < n_slabs) load_slab(s1, kbeg + (kt + 2) * BK);
    } else {
      store_slab(s0, stage_base);
      if (kt + 2 < n_slabs) load_slab(s0, kbeg + (kt + 2) * BK);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = stage_base, a_lo = stage_base + kATile;
      const uint32_t b_hi = stage_base + 2 * kATile, b_lo = b_hi + kBTile;
      const uint32_t d = tmem_d + (uint32_t)((c & 1) * 256);
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {
        // k-step j = 16-byte k-chunks 2j, 2j+1: LBO 128 (chunk stride), SBO 512 (8-row group stride)
        const uint64_t dah = make_desc(a_hi + j * 256, 128, 512), dal = make_desc(a_lo + j * 256, 128, 512);
        const uint64_t dbh = make_desc(b_hi + j * 256, 128, 512), dbl = make_desc(b_lo + j * 256, 128, 512);
        mma_tf32(d, dal, dbh, idesc, (!first_in_chunk || j != 0) ? 1u : 0u);
        mma_tf32(d, dah, dbl, idesc, 1);
        mma_tf32(d, dah, dbh, idesc, 1);
      }
      mma_commit(stage_bar0 + 8 * st);
      if (last_in_chunk) mma_commit(region_bar0 + 8 * (c & 1));
    }
    // add the previous chunk into the register accumulators while this one multiplies
    if (last_in_chunk && c >= 1) drain(c - 1);
  }
  drain(n_chunks - 1);

  if (SPLITK && a.colsum && blockIdx.x == 0) {
    const int mr = m0 + (tid & 127);
    if (mr < M && csum != 0.f) atomicAdd(a.colsum + mr, csum);
  }
  // ---- epilogue from registers ----
  const int m = m0 + (warp & 3) * 32 + lane;
  if (m < M) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c0 = col_half + 32 * q;
      if (c0 < BN) {
        float* __restrict__ crow = a.C + (size_t)m * a.ldc + n0 + c0;
        const float* __restrict__ mrow = a.relu_mask ? a.relu_mask + (size_t)m * a.ld_mask + n0 + c0 : nullptr;
        const int nvalid = min(32, a.N - n0 - c0);
        if (SPLITK) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) atomicAdd(crow + j, acc[32 * q + j]);
        } else {
          const bool vec = (nvalid == 32) && ((a.ldc & 3) == 0) && (!mrow || (a.ld_mask & 3) == 0);
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = acc[32 * q + j4 + e];
              if (a.bias && j4 + e < nvalid) x += __ldg(a.bias + n0 + c0 + j4 + e);
              if (a.relu) x = fmaxf(x, 0.f);
              v[e] = x;
            }
            if (vec) {
              if (mrow) {
                const float4 mk = ldg4(mrow + j4);
                if (!(mk.x > 0.f)) v[0] = 0.f;
                if (!(mk.y > 0.f)) v[1] = 0.f;
                if (!(mk.z > 0.f)) v[2] = 0.f;
                if (!(mk.w > 0.f)) v[3] = 0.f;
              }
              *reinterpret_cast<float4*>(crow + j4) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (j4 + e < nvalid) {
                  float x = v[e];
                  if (mrow && !(mrow[j4 + e] > 0.f)) x = 0.f;
                  crow[j4 + e] = x;
                }
              }
            }
          }
        }
      }
    }
  }

  // ---- teardown ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
  }
}

template <bool A_MN, bool B_MN, bool SPLITK>
int launch(const TcArgs& a, dim3 grid, cudaStream_t st) {
  static bool configured = false;
  auto kern = gemm_tc_kernel<A_MN, B_MN, SPLITK>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return fgs::check_cuda(e, "gemm_tc smem attribute");
    configured = true;
  }
  kern<<<grid, kThreads, kSmemBytes, st>>>(a);
  return fgs::check_launch("gemm_tc", st);
}

}  // namespace

namespace fine {

// Same contract as fine::gemm (gemm.cu); additionally K-major operands (A in modes 0/1, B in mode
// 0) must be 16-byte aligned with a leading dimension that is a multiple of 4 floats and readable
// up to the next multiple of 4 along k (the fused step pads its rows with zeros).
int gemm_tc(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
            int ldc, const float* bias, int relu, const float* relu_mask, int ld_mask,
            const int32_t* rows_dev, cudaStream_t st, float* colsum) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const bool a_vec = (mode != 2), b_vec = (mode == 0);   // operands read with 128-bit loads
  if ((a_vec && ((lda & 3) || ((uintptr_t)A & 15))) || (b_vec && ((ldb & 3) || ((uintptr_t)B & 15))))
    return fgs::fail_arg("gemm_tc needs 16-byte aligned K-major operands with leading dimension % 4 == 0");
  TcArgs a{A, B, C, bias, relu_mask, rows_dev, M, N, K, lda, ldb, ldc, ld_mask, relu, colsum};
  dim3 grid(fgs::cdiv(N, 256), fgs::cdiv(M, BM), 1);
  switch (mode) {
    case 0: return launch<false, false, false>(a, grid, st);
    case 1: return launch<false, true, false>(a, grid, st);
    case 2: {
      const int tiles = grid.x * grid.y;
      int splits = (FGS_NUM_SMS + tiles - 1) / tiles;
      const int max_splits = fgs::cdiv(K, 8 * BK);
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
      grid.z = splits;
      return launch<true, true, true>(a, grid, st);
    }
    default: return fgs::fail_arg("gemm mode must be 0,1,2");
  }
}

}  // namespace fine

extern "C" int fgs_gemm_tc(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                           float* C, int ldc, const float* bias, int relu, const float* relu_mask,
                           int ld_mask, const int32_t* rows_dev, fgs_stream_t stream) {
  FGS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
  return fine::gemm_tc(mode, M, N, K, A, lda, B, ldb, C, ldc, bias, relu, relu_mask, ld_mask, rows_dev,
                       (cudaStream_t)stream, nullptr);
}
