// Tensor-core GEMMs "activations x weights" of the colour MLPs (forward  Y = X W^T + b  and
// backward-data  dX = dZ W): the weight operand arrives PRE-SPLIT and by TMA (sm_100a).
//
// Same numerics as gemm_tc.cu (3xTF32 split, 32-k accumulation chunks in rotating TMEM regions,
// fp32 register accumulators), different data movement -- each choice below answers a measured
// cost of an earlier version (tools/gemm_phase_probe.py, per-phase clock64 of one CTA):
//   * the 0.45 M MLP weights are split ONCE per step into hi / lo TF32 planes, straight ([out][in],
//     forward) and transposed ([in][out], dZ*W) -- weight_split_kernel, one launch -- instead of
//     being re-split from fp32 by every CTA for every k-slab;
//   * the B (weight) tiles are fetched by TMA: one elected thread issues two 2-D
//     cp.async.bulk.tensor per k-slab (hi, lo planes; box 32 floats x 128 rows, SWIZZLE_128B), no
//     registers, no LSU instructions, completion on an mbarrier (expect_tx), prefetched two slabs
//     ahead across tile boundaries;
//   * THREE roles that only meet at mbarriers: 4 producer warps (activation slab: 128-bit loads, two
//     slabs in flight -> hi / lo split -> XOR-swizzled st.shared), one issuer thread (TMA + all
//     tcgen05.mma), 8 consumer warps (drain each finished chunk into registers, then the epilogue).
//     With producers == consumers (the previous version) the tensor core idled during every drain and
//     every epilogue: 42 % MMA occupancy inside a tile, 53 us per 27 K x 256 x 256 GEMM;
//   * four 128-column TMEM regions rotate, so the tensor core runs up to 4 chunks ahead of the drain
//     and the epilogue of tile t overlaps the MMAs of tile t + 1;
//   * persistent CTAs (one per SM) loop over the 128-row tiles of the sample dimension, whose
//     extent is read on the device; TMEM (512 columns) and mbarriers are set up once per CTA;
//   * ReLU masks travel as BITS: the forward epilogue packs (x > 0) of its 32 columns into one
//     word per thread, the backward epilogue reads 4 words instead of 128 floats of H;
//   * the epilogue leaves through TMA: each consumer warp writes its 32 rows x 32 columns as swizzled
//     128-byte rows (conflict-free 16-byte st.shared) and one lane issues a bulk tensor store
//     (row-per-thread 16-byte global stores 1 KB apart cost 11 us per tile; a smem transpose with
//     scalar readback cost 15 K cycles per tile at one warp per scheduler).
#include <cuda.h>

#include <cstring>
#include <map>
#include <tuple>

#include "common.cuh"
#include "tc_api.cuh"

namespace {

constexpr int BM = 128, BK = 32;                    // one k-slab = 32 floats = one 128-byte swizzle row
constexpr int TN = 128;                              // output columns per CTA (N = 256 -> two column CTAs)
constexpr int kTmemCols = 512;
constexpr int kRegions = 4;                          // 4 accumulation regions of 128 TMEM columns
constexpr uint32_t kATile = BM * BK * 4;             // 16 KB
constexpr uint32_t kBTile = TN * BK * 4;             // 16 KB
constexpr uint32_t kStage = 2 * kATile + 2 * kBTile; // A_hi, A_lo, B_hi, B_lo = 64 KB
constexpr int kStages = 3;
constexpr uint32_t kStagingOff = kStages * kStage;   // 8 consumer warps x (32 rows x 128 bytes), 1024-byte aligned
constexpr uint32_t kStagingWarp = 32 * 128;
constexpr uint32_t kStagingBytes = 8 * kStagingWarp;
constexpr uint32_t kBarOff = kStagingOff + kStagingBytes;  // mbarriers + tmem slot, then the bias slice
constexpr uint32_t kBiasOff = kBarOff + 256;
constexpr uint32_t kSmemBytes = kBiasOff + TN * 4;
constexpr int kAcc = TN / 2;                         // accumulator registers per consumer thread (half a row)

// Thread roles (13 warps, 128 registers per thread; ptxas does not raise a role's allocation above the
// kernel-wide cap after setmaxnreg, so the roles are sized to fit it):
//   warps 0-7    consumers (warp w: TMEM lanes 32 (w % 4).., columns 64 (w / 4)..): drain every finished
//                32-k chunk from TMEM into fp32 register accumulators (round-to-nearest adds), then the
//                bias / ReLU / mask epilogue, swizzled st.shared and one TMA store per 32x32 block
//   warps 8-11   producers: activation slab global -> registers (two slabs in flight) -> hi / lo TF32
//                split -> swizzled st.shared
//   warp 12      lane 0 issues TMA (weight planes) and every tcgen05.mma
constexpr int kConsumers = 256;
constexpr int kProducers = 128;
constexpr int kIssuerWarp = (kConsumers + kProducers) / 32;
constexpr int kThreadsWS = kConsumers + kProducers + 32;

struct WArgs {
  const float* A; float* C;
  const float* bias; const uint32_t* bits_in; uint32_t* bits_out; const int* rows_dev;
  int M, N, K, lda, ldc, relu, c_tma;
  long long* dbg;  // optional per-role phase timing of CTA (0,0) (tools/gemm_phase_probe.py), 10 values
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  if (!done) __trap();  // a phase bug must surface as an error, never as a hung GPU
}
// The same wait for the roles that are not on the critical hand-off path (producers waiting for a free stage,
// consumers waiting for a finished accumulator): back off between polls, so the waiting warps do not take issue
// slots from the working warps of their scheduler (a producer warp sharing its scheduler with two spinning
// consumer warps needed 1.8 K cycles for an EMPTY slab iteration).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    __nanosleep(40);
  }
  __trap();
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row (1024-byte) swizzle
// atoms (SBO 1024), 16-byte chunk c of row r stored at chunk position c ^ (r % 8); layout_type = 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
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
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Round-to-nearest (ties away from zero) to TF32 = cvt.rna.tf32.f32, but on the integer pipe: the conversion
// instruction issues at 16 lanes / clock / SM and made the operand producers the bottleneck of every GEMM here
// (2 cvt per element: 2 K cycles per 64 KB slab against 1.3 K cycles of MMAs); add-half-ulp-and-mask is two
// full-rate ALU ops and bit-identical for finite inputs.
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// streamed operands bypass L1: with ~210 KB of the SM's 256 KB carved out as shared memory the remaining L1 tracks too
// few outstanding lines to keep 32-64 KB of loads in flight (ld.global.nc ran the dW producers at 4 K cycles per slab)
__device__ __forceinline__ float4 ldcg4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// hi / lo TF32 planes of a weight matrix W [rows][ld] (cols valid), straight or transposed.
struct SplitJob {
  const float* W; float* hi; float* lo;
  int rows, cols, ld, ld_out, transpose;  // transpose: out[c][r] (ld_out >= rows) else out[r][c] (ld_out >= cols)
  int bf16;                               // planes are bf16 hi / mid (ld_out in elements) instead of TF32 hi / lo
};
struct SplitJobs {
  SplitJob j[16];
  int n;
};

// One CTA (32 x 8 threads) per 32 x 32 tile of a job, four rows per thread: the reads are coalesced along the columns
// of W; a straight job writes the same way, a transposed job goes through a padded shared-memory tile so that its
// writes are coalesced along the rows of W (the first version indexed the flat element range with a division per
// element and wrote the transposed bf16 planes as scattered 2-byte stores: 18 us per step for 7 MB).
__global__ void __launch_bounds__(256) weight_split_kernel(const __grid_constant__ SplitJobs jobs) {
  const SplitJob& J = jobs.j[blockIdx.y];
  __shared__ uint32_t s_hi[32][33], s_lo[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tiles_c = (J.cols + 31) >> 5, tiles_r = (J.rows + 31) >> 5;
  for (int t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const int r0 = (t / tiles_c) << 5, c0 = (t - (t / tiles_c) * tiles_c) << 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + 8 * k, c = c0 + tx;
      uint32_t h = 0, l = 0;
      if (r < J.rows && c < J.cols) {
        const float x = J.W[(size_t)r * J.ld + c];
        if (J.bf16) {
          h = (__float_as_uint(x) + 0x8000u) & 0xFFFF0000u;
          l = __float_as_uint(x - __uint_as_float(h)) + 0x8000u;
        } else {
          const float hf = rn_tf32(x);
          h = __float_as_uint(hf);
          l = __float_as_uint(rn_tf32(x - hf));
        }
        if (!J.transpose) {
          const size_t o = (size_t)r * J.ld_out + c;
          if (J.bf16) {
            reinterpret_cast<uint16_t*>(J.hi)[o] = (uint16_t)(h >> 16);
            reinterpret_cast<uint16_t*>(J.lo)[o] = (uint16_t)(l >> 16);
          } else {
            J.hi[o] = __uint_as_float(h);
            J.lo[o] = __uint_as_float(l);
          }
        }
      }
      if (J.transpose) { s_hi[ty + 8 * k][tx] = h; s_lo[ty + 8 * k][tx] = l; }
    }
    if (J.transpose) {                                   // (uniform per CTA)
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;     // out[c][r]: consecutive lanes -> consecutive r
        if (r < J.rows && c < J.cols) {
          const uint32_t h = s_hi[tx][ty + 8 * k], l = s_lo[tx][ty + 8 * k];
          const size_t o = (size_t)c * J.ld_out + r;
          if (J.bf16) {
            reinterpret_cast<uint16_t*>(J.hi)[o] = (uint16_t)(h >> 16);
            reinterpret_cast<uint16_t*>(J.lo)[o] = (uint16_t)(l >> 16);
          } else {
            J.hi[o] = __uint_as_float(h);
            J.lo[o] = __uint_as_float(l);
          }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Consumer side of the accumulation: chunk c (one 32-k slab = 12 MMAs) has been committed into TMEM region
// c % kRegions; add this warp's 64 columns of it into the thread's fp32 register accumulators (its row) and
// hand the region back.  tmem_row = base + (lane quarter << 16) + column half.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr);
// rempty_cluster0 != 0: r_empty lives in the leader CTA of a pair (shared::cluster address), arrive there
template <bool PAIRED>
__device__ __forceinline__ void drain_chunk(float (&acc)[kAcc], uint32_t tmem_row, int col_half, uint32_t c, int BN,
                                            uint32_t rfull_bar0, uint32_t rempty_bar0, uint32_t rempty_cluster0 = 0) {
  const uint32_t rg = c % kRegions;
  mbar_wait_relaxed(rfull_bar0 + 8 * rg, (c / kRegions) & 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
  for (int q = 0; q < kAcc / 32; ++q) {
    if (col_half + 32 * q < BN) {
      if (PAIRED) {                         // two 16-column loads in flight
        uint32_t r0[16], r1[16];
        tmem_ld16(tmem_row + rg * TN + 32 * q, r0);
        tmem_ld16(tmem_row + rg * TN + 32 * q + 16, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[32 * q + j] = __fadd_rn(acc[32 * q + j], __uint_as_float(r0[j]));
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[32 * q + 16 + j] = __fadd_rn(acc[32 * q + 16 + j], __uint_as_float(r1[j]));
      } else {                              // tighter register budget: one load at a time
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r0[16];
          tmem_ld16(tmem_row + rg * TN + 32 * q + 16 * h, r0);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            acc[32 * q + 16 * h + j] = __fadd_rn(acc[32 * q + 16 * h + j], __uint_as_float(r0[j]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {   // the region may be overwritten by chunk c + kRegions
    if (rempty_cluster0) mbar_arrive_cluster(rempty_cluster0 + 8 * rg);
    else mbar_arrive(rempty_bar0 + 8 * rg);
  }
}

// TMA store of one swizzled 32-row x 128-byte block (issued by one lane; completion tracked per thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// hi / lo TF32 split of one 16-byte chunk, stored at the same (swizzled) offset of the hi and lo tiles.
// hi is rounded to nearest; lo = x - hi is stored as it is (exact, <= 13 significant bits): the tensor core reads
// the upper 19 bits of a kind::tf32 operand, i.e. truncates lo by < 2^-10 |lo| <= 2^-21 |x| -- below the 2^-21..-22
// the rounded version reached, and 8 ALU ops per chunk less in producers that are issue-bound.
__device__ __forceinline__ void split_store(uint32_t hi_addr, uint32_t lo_addr, const float4& v) {
  float4 hi, lo;
  hi.x = rn_tf32(v.x); hi.y = rn_tf32(v.y); hi.z = rn_tf32(v.z); hi.w = rn_tf32(v.w);
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
  st_shared_v4(hi_addr, hi);
  st_shared_v4(lo_addr, lo);
}

// ---- 3xBF16 operands (backward kernels) ----
// x = hi + mid + O(2^-17 |x|) with hi = bf16(x), mid = bf16(x - hi); the products hi*hi + hi*mid + mid*hi leave
// a relative error of ~2^-16 per product (the mid*mid term), measured end to end at <= 1.2e-5 on every gradient
//  -- inside the 1e-4 bar -- at HALF the tensor-core instructions of 3xTF32: a kind::f16 MMA covers
// K = 16 in the 32 bytes per row a kind::tf32 MMA needs for K = 8.  Rounding on the integer pipe (add half an ulp,
// keep the upper 16 bits), two values packed per word with one PRMT.
__device__ __forceinline__ void bf16_split2(float x0, float x1, uint32_t& hi, uint32_t& mid) {
  const uint32_t h0 = (__float_as_uint(x0) + 0x8000u) & 0xFFFF0000u, h1 = (__float_as_uint(x1) + 0x8000u) & 0xFFFF0000u;
  const uint32_t m0 = __float_as_uint(x0 - __uint_as_float(h0)) + 0x8000u;
  const uint32_t m1 = __float_as_uint(x1 - __uint_as_float(h1)) + 0x8000u;
  hi = __byte_perm(h0, h1, 0x7632);      // element 0 in the low half
  mid = __byte_perm(m0, m1, 0x7632);
}
__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_v2u(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// 4 consecutive MN elements of one k row (MN-major tiles of the weight-gradient kernel) -> 8 bytes per tile
__device__ __forceinline__ void split_store_bf4(uint32_t hi_addr, uint32_t mid_addr, const float4& a) {
  uint32_t h[2], m[2];
  bf16_split2(a.x, a.y, h[0], m[0]);
  bf16_split2(a.z, a.w, h[1], m[1]);
  st_shared_v2u(hi_addr, h[0], h[1]);
  st_shared_v2u(mid_addr, m[0], m[1]);
}

struct Bars {
  uint32_t free0, bfull0, afull0, rfull0, rempty0;
};
__device__ __forceinline__ Bars make_bars(uint32_t smem) {
  Bars b;
  b.free0 = smem + kBarOff;
  b.bfull0 = b.free0 + 8 * kStages;
  b.afull0 = b.bfull0 + 8 * kStages;
  b.rfull0 = b.afull0 + 8 * kStages;
  b.rempty0 = b.rfull0 + 8 * kRegions;
  return b;
}
constexpr uint32_t kTmemSlotOff = kBarOff + 8 * (3 * kStages + 2 * kRegions) + 8;

// TMEM allocation + mbarrier initialisation, common to both kernels; returns the TMEM base address
__device__ __forceinline__ uint32_t cta_setup(unsigned char* smem_raw, const Bars& B, int warp, int tid,
                                               int n_producers) {
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + kTmemSlotOff);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(B.free0 + 8 * i, 1);
      mbar_init(B.bfull0 + 8 * i, 1);
      mbar_init(B.afull0 + 8 * i, n_producers / 32);       // one arrival per producer warp
    }
    for (int i = 0; i < kRegions; ++i) {
      mbar_init(B.rfull0 + 8 * i, 1);
      mbar_init(B.rempty0 + 8 * i, kConsumers / 32);       // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *tmem_slot;
}
__device__ __forceinline__ void cta_teardown(uint32_t tmem_d, int warp) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
  }
}

// 12 MMAs of one 32-k slab: D (+)= A_lo B_hi + A_hi B_lo + A_hi B_hi per 8-wide k-step, small terms first
__device__ __forceinline__ void mma_slab(uint32_t d, uint32_t stage_base, uint32_t idesc) {
  const uint32_t a_hi = stage_base, a_lo = stage_base + kATile;
  const uint32_t b_hi = stage_base + 2 * kATile, b_lo = b_hi + kBTile;
#pragma unroll
  for (int j = 0; j < BK / 8; ++j) {
    // k-step j = bytes [32 j, 32 j + 32) of every 128-byte row (the swizzle is a function of the address)
    const uint64_t dah = make_desc(a_hi + j * 32), dal = make_desc(a_lo + j * 32);
    const uint64_t dbh = make_desc(b_hi + j * 32), dbl = make_desc(b_lo + j * 32);
    mma_tf32(d, dal, dbh, idesc, j != 0 ? 1u : 0u);   // every slab starts a fresh accumulation region
    mma_tf32(d, dah, dbl, idesc, 1);
    mma_tf32(d, dah, dbh, idesc, 1);
  }
}

// Epilogue arithmetic on one 32-column block of a consumer thread's row, in place.
template <bool RELU, bool MASK, int NV = 32>
__device__ __forceinline__ uint32_t epilogue_block(float* v, uint32_t bias_addr, bool has_bias, uint32_t in_bits) {
  uint32_t out_bits = 0;
#pragma unroll
  for (int j4 = 0; j4 < NV / 4; ++j4) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_bias) b = lds4(bias_addr + 16 * j4);           // same address in every lane: a broadcast
    float x[4] = {v[4 * j4] + b.x, v[4 * j4 + 1] + b.y, v[4 * j4 + 2] + b.z, v[4 * j4 + 3] + b.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = 4 * j4 + t;
      if (RELU) {
        if (x[t] > 0.f) out_bits |= 1u << j;
        x[t] = fmaxf(x[t], 0.f);
      }
      if (MASK && !((in_bits >> j) & 1u)) x[t] = 0.f;
      v[j] = x[t];
    }
  }
  return out_bits;
}

// ---- fast variant: whole-K accumulation in TMEM, 256-column tiles --------------------------------------
// gemm_w_kernel keeps the tensor core's accumulation chains short (12 MMAs, then a round-to-nearest add on
// the CUDA cores) because the TMEM accumulator truncates at every MMA; that costs 128-column tiles (the
// register accumulators), a TMEM drain per 32 k and half-rate-sized MMAs (M128 N128 K8 takes 107 cycles,
// N256 171: tools/umma_bench).  This variant lets one accumulator take all K (96-120 MMAs: error ~3e-6
// of max|out| instead of 3e-7, both far inside the 1e-4 bar), which frees the structure:
//   * 128 x 256 output tiles, one M128 N256 K8 MMA per product: 25 % fewer tensor-core cycles per FLOP and
//     the activation slab is loaded / split once instead of once per column half;
//   * two 256-column TMEM regions ping-pong between TILES: the consumers' only job is the epilogue of
//     tile t (TMEM -> bias / ReLU / mask -> swizzled smem -> TMA store) while tile t + 1 multiplies;
//   * 8 producer warps, 8 consumer warps, 1 issuer warp (96 registers each, nobody holds accumulators).
constexpr int TNF = 256;
constexpr uint32_t kBTileF = TNF * BK * 4;                  // 32 KB
constexpr uint32_t kStageF = 2 * kATile + 2 * kBTileF;      // 96 KB
constexpr int kStagesF = 2;
static_assert(kStagesF * kStageF == kStagingOff, "staging / barrier offsets are shared with gemm_w_kernel");
constexpr uint32_t kSmemBytesF = kBiasOff + TNF * 4;
constexpr int kProducersF = 256;
constexpr int kIssuerWarpF = (kConsumers + kProducersF) / 32;
constexpr int kThreadsF = kConsumers + kProducersF + 32;

__global__ void __launch_bounds__(kThreadsF, 1)
gemm_wf_kernel(const WArgs a, const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const __grid_constant__ CUtensorMap tm_c) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  const Bars B = make_bars(smem);

  int M = a.M;
  if (a.rows_dev) M = min(M, *a.rows_dev);
  const int n0 = blockIdx.x * TNF;
  const int n_tiles = (M + BM - 1) / BM;
  if ((int)blockIdx.y >= n_tiles) return;
  const int K = a.K;
  const int BN = min(TNF, (a.N - n0 + 15) & ~15);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid < TNF) reinterpret_cast<float*>(smem_raw + kBiasOff)[tid] = (a.bias && n0 + tid < a.N) ? a.bias[n0 + tid] : 0.f;
  const uint32_t tmem_d = cta_setup(smem_raw, B, warp, tid, kProducersF);
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                         ((uint32_t)(BM >> 4) << 24);

  const uint32_t n_slabs = (uint32_t)((K + BK - 1) / BK);
  const int my_tiles = (n_tiles - 1 - (int)blockIdx.y) / (int)gridDim.y + 1;
  const uint32_t total_slabs = (uint32_t)my_tiles * n_slabs;
  const bool dbg_cta = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;

  if (warp >= kIssuerWarpF) {
    // ================= MMA + TMA issuer =================
    if (warp == kIssuerWarpF && lane == 0) {
      long long t_a = 0, t_b = 0, t_r = 0, t_f = 0, t_all = 0, t0 = 0;
      if (dbg_cta) t_all = clock64();
      const int n_boxes = BN > 128 ? 2 : 1;               // weight rows arrive in boxes of 128
      auto issue_b = [&](uint32_t gs) {
        const uint32_t st = gs % kStagesF;
        const uint32_t b_hi = smem + st * kStageF + 2 * kATile, b_lo = b_hi + kBTileF;
        const uint32_t bar = B.bfull0 + 8 * st;
        const int k0 = (int)(gs % n_slabs) * BK;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n_boxes * 2 * kBTile) : "memory");
        for (int bx = 0; bx < n_boxes; ++bx) {
          tma_load_2d(b_hi + bx * kBTile, &tm_hi, bar, k0, n0 + 128 * bx);
          tma_load_2d(b_lo + bx * kBTile, &tm_lo, bar, k0, n0 + 128 * bx);
        }
      };
      issue_b(0);
      for (uint32_t g = 0; g < total_slabs; ++g) {
        const uint32_t st = g % kStagesF;
        const uint32_t par = (g / kStagesF) & 1;
        const uint32_t lt = g / n_slabs, kt = g % n_slabs, rg = lt & 1;   // local tile, slab, TMEM region
        if (dbg_cta) t0 = clock64();
        mbar_wait(B.afull0 + 8 * st, par);
        if (dbg_cta) { const long long t1 = clock64(); t_a += t1 - t0; t0 = t1; }
        mbar_wait(B.bfull0 + 8 * st, par);
        if (dbg_cta) { const long long t1 = clock64(); t_b += t1 - t0; t0 = t1; }
        if (kt == 0 && lt >= 2) mbar_wait(B.rempty0 + 8 * rg, ((lt >> 1) - 1) & 1);
        if (dbg_cta) { const long long t1 = clock64(); t_r += t1 - t0; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          const uint32_t stage_base = smem + st * kStageF, d = tmem_d + rg * TNF;
          const uint32_t a_hi = stage_base, a_lo = stage_base + kATile;
          const uint32_t b_hi = stage_base + 2 * kATile, b_lo = b_hi + kBTileF;
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t dah = make_desc(a_hi + j * 32), dal = make_desc(a_lo + j * 32);
            const uint64_t dbh = make_desc(b_hi + j * 32), dbl = make_desc(b_lo + j * 32);
            mma_tf32(d, dal, dbh, idesc, (kt != 0 || j != 0) ? 1u : 0u);   // a tile starts a fresh accumulator
            mma_tf32(d, dah, dbl, idesc, 1);
            mma_tf32(d, dah, dbh, idesc, 1);
          }
        }
        mma_commit(B.free0 + 8 * st);
        if (kt == n_slabs - 1) mma_commit(B.rfull0 + 8 * rg);
        const uint32_t gs = g + 1;     // next slab's weights; its stage was read by slab g - 1
        if (gs < total_slabs) {
          if (dbg_cta) t0 = clock64();
          if (gs >= (uint32_t)kStagesF) mbar_wait(B.free0 + 8 * (gs % kStagesF), ((gs / kStagesF) - 1) & 1);
          if (dbg_cta) t_f += clock64() - t0;
          issue_b(gs);
        }
      }
      if (dbg_cta) {
        a.dbg[5] = t_a; a.dbg[6] = t_b; a.dbg[7] = t_r; a.dbg[8] = t_f; a.dbg[9] = clock64() - t_all;
      }
    }
  } else if (warp >= kConsumers / 32) {
    // ================= producers: activation slabs -> hi / lo TF32 tiles =================
    // warp pw owns tile rows [16 pw, 16 pw + 16); load i (0..3) covers rows 4 i + lane / 8, lane % 8 = k-chunk
    const int pw = warp - kConsumers / 32;
    const int prow = pw * 16 + (lane >> 3), pc = lane & 7;
    const bool dbg_on = dbg_cta && tid == kConsumers;
    long long t_wait = 0, t_work = 0, t0 = 0;
    auto load_a = [&](float4 (&v)[4], uint32_t gs) {
      if (gs >= total_slabs) return;
      const int tile = (int)blockIdx.y + (int)(gs / n_slabs) * (int)gridDim.y;
      const int k = (int)(gs % n_slabs) * BK + 4 * pc;
      const float* __restrict__ src = a.A + (size_t)(tile * BM + prow) * a.lda + k;
      const int mleft = M - (tile * BM + prow);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        v[i] = (4 * i < mleft && k < K) ? ldcg4(src + (size_t)(4 * i) * a.lda) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto produce = [&](float4 (&v)[4], uint32_t g) {
      const uint32_t st = g % kStagesF;
      const uint32_t stage_base = smem + st * kStageF;
      if (dbg_on) t0 = clock64();
      if (g >= (uint32_t)kStagesF) mbar_wait_relaxed(B.free0 + 8 * st, ((g / kStagesF) - 1) & 1);
      if (dbg_on) { const long long t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = prow + 4 * i;
        const uint32_t off = (uint32_t)(r * 128 + ((pc ^ (r & 7)) * 16));
        split_store(stage_base + off, stage_base + kATile + off, v[i]);
      }
      load_a(v, g + 2);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(B.afull0 + 8 * st);   // one arrival per producer warp
      if (dbg_on) t_work += clock64() - t0;
    };
    float4 v0[4], v1[4];
    load_a(v0, 0);
    load_a(v1, 1);
    for (uint32_t g = 0; g < total_slabs; g += 2) {
      produce(v0, g);
      if (g + 1 < total_slabs) produce(v1, g + 1);
    }
    if (dbg_on) { a.dbg[3] = t_wait; a.dbg[4] = t_work; }
  } else {
    // ================= consumers: epilogue of tile t while tile t + 1 multiplies =================
    const int quarter = warp & 3, col_half = (warp >> 2) * (TNF / 2);
    const uint32_t tmem_row = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col_half;
    const uint32_t stg = smem + kStagingOff + (uint32_t)warp * kStagingWarp;
    const uint32_t bias_s = smem + kBiasOff;
    long long t_main = 0, t_epi = 0, t0 = 0;
    const bool dbg_on = dbg_cta && tid == 0;
    uint32_t lt = 0;
    for (int tile = blockIdx.y; tile < n_tiles; tile += gridDim.y, ++lt) {
      const int m0 = tile * BM;
      const uint32_t rg = lt & 1;
      if (dbg_on) t0 = clock64();
      mbar_wait_relaxed(B.rfull0 + 8 * rg, (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (dbg_on) { const long long t1 = clock64(); t_main += t1 - t0; t0 = t1; }
      const int mr0 = m0 + quarter * 32;
      const int m = mr0 + lane;
#pragma unroll 1
      for (int q = 0; q < TNF / 2 / 32; ++q) {
        const int c0 = col_half + 32 * q;
        if (c0 < BN) {
          float v[32];
          {
            uint32_t r0[16], r1[16];
            tmem_ld16(tmem_row + rg * TNF + 32 * q, r0);
            tmem_ld16(tmem_row + rg * TNF + 32 * q + 16, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = __uint_as_float(r0[j]); v[16 + j] = __uint_as_float(r1[j]); }
          }
          const int word = (n0 + c0) >> 5;
          uint32_t in_bits = 0xffffffffu, out_bits = 0;
          if (a.bits_in && m < M) in_bits = a.bits_in[(size_t)m * 8 + word];
          const uint32_t baddr = bias_s + 4 * c0;
          const bool hb = a.bias != nullptr;
          if (a.relu) out_bits = epilogue_block<true, false>(v, baddr, hb, in_bits);
          else if (a.bits_in) out_bits = epilogue_block<false, true>(v, baddr, hb, in_bits);
          else out_bits = epilogue_block<false, false>(v, baddr, hb, in_bits);
          if (a.bits_out && m < M) a.bits_out[(size_t)m * 8 + word] = out_bits;
          if (a.c_tma && mr0 + 32 <= M) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              st_shared_v4(stg + (uint32_t)(lane * 128 + ((j4 ^ (lane & 7)) * 16)),
                           make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tma_store_2d(&tm_c, stg, n0 + c0, mr0);
          } else if (m < M) {
            float* __restrict__ cp = a.C + (size_t)m * a.ldc + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < a.N) cp[j] = v[j];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(B.rempty0 + 8 * rg);                 // tile t + 2 may overwrite the region
      if (dbg_on) { const long long t1 = clock64(); t_epi += t1 - t0; }
    }
    if (lane == 0) tma_store_wait_all();
    if (dbg_on) { a.dbg[0] = t_main; a.dbg[1] = t_epi; a.dbg[2] = 0; }
  }
  cta_teardown(tmem_d, warp);
}

// ---- fused MLP chain: all layers of one direction in ONE persistent launch ------------------------------
// A layer's 128-row output tile depends only on the same 128 rows of the previous layer, so a CTA can
// walk a row tile through the whole chain (forward: X0 -> H0 -> H1 -> H2 -> XR -> H3 -> H4 -> H5 -> logit;
// backward-data: dlogit -> dZ6 -> ... -> dX0) without any grid-wide dependency.  Per launch of gemm_wf_kernel
// about 10 us are not tensor-core time (launch, TMEM / barrier setup, pipeline fill, the last epilogue);
// 16 such launches per step are replaced by 2.  Work items = (step, local tile), step-major over the CTA's
// tiles, so the epilogue + TMA store of (tile, step) overlaps the MMAs of the CTA's other tile(s) before
// (tile, step + 1) needs it; completion travels through one mbarrier per local tile (c_done: the storing
// lanes arrive after cp.async.bulk.wait_group 0), the dependent loads are ld.global.cg.
constexpr int kChainMaxSteps = 10;
constexpr int kChainMaxTiles = 16;                         // local tiles per CTA (c_done barriers)
constexpr uint32_t kCDoneOff = kBiasOff + 2 * TNF * 4;     // bias slices are double-buffered by item parity
constexpr uint32_t kSmemBytesChain = kCDoneOff + 8 * kChainMaxTiles;

struct ChainStep {
  const float* A; float* C; const float* bias; const uint32_t* bits_in; uint32_t* bits_out;
  int lda, ldc, K, N, n0, relu, c_tma, pad_;
};
struct ChainArgs {
  ChainStep s[kChainMaxSteps];
  const int* rows_dev;
  long long* dbg;
  int n_steps, M;
  int row_off;     // this launch covers rows [row_off, row_off + M) of the problem (device-side count: *rows_dev - row_off)
};
struct ChainMaps {
  CUtensorMap hi[kChainMaxSteps], lo[kChainMaxSteps], c[kChainMaxSteps];
};

// ---- the same chain on CTA PAIRS (tcgen05 cta_group::2): M = 256 per MMA, each SM holds half of the weights -----
// In the single-CTA chain every SM streams the whole 64 KB weight slab per 32 k and its tensor core reads 12 KB
// of operands per MMA through one shared-memory port (measured: 178-200 cycles per M128 N256 K8 MMA against 138
// when less is read, and ~600 cycles per slab waiting for the weight TMA with room for only two stages).  A CTA
// pair multiplies a 256-row tile with ONE MMA stream issued by the leader CTA: each CTA contributes its own 128
// activation rows (A) and HALF of the weight rows (B: 128 of 256), so per SM the weight traffic, the weight
// footprint (32 KB per stage -> three stages fit) and the operand reads per MMA (8 KB) all drop.
// Cross-CTA signalling: tcgen05.commit multicasts `free` / `r_full` to both CTAs; the follower's producers and
// consumers arrive on the LEADER's a_full / r_empty barriers (mapa + mbarrier.arrive.release.cluster), and its half of
// the weight slab completes its bytes directly on the leader's b_full (cp.async.bulk.tensor .cta_group::2).
// Measured (tools/chain_phase_probe.py, M2 = 42.5 K): the pair issues an M256 N256 K8 MMA every ~108 cycles (178 for
// the single-CTA M128 MMA), b_full waits drop from 620 to 210 cycles per slab; what remains is the follower's
// a_full round trip (commit multicast -> fill -> remote arrive, ~550 cycles per slab with three stages).  Everything else (producers, epilogue, per-tile completion barriers) is
// per CTA and identical to mlp_chain_kernel.
constexpr int kStages2 = 3;
constexpr uint32_t kBHalf2 = 128 * BK * 4;                       // 16 KB: one plane of this CTA's half of a slab
constexpr uint32_t kStage2 = 2 * kATile + 2 * kBHalf2;           // A_hi, A_lo, B_hi, B_lo = 64 KB
static_assert(kStages2 * kStage2 == kStagingOff, "staging / barrier offsets are shared");
constexpr uint32_t kSmemBytesChain2 = kCDoneOff + 8 * kChainMaxTiles;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  if (!done) __trap();
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool BF>
__device__ __forceinline__ void mma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (BF) mma2_bf16(tmem_d, da, db, idesc, accumulate);
  else mma2_tf32(tmem_d, da, db, idesc, accumulate);
}
// instruction descriptor: fp32 accumulator, A / B format TF32 (2) or BF16 (1), N, M
__device__ __forceinline__ uint32_t make_idesc(bool bf, int N, int M, bool mn_major) {
  const uint32_t fmt = bf ? 1u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma2_commit(uint32_t bar) {    // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3)
               : "memory");
}

// (The forward chain: TF32 hi / lo planes, kind::tf32, 32-k slabs.  The backward-data chain is mlp_bwd_planes_kernel.)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsF, 1)
mlp_chain2_kernel(const __grid_constant__ ChainArgs a, const __grid_constant__ ChainMaps maps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  const Bars B = make_bars(smem);
  const uint32_t cdone0 = smem + kCDoneOff;
  constexpr int SK = BK;                                       // k per slab (one 128-byte swizzle row per operand row)
  constexpr bool BF = false;

  int M = a.M;
  if (a.rows_dev) M = min(M, *a.rows_dev - a.row_off);
  if (M <= 0) return;
  const int n_pair_tiles = (M + 2 * BM - 1) / (2 * BM);        // 256-row tiles
  const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
  if (pair >= n_pair_tiles) return;                            // both CTAs of the pair leave together
  const uint32_t rank = cluster_ctarank();                     // 0 = leader (issues the MMAs)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- set-up: barriers (the leader's a_full / b_full / r_empty also count the follower's arrivals), TMEM ----
  // a_full: every CTA's producers arrive on their OWN barrier (CTA scope); the follower's otherwise idle issuer
  // thread forwards each completed phase to the leader.  A cluster-scope release issued by a producer itself waits
  // for that thread's prefetch loads in flight -- one memory latency per slab.
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + kTmemSlotOff);
  if (tid == 0) {
    for (int i = 0; i < kChainMaxTiles; ++i) mbar_init(cdone0 + 8 * i, 8);
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(B.free0 + 8 * i, 1);                                     // one multicast commit
      mbar_init(B.bfull0 + 8 * i, 1);                                    // the leader's expect_tx arrive; bytes from both CTAs
      mbar_init(B.afull0 + 8 * i, kProducersF / 32 + (rank == 0 ? 1 : 0));   // own producer warps (+ the follower's relay)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(B.rfull0 + 8 * i, 1);
      mbar_init(B.rempty0 + 8 * i, 2 * (kConsumers / 32));               // consumer warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();                                          // barriers of both CTAs exist before anyone signals
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  const int my_tiles = (n_pair_tiles - 1 - pair) / n_pairs + 1;
  const int n_items = a.n_steps * my_tiles;                    // item w: step w / my_tiles, local tile w % my_tiles
  const bool dbg_cta = a.dbg != nullptr && blockIdx.x == 0;
  // leader-side addresses of the barriers the follower signals
  const uint32_t l_afull0 = map_to_cta(B.afull0, 0), l_bfull0 = map_to_cta(B.bfull0, 0), l_rempty0 = map_to_cta(B.rempty0, 0);
  auto tile_row0 = [&](int lt) { return ((pair + lt * n_pairs) * 2 + (int)rank) * BM; };   // this CTA's 128 rows

  if (warp >= kIssuerWarpF) {
    if (warp == kIssuerWarpF && lane == 0) {
      // both CTAs: TMA of THIS CTA's half of the weight rows; leader: also every MMA
      long long t_a = 0, t_b = 0, t_r = 0, t_f = 0, t_all = 0, t0 = 0;
      if (dbg_cta) t_all = clock64();
      struct Cursor { int w, step, kt, n_slabs; };
      auto cur_next = [&](Cursor& c) {
        if (++c.kt == c.n_slabs) {
          c.kt = 0;
          ++c.w;
          if (c.w < n_items) { c.step = c.w / my_tiles; c.n_slabs = (a.s[c.step].K + SK - 1) / SK; }
        }
      };
      auto issue_b = [&](const Cursor& c, uint32_t gs) {
        const ChainStep& S = a.s[c.step];
        const int BN = min(TNF, (S.N - S.n0 + 15) & ~15);
        const uint32_t st = gs % kStages2;
        const uint32_t b_hi = smem + st * kStage2 + 2 * kATile, b_lo = b_hi + kBHalf2;
        // both halves complete their bytes on the LEADER's b_full (cta_group::2 form of the bulk tensor copy: the
        // mbarrier may live in the peer CTA), which the leader arms for the whole 64 KB
        if (rank == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(B.bfull0 + 8 * st), "r"(4 * kBHalf2) : "memory");
        const uint32_t bar = l_bfull0 + 8 * st;
        const int row0 = S.n0 + (int)rank * (BN >> 1);          // B is split along N between the two CTAs
        tma2_load_2d(b_hi, &maps.hi[c.step], bar, c.kt * SK, row0);
        tma2_load_2d(b_lo, &maps.lo[c.step], bar, c.kt * SK, row0);
      };
      Cursor lc{0, 0, 0, (a.s[0].K + SK - 1) / SK};
      uint32_t lg = 0;
      constexpr uint32_t kPf = kStages2 - 1;                   // weight slabs in flight ahead of the MMAs
      for (; lg < kPf && lc.w < n_items; ++lg) { issue_b(lc, lg); cur_next(lc); }
      uint32_t g = 0;
      for (int w = 0; w < n_items; ++w) {
        const int step = w / my_tiles;
        const ChainStep& S = a.s[step];
        const int BN = min(TNF, (S.N - S.n0 + 15) & ~15);
        const int n_slabs = (S.K + SK - 1) / SK;
        const uint32_t idesc = make_idesc(BF, BN, 2 * BM, false);          // M = 256 across the pair
        const uint32_t rg = (uint32_t)w & 1;
        for (int kt = 0; kt < n_slabs; ++kt, ++g) {
          const uint32_t st = g % kStages2;
          const uint32_t par = (g / kStages2) & 1;
          if (rank == 0) {
            if (dbg_cta) t0 = clock64();
            mbar_wait_cluster(B.afull0 + 8 * st, par);
            if (dbg_cta) { const long long t1 = clock64(); t_a += t1 - t0; t0 = t1; }
            mbar_wait_cluster(B.bfull0 + 8 * st, par);
            if (dbg_cta) { const long long t1 = clock64(); t_b += t1 - t0; t0 = t1; }
            if (kt == 0 && w >= 2) mbar_wait_cluster(B.rempty0 + 8 * rg, (((uint32_t)w >> 1) - 1) & 1);
            if (dbg_cta) { const long long t1 = clock64(); t_r += t1 - t0; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t stage_base = smem + st * kStage2, d = tmem_d + rg * TNF;
            const uint32_t a_hi = stage_base, a_lo = stage_base + kATile;
            const uint32_t b_hi = stage_base + 2 * kATile, b_lo = b_hi + kBHalf2;
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
              const uint64_t dah = make_desc(a_hi + j * 32), dal = make_desc(a_lo + j * 32);
              const uint64_t dbh = make_desc(b_hi + j * 32), dbl = make_desc(b_lo + j * 32);
              mma2_tf32(d, dal, dbh, idesc, (kt != 0 || j != 0) ? 1u : 0u);   // (32 bytes per row per k-step either way)
              mma2_tf32(d, dah, dbl, idesc, 1);
              mma2_tf32(d, dah, dbh, idesc, 1);
            }
            mma2_commit(B.free0 + 8 * st);
            if (kt == n_slabs - 1) mma2_commit(B.rfull0 + 8 * rg);
          } else {
            mbar_wait(B.afull0 + 8 * st, par);                   // this CTA's activation slab g is in place
            mbar_arrive_cluster(l_afull0 + 8 * st);              // -> tell the leader
          }
          // next weight slab of this CTA's half (its stage was last read by slab lg - kStages2)
          if (lc.w < n_items) {
            if (dbg_cta) t0 = clock64();
            if (lg >= (uint32_t)kStages2) mbar_wait(B.free0 + 8 * (lg % kStages2), ((lg / kStages2) - 1) & 1);
            if (dbg_cta) t_f += clock64() - t0;
            issue_b(lc, lg);
            cur_next(lc);
            ++lg;
          }
        }
      }
      if (dbg_cta) {
        a.dbg[5] = t_a; a.dbg[6] = t_b; a.dbg[7] = t_r; a.dbg[8] = t_f; a.dbg[9] = clock64() - t_all;
      }
    }
  } else if (warp >= kConsumers / 32) {
    // ================= producers: this CTA's 128 activation rows -> hi / lo operand tiles =================
    const int pw = warp - kConsumers / 32;
    const bool dbg_on = dbg_cta && tid == kConsumers;
    long long t_wait = 0, t_work = 0, t_dep = 0, t0 = 0;
    uint32_t g = 0;
    for (int w = 0; w < n_items; ++w) {
      const int step = w / my_tiles, lt = w - step * my_tiles;
      const ChainStep& S = a.s[step];
      const int K = S.K, lda = S.lda;
      const int n_slabs = (K + SK - 1) / SK;
      if (step > 0) {
        if (dbg_on) t0 = clock64();
        mbar_wait_relaxed(cdone0 + 8 * lt, (uint32_t)(step - 1) & 1);
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (dbg_on) t_dep += clock64() - t0;
      }
      // warp pw owns tile rows [16 pw, 16 pw + 16); load i (0..3) covers rows 4 i + lane / 8, lane % 8 = 16-byte k-chunk
      const int prow = pw * 16 + (lane >> 3), pc = lane & 7;
      const int row = tile_row0(lt) + prow;
      const float* __restrict__ src = S.A + (size_t)row * lda + 4 * pc;
      const int mleft = M - row;
      auto load_a = [&](float4 (&v)[4], int kt) {
        if (kt >= n_slabs) return;
        const int k = kt * BK + 4 * pc;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          v[i] = (4 * i < mleft && k < K) ? ldcg4(src + (size_t)(4 * i) * lda + kt * BK) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      auto produce = [&](float4 (&v)[4], int kt) {
        const uint32_t st = g % kStages2;
        const uint32_t stage_base = smem + st * kStage2;
        if (dbg_on) t0 = clock64();
        if (g >= (uint32_t)kStages2) mbar_wait_relaxed(B.free0 + 8 * st, ((g / kStages2) - 1) & 1);
        if (dbg_on) { const long long t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = prow + 4 * i;
          const uint32_t off = (uint32_t)(r * 128 + ((pc ^ (r & 7)) * 16));
          split_store(stage_base + off, stage_base + kATile + off, v[i]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(B.afull0 + 8 * st);            // own CTA's barrier (the follower's is relayed)
        load_a(v, kt + 2);
        ++g;
        if (dbg_on) t_work += clock64() - t0;
      };
      float4 v0[4], v1[4];
      load_a(v0, 0);
      load_a(v1, 1);
      for (int kt = 0; kt < n_slabs; kt += 2) {
        produce(v0, kt);
        if (kt + 1 < n_slabs) produce(v1, kt + 1);
      }
    }
    if (dbg_on) { a.dbg[2] = t_dep; a.dbg[3] = t_wait; a.dbg[4] = t_work; }
  } else {
    // ================= consumers: epilogue of this CTA's 128 rows of item w =================
    const int quarter = warp & 3, col_half = (warp >> 2) * (TNF / 2);
    const uint32_t tmem_row = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col_half;
    const uint32_t stg = smem + kStagingOff + (uint32_t)warp * kStagingWarp;
    long long t_main = 0, t_epi = 0, t0 = 0, t_bias = 0, t_tm = 0, t_ma = 0, t_st = 0, t_fe = 0, tq = 0;
    const bool dbg_on = dbg_cta && tid == 0;
    for (int w = 0; w < n_items; ++w) {
      const int step = w / my_tiles, lt = w - step * my_tiles;
      const ChainStep& S = a.s[step];
      const int n0 = S.n0, N = S.N;
      if (dbg_on) tq = clock64();
      const int BN = min(TNF, (N - n0 + 15) & ~15);
      const uint32_t rg = (uint32_t)w & 1;
      const uint32_t bias_s = smem + kBiasOff + rg * (TNF * 4);
      reinterpret_cast<float*>(smem_raw + kBiasOff)[rg * TNF + tid] = (S.bias && n0 + tid < N) ? S.bias[n0 + tid] : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (dbg_on) { t0 = clock64(); t_bias += t0 - tq; }
      const int m0 = tile_row0(lt);
      const int mr0 = m0 + quarter * 32;
      const int m = mr0 + lane;
      // the ReLU-mask words of this thread's row (backward-data steps): in flight while the accumulator is still being
      // computed (loaded inside the block loop they cost one L2 round trip per block: 44 of the epilogue's 190 us)
      uint32_t mask_w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (S.bits_in && m < M) {
        const uint4 mw = *reinterpret_cast<const uint4*>(S.bits_in + (size_t)m * 8 + ((n0 + col_half) >> 5));
        mask_w[0] = mw.x; mask_w[1] = mw.y; mask_w[2] = mw.z; mask_w[3] = mw.w;
      }
      mbar_wait_relaxed(B.rfull0 + 8 * rg, ((uint32_t)w >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (dbg_on) { const long long t1 = clock64(); t_main += t1 - t0; t0 = t1; }
      // The 4 column blocks of this warp, in halves of 16 columns: the TMEM load of half h + 1 is in flight while
      // half h goes through the arithmetic and into the staging tile (tcgen05.ld needs several hundred cycles while the
      // tensor core accumulates into the other region; issued and awaited block by block that latency was exposed four
      // times per item), and the region goes back to the issuer as soon as the LAST load has landed, not after the
      // last store.  (Two whole 32-column blocks in flight do not fit the 96-register budget: 2 KB of spills.)
      const int nblk = max(0, min(4, (BN - col_half + 31) >> 5));
      const bool hb = S.bias != nullptr;
      const int relu = S.relu;
      const bool has_mask = S.bits_in != nullptr;
      const bool tma_out = S.c_tma && mr0 + 32 <= M;
      auto release_region = [&]() {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {                                        // the leader may overwrite the region in BOTH CTAs
          if (rank == 0) mbar_arrive(B.rempty0 + 8 * rg);
          else mbar_arrive_cluster(l_rempty0 + 8 * rg);
        }
      };
      // half hf (0..7) = columns col_half + 16 hf .. + 16 of the item; returns its 16 mask bits (ReLU forward)
      auto process_half = [&](uint32_t (&r)[16], int hf, uint32_t in_bits16) -> uint32_t {
        const int c0 = col_half + 16 * hf;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        uint32_t out_bits = 0;
        const uint32_t baddr = bias_s + 4 * c0;
        if (relu) out_bits = epilogue_block<true, false, 16>(v, baddr, hb, in_bits16);
        else if (has_mask) out_bits = epilogue_block<false, true, 16>(v, baddr, hb, in_bits16);
        else out_bits = epilogue_block<false, false, 16>(v, baddr, hb, in_bits16);
        if (tma_out) {
          if ((hf & 1) == 0) {                                  // first half of a 32-column block: the tile must be free
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            st_shared_v4(stg + (uint32_t)(lane * 128 + (((4 * (hf & 1) + j4) ^ (lane & 7)) * 16)),
                         make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
          if (hf & 1) {
            // (TMA store: the copy engine moves the 32 x 32 block.  A version that read the staging tile back and stored
            // it with 128-bit st.global was slower, 0.230 -> 0.266 ms for the backward chain.)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tma_store_2d(&maps.c[step], stg, n0 + c0 - 16, mr0);
          }
        } else if (m < M) {
          float* __restrict__ cp = S.C + (size_t)m * S.ldc + n0 + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < N) cp[j] = v[j];
        }
        return out_bits;
      };
      {
        uint32_t ra[16], rb[16];
        const int nhalf = 2 * nblk;
        if (nhalf > 0) tmem_ld16(tmem_row + rg * TNF, ra);
        else release_region();
        uint32_t bits_lo = 0;
#pragma unroll
        for (int hf = 0; hf < 8; ++hf) {
          if (hf < nhalf) {
            tmem_ld_wait();                                     // half hf is in registers
            if (hf + 1 < nhalf) {
              if (hf & 1) tmem_ld16(tmem_row + rg * TNF + 16 * (hf + 1), ra);
              else tmem_ld16(tmem_row + rg * TNF + 16 * (hf + 1), rb);
            } else {
              release_region();
            }
            const uint32_t mw = mask_w[hf >> 1] >> (16 * (hf & 1));
            const uint32_t ob = (hf & 1) ? process_half(rb, hf, mw) : process_half(ra, hf, mw);
            if (hf & 1) {
              if (S.bits_out && m < M) S.bits_out[(size_t)m * 8 + ((n0 + col_half) >> 5) + (hf >> 1)] = bits_lo | (ob << 16);
            } else {
              bits_lo = ob;
            }
          }
        }
      }
      if (dbg_on) { const long long t1 = clock64(); t_st += t1 - t0 - 0; tq = t1; }
      // the tile's rows of this step are complete in global memory -> (tile, step + 1) may load them
      __syncwarp();
      if (lane == 0) {
        tma_store_wait_all();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __threadfence();
        mbar_arrive(cdone0 + 8 * lt);
      }
      if (dbg_on) { const long long t1 = clock64(); t_epi += t1 - t0; t_fe += t1 - tq; }
    }
    if (dbg_on) {
      a.dbg[0] = t_main; a.dbg[1] = t_epi;
      // finer split of the epilogue (slots 20..24): bias fill + consumer barrier, TMEM loads, bits + arithmetic,
      // whole q loop incl. stores (from r_full), fence + arrivals
      a.dbg[20] = t_bias; a.dbg[21] = t_tm; a.dbg[22] = t_ma; a.dbg[23] = t_st; a.dbg[24] = t_fe;
    }
  }
  // ---- teardown: nobody may leave (or free TMEM) while the partner can still signal it or multiply into it ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
  }
}

// ---- backward-weight GEMMs:  C_l[M_l,N_l] += A_l[K,M_l]^T B_l[K,N_l]  (dW_l = dZ_l^T X_l) for ALL layers in one launch --
// The reduction index is the SAMPLE dimension, the slow dimension of both row-major operands, i.e. both tiles
// are MN-major.  kind::tf32 reads MN-major tiles in the SWIZZLE_128B_BASE32B layout (tools/umma_probe.cu):
//   addr(mn, k) = (mn / 32) * LBO + (k / 4) * SBO + (k % 4) * 128 + (((mn % 32) / 8) ^ (k % 4)) * 32 + (mn % 8) * 4
// With LBO = 4096, SBO = 512 a slab is four blocks of [32 samples][32 features], one 128-byte row per sample:
// a producer warp reads one sample row of 128 features with one coalesced 128-bit load per lane, splits it
// into hi / lo and stores 16 bytes per lane -- no transposition anywhere (an earlier version transposed with
// 32-bit loads and ran at 2.3x the tensor-core time).  The bias gradient (column sums of dZ) rides along.
// Work is the flattened (tile, 32-sample slab) space of all 8 layers' 128 x 128 output tiles, cut into equal
// contiguous ranges, one per CTA (stream-K): every SM gets the same number of MMAs, a CTA crosses at most a
// few tile boundaries, each costing one 128 x 128 red.global.add epilogue.  Accumulation: 32-k chunks in
// rotating TMEM regions, fp32 register accumulators (thousands of MMAs per range -- the TMEM accumulator's
// per-MMA truncation would reach 1e-4 if it took them all).
constexpr int kDwMaxProblems = 8, kDwMaxTiles = 32;
struct DwProblem {
  const float* A; const float* B; float* C; float* colsum;
  int M, N, lda, ldb, ldc, pad_;
};
struct DwArgs {
  DwProblem p[kDwMaxProblems];
  unsigned char tile_prob[kDwMaxTiles], tile_m[kDwMaxTiles], tile_n[kDwMaxTiles];
  const int* rows_dev;
  long long* dbg;
  int n_tiles, K;
};
constexpr int kProducersTn = 256;
constexpr int kIssuerWarpTn = (kConsumers + kProducersTn) / 32;
constexpr int kThreadsTn = kConsumers + kProducersTn + 32;

// MN-major SWIZZLE_128B_BASE32B descriptor (layout_type 1): LBO 4096 between 32-wide MN blocks, SBO 512
// between groups of 4 k-rows
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
         (1ull << 61);
}

// ---- the same on CTA pairs (cta_group::2): a pair owns a 256 (all out-features) x 128 output tile ----------------
// Each CTA produces the MN-major tile of ITS 128 out-features of dZ (A) and of ITS half of the tile's input features
// of X (B): per slab 48 KB of st.shared per SM instead of 64 KB, for twice the output, and one M256 N128 K8 MMA
// stream issued by the leader (~92 cycles per MMA).  Signalling as in mlp_chain2_kernel, four 48-KB stages.
// Measured: 230 -> 187 us per step once the follower's a_full went through the relay thread (with its producers
// arriving on the leader's barrier themselves -- a cluster-scope release behind their prefetch loads -- it was
// SLOWER than the single-CTA kernel: 259 us).  At 187 us the kernel reads its ~1 GB of operands at 5.6 TB/s.
// stage = A_hi, A_lo (16 KB each: this CTA's 128 out-features) + B_hi, B_lo (8 KB each: its <= 64 input features):
// 48 KB, so FOUR stages fit -- the follower's producers sit on a longer loop than in the single-CTA kernel
constexpr int kStagesD2 = 4;
constexpr uint32_t kBHalfD2 = 64 * BK * 4;                       // 8 KB
constexpr uint32_t kStageD2 = 2 * kATile + 2 * kBHalfD2;         // 48 KB
static_assert(kStagesD2 * kStageD2 == kStagingOff, "staging / barrier offsets are shared");
constexpr uint32_t kBarD2Off = kCDoneOff;                        // its own barrier block (4 stages do not fit Bars)
constexpr uint32_t kSmemBytesDw2 = kBarD2Off + 8 * (2 * kStagesD2 + 2 * kRegions);

constexpr int kDwMaxPairs = FGS_NUM_SMS / 2;
struct Dw2Args {
  DwProblem p[kDwMaxProblems];
  unsigned char tile_prob[kDwMaxTiles], tile_n[kDwMaxTiles];
  // aligned != 0: pair i works on ONE tile, pair_tile[i], and owns the pair_rank[i]-th of pair_q[i] equal slab ranges of
  // it.  All tiles of a problem get the same q, so the pairs that read the same rows of dZ_l (one per 128-wide block
  // of input features) walk them at the same time and the second reader finds them in L2: with the tile-major
  // stream-K cut (aligned == 0, kept for launches with fewer pairs than tiles) every pair sat in a different layer
  // and dZ_l came from DRAM once per block (1.05 GB per launch against 0.64 GB of operands, L2 hit rate 16 %).
  unsigned char pair_tile[kDwMaxPairs], pair_rank[kDwMaxPairs], pair_q[kDwMaxPairs];
  const int* rows_dev;
  long long* dbg;
  int n_tiles, K, aligned;
};

// MN-major SWIZZLE_128B descriptor for 16-bit elements (the canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in units of
// 16 bytes): 64 MN elements = one 128-byte row per k, 8 k-rows = one 1024-byte swizzle atom (chunk ^ k % 8), SBO 1024
// between 8-k groups, LBO 8192 between 64-wide MN blocks (64 k-rows of a slab each); layout_type 2.
__device__ __forceinline__ uint64_t make_desc_mn16(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

// BF = false: TF32 hi / lo tiles, 32-sample slabs (kept for the stand-alone fp32-grade entry fgs_gemm_tn).
// BF = true: bf16 hi / mid tiles, 64-sample slabs in the same bytes, kind::f16: the weight gradients of the fused step.
// bf16 variant (the fused step): NO conversion of the dZ operand and no register staging.
//   A (dZ_l, 2/3 of the operand elements) is read as the bf16 hi / mid planes that the backward chain's epilogue wrote:
//     per slab and CTA four bulk-tensor loads (2 planes x 2 blocks of 64 out-features x 64 samples) land directly in
//     the MN-major SWIZZLE_128B operand tile -- a row of the plane IS a row of the tile;
//   B (X_l, fp32 activations of the forward) still has to be split, by the 8 producer warps, but reaches them through
//     a per-warp TMA ring (4 sample rows x 64 features = 1 KB per unit, kRawSlots units in flight, the warp's own
//     mbarriers, no registers, no cross-warp synchronisation).
// History: with both operands converted from registers the issuer waited for operands 60 % of the time; a ring for both
// (96 KB in flight instead of 48 KB) changed nothing -- the conversion (5 integer ops per element, 12 K elements per
// slab), not the loads, was the bound (ncu: issue slots 53 % busy with 17 warps, tensor pipe 22 %).
struct DwMaps {
  CUtensorMap a_hi[kDwMaxProblems], a_mid[kDwMaxProblems];   // bf16 planes [samples][features]: box (64 features, 64 samples), SWIZZLE_128B
  CUtensorMap b[kDwMaxProblems];                             // fp32 [samples][features]: box (64 features, 8 samples), no swizzle
};
constexpr int kRawSlots = 3;
constexpr uint32_t kRawUnit = 8 * 64 * 4;                                                      // 2 KB: 8 samples x 64 features
constexpr int kStagesD2B = 3;
constexpr uint32_t kRawOff = kStagesD2B * kStageD2;                                             // 144 KB
static_assert(kRawOff + 8 * kRawSlots * kRawUnit <= kStagingOff, "raw ring overlaps the epilogue staging");
constexpr uint32_t kRawBarOff = kBarD2Off + 8 * (2 * kStagesD2 + 2 * kRegions);                 // 8 warps x kRawSlots barriers
constexpr uint32_t kSmemBytesDw2B = kRawBarOff + 8 * 8 * kRawSlots;
static_assert(kSmemBytesDw2B <= 232448, "shared memory budget");

template <bool BF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsTn, 1)
mlp_dw2_kernel(const __grid_constant__ Dw2Args a, const __grid_constant__ DwMaps maps) {
  constexpr int NST = BF ? kStagesD2B : kStagesD2;              // operand-tile stages
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  Bars B;                                                       // free[4], afull[4], rfull[4], rempty[4]
  B.free0 = smem + kBarD2Off;
  B.afull0 = B.free0 + 8 * kStagesD2;          // (barrier block laid out for the larger stage count)
  B.rfull0 = B.afull0 + 8 * kStagesD2;
  B.rempty0 = B.rfull0 + 8 * kRegions;
  B.bfull0 = 0;

  int K = a.K;
  if (a.rows_dev) K = min(K, *a.rows_dev);
  constexpr int SK = BF ? 2 * BK : BK;                          // samples per slab
  const int n_slabs = (K + SK - 1) / SK;
  const long long units = (long long)a.n_tiles * n_slabs;
  const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
  int u_beg, u_end;
  if (a.aligned) {
    const int t = a.pair_tile[pair], r = a.pair_rank[pair], q = a.pair_q[pair];
    u_beg = t * n_slabs + (int)((long long)n_slabs * r / q);
    u_end = t * n_slabs + (int)((long long)n_slabs * (r + 1) / q);
  } else {
    u_beg = (int)(units * pair / n_pairs);
    u_end = (int)(units * (pair + 1) / n_pairs);
  }
  if (u_beg >= u_end) return;                                  // both CTAs of the pair
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + kTmemSlotOff);
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(B.free0 + 8 * i, 1);
      mbar_init(B.afull0 + 8 * i, kProducersTn / 32 + (rank == 0 ? (BF ? 2 : 1) : 0));   // own producer warps (+ the follower's relay, + the leader's expect_tx of the A planes)
    }
    if (BF)
      for (int i = 0; i < 8 * kRawSlots; ++i) mbar_init(smem + kRawBarOff + 8 * i, 1);   // one expect_tx arrival each
    for (int i = 0; i < kRegions; ++i) {
      mbar_init(B.rfull0 + 8 * i, 1);
      mbar_init(B.rempty0 + 8 * i, 2 * (kConsumers / 32));     // consumer warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  const bool dbg_cta = a.dbg != nullptr && blockIdx.x == 0;
  const uint32_t l_afull0 = map_to_cta(B.afull0, 0), l_rempty0 = map_to_cta(B.rempty0, 0);

  if (warp >= kIssuerWarpTn) {
    // A planes (bf16 variant): both CTAs load their own 128 out-features of slab g into stage g % NST, completing on the
    // LEADER's a_full (armed by the leader for both CTAs' 2 x 32 KB)
    long long t_ia = 0;
    auto issue_a = [&](uint32_t g, int u) {
      const long long ti0 = dbg_cta ? clock64() : 0;
      const int tile = u / n_slabs, s = u - tile * n_slabs;
      const int prob = a.tile_prob[tile];
      const uint32_t st = g % NST, base = smem + st * kStageD2;
      if (g >= (uint32_t)NST) mbar_wait(B.free0 + 8 * st, ((g / NST) - 1) & 1);
      if (rank == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(B.afull0 + 8 * st), "r"(4 * kATile) : "memory");
      const uint32_t bar = l_afull0 + 8 * st;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        tma2_load_2d(base + blk * 8192, &maps.a_hi[prob], bar, (int)rank * BM + 64 * blk, s * 64);
        tma2_load_2d(base + kATile + blk * 8192, &maps.a_mid[prob], bar, (int)rank * BM + 64 * blk, s * 64);
      }
      if (dbg_cta) t_ia += clock64() - ti0;
    };
    if (rank != 0 && warp == kIssuerWarpTn && lane == 0) {
      // follower: forward "my operand tiles of slab g are in place" to the leader (see mlp_chain2_kernel)
      const uint32_t total = (uint32_t)(u_end - u_beg);
      if (BF) {
        issue_a(0, u_beg);
        if (total > 1) issue_a(1, u_beg + 1);
      }
      for (uint32_t g = 0; g < total; ++g) {
        const uint32_t st = g % NST;
        mbar_wait(B.afull0 + 8 * st, (g / NST) & 1);
        mbar_arrive_cluster(l_afull0 + 8 * st);
        if (BF && g + 2 < total) issue_a(g + 2, u_beg + (int)g + 2);     // two slabs ahead (NST = 3 stages)
      }
    }
    if (rank == 0 && warp == kIssuerWarpTn && lane == 0) {
      uint32_t g = 0;
      long long t_a = 0, t_r = 0, t_all = 0, t0 = 0;
      if (dbg_cta) t_all = clock64();
      if (BF) {
        // the A planes run TWO slabs ahead of the MMAs (three stages): issued one slab ahead, right before waiting for
        // them, the whole load latency (~1 us) was exposed every slab -- 1.7 us per slab for 0.73 us of MMAs
        issue_a(0, u_beg);
        if (u_beg + 1 < u_end) issue_a(1, u_beg + 1);
      }
      for (int u = u_beg; u < u_end;) {
        const int tile = u / n_slabs, s0 = u - tile * n_slabs;
        const int len = min(n_slabs - s0, u_end - u);
        const DwProblem& P = a.p[a.tile_prob[tile]];
        const int BN = min(TN, (P.N - a.tile_n[tile] * TN + 15) & ~15);
        const uint32_t idesc = make_idesc(BF, BN, 2 * BM, true);                            // M = 256, MN-major
        for (int i = 0; i < len; ++i, ++g) {
          const uint32_t st = g % NST;
          if (dbg_cta) t0 = clock64();
          mbar_wait_cluster(B.afull0 + 8 * st, (g / NST) & 1);
          if (dbg_cta) { const long long t1 = clock64(); t_a += t1 - t0; t0 = t1; }
          if (g >= (uint32_t)kRegions) mbar_wait_cluster(B.rempty0 + 8 * (g % kRegions), ((g / kRegions) - 1) & 1);
          if (dbg_cta) t_r += clock64() - t0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t stage_base = smem + st * kStageD2, d = tmem_d + (g % kRegions) * TN;
          const uint32_t a_hi = stage_base, a_lo = stage_base + kATile;
          const uint32_t b_hi = stage_base + 2 * kATile, b_lo = b_hi + kBHalfD2;
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            // k-step j: TF32 = sample rows [8 j, 8 j + 8) of every block (1024 bytes on); bf16 = rows [16 j, 16 j + 16)
            // = two 8-row groups (2048 bytes on)
            const uint32_t ko = BF ? j * 2048 : j * 1024;
            const uint64_t dah = BF ? make_desc_mn16(a_hi + ko) : make_desc_mn(a_hi + ko);
            const uint64_t dal = BF ? make_desc_mn16(a_lo + ko) : make_desc_mn(a_lo + ko);
            const uint64_t dbh = BF ? make_desc_mn16(b_hi + ko) : make_desc_mn(b_hi + ko);
            const uint64_t dbl = BF ? make_desc_mn16(b_lo + ko) : make_desc_mn(b_lo + ko);
            mma2<BF>(d, dal, dbh, idesc, j != 0 ? 1u : 0u);
            mma2<BF>(d, dah, dbl, idesc, 1);
            mma2<BF>(d, dah, dbh, idesc, 1);
          }
          mma2_commit(B.free0 + 8 * st);
          mma2_commit(B.rfull0 + 8 * (g % kRegions));
          if (BF && u + i + 2 < u_end) issue_a(g + 2, u + i + 2);      // the A planes two slabs on (waits for MMA g - 1)
        }
        u += len;
      }
      if (dbg_cta) { a.dbg[5] = t_a; a.dbg[6] = t_ia; a.dbg[7] = t_r; a.dbg[8] = g; a.dbg[9] = clock64() - t_all; }
    }
  } else if (warp >= kConsumers / 32) {
    // ================= producers =================
    // A: warp pw owns sample rows 4 pw .. 4 pw + 3; lane l owns out-features 128 rank + 4 l .. + 3
    // B: this CTA's half (h = BN / 2 <= 64 input features): 16 lanes per sample row, two rows per load
    // bf16: the same loads per 32 samples; a 64-sample slab is produced as two halves (rows 32 h + ...), 8-byte stores:
    // lane l's 4 features sit in block l / 16 at chunk (l % 16) / 2 (^ row % 8), byte (l % 2) * 8 of the 128-byte row
    const int pw = warp - kConsumers / 32;
    const uint32_t a_blk = BF ? (uint32_t)(lane >> 4) * 8192u + (uint32_t)(lane & 1) * 8u
                              : (uint32_t)(lane >> 3) * 4096u + (uint32_t)(lane & 1) * 16u;
    const uint32_t a_c32 = BF ? (uint32_t)((lane & 15) >> 1) : (uint32_t)((lane & 7) >> 1);
    const int bq = lane & 15, bsub = lane >> 4;                 // feature quad and row-in-pair of the B loads
    const uint32_t b_blk = BF ? (uint32_t)(bq & 1) * 8u : (uint32_t)(bq >> 3) * 4096u + (uint32_t)(bq & 1) * 16u;
    const uint32_t b_c32 = BF ? (uint32_t)(bq >> 1) : (uint32_t)((bq & 7) >> 1);
    uint32_t g = 0;
    if constexpr (BF) {
      // ---- B only: per-warp TMA ring of raw fp32 rows of X -> bf16 hi / mid operand tile ----
      // unit = this warp's 8 sample rows of one 64-sample slab (2 KB): one barrier wait, one TMA issue, one proxy fence
      // and one arrival per SLAB (with 4-row half-slab units those fixed latencies, ~670 cycles per unit, made the
      // conversion of a slab as long as its 12 MMAs although only a third of the operand elements is converted here)
      const uint32_t raw0 = smem + kRawOff + (uint32_t)pw * (kRawSlots * kRawUnit);
      const uint32_t rbar0 = smem + kRawBarOff + (uint32_t)pw * (8 * kRawSlots);
      auto issue = [&](int u, uint32_t n) {                     // lane 0: one bulk-tensor load of unit n into slot n % kRawSlots
        if (u >= u_end) return;
        const int tile = u / n_slabs, sl = u - tile * n_slabs;
        const int prob = a.tile_prob[tile];
        const DwProblem& P = a.p[prob];
        const int n0 = a.tile_n[tile] * TN;
        const int h = min(TN, (P.N - n0 + 15) & ~15) >> 1;
        const uint32_t slot = n % kRawSlots, bar = rbar0 + 8 * slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kRawUnit) : "memory");
        tma_load_2d(raw0 + slot * kRawUnit, &maps.b[prob], bar, n0 + (int)rank * h, sl * 64 + 8 * pw);
      };
      if (lane == 0)
        for (uint32_t n_iss = 0; n_iss < (uint32_t)kRawSlots; ++n_iss) issue(u_beg + (int)n_iss, n_iss);
      const bool dbg_p = dbg_cta && tid == kConsumers;
      long long tp_raw = 0, tp_free = 0, tp_work = 0, tp0 = 0;
      uint32_t n = 0;
      for (int u = u_beg; u < u_end; ++u, ++n) {
        if (dbg_p) tp0 = clock64();
        const int tile = u / n_slabs, sl = u - tile * n_slabs;
        const DwProblem& P = a.p[a.tile_prob[tile]];
        const int n0 = a.tile_n[tile] * TN;
        const int h = min(TN, (P.N - n0 + 15) & ~15) >> 1;
        const bool in_b = 4 * bq < h && n0 + (int)rank * h + 4 * bq < P.N;
        const int k0 = sl * 64 + 8 * pw;
        const uint32_t slot = n % kRawSlots, src = raw0 + slot * kRawUnit;
        mbar_wait_relaxed(rbar0 + 8 * slot, (n / kRawSlots) & 1);
        if (dbg_p) { const long long t1 = clock64(); tp_raw += t1 - tp0; tp0 = t1; }
        float4 vb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 2 * i + bsub;
          const float4 y = lds4(src + (uint32_t)(r * 256 + bq * 16));
          vb[i] = (k0 + r < K && in_b) ? y : make_float4(0.f, 0.f, 0.f, 0.f);   // samples beyond the count multiply as 0
        }
        __syncwarp();                                           // every lane has read the slot: it may be refilled
        if (lane == 0) issue(u + kRawSlots, n + kRawSlots);
        const uint32_t st = g % NST;
        const uint32_t stage_base = smem + st * kStageD2;
        if (dbg_p) { const long long t1 = clock64(); tp_work += t1 - tp0; tp0 = t1; }
        if (g >= (uint32_t)NST) mbar_wait_relaxed(B.free0 + 8 * st, ((g / NST) - 1) & 1);
        if (dbg_p) { const long long t1 = clock64(); tp_free += t1 - tp0; tp0 = t1; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t r = (uint32_t)(8 * pw + 2 * i + bsub);
          const uint32_t off = b_blk + r * 128u + ((b_c32 ^ (r & 7u)) * 16u);
          split_store_bf4(stage_base + 2 * kATile + off, stage_base + 2 * kATile + kBHalfD2 + off, vb[i]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(B.afull0 + 8 * st);         // own CTA's barrier (the follower's is relayed)
        ++g;
        if (dbg_p) tp_work += clock64() - tp0;
      }
      if (dbg_p) { a.dbg[2] = tp_raw; a.dbg[3] = tp_free; a.dbg[4] = tp_work; }
    } else {
    // cursor over the CTA's units; bf16 walks HALF slabs (h = 0, 1 inside slab s of `tile`)
    struct Cur { int u, tile, s, h; };
    auto advance = [&](Cur& c) {
      if (BF && c.h == 0) { c.h = 1; return; }
      c.h = 0;
      ++c.u; ++c.s;
      if (c.s == n_slabs) { c.s = 0; ++c.tile; }
    };
    auto load_ab = [&](float4 (&va)[4], float4 (&vb)[2], const Cur& c) {
      if (c.u >= u_end) return;
      const DwProblem& P = a.p[a.tile_prob[c.tile]];
      const int n0 = a.tile_n[c.tile] * TN;
      const int h = min(TN, (P.N - n0 + 15) & ~15) >> 1;        // input features per CTA
      const int fa = (int)rank * BM + 4 * lane, fb = n0 + (int)rank * h + 4 * bq;
      const int k0 = (BF ? (2 * c.s + c.h) : c.s) * BK + 4 * pw;
      const bool in_a = fa < P.M, in_b = 4 * bq < h && fb < P.N;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + i < K && in_a) x = ldcg4(P.A + (size_t)(k0 + i) * P.lda + fa);
        va[i] = x;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = k0 + 2 * i + bsub;
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K && in_b) y = ldcg4(P.B + (size_t)k * P.ldb + fb);
        vb[i] = y;
      }
    };
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush_colsum = [&](int tile) {
      const DwProblem& P = a.p[a.tile_prob[tile]];
      if (P.colsum && a.tile_n[tile] == 0) {
        const int f = (int)rank * BM + 4 * lane;
        if (f < P.M && cs.x != 0.f) atomicAdd(P.colsum + f, cs.x);
        if (f + 1 < P.M && cs.y != 0.f) atomicAdd(P.colsum + f + 1, cs.y);
        if (f + 2 < P.M && cs.z != 0.f) atomicAdd(P.colsum + f + 2, cs.z);
        if (f + 3 < P.M && cs.w != 0.f) atomicAdd(P.colsum + f + 3, cs.w);
      }
      cs = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto produce = [&](float4 (&va)[4], float4 (&vb)[2], const Cur& c, Cur& pf) {
      const uint32_t st = g % NST;
      const uint32_t stage_base = smem + st * kStageD2;
      const bool first = !BF || c.h == 0, last = !BF || c.h == 1;      // of this slab's units
      if (first && g >= (uint32_t)NST) mbar_wait_relaxed(B.free0 + 8 * st, ((g / NST) - 1) & 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cs.x += va[i].x; cs.y += va[i].y; cs.z += va[i].z; cs.w += va[i].w;
        if (BF) {
          const uint32_t r = (uint32_t)(32 * c.h + 4 * pw + i);
          const uint32_t off = a_blk + r * 128u + ((a_c32 ^ (r & 7u)) * 16u);
          split_store_bf4(stage_base + off, stage_base + kATile + off, va[i]);
        } else {
          const uint32_t r = (uint32_t)(4 * pw + i);
          const uint32_t off = a_blk + r * 128u + ((a_c32 ^ (r & 3u)) * 32u);
          split_store(stage_base + off, stage_base + kATile + off, va[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (BF) {
          const uint32_t r = (uint32_t)(32 * c.h + 4 * pw + 2 * i + bsub);
          const uint32_t off = b_blk + r * 128u + ((b_c32 ^ (r & 7u)) * 16u);
          split_store_bf4(stage_base + 2 * kATile + off, stage_base + 2 * kATile + kBHalfD2 + off, vb[i]);
        } else {
          const uint32_t r = (uint32_t)(4 * pw + 2 * i + bsub);
          const uint32_t off = b_blk + r * 128u + ((b_c32 ^ (r & 3u)) * 32u);
          split_store(stage_base + 2 * kATile + off, stage_base + 2 * kATile + kBHalfD2 + off, vb[i]);
        }
      }
      if (last) {
        if (c.s == n_slabs - 1 || c.u == u_end - 1) flush_colsum(c.tile);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(B.afull0 + 8 * st);   // own CTA's barrier (the follower's is relayed)
        ++g;
      }
      load_ab(va, vb, pf);
      advance(pf);
    };
    Cur c{u_beg, u_beg / n_slabs, u_beg % n_slabs, 0}, pf = c;
    float4 va0[4], vb0[2], va1[4], vb1[2];
    load_ab(va0, vb0, pf); advance(pf);
    load_ab(va1, vb1, pf); advance(pf);
    while (c.u < u_end) {
      produce(va0, vb0, c, pf); advance(c);
      if (c.u < u_end) { produce(va1, vb1, c, pf); advance(c); }
    }
    }
  } else {
    // ================= consumers: this CTA's 128 out-features x BN =================
    const int quarter = warp & 3, col_half = (warp >> 2) * kAcc;
    const uint32_t tmem_row = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col_half;
    float* stg = reinterpret_cast<float*>(smem_raw + kStagingOff + warp * kStagingWarp);
    uint32_t g = 0;
    const bool dbg_c = dbg_cta && tid == 0;
    long long t_main = 0, t_epi = 0, t0c = 0;
    for (int u = u_beg; u < u_end;) {
      const int tile = u / n_slabs, s0 = u - tile * n_slabs;
      const int len = min(n_slabs - s0, u_end - u);
      const DwProblem& P = a.p[a.tile_prob[tile]];
      const int m0 = (int)rank * BM, n0 = a.tile_n[tile] * TN;
      const int BN = min(TN, (P.N - n0 + 15) & ~15);
      float acc[kAcc];
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;
      if (dbg_c) t0c = clock64();
      for (int i = 0; i < len; ++i, ++g)
        drain_chunk<false>(acc, tmem_row, col_half, g, BN, B.rfull0, B.rempty0, rank == 0 ? 0u : l_rempty0);
      if (dbg_c) { const long long t1 = clock64(); t_main += t1 - t0c; t0c = t1; }
#pragma unroll
      for (int q = 0; q < kAcc / 32; ++q) {
        const int c0 = col_half + 32 * q;
        if (c0 < BN) {
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 32 + (j ^ lane)] = acc[32 * q + j];
          __syncwarp();
          const int col = n0 + c0 + lane;
          if (col < P.N) {
            const int mr0 = m0 + quarter * 32;
            float* __restrict__ cp = P.C + (size_t)mr0 * P.ldc + col;
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
              if (mr0 + r < P.M) atomicAdd(cp + (size_t)r * P.ldc, stg[r * 32 + (lane ^ r)]);
          }
          __syncwarp();
        }
      }
      u += len;
      if (dbg_c) t_epi += clock64() - t0c;
    }
    if (dbg_c) { a.dbg[0] = t_main; a.dbg[1] = t_epi; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
  }
}

// ---- backward-data chain on bf16 PLANES: no operand producers at all ---------------------------------------------
// Measured on the 3xBF16 chain above (tools/chain_phase_probe.py, ncu): once the tensor core needs half the time, the
// kernel is bound by the CUDA cores -- the producers split every activation into hi / mid again (5 integer ops per
// element) right after the epilogue of the previous step wrote it as fp32, the weight-gradient kernel splits the same
// dZ a second time, and 8 epilogue warps cannot keep up with 12 MMAs per 64 k (issuer waiting for a free accumulator
// 30-40 % of the time).  Here every gradient tensor that feeds another GEMM lives in HBM / L2 as two bf16 planes
// ([rows][256] hi, then mid: the same 4 bytes per element as fp32), written ONCE by the epilogue that computes it and
// read by TMA straight into the operand tiles of the next step and of the weight-gradient kernel:
//   * no producer warps: one thread per CTA issues four bulk-tensor loads per 64-k slab (its 128 rows of A hi / mid
//     and its half of the weight rows hi / mid), all completing on the leader's barrier;
//   * 16 epilogue warps (4 per TMEM lane quarter, 64 columns each) instead of 8: TMEM -> mask -> column sums (the
//     bias gradient, which the weight-gradient kernel can no longer take from fp32 operands) -> hi / mid split ->
//     swizzled staging tile -> one TMA store per plane and 32-column block;
//   * the steps whose output leaves the GEMMs (dX0, the reflection part of dXR -> feat_bwd_kernel) store fp32.
constexpr int kBwConsumers = 512;
constexpr int kBwIssuerWarp = kBwConsumers / 32;
constexpr int kBwThreads = kBwConsumers + 32;
constexpr int kBwStages = 3;
constexpr uint32_t kBwPlane = 128 * 128;                          // 16 KB: 128 rows x 64 bf16
constexpr uint32_t kBwStage = 4 * kBwPlane;                       // A_hi, A_mid, B_hi, B_mid (this CTA's half) = 64 KB
constexpr uint32_t kBwStagingOff = kBwStages * kBwStage;          // 192 KB
constexpr uint32_t kBwStagingWarp = 2048;                         // 16 columns: [hi 32 x 32 B | mid 32 x 32 B] or fp32 32 x 64 B
constexpr uint32_t kBwBarOff = kBwStagingOff + 16 * kBwStagingWarp;   // 224 KB: free[3] full[3] rfull[2] rempty[2]
constexpr uint32_t kBwCDoneOff = kBwBarOff + 8 * 10;
static_assert(kBwCDoneOff + 8 * kChainMaxTiles + 16 + 256 * 4 <= 232448, "shared memory budget");
constexpr uint32_t kBwTmemSlotOff = kBwCDoneOff + 8 * kChainMaxTiles;
constexpr uint32_t kBwColsumOff = kBwTmemSlotOff + 16;
constexpr uint32_t kSmemBytesBw = kBwColsumOff + 256 * 4;

struct BwStep {
  const uint32_t* bits_in;      // ReLU mask of the layer this gradient flows into ([rows][8] words), NULL: none
  float* C;                     // fp32 output [rows][ldc] (out_planes == 0)
  uint16_t* Chi; uint16_t* Cmid;   // bf16 planes [rows][256] (out_planes != 0)
  float* colsum;                // += column sums of the masked output = the bias gradient of that layer, NULL: none
  int ldc, K, N, n0, out_planes, pad_;
};
struct BwArgs {
  BwStep s[kChainMaxSteps];
  const int* rows_dev;
  long long* dbg;
  int n_steps, M, row_off;
};
struct BwMaps {
  CUtensorMap a_hi[kChainMaxSteps], a_mid[kChainMaxSteps], b_hi[kChainMaxSteps], b_mid[kChainMaxSteps],
      c_hi[kChainMaxSteps], c_mid[kChainMaxSteps], c32[kChainMaxSteps];
};

// sum over the 32 lanes of a warp of 16 per-lane values: butterfly that halves the values per round (8 + 4 + 2 + 1 + 1
// shuffles); afterwards the lane pair (2 c, 2 c + 1) holds the total of column c_of_lane
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane, int& col) {
  float a8[8], a4[4], a2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? v[i] : v[i + 8];
    a8[i] = (b4 ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? a8[i] : a8[i + 4];
    a4[i] = (b3 ? a8[i + 4] : a8[i]) + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? a4[i] : a4[i + 2];
    a2[i] = (b2 ? a4[i + 2] : a4[i]) + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? a2[0] : a2[1];
  float r = (b1 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  col = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
  return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBwThreads, 1)
mlp_bwd_planes_kernel(const __grid_constant__ BwArgs a, const __grid_constant__ BwMaps maps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  const uint32_t free0 = smem + kBwBarOff, full0 = free0 + 8 * kBwStages, rfull0 = full0 + 8 * kBwStages, rempty0 = rfull0 + 16;
  const uint32_t cdone0 = smem + kBwCDoneOff;
  float* s_colsum = reinterpret_cast<float*>(smem_raw + kBwColsumOff);

  int M = a.M;
  if (a.rows_dev) M = min(M, *a.rows_dev - a.row_off);
  if (M <= 0) return;
  const int n_pair_tiles = (M + 2 * BM - 1) / (2 * BM);
  const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
  if (pair >= n_pair_tiles) return;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + kBwTmemSlotOff);
  if (tid == 0) {
    for (int i = 0; i < kChainMaxTiles; ++i) mbar_init(cdone0 + 8 * i, kBwConsumers / 32);
    for (int i = 0; i < kBwStages; ++i) {
      mbar_init(free0 + 8 * i, 1);                               // one multicast commit
      mbar_init(full0 + 8 * i, 1);                               // the leader's expect_tx arrive; bytes from both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(rfull0 + 8 * i, 1);
      mbar_init(rempty0 + 8 * i, 2 * (kBwConsumers / 32));       // consumer warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 256) s_colsum[tid] = 0.f;
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  const int my_tiles = (n_pair_tiles - 1 - pair) / n_pairs + 1;
  const int n_items = a.n_steps * my_tiles;                    // item w: step w / my_tiles, local tile w % my_tiles
  const uint32_t l_full0 = map_to_cta(full0, 0), l_rempty0 = map_to_cta(rempty0, 0);
  const bool dbg_cta = a.dbg != nullptr && blockIdx.x == 0;
  auto tile_row0 = [&](int lt) { return ((pair + lt * n_pairs) * 2 + (int)rank) * BM; };   // this CTA's 128 rows

  if (warp == kBwIssuerWarp) {
    if (lane == 0) {
      // both CTAs: TMA of THIS CTA's rows of A and its half of the weight rows; leader: also every MMA
      struct Cursor { int w, step, lt, kt, n_slabs; };
      auto cur_init = [&](Cursor& c) {
        c.w = 0; c.step = 0; c.lt = 0; c.kt = 0; c.n_slabs = (a.s[0].K + 63) / 64;
      };
      auto cur_next = [&](Cursor& c) {
        if (++c.kt == c.n_slabs) {
          c.kt = 0;
          ++c.w;
          if (c.w < n_items) { c.step = c.w / my_tiles; c.lt = c.w - c.step * my_tiles; c.n_slabs = (a.s[c.step].K + 63) / 64; }
        }
      };
      long long t_iss = 0, t_full = 0, t_rem = 0, t_all = dbg_cta ? clock64() : 0, t0 = 0;
      auto issue = [&](const Cursor& c, uint32_t g) {
        const long long ti0 = dbg_cta ? clock64() : 0;
        const BwStep& S = a.s[c.step];
        const int BN = min(TNF, (S.N - S.n0 + 15) & ~15);
        const uint32_t st = g % kBwStages, base = smem + st * kBwStage;
        if (g >= (uint32_t)kBwStages) mbar_wait(free0 + 8 * st, ((g / kBwStages) - 1) & 1);
        if (c.kt == 0 && c.step > 0) {
          // this tile's rows of the step's input were stored by the epilogue of (tile, step - 1)
          mbar_wait(cdone0 + 8 * c.lt, (uint32_t)(c.step - 1) & 1);
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        if (rank == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8 * st), "r"(8 * kBwPlane) : "memory");
        const uint32_t bar = l_full0 + 8 * st;
        const int row0 = tile_row0(c.lt), brow0 = S.n0 + (int)rank * (BN >> 1);
        tma2_load_2d(base, &maps.a_hi[c.step], bar, c.kt * 64, row0);
        tma2_load_2d(base + kBwPlane, &maps.a_mid[c.step], bar, c.kt * 64, row0);
        tma2_load_2d(base + 2 * kBwPlane, &maps.b_hi[c.step], bar, c.kt * 64, brow0);
        tma2_load_2d(base + 3 * kBwPlane, &maps.b_mid[c.step], bar, c.kt * 64, brow0);
        if (dbg_cta) t_iss += clock64() - ti0;
      };
      // operand loads run kPf slabs ahead of the MMAs.  Two (three stages) where the dependency allows it: the loads of
      // an item's first slab wait for the epilogue of (tile, step - 1), i.e. of the item my_tiles places earlier, whose
      // MMAs must have been ISSUED before this thread may block on it -- true two slabs ahead iff my_tiles >= 2.
      const uint32_t kPf = my_tiles >= 2 ? 2u : 1u;
      Cursor lc;
      cur_init(lc);
      uint32_t lg = 0;
      for (; lg < kPf && lc.w < n_items; ++lg) { issue(lc, lg); cur_next(lc); }
      uint32_t g = 0;
      for (int w = 0; w < n_items; ++w) {
        const int step = w / my_tiles;
        const BwStep& S = a.s[step];
        const int BN = min(TNF, (S.N - S.n0 + 15) & ~15);
        const int n_slabs = (S.K + 63) / 64;
        const uint32_t idesc = make_idesc(true, BN, 2 * BM, false);
        const uint32_t rg = (uint32_t)w & 1;
        for (int kt = 0; kt < n_slabs; ++kt, ++g) {
          const uint32_t st = g % kBwStages;
          if (rank == 0) {
            if (dbg_cta) t0 = clock64();
            mbar_wait_cluster(full0 + 8 * st, (g / kBwStages) & 1);
            if (dbg_cta) { const long long t1 = clock64(); t_full += t1 - t0; t0 = t1; }
            if (kt == 0 && w >= 2) mbar_wait_cluster(rempty0 + 8 * rg, (((uint32_t)w >> 1) - 1) & 1);
            if (dbg_cta) t_rem += clock64() - t0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem + st * kBwStage, d = tmem_d + rg * TNF;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t dah = make_desc(base + j * 32), dam = make_desc(base + kBwPlane + j * 32);
              const uint64_t dbh = make_desc(base + 2 * kBwPlane + j * 32), dbm = make_desc(base + 3 * kBwPlane + j * 32);
              mma2_bf16(d, dam, dbh, idesc, (kt != 0 || j != 0) ? 1u : 0u);
              mma2_bf16(d, dah, dbm, idesc, 1);
              mma2_bf16(d, dah, dbh, idesc, 1);
            }
            mma2_commit(free0 + 8 * st);
            if (kt == n_slabs - 1) mma2_commit(rfull0 + 8 * rg);
          }
          // operands kPf slabs on (their stage was read by slab g + kPf - 3, whose MMAs are queued in front of slab g's)
          if (lc.w < n_items) {
            issue(lc, lg);
            cur_next(lc);
            ++lg;
          }
        }
      }
      // (slots as mlp_chain2_kernel: 5 = operands not there yet, 7 = no free accumulator, 8 = inside issue(): stage /
      // dependency waits + 4 TMA, 9 = total)
      if (dbg_cta) { a.dbg[5] = t_full; a.dbg[6] = 0; a.dbg[7] = t_rem; a.dbg[8] = t_iss; a.dbg[9] = clock64() - t_all; }
    }
  } else {
    // ================= 16 epilogue warps =================
    const int quarter = warp & 3, cg = warp >> 2;                // TMEM lanes 32 quarter.., columns 64 cg..
    const uint32_t stg = smem + kBwStagingOff + (uint32_t)warp * kBwStagingWarp;
    const uint32_t swz64 = (uint32_t)((lane >> 1) & 3);        // SWIZZLE_64B: 16-byte chunk ^= (row / 2) % 4
    const uint32_t swz32 = (uint32_t)((lane >> 2) & 1);        // SWIZZLE_32B: 16-byte chunk ^= (row / 4) % 2
    const bool dbg_on = dbg_cta && tid == 0;
    long long t_main = 0, t_epi = 0, t_fe = 0, tc0 = 0;
    for (int w = 0; w < n_items; ++w) {
      const int step = w / my_tiles, lt = w - step * my_tiles;
      const BwStep& S = a.s[step];
      const int n0 = S.n0, N = S.N;
      const int BN = min(TNF, (N - n0 + 15) & ~15);
      const uint32_t rg = (uint32_t)w & 1;
      const int m0 = tile_row0(lt), mr0 = m0 + quarter * 32, m = mr0 + lane;
      const uint32_t tmem_row = tmem_d + ((uint32_t)(quarter * 32) << 16) + rg * TNF + 64 * cg;
      const int nhalf = max(0, min(4, (BN - 64 * cg + 15) >> 4));
      const bool has_mask = S.bits_in != nullptr, planes = S.out_planes != 0, do_colsum = S.colsum != nullptr;
      const bool tma_out = mr0 + 32 <= M;
      uint32_t mw0 = 0xffffffffu, mw1 = 0xffffffffu;
      if (has_mask && m < M) {
        const uint2 t = *reinterpret_cast<const uint2*>(S.bits_in + (size_t)m * 8 + 2 * cg);
        mw0 = t.x; mw1 = t.y;
      }
      if (dbg_on) tc0 = clock64();
      mbar_wait_relaxed(rfull0 + 8 * rg, ((uint32_t)w >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (dbg_on) { const long long t1 = clock64(); t_main += t1 - tc0; tc0 = t1; }
      auto release_region = [&]() {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(rempty0 + 8 * rg);
          else mbar_arrive_cluster(l_rempty0 + 8 * rg);
        }
      };
      // half hf (0..3) = columns 64 cg + 16 hf .. + 16 of the item
      auto process_half = [&](uint32_t (&r)[16], int hf, uint32_t bits16) {
        const int c0 = 64 * cg + 16 * hf;                      // column within the item's 256-wide block
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // mask bit j of bits16 -> all-ones / all-zeros by sign extension of a 1-bit field
          const int keep = has_mask ? ((int)(bits16 << (31 - j)) >> 31) : -1;
          v[j] = __uint_as_float(r[j] & (uint32_t)keep);
        }
        if (m >= M) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;             // rows beyond the sample count: keep them out of the sums
        }
        if (do_colsum) {
          int col;
          const float tot = warp_colsum16(v, lane, col);
          if ((lane & 1) == 0 && tot != 0.f) atomicAdd(s_colsum + c0 + col, tot);
        }
        if (planes) {
          uint32_t h[8], md[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bf16_split2(v[2 * j], v[2 * j + 1], h[j], md[j]);
          if (tma_out) {
            // staging tile of this warp: [hi 32 rows x 32 B | mid 32 rows x 32 B], one TMA store per plane and half
            if (lane == 0) tma_store_wait_read();              // (the previous half's stores have read the tile; the
            __syncwarp();                                      //  arithmetic above ran while they did)
            const uint32_t row = stg + (uint32_t)lane * 32;
            st_shared_v4u(row + ((0u ^ swz32) * 16), h[0], h[1], h[2], h[3]);
            st_shared_v4u(row + ((1u ^ swz32) * 16), h[4], h[5], h[6], h[7]);
            st_shared_v4u(row + 1024 + ((0u ^ swz32) * 16), md[0], md[1], md[2], md[3]);
            st_shared_v4u(row + 1024 + ((1u ^ swz32) * 16), md[4], md[5], md[6], md[7]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&maps.c_hi[step], stg, n0 + c0, mr0);
              tma_store_2d(&maps.c_mid[step], stg + 1024, n0 + c0, mr0);
            }
          } else if (m < M) {
            uint4* ph = reinterpret_cast<uint4*>(S.Chi + (size_t)m * 256 + n0 + c0);
            uint4* pm = reinterpret_cast<uint4*>(S.Cmid + (size_t)m * 256 + n0 + c0);
            ph[0] = make_uint4(h[0], h[1], h[2], h[3]); ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
            pm[0] = make_uint4(md[0], md[1], md[2], md[3]); pm[1] = make_uint4(md[4], md[5], md[6], md[7]);
          }
        } else {
          if (tma_out) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              st_shared_v4(stg + (uint32_t)(lane * 64 + (((uint32_t)j4 ^ swz64) * 16)),
                           make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tma_store_2d(&maps.c32[step], stg, n0 + c0, mr0);
          } else if (m < M) {
            float* __restrict__ cp = S.C + (size_t)m * S.ldc + n0 + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c0 + j < N) cp[j] = v[j];
          }
        }
      };
      {
        uint32_t ra[16], rb[16];
        if (nhalf > 0) tmem_ld16(tmem_row, ra);
        else release_region();
#pragma unroll
        for (int hf = 0; hf < 4; ++hf) {
          if (hf < nhalf) {
            tmem_ld_wait();
            if (hf + 1 < nhalf) {
              if (hf & 1) tmem_ld16(tmem_row + 16 * (hf + 1), ra);
              else tmem_ld16(tmem_row + 16 * (hf + 1), rb);
            } else {
              release_region();
            }
            const uint32_t bits16 = (hf < 2 ? mw0 : mw1) >> (16 * (hf & 1));
            if (hf & 1) process_half(rb, hf, bits16); else process_half(ra, hf, bits16);
          }
        }
      }
      // the tile's rows of this step are complete in global memory -> the issuer may load them for (tile, step + 1)
      __syncwarp();
      if (dbg_on) { const long long t1 = clock64(); t_epi += t1 - tc0; tc0 = t1; }
      if (lane == 0) {
        tma_store_wait_all();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __threadfence();
        mbar_arrive(cdone0 + 8 * lt);
      }
      // bias gradient of this step: the CTA's partial column sums leave once its last tile of the step is done
      if (do_colsum && lt == my_tiles - 1) {
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (tid < 256) {
          const float t = s_colsum[tid];
          if (t != 0.f && n0 + tid < N) atomicAdd(S.colsum + n0 + tid, t);
          s_colsum[tid] = 0.f;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      if (dbg_on) t_fe += clock64() - tc0;
    }
    if (dbg_on) { a.dbg[0] = t_main; a.dbg[1] = t_epi; a.dbg[24] = t_fe; }
  }
  // ---- teardown: nobody may leave (or free TMEM) while the partner can still signal it or multiply into it ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
  }
}

// fp32 rows -> bf16 hi / mid planes (tests and stand-alone entries; inside the fused step the planes are written by the
// kernels that compute the values).  One thread per 2 elements; column sums by atomics (small matrices only).
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ X, int ld, int rows, int cols, uint16_t* __restrict__ hi,
                    uint16_t* __restrict__ mid, int ldp, float* __restrict__ colsum, const int* __restrict__ rows_dev) {
  int valid = rows;
  if (rows_dev) valid = min(valid, *rows_dev);
  const int half = ldp >> 1;
  const long long total = (long long)rows * half;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / half), c = 2 * (int)(i - (long long)r * half);
    const float x0 = (c < cols) ? X[(size_t)r * ld + c] : 0.f, x1 = (c + 1 < cols) ? X[(size_t)r * ld + c + 1] : 0.f;
    uint32_t h, m;
    bf16_split2(x0, x1, h, m);
    reinterpret_cast<uint32_t*>(hi + (size_t)r * ldp)[c >> 1] = h;
    reinterpret_cast<uint32_t*>(mid + (size_t)r * ldp)[c >> 1] = m;
    if (colsum && r < valid) {
      if (c < cols && x0 != 0.f) atomicAdd(colsum + c, x0);
      if (c + 1 < cols && x1 != 0.f) atomicAdd(colsum + c + 1, x1);
    }
  }
}

// Weight gradient of a layer with a handful of out-features (the last refnet layer: 3): C[m][n] += sum_k A[k][m] B[k][n],
// A fp32 [K][lda] (lda >= 4, m < 4), B fp32 [K][ldb].  On the tensor-core kernel these problems cost as much as a
// 256-out-feature layer (the MMA is 256 rows tall: 2 of 16 tiles, 12 % of the kernel, for 0.3 % of the FLOPs); here one
// thread per input feature streams B once (coalesced rows) with the A row broadcast from shared memory.
// (16 rows per chunk: the grid is sized from the CAPACITY -- 592 CTAs -- while the row count lives on the device, so
// with 64-row chunks 44 K rows were 693 chunks = two rounds with the second 17 % full; 16-row chunks balance to 6 %)
constexpr int kDwSmallRows = 16;
__global__ void __launch_bounds__(256)
dw_small_kernel(const float* __restrict__ A, int lda, int M, const float* __restrict__ B, int ldb, int N,
                float* __restrict__ C, int ldc, int K, const int* __restrict__ rows_dev) {
  __shared__ float4 s_a[kDwSmallRows];
  if (rows_dev) K = min(K, *rows_dev);
  const int n = blockIdx.y * 256 + threadIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = blockIdx.x * kDwSmallRows; k0 < K; k0 += gridDim.x * kDwSmallRows) {
    // the chunk's 16 B rows first (independent loads, all in flight), THEN the A rows through shared memory: the A load's
    // latency hides behind the B loads instead of in front of them
    float b[kDwSmallRows];
#pragma unroll
    for (int j = 0; j < kDwSmallRows; ++j) b[j] = (n < N && k0 + j < K) ? __ldg(B + (size_t)(k0 + j) * ldb + n) : 0.f;
    __syncthreads();                                   // the previous chunk's readers of s_a are done
    if (threadIdx.x < kDwSmallRows) {
      const int k = k0 + threadIdx.x;
      s_a[threadIdx.x] = k < K ? *reinterpret_cast<const float4*>(A + (size_t)k * lda) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kDwSmallRows; ++j) {
      const float4 a4 = s_a[j];
      acc[0] = fmaf(a4.x, b[j], acc[0]); acc[1] = fmaf(a4.y, b[j], acc[1]);
      acc[2] = fmaf(a4.z, b[j], acc[2]); acc[3] = fmaf(a4.w, b[j], acc[3]);
    }
  }
  if (n < N)
    for (int m = 0; m < M && m < 4; ++m)
      if (acc[m] != 0.f) atomicAdd(C + (size_t)m * ldc + n, acc[m]);
}

long long* g_dbg = nullptr;
long long* g_dbg_dw = nullptr;
int g_dw_sm_reserve = 0;

template <typename Kern>
int configure_once(Kern kern, bool& configured, const char* what, uint32_t smem_bytes) {
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return fgs::check_cuda(e, what);
    configured = true;
  }
  return 0;
}

// 2-D tensor map of a K-major fp32 plane [rows_alloc][ld]; box = (32 floats, 128 rows) = one k-slab of
// 128 rows, written to shared memory as 128-byte rows with the 128B swizzle -- the UMMA K-major
// SWIZZLE_128B layout.  (A first version used a 4-D map producing the no-swizzle core-matrix layout;
// its 16-byte inner box made TMA the bottleneck: ~1.4 us per 32 KB slab.)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::map<std::tuple<const void*, int, int, int, int, int>, CUtensorMap> g_maps;

// 2-D map of a row-major fp32 matrix: `cols` valid columns (dim 0), `rows` rows of pitch ld floats; box =
// (32 floats, box_rows rows), SWIZZLE_128B.  Out-of-range box elements read as zero / are not written.
// bf16: the plane holds 2-byte elements (cols / ld in elements), box = (64 elements, box_rows rows): the same 128 bytes.
// raw_box_cols > 0: a plain (unswizzled) box of raw_box_cols x box_rows (the raw operand ring of the weight-gradient kernel);
// raw_box_cols < 0: a box of -raw_box_cols columns with SWIZZLE_32B / SWIZZLE_64B (32- / 64-byte rows: the staging tiles of
// the plane epilogue).
int get_map(const void* plane, int cols, int rows, int ld, int box_rows, CUtensorMap* out, bool bf16 = false,
            int raw_box_cols = 0) {
  auto key = std::make_tuple((const void*)plane, cols, rows, ld, box_rows, (int)bf16 + 2 * raw_box_cols);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return 0;
  }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn)
      return fgs::check_cuda(e != cudaSuccess ? e : cudaErrorUnknown, "cuTensorMapEncodeTiled lookup");
    g_encode = (EncodeTiledFn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * (bf16 ? 2 : 4)};              // bytes, dim 1
  const cuuint32_t box[2] = {(cuuint32_t)(raw_box_cols > 0 ? raw_box_cols : raw_box_cols < 0 ? -raw_box_cols : (bf16 ? 2 * BK : BK)),
                             (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = g_encode(&tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)plane, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        raw_box_cols > 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                        : raw_box_cols < 0 ? ((-raw_box_cols) * (bf16 ? 2 : 4) == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B)
                                           : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(fgs::g_err, sizeof(fgs::g_err), "fgs_b200: cuTensorMapEncodeTiled failed (%d) cols %d rows %d ld %d",
             (int)r, cols, rows, ld);
    return FGS_E_BADARG;
  }
  if (g_maps.size() > 4096) g_maps.clear();   // the fused step uses ~70 distinct maps; tests may churn through more
  g_maps[key] = tm;
  *out = tm;
  return 0;
}

float* g_scratch = nullptr;
size_t g_scratch_floats = 0;

}  // namespace

namespace fine {

// C[M,N] = A[M,K] * (Bh + Bl)[N,K]^T (+bias, ReLU -> bits_out, masked by bits_in).  Bh/Bl: K-major
// planes [round_up(N,8)][ldb] (rows >= N and columns >= K zero), 16-byte aligned, ldb % 4 == 0.
// bits_*: [M][8] words (bit n of row m).
int gemm_w(int M, int N, int K, const float* A, int lda, const float* Bh, const float* Bl, int ldb, float* C,
           int ldc, const float* bias, int relu, const uint32_t* bits_in, uint32_t* bits_out,
           const int32_t* rows_dev, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((lda & 3) || (ldb & 3) || (((uintptr_t)A | (uintptr_t)Bh | (uintptr_t)Bl) & 15))
    return fgs::fail_arg("gemm_w needs 16-byte aligned operands with leading dimensions % 4 == 0");
  if (K > ldb) return fgs::fail_arg("gemm_w: K exceeds the weight plane's row length");
  const int n_tiles = fgs::cdiv(M, BM);
  static bool configured_f = false;
  int rc = configure_once(gemm_wf_kernel, configured_f, "gemm_wf smem attribute", kSmemBytesF);
  if (rc) return rc;
  CUtensorMap tm_hi, tm_lo;
  const int rows_alloc = (N + 7) & ~7;
  if ((rc = get_map(Bh, ldb, rows_alloc, ldb, TN, &tm_hi))) return rc;
  if ((rc = get_map(Bl, ldb, rows_alloc, ldb, TN, &tm_lo))) return rc;
  // C leaves through TMA when its rows are 16-byte aligned (always, inside the fused step)
  const int c_tma = ((ldc & 3) == 0 && ((uintptr_t)C & 15) == 0) ? 1 : 0;
  CUtensorMap tm_c = tm_hi;
  if (c_tma && (rc = get_map(C, N, M, ldc, 32, &tm_c))) return rc;
  WArgs a{A, C, bias, bits_in, bits_out, rows_dev, M, N, K, lda, ldc, relu, c_tma, g_dbg};
  // persistent CTAs: one per SM, nx column blocks x ny row-tile walkers
  const int nx = fgs::cdiv(N, TNF);
  int ny = FGS_NUM_SMS / nx;
  if (ny > n_tiles) ny = n_tiles;
  if (ny < 1) ny = 1;
  gemm_wf_kernel<<<dim3(nx, ny, 1), kThreadsF, kSmemBytesF, st>>>(a, tm_hi, tm_lo, tm_c);
  return fgs::check_launch("gemm_w", st);
}

// The MLP chain in one launch (see mlp_chain_kernel).  steps[i]: C_i = act(A_i W_i^T + b_i) on the column
// block [n0, n0 + 256) with pre-split weight planes Bh / Bl ([round_up(N,8)][ldb], K-major).
int mlp_chain(const ChainStepHost* steps, int n_steps, int M, const int32_t* rows_dev, cudaStream_t st) {
  constexpr bool bf16 = false;
  if (M <= 0 || n_steps <= 0) return 0;
  if (n_steps > kChainMaxSteps) return fgs::fail_arg("mlp_chain: too many steps");
  static bool configured2 = false;
  int rc = configure_once(mlp_chain2_kernel, configured2, "mlp_chain2 smem attribute", kSmemBytesChain2);
  if (rc) return rc;
  // a CTA pair walks at most kChainMaxTiles 256-row tiles (one completion barrier each): larger workspaces (the host
  // only knows the CAPACITY; the sample count lives on the device) are covered by several launches over row ranges,
  // of which all but the first exit at once when the device-side count is small
  const int pairs_all = FGS_NUM_SMS / 2;
  const int max_rows = kChainMaxTiles * pairs_all * 2 * BM;
  for (int row_off = 0; row_off < M; row_off += max_rows) {
    const int Mc = M - row_off < max_rows ? M - row_off : max_rows;
    ChainArgs a;
    ChainMaps maps;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < n_steps; ++i) {
      const ChainStepHost& h = steps[i];
      if ((h.lda & 3) || (h.ldb & (bf16 ? 7 : 3)) || (((uintptr_t)h.A | (uintptr_t)h.Bh | (uintptr_t)h.Bl) & 15))
        return fgs::fail_arg("mlp_chain needs 16-byte aligned operands and rows");
      if (h.K > h.ldb) return fgs::fail_arg("mlp_chain: K exceeds the weight plane's row length");
      const int rows_alloc = (h.N + 7) & ~7;
      if ((rc = get_map(h.Bh, h.ldb, rows_alloc, h.ldb, TN, &maps.hi[i], bf16))) return rc;
      if ((rc = get_map(h.Bl, h.ldb, rows_alloc, h.ldb, TN, &maps.lo[i], bf16))) return rc;
      float* Cc = h.C + (size_t)row_off * h.ldc;
      const int c_tma = ((h.ldc & 3) == 0 && ((uintptr_t)Cc & 15) == 0) ? 1 : 0;
      maps.c[i] = maps.hi[i];
      if (c_tma && (rc = get_map(Cc, h.N, Mc, h.ldc, 32, &maps.c[i]))) return rc;
      ChainStep& s = a.s[i];
      s.A = h.A + (size_t)row_off * h.lda; s.C = Cc; s.bias = h.bias;
      s.bits_in = h.bits_in ? h.bits_in + (size_t)row_off * 8 : nullptr;
      s.bits_out = h.bits_out ? h.bits_out + (size_t)row_off * 8 : nullptr;
      s.lda = h.lda; s.ldc = h.ldc; s.K = h.K; s.N = h.N; s.n0 = h.n0; s.relu = h.relu; s.c_tma = c_tma;
    }
    for (int i = n_steps; i < kChainMaxSteps; ++i) { maps.hi[i] = maps.hi[0]; maps.lo[i] = maps.lo[0]; maps.c[i] = maps.c[0]; }
    a.rows_dev = rows_dev; a.dbg = g_dbg; a.n_steps = n_steps; a.M = Mc; a.row_off = row_off;
    // CTA pairs (static cluster size 2): the grid is a whole number of pairs
    const int n_pair_tiles = fgs::cdiv(Mc, 2 * BM);
    int pairs_n = pairs_all;
    if (pairs_n > n_pair_tiles) pairs_n = n_pair_tiles;
    mlp_chain2_kernel<<<2 * pairs_n, kThreadsF, kSmemBytesChain2, st>>>(a, maps);
    if ((rc = fgs::check_launch(steps[0].bias ? "mlp_chain_fwd" : "mlp_chain_bwd", st))) return rc;   // (only the forward adds biases)
  }
  return 0;
}

// The backward-data chain on bf16 planes (mlp_bwd_planes_kernel), large workspaces in row ranges like mlp_chain.
int mlp_bwd_planes(const BwStepHost* steps, int n_steps, int M, const int32_t* rows_dev, cudaStream_t st) {
  if (M <= 0 || n_steps <= 0) return 0;
  if (n_steps > kChainMaxSteps) return fgs::fail_arg("mlp_bwd_planes: too many steps");
  static bool configured = false;
  int rc = configure_once(mlp_bwd_planes_kernel, configured, "mlp_bwd_planes smem attribute", kSmemBytesBw);
  if (rc) return rc;
  const int pairs_all = FGS_NUM_SMS / 2;
  const int max_rows = kChainMaxTiles * pairs_all * 2 * BM;
  for (int row_off = 0; row_off < M; row_off += max_rows) {
    const int Mc = M - row_off < max_rows ? M - row_off : max_rows;
    BwArgs a;
    BwMaps maps;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < n_steps; ++i) {
      const BwStepHost& h = steps[i];
      if ((h.lda & 7) || (h.ldb & 7) || (((uintptr_t)h.Ahi | (uintptr_t)h.Amid | (uintptr_t)h.Bh | (uintptr_t)h.Bl) & 15))
        return fgs::fail_arg("mlp_bwd_planes needs 16-byte aligned planes with rows of whole 16 bytes");
      if (h.K > h.ldb || h.K > h.lda) return fgs::fail_arg("mlp_bwd_planes: K exceeds a plane's row length");
      const uint16_t* ah = h.Ahi + (size_t)row_off * h.lda;
      const uint16_t* am = h.Amid + (size_t)row_off * h.lda;
      if ((rc = get_map(ah, h.lda, Mc, h.lda, 128, &maps.a_hi[i], true))) return rc;
      if ((rc = get_map(am, h.lda, Mc, h.lda, 128, &maps.a_mid[i], true))) return rc;
      const int rows_alloc = (h.N + 7) & ~7;
      if ((rc = get_map(h.Bh, h.ldb, rows_alloc, h.ldb, TN, &maps.b_hi[i], true))) return rc;
      if ((rc = get_map(h.Bl, h.ldb, rows_alloc, h.ldb, TN, &maps.b_mid[i], true))) return rc;
      BwStep& s = a.s[i];
      s.bits_in = h.bits_in ? h.bits_in + (size_t)row_off * 8 : nullptr;
      s.colsum = h.colsum;
      s.K = h.K; s.N = h.N; s.n0 = h.n0; s.out_planes = h.out_planes; s.ldc = h.ldc;
      maps.c_hi[i] = maps.a_hi[i]; maps.c_mid[i] = maps.a_hi[i]; maps.c32[i] = maps.a_hi[i];
      if (h.out_planes) {
        if (h.N > 256 || h.n0 != 0) return fgs::fail_arg("mlp_bwd_planes: plane outputs are [rows][256]");
        s.Chi = h.Chi + (size_t)row_off * 256; s.Cmid = h.Cmid + (size_t)row_off * 256;
        if ((((uintptr_t)s.Chi | (uintptr_t)s.Cmid) & 15)) return fgs::fail_arg("mlp_bwd_planes: unaligned output planes");
        // box 16 columns x 32 rows of 2-byte elements = 32-byte rows, SWIZZLE_32B (the staging tile's layout)
        if ((rc = get_map(s.Chi, 256, Mc, 256, 32, &maps.c_hi[i], true, -16))) return rc;
        if ((rc = get_map(s.Cmid, 256, Mc, 256, 32, &maps.c_mid[i], true, -16))) return rc;
      } else {
        s.C = h.C + (size_t)row_off * h.ldc;
        if ((h.ldc & 3) || (((uintptr_t)s.C) & 15)) return fgs::fail_arg("mlp_bwd_planes: fp32 output rows must be 16-byte aligned");
        if ((rc = get_map(s.C, h.N, Mc, h.ldc, 32, &maps.c32[i], false, -16))) return rc;   // 16 fp32 columns = 64-byte rows
      }
    }
    for (int i = n_steps; i < kChainMaxSteps; ++i) {
      maps.a_hi[i] = maps.a_hi[0]; maps.a_mid[i] = maps.a_hi[0]; maps.b_hi[i] = maps.a_hi[0]; maps.b_mid[i] = maps.a_hi[0];
      maps.c_hi[i] = maps.a_hi[0]; maps.c_mid[i] = maps.a_hi[0]; maps.c32[i] = maps.a_hi[0];
    }
    a.rows_dev = rows_dev; a.dbg = g_dbg; a.n_steps = n_steps; a.M = Mc; a.row_off = row_off;
    const int n_pair_tiles = fgs::cdiv(Mc, 2 * BM);
    int pairs_n = pairs_all;
    if (pairs_n > n_pair_tiles) pairs_n = n_pair_tiles;
    mlp_bwd_planes_kernel<<<2 * pairs_n, kBwThreads, kSmemBytesBw, st>>>(a, maps);
    if ((rc = fgs::check_launch("mlp_chain_bwd", st))) return rc;
  }
  return 0;
}

int split_planes(const float* X, int ld, int rows, int cols, uint16_t* hi, uint16_t* mid, int ldp, float* colsum,
                 const int32_t* rows_dev, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  if ((ldp & 1) || cols > ldp) return fgs::fail_arg("split_planes: ldp must be even and >= cols");
  const int grid = fgs::stream_grid((long long)rows * (ldp >> 1), 256, 8);
  split_planes_kernel<<<grid, 256, 0, st>>>(X, ld, rows, cols, hi, mid, ldp, colsum, rows_dev);
  return fgs::check_launch("split_planes", st);
}

// C[M <= 4][N] += A[K][M]^T B[K][N] in fp32 on the CUDA cores (A rows read as float4: lda % 4 == 0)
int dw_small(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             const int32_t* rows_dev, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (M > 4 || (lda & 3) || (((uintptr_t)A) & 15)) return fgs::fail_arg("dw_small: M <= 4 and 16-byte aligned A rows");
  const int gx = fgs::stream_grid(fgs::cdiv(K, kDwSmallRows), 1, 4);
  dw_small_kernel<<<dim3(gx, fgs::cdiv(N, 256)), 256, 0, st>>>(A, lda, M, B, ldb, N, C, ldc, K, rows_dev);
  return fgs::check_launch("mlp_dw_small", st);
}

// All weight / bias gradients of the step in one launch: C_i[M_i,N_i] += A_i[K,M_i]^T B_i[K,N_i] over the first
// min(K, *rows_dev) rows, colsum_i[m] += sum_k A_i[k][m] (may be NULL).  Rows of A / B must be readable in whole
// float4s (lda, ldb % 4 == 0, 16-byte aligned bases).
int mlp_dw(const DwProblemHost* probs, int n, int K, const int32_t* rows_dev, cudaStream_t st, bool bf16) {
  if (n <= 0 || K <= 0) return 0;
  if (n > kDwMaxProblems) return fgs::fail_arg("mlp_dw: too many problems");
  // CTA pairs: a pair owns all (<= 256) out-features x one 128-wide block of input features
  Dw2Args b;
  memset(&b, 0, sizeof(b));
  int nt2 = 0;
  for (int i = 0; i < n; ++i) {
    const DwProblemHost& h = probs[i];
    if (h.M <= 0 || h.N <= 0) continue;
    if (!bf16 && ((h.lda & 3) || (((uintptr_t)h.A) & 15) || ((h.M + 3) & ~3) > h.lda))
      return fgs::fail_arg("mlp_dw needs 16-byte aligned operands with leading dimensions % 4 == 0");
    if ((h.ldb & 3) || (((uintptr_t)h.B) & 15) || ((h.N + 3) & ~3) > h.ldb)
      return fgs::fail_arg("mlp_dw: operand rows must be padded to 4 floats");
    if (h.M > 2 * BM) return fgs::fail_arg("mlp_dw: more than 256 out-features");
    DwProblem& p = b.p[i];
    p.A = h.A; p.B = h.B; p.C = h.C; p.colsum = h.colsum;
    p.M = h.M; p.N = h.N; p.lda = h.lda; p.ldb = h.ldb; p.ldc = h.ldc;
    for (int c = 0; c < fgs::cdiv(h.N, TN); ++c) {
      if (nt2 >= kDwMaxTiles) return fgs::fail_arg("mlp_dw: more than 32 output tiles");
      b.tile_prob[nt2] = (unsigned char)i; b.tile_n[nt2] = (unsigned char)c;
      ++nt2;
    }
  }
  if (nt2 == 0) return 0;
  b.n_tiles = nt2; b.K = K; b.rows_dev = rows_dev; b.dbg = g_dbg_dw;
  static bool configured2 = false, configured2b = false;
  int rc2 = bf16 ? configure_once(mlp_dw2_kernel<true>, configured2b, "mlp_dw2 (bf16) smem attribute", kSmemBytesDw2B)
                 : configure_once(mlp_dw2_kernel<false>, configured2, "mlp_dw2 smem attribute", kSmemBytesDw2);
  if (rc2) return rc2;
  DwMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (bf16) {
    // A: bf16 planes of dZ, boxes of 64 out-features x 64 samples straight into the MN-major operand tile;
    // B: raw fp32 rows of X for the per-warp rings (4 samples x 64 features).  Columns beyond the matrix and rows beyond
    // the allocation read as zero.
    int first = -1;
    for (int i = 0; i < n; ++i) {
      const DwProblemHost& h = probs[i];
      if (h.M <= 0 || h.N <= 0) continue;
      if (!h.Ahi || !h.Amid || (h.ldp & 7) || h.colsum)
        return fgs::fail_arg("mlp_dw (bf16): A must be given as bf16 planes (ldp % 8 == 0) and colsum must be NULL");
      if ((rc2 = get_map(h.Ahi, h.ldp, K, h.ldp, 64, &maps.a_hi[i], true))) return rc2;
      if ((rc2 = get_map(h.Amid, h.ldp, K, h.ldp, 64, &maps.a_mid[i], true))) return rc2;
      if ((rc2 = get_map(h.B, h.ldb, K, h.ldb, 8, &maps.b[i], false, 64))) return rc2;
      if (first < 0) first = i;
    }
    for (int i = 0; i < kDwMaxProblems; ++i)
      if (i >= n || probs[i].M <= 0 || probs[i].N <= 0) { maps.a_hi[i] = maps.a_hi[first]; maps.a_mid[i] = maps.a_mid[first]; maps.b[i] = maps.b[first]; }
  }
  // every pair should own at least ~8 slabs, else the epilogues dominate
  const long long units2 = (long long)nt2 * fgs::cdiv(K, bf16 ? 2 * BK : BK);
  int pairs_n = (FGS_NUM_SMS - g_dw_sm_reserve) / 2;
  if (units2 < 8LL * pairs_n) pairs_n = (int)((units2 + 7) / 8);
  if (pairs_n < 1) pairs_n = 1;
  if (pairs_n >= 2 * nt2) {
    // sibling-aligned cut: q pairs per tile, the same q for all tiles of a problem; the pairs left over go, one per
    // tile, to the problems with the fewest pairs per tile that still fit
    int q[kDwMaxProblems] = {0}, tiles_of[kDwMaxProblems] = {0};
    for (int t = 0; t < nt2; ++t) ++tiles_of[b.tile_prob[t]];
    const int base = pairs_n / nt2;
    int left = pairs_n - base * nt2;
    for (int i = 0; i < n; ++i) q[i] = tiles_of[i] ? base : 0;
    for (bool again = true; again && left > 0;) {
      again = false;
      for (int i = 0; i < n && left > 0; ++i)
        if (tiles_of[i] && q[i] == base && tiles_of[i] <= left) { ++q[i]; left -= tiles_of[i]; again = true; }
    }
    int np = 0;
    for (int t = 0; t < nt2; ++t)
      for (int r = 0; r < q[b.tile_prob[t]]; ++r) {
        b.pair_tile[np] = (unsigned char)t; b.pair_rank[np] = (unsigned char)r; b.pair_q[np] = (unsigned char)q[b.tile_prob[t]];
        ++np;
      }
    pairs_n = np;
    b.aligned = 1;
  }
  if (bf16) mlp_dw2_kernel<true><<<2 * pairs_n, kThreadsTn, kSmemBytesDw2B, st>>>(b, maps);
  else mlp_dw2_kernel<false><<<2 * pairs_n, kThreadsTn, kSmemBytesDw2, st>>>(b, maps);
  return fgs::check_launch("mlp_dw", st);
}

// C[M,N] += A[K,M]^T B[K,N] over the first min(K, *rows_dev) rows; colsum[m] += sum_k A[k][m] (may be NULL)
int gemm_tn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            float* colsum, const int32_t* rows_dev, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  DwProblemHost p{A, B, C, colsum, M, N, lda, ldb, ldc};
  return mlp_dw(&p, 1, K, rows_dev, st);
}

void set_dw_sm_reserve(int n_sms) { g_dw_sm_reserve = n_sms; }

// ---- weight splitting ------------------------------------------------------------------------
int weight_split_jobs(const SplitJobHost* jobs, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n > 16) return fgs::fail_arg("at most 16 weight split jobs per launch");
  SplitJobs J;
  J.n = n;
  for (int i = 0; i < n; ++i) {
    J.j[i].W = jobs[i].W; J.j[i].hi = jobs[i].hi; J.j[i].lo = jobs[i].lo;
    J.j[i].rows = jobs[i].rows; J.j[i].cols = jobs[i].cols; J.j[i].ld = jobs[i].ld;
    J.j[i].ld_out = jobs[i].ld_out; J.j[i].transpose = jobs[i].transpose; J.j[i].bf16 = jobs[i].bf16;
  }
  int tiles = 1;
  for (int i = 0; i < n; ++i) tiles = max(tiles, ((jobs[i].rows + 31) / 32) * ((jobs[i].cols + 31) / 32));
  weight_split_kernel<<<dim3(min(tiles, 4 * FGS_NUM_SMS), n), 256, 0, st>>>(J);
  return fgs::check_launch("weight_split", st);
}

}  // namespace fine

// Stand-alone entry for tests: B is a plain fp32 matrix ([N][ldb] for mode 0, [K][ldb] for mode 1);
// it is split into an internal scratch buffer first.
extern "C" int fgs_gemm_tcw(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                            float* C, int ldc, const float* bias, int relu, const uint32_t* bits_in,
                            uint32_t* bits_out, const int32_t* rows_dev, fgs_stream_t stream) {
  FGS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
  FGS_REQUIRE(mode == 0 || mode == 1, "fgs_gemm_tcw: mode must be 0 (B [N][K]) or 1 (B [K][N])");
  if (M == 0 || N == 0 || K == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int kp = (K + 3) & ~3, np = (N + 7) & ~7;
  const size_t need = 2 * (size_t)np * kp;
  if (need > g_scratch_floats) {
    if (g_scratch) {
      cudaStreamSynchronize(st);
      cudaFree(g_scratch);
      g_maps.clear();
    }
    int rc = fgs::check_cuda(cudaMalloc(&g_scratch, need * sizeof(float)), "gemm_tcw scratch");
    if (rc) return rc;
    g_scratch_floats = need;
  }
  int rc = fgs::check_cuda(cudaMemsetAsync(g_scratch, 0, need * sizeof(float), st), "gemm_tcw scratch memset");
  if (rc) return rc;
  fine::SplitJobHost job;
  job.W = B; job.hi = g_scratch; job.lo = g_scratch + (size_t)np * kp;
  job.ld = ldb; job.ld_out = kp; job.bf16 = 0;
  if (mode == 0) { job.rows = N; job.cols = K; job.transpose = 0; }
  else { job.rows = K; job.cols = N; job.transpose = 1; }
  if ((rc = fine::weight_split_jobs(&job, 1, st))) return rc;
  return fine::gemm_w(M, N, K, A, lda, job.hi, job.lo, kp, C, ldc, bias, relu, bits_in, bits_out, rows_dev, st);
}

// Stand-alone entries of the 3xBF16 kernels the fused step's BACKWARD runs (tests): one step of the chain kernel
// (C = A B^T for mode 0 / A B for mode 1, optional mask bits) and the weight-gradient product.
namespace {
// scratch for the stand-alone bf16 entries: weight planes + operand planes + output planes
void* g_scratch2 = nullptr;
size_t g_scratch2_bytes = 0;
int ensure_scratch2(size_t bytes, cudaStream_t st) {
  if (bytes <= g_scratch2_bytes) return 0;
  if (g_scratch2) {
    cudaStreamSynchronize(st);
    cudaFree(g_scratch2);
    g_maps.clear();
  }
  int rc = fgs::check_cuda(cudaMalloc(&g_scratch2, bytes), "bf16 scratch");
  if (rc) return rc;
  g_scratch2_bytes = bytes;
  return 0;
}
__global__ void __launch_bounds__(256)
planes_to_fp32_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ mid, int ldp, int rows, int cols,
                      float* __restrict__ C, int ldc, const int* __restrict__ rows_dev) {
  int valid = rows;
  if (rows_dev) valid = min(valid, *rows_dev);
  const long long total = (long long)valid * cols;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    C[(size_t)r * ldc + c] = __uint_as_float((uint32_t)hi[(size_t)r * ldp + c] << 16) + __uint_as_float((uint32_t)mid[(size_t)r * ldp + c] << 16);
  }
}
}  // namespace

// Stand-alone entries of the 3xBF16 kernels the fused step's BACKWARD runs (tests): one step of the planes chain
// (C = A B^T for mode 0 / A B for mode 1, optional mask bits, fp32 or plane output) and the weight-gradient product.
// A is split into bf16 planes by a pre-pass (inside the fused step the producing kernels write the planes).
extern "C" int fgs_gemm_tcw_bf16(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                                 float* C, int ldc, const uint32_t* bits_in, float* colsum, int via_planes,
                                 const int32_t* rows_dev, fgs_stream_t stream) {
  FGS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
  FGS_REQUIRE(mode == 0 || mode == 1, "fgs_gemm_tcw_bf16: mode must be 0 (B [N][K]) or 1 (B [K][N])");
  FGS_REQUIRE(N <= 256, "fgs_gemm_tcw_bf16: one 256-column block");
  FGS_REQUIRE(!via_planes || (ldc >= N), "bad ldc");
  if (M == 0 || N == 0 || K == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int kp = (K + 7) & ~7, np = (N + 7) & ~7;
  const size_t wbytes = (size_t)np * kp * 4, abytes = (size_t)M * kp * 4, cbytes = (size_t)M * 256 * 4;
  int rc = ensure_scratch2(wbytes + abytes + cbytes + 1024, st);
  if (rc) return rc;
  if ((rc = fgs::check_cuda(cudaMemsetAsync(g_scratch2, 0, wbytes + abytes + cbytes, st), "bf16 scratch memset"))) return rc;
  char* base = (char*)g_scratch2;
  fine::SplitJobHost job;
  job.W = B; job.hi = (float*)base; job.lo = (float*)(base + wbytes / 2);
  job.ld = ldb; job.ld_out = kp; job.bf16 = 1;
  if (mode == 0) { job.rows = N; job.cols = K; job.transpose = 0; }
  else { job.rows = K; job.cols = N; job.transpose = 1; }
  if ((rc = fine::weight_split_jobs(&job, 1, st))) return rc;
  uint16_t* ahi = (uint16_t*)(base + wbytes);
  uint16_t* amid = (uint16_t*)(base + wbytes + abytes / 2);
  if ((rc = fine::split_planes(A, lda, M, K, ahi, amid, kp, nullptr, nullptr, st))) return rc;
  uint16_t* chi = (uint16_t*)(base + wbytes + abytes);
  uint16_t* cmid = (uint16_t*)(base + wbytes + abytes + cbytes / 2);
  fine::BwStepHost s{};
  s.Ahi = ahi; s.Amid = amid; s.lda = kp; s.K = K; s.Bh = job.hi; s.Bl = job.lo; s.ldb = kp;
  s.C = C; s.ldc = ldc; s.Chi = chi; s.Cmid = cmid; s.out_planes = via_planes ? 1 : 0;
  s.N = N; s.n0 = 0; s.bits_in = bits_in; s.colsum = colsum;
  if ((rc = fine::mlp_bwd_planes(&s, 1, M, rows_dev, st))) return rc;
  if (via_planes) {
    planes_to_fp32_kernel<<<fgs::stream_grid((long long)M * N, 256, 8), 256, 0, st>>>(chi, cmid, 256, M, N, C, ldc, rows_dev);
    return fgs::check_launch("planes_to_fp32", st);
  }
  return 0;
}

extern "C" int fgs_gemm_tn_bf16(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                                int ldc, float* colsum, const int32_t* rows_dev, fgs_stream_t stream) {
  FGS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int mp = (M + 7) & ~7;
  const size_t abytes = (size_t)K * mp * 4;
  int rc = ensure_scratch2(abytes + 1024, st);
  if (rc) return rc;
  uint16_t* ahi = (uint16_t*)g_scratch2;
  uint16_t* amid = (uint16_t*)((char*)g_scratch2 + abytes / 2);
  if ((rc = fine::split_planes(A, lda, K, M, ahi, amid, mp, colsum, rows_dev, st))) return rc;
  fine::DwProblemHost p{A, B, C, nullptr, M, N, lda, ldb, ldc};
  p.Ahi = ahi; p.Amid = amid; p.ldp = mp;
  return fine::mlp_dw(&p, 1, K, rows_dev, st, true);
}

// Debug: device buffer of 2 int64 that the next gemm_w launches fill with the clock cycles consumer 0 of
// CTA (0,0) spent draining chunks (incl. waiting for the tensor core) and in epilogues (NULL switches it off).
extern "C" int fgs_debug_gemm_phases(void* buf) {
  g_dbg = (long long*)buf;
  g_dbg_dw = buf ? (long long*)buf + 10 : nullptr;    // slots 10..19: the weight-gradient kernel
  return 0;
}

// Stand-alone entry for tests: C[M,N] += A[K,M]^T B[K,N] (warp-specialised split-K kernel).
extern "C" int fgs_gemm_tn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                           int ldc, float* colsum, const int32_t* rows_dev, fgs_stream_t stream) {
  FGS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
  return fine::gemm_tn(M, N, K, A, lda, B, ldb, C, ldc, colsum, rows_dev, (cudaStream_t)stream);
}
