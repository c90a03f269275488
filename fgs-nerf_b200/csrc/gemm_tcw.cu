// Tensor-core GEMMs "activations x weights" of the colour MLPs (forward  Y = X W^T + b  and
// backward-data  dX = dZ W): the weight operand arrives PRE-SPLIT and by TMA (sm_100a).
//
// Same numerics as gemm_tc.cu (3xTF32 split, 32-k accumulation chunks in alternating TMEM regions,
// fp32 register accumulators), different data movement -- each choice below answers a measured
// cost of the previous version (tools/gemm_phase_probe.py, per-phase clock64 of one CTA):
//   * the 0.45 M MLP weights are split ONCE per step into hi / lo TF32 planes, straight ([out][in],
//     forward) and transposed ([in][out], dZ*W) -- weight_split_kernel, one launch -- instead of
//     being re-split from fp32 by every CTA for every k-slab;
//   * the B (weight) tiles are fetched by TMA: one elected thread issues two
//     cp.async.bulk.tensor.4d per k-slab (hi, lo).  The 4-D tensor map views a [rows][ld] plane as
//     (4 floats, 8 rows, k-chunks, row-groups), so the box lands in shared memory exactly in the
//     canonical no-swizzle K-major UMMA layout (8 x 16-byte core matrices, LBO 128, SBO 512): no
//     registers, no LSU instructions (2048 16-byte cp.async per slab cost 0.65 us of issue time),
//     completion on an mbarrier (expect_tx), prefetched two slabs ahead across tile boundaries;
//   * only the A (activation) slab -- one third of the bytes -- takes the load / split / store
//     path, with a lane-per-row mapping whose 128-bit stores are bank-conflict free;
//   * persistent CTAs (one per SM) loop over the 128-row tiles of the sample dimension, whose
//     extent is read on the device; TMEM (512 columns) and mbarriers are set up once per CTA;
//   * ReLU masks travel as BITS: the forward epilogue packs (x > 0) of its 32 columns into one
//     word per thread, the backward epilogue reads 4 words instead of 128 floats of H;
//   * the epilogue is transposed through shared memory (32x33 floats per warp) so global stores
//     are full 128-byte lines (row-per-thread 16-byte stores 1 KB apart cost 11 us per tile).
#include <cuda.h>

#include <map>
#include <tuple>

#include "common.cuh"
#include "tc_api.cuh"

namespace {

constexpr int BM = 128, BK = 32;                    // one k-slab = 32 floats = one 128-byte swizzle row
constexpr int kTmemCols = 512;
constexpr uint32_t kATile = BM * BK * 4;             // 16 KB
// Tile width TN (output columns per CTA) is a template parameter: 256 (2 stages of 96 KB) or 128
// (3 stages of 64 KB, 64 accumulator registers per thread, 2x the CTAs -> finer wave quantisation).
constexpr int kChunkSlabs = 1;  // accumulation chunk = one 32-k slab (12 MMAs) per TMEM region

template <int TN>
struct Cfg {
  static constexpr uint32_t kBTile = TN * BK * 4;
  static constexpr uint32_t kStage = 2 * kATile + 2 * kBTile;      // A_hi, A_lo, B_hi, B_lo
  static constexpr int kStages = (TN == 256) ? 2 : 3;
  static constexpr uint32_t kStagingOff = kStages * kStage;       // 8 warps x 32 x 33 floats
  static constexpr uint32_t kStagingBytes = 8 * 32 * 33 * 4;
  static constexpr uint32_t kBarOff = kStagingOff + kStagingBytes;  // mbarriers + tmem slot
  static constexpr uint32_t kSmemBytes = kBarOff + 256;
  static constexpr int kAcc = TN / 2;                               // accumulator registers per worker thread
};


struct WArgs {
  const float* A; float* C;
  const float* bias; const uint32_t* bits_in; uint32_t* bits_out; const int* rows_dev;
  int M, N, K, lda, ldc, relu;
  long long* dbg;  // optional phase timing of CTA (0,0) thread 0 (tools/gemm_phase_probe.py)
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
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
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
};
struct SplitJobs {
  SplitJob j[16];
  int n;
};

__global__ void __launch_bounds__(256) weight_split_kernel(const __grid_constant__ SplitJobs jobs) {
  const SplitJob& J = jobs.j[blockIdx.y];
  const int total = J.rows * J.cols;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int r = i / J.cols, c = i - r * J.cols;
    const float x = J.W[(size_t)r * J.ld + c];
    const float h = rn_tf32(x), l = rn_tf32(x - h);
    const size_t o = J.transpose ? (size_t)c * J.ld_out + r : (size_t)r * J.ld_out + c;
    J.hi[o] = h;
    J.lo[o] = l;
  }
}

constexpr int kWorkers = 256;            // warps 0-7: A producers, TMEM drain, epilogue
constexpr int kThreadsWS = kWorkers + 128; // third warpgroup: warp 8 lane 0 issues TMA (B) and tcgen05.mma;
                                           // registers are re-balanced with setmaxnreg (40 vs 232 per thread)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Warp-specialised: the MMA/TMA thread and the 256 workers only meet at mbarriers
//   a_full[st]   workers -> MMA   (256 arrivals: the A hi/lo slab is stored and proxy-fenced)
//   b_full[st]   TMA     -> MMA   (expect_tx: B hi/lo bytes landed)
//   free[st]     MMA     -> all   (tcgen05.commit: the MMAs reading the stage are done)
//   r_full[r]    MMA     -> workers (tcgen05.commit at the end of a 32-k chunk)
//   r_empty[r]   workers -> MMA   (256 arrivals: the chunk has been added into registers)
// so the blocking MMA issue never stalls the producers, and the producers run up to kStages slabs
// ahead of the tensor core.
template <int TN>
__global__ void __launch_bounds__(kThreadsWS, 1)
gemm_w_kernel(const WArgs a, const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo) {
  constexpr uint32_t kBTile = Cfg<TN>::kBTile, kStage = Cfg<TN>::kStage, kStagingOff = Cfg<TN>::kStagingOff,
                     kBarOff = Cfg<TN>::kBarOff;
  constexpr int kStages = Cfg<TN>::kStages, kAcc = Cfg<TN>::kAcc, kPfB = kStages - 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  const uint32_t free_bar0 = smem + kBarOff;
  const uint32_t bfull_bar0 = free_bar0 + 8 * kStages;
  const uint32_t afull_bar0 = bfull_bar0 + 8 * kStages;
  const uint32_t rfull_bar0 = afull_bar0 + 8 * kStages;
  const uint32_t rempty_bar0 = rfull_bar0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + kBarOff + 8 * (3 * kStages + 4) + 8);
  float* staging = reinterpret_cast<float*>(smem_raw + kStagingOff);

  int M = a.M;
  if (a.rows_dev) M = min(M, *a.rows_dev);
  const int n0 = blockIdx.x * TN;
  const int n_tiles = (M + BM - 1) / BM;
  if ((int)blockIdx.y >= n_tiles) return;
  const int K = a.K;
  const int BN = min(TN, (a.N - n0 + 15) & ~15);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(free_bar0 + 8 * i, 1);
      mbar_init(bfull_bar0 + 8 * i, 1);
      mbar_init(afull_bar0 + 8 * i, kWorkers);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(rfull_bar0 + 8 * i, 1);
      mbar_init(rempty_bar0 + 8 * i, kWorkers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                         ((uint32_t)(BM >> 4) << 24);

  const int n_slabs = (K + BK - 1) / BK;
  const int n_chunks = (n_slabs + kChunkSlabs - 1) / kChunkSlabs;
  const int my_tiles = (n_tiles - 1 - (int)blockIdx.y) / (int)gridDim.y + 1;
  const uint32_t total_slabs = (uint32_t)my_tiles * (uint32_t)n_slabs;

  if (warp >= kWorkers / 32) {
    // ================= MMA + TMA issuer (one thread of the third warpgroup) =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == kWorkers / 32 && lane == 0) {
      auto issue_b = [&](uint32_t gs) {
        const uint32_t st = gs % kStages;
        const uint32_t b_hi = smem + st * kStage + 2 * kATile, b_lo = b_hi + kBTile;
        const uint32_t bar = bfull_bar0 + 8 * st;
        const int k0 = (int)(gs % (uint32_t)n_slabs) * BK;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * kBTile) : "memory");
        tma_load_2d(b_hi, &tm_hi, bar, k0, n0);
        tma_load_2d(b_lo, &tm_lo, bar, k0, n0);
      };
      for (uint32_t gs = 0; gs < (uint32_t)kPfB && gs < total_slabs; ++gs) issue_b(gs);
      for (uint32_t g = 0; g < total_slabs; ++g) {
        const uint32_t st = g % kStages;
        const uint32_t par = (g / kStages) & 1;
        const uint32_t kt = g % (uint32_t)n_slabs;
        const uint32_t cg = (g / (uint32_t)n_slabs) * (uint32_t)n_chunks + kt / kChunkSlabs;  // global chunk
        const bool first_in_chunk = (kt % kChunkSlabs) == 0;
        const bool last_in_chunk = (kt % kChunkSlabs) == kChunkSlabs - 1 || kt == (uint32_t)n_slabs - 1;
        mbar_wait(afull_bar0 + 8 * st, par);
        mbar_wait(bfull_bar0 + 8 * st, par);
        if (first_in_chunk && cg >= 2) mbar_wait(rempty_bar0 + 8 * (cg & 1), ((cg >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stage_base = smem + st * kStage;
        const uint32_t a_hi = stage_base, a_lo = stage_base + kATile;
        const uint32_t b_hi = stage_base + 2 * kATile, b_lo = b_hi + kBTile;
        const uint32_t d = tmem_d + (cg & 1) * 256;
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          // k-step j = bytes [32 j, 32 j + 32) of every 128-byte row (the swizzle is a function of the address)
          const uint64_t dah = make_desc(a_hi + j * 32), dal = make_desc(a_lo + j * 32);
          const uint64_t dbh = make_desc(b_hi + j * 32), dbl = make_desc(b_lo + j * 32);
          mma_tf32(d, dal, dbh, idesc, (!first_in_chunk || j != 0) ? 1u : 0u);
          mma_tf32(d, dah, dbl, idesc, 1);
          mma_tf32(d, dah, dbh, idesc, 1);
        }
        mma_commit(free_bar0 + 8 * st);
        if (last_in_chunk) mma_commit(rfull_bar0 + 8 * (cg & 1));
        const uint32_t gs = g + kPfB;   // B two slabs ahead; its stage was read by slab gs - kStages
        if (gs < total_slabs) {
          if (gs >= (uint32_t)kStages) mbar_wait(free_bar0 + 8 * (gs % kStages), ((gs / kStages) - 1) & 1);
          issue_b(gs);
        }
      }
    }
  } else {
    // ================= workers: A slabs, TMEM drain, epilogue =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
    const int col_half = (warp >> 2) * kAcc;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // A slots: thread -> row ar, 16-byte k-chunks ac, ac + 2, ac + 4, ac + 6 of the 128-byte row; chunk c
    // of row r lives at chunk position c ^ (r % 8), so the 8 rows of a quarter-warp hit 8 distinct
    // bank groups
    const int ar = tid & 127, ac = tid >> 7;
    const uint32_t arow = (uint32_t)(ar * 128);
    uint32_t g = 0, gc = 0;
    long long t_main = 0, t_epi = 0, t0 = 0;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;

    for (int tile = blockIdx.y; tile < n_tiles; tile += gridDim.y) {
      const int m0 = tile * BM;
      if (dbg_on) t0 = clock64();
      float acc[kAcc];
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;

      auto load_a = [&](float4 (&v)[4], int k0) {
        const int m = m0 + ar;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = k0 + 4 * (ac + 2 * i);
          v[i] = (m < M && k < K) ? ldg4(a.A + (size_t)m * a.lda + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto store_a = [&](const float4 (&v)[4], uint32_t stage_base) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = arow + (uint32_t)(((ac + 2 * i) ^ (ar & 7)) * 16);
          float4 hi, lo;
          hi.x = rn_tf32(v[i].x); hi.y = rn_tf32(v[i].y); hi.z = rn_tf32(v[i].z); hi.w = rn_tf32(v[i].w);
          lo.x = rn_tf32(v[i].x - hi.x); lo.y = rn_tf32(v[i].y - hi.y);
          lo.z = rn_tf32(v[i].z - hi.z); lo.w = rn_tf32(v[i].w - hi.w);
          st_shared_v4(stage_base + off, hi);
          st_shared_v4(stage_base + kATile + off, lo);
        }
      };
      auto drain = [&](uint32_t c) {   // c = global chunk index
        mbar_wait(rfull_bar0 + 8 * (c & 1), (c >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int q = 0; q < kAcc / 32; ++q) {
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
        mbar_arrive(rempty_bar0 + 8 * (c & 1));   // the region may be overwritten by chunk c + 2
      };

      float4 a0[4];
      load_a(a0, 0);
      for (int kt = 0; kt < n_slabs; ++kt, ++g) {
        const uint32_t st = g % kStages;
        const uint32_t stage_base = smem + st * kStage;
        const int cl = kt / kChunkSlabs;
        const bool last_in_chunk = (kt % kChunkSlabs) == kChunkSlabs - 1 || kt == n_slabs - 1;
        if (g >= (uint32_t)kStages) mbar_wait(free_bar0 + 8 * st, ((g / kStages) - 1) & 1);
        store_a(a0, stage_base);
        if (kt + 1 < n_slabs) load_a(a0, (kt + 1) * BK);   // in flight during the drain below
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(afull_bar0 + 8 * st);
        if (last_in_chunk && cl >= 1) drain(gc + cl - 1);
      }
      drain(gc + n_chunks - 1);
      gc += n_chunks;
      if (dbg_on) { const long long t1 = clock64(); t_main += t1 - t0; t0 = t1; }

      // ---- epilogue: registers -> (bias, ReLU, mask) -> 32x33 smem transpose -> 128-byte row stores ----
      {
        const int m = m0 + (warp & 3) * 32 + lane;               // the row this thread accumulated
        float* stg = staging + warp * (32 * 33);
#pragma unroll
        for (int q = 0; q < kAcc / 32; ++q) {
          const int c0 = col_half + 32 * q;
          if (c0 < BN) {
            const int col = n0 + c0 + lane;                      // the column this lane stores
            const int word = (n0 + c0) >> 5;
            const float bcol = (a.bias && col < a.N) ? __ldg(a.bias + col) : 0.f;
            uint32_t in_bits = 0xffffffffu, out_bits = 0;
            if (a.bits_in && m < M) in_bits = a.bits_in[(size_t)m * 8 + word];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = acc[32 * q + j] + __shfl_sync(0xffffffffu, bcol, j);
              if (a.relu) {
                if (x > 0.f) out_bits |= 1u << j;
                x = fmaxf(x, 0.f);
              }
              if (!((in_bits >> j) & 1u)) x = 0.f;
              stg[lane * 33 + j] = x;
            }
            if (a.bits_out && m < M) a.bits_out[(size_t)m * 8 + word] = out_bits;
            __syncwarp();
            if (col < a.N) {
              const int mr0 = m0 + (warp & 3) * 32;
              float* __restrict__ cp = a.C + (size_t)mr0 * a.ldc + col;
#pragma unroll 8
              for (int r = 0; r < 32; ++r)
                if (mr0 + r < M) cp[(size_t)r * a.ldc] = stg[r * 33 + lane];
            }
            __syncwarp();
          }
        }
      }
      if (dbg_on) { const long long t1 = clock64(); t_epi += t1 - t0; }
    }
    if (dbg_on) { a.dbg[0] = t_main; a.dbg[1] = t_epi; }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
  }
}

// ---- backward-weight GEMM:  C[M,N] += A[K,M]^T B[K,N]  (dW = dZ^T X), split over the samples ----------
// Same machinery (warp-specialised, SWIZZLE_128B tiles, 32-k accumulation chunks, register
// accumulators), but the reduction index is the SAMPLE dimension, which is the slow dimension of both
// row-major operands.  kind::tf32 cannot read MN-major tiles without the 128B_BASE32B swizzle, so the
// 256 workers transpose both operands on the fly: coalesced 32-bit loads along the feature dimension
// (a warp reads 128 contiguous bytes of one sample row), four consecutive samples of one feature form
// the 16-byte chunk that is split into hi / lo and stored to the swizzled K-major tile.  The bias
// gradient (column sums of dZ) rides along with the A loads.  Partial products of the K-splits are
// combined with 128-byte red.global.add rows (epilogue transposed through shared memory).
struct TnArgs {
  const float* A; const float* B; float* C; float* colsum; const int* rows_dev;
  int M, N, K, lda, ldb, ldc;
};

__global__ void __launch_bounds__(kThreadsWS, 1) gemm_tn_kernel(const TnArgs a) {
  constexpr int TN = 128;
  constexpr uint32_t kBTile = Cfg<TN>::kBTile, kStage = Cfg<TN>::kStage, kStagingOff = Cfg<TN>::kStagingOff,
                     kBarOff = Cfg<TN>::kBarOff;
  constexpr int kStages = Cfg<TN>::kStages, kAcc = Cfg<TN>::kAcc;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  const uint32_t free_bar0 = smem + kBarOff;
  const uint32_t afull_bar0 = free_bar0 + 8 * kStages;
  const uint32_t rfull_bar0 = afull_bar0 + 8 * kStages;
  const uint32_t rempty_bar0 = rfull_bar0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + kBarOff + 8 * (2 * kStages + 4) + 8);
  float* staging = reinterpret_cast<float*>(smem_raw + kStagingOff);

  int K = a.K;
  if (a.rows_dev) K = min(K, *a.rows_dev);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * TN;
  int chunk = (K + gridDim.z - 1) / gridDim.z;
  chunk = (chunk + BK - 1) / BK * BK;
  const int kbeg = blockIdx.z * chunk;
  const int kend = min(K, kbeg + chunk);
  if (kbeg >= kend) return;
  const int BN = min(TN, (a.N - n0 + 15) & ~15);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(free_bar0 + 8 * i, 1);
      mbar_init(afull_bar0 + 8 * i, kWorkers);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(rfull_bar0 + 8 * i, 1);
      mbar_init(rempty_bar0 + 8 * i, kWorkers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                         ((uint32_t)(BM >> 4) << 24);
  const int n_slabs = (kend - kbeg + BK - 1) / BK;

  if (warp >= kWorkers / 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == kWorkers / 32 && lane == 0) {
      for (int kt = 0; kt < n_slabs; ++kt) {
        const uint32_t st = (uint32_t)kt % kStages;
        mbar_wait(afull_bar0 + 8 * st, ((uint32_t)kt / kStages) & 1);
        if (kt >= 2) mbar_wait(rempty_bar0 + 8 * (kt & 1), ((kt >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stage_base = smem + st * kStage;
        const uint32_t a_hi = stage_base, a_lo = stage_base + kATile;
        const uint32_t b_hi = stage_base + 2 * kATile, b_lo = b_hi + kBTile;
        const uint32_t d = tmem_d + (uint32_t)(kt & 1) * 256;
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint64_t dah = make_desc(a_hi + j * 32), dal = make_desc(a_lo + j * 32);
          const uint64_t dbh = make_desc(b_hi + j * 32), dbl = make_desc(b_lo + j * 32);
          mma_tf32(d, dal, dbh, idesc, j != 0 ? 1u : 0u);
          mma_tf32(d, dah, dbl, idesc, 1);
          mma_tf32(d, dah, dbh, idesc, 1);
        }
        mma_commit(free_bar0 + 8 * st);
        mma_commit(rfull_bar0 + 8 * (kt & 1));
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
    const int col_half = (warp >> 2) * kAcc;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int ar = tid & 127, ac = tid >> 7;           // tile row (feature) and first 16-byte chunk
    const uint32_t arow = (uint32_t)(ar * 128);
    float acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;
    float csum = 0.f;

    // 4 consecutive samples k .. k+3 of feature (f0 + ar): one 16-byte chunk of the K-major tile
    auto load_t = [&](float4 (&v)[4], const float* __restrict__ src, int ld, int f0, int fmax, int k0) {
      const int f = f0 + ar;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + 4 * (ac + 2 * i);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < fmax) {
          const float* p = src + (size_t)k * ld + f;
          if (k < kend) x.x = __ldg(p);
          if (k + 1 < kend) x.y = __ldg(p + ld);
          if (k + 2 < kend) x.z = __ldg(p + 2 * (size_t)ld);
          if (k + 3 < kend) x.w = __ldg(p + 3 * (size_t)ld);
        }
        v[i] = x;
      }
    };
    auto store_t = [&](const float4 (&v)[4], uint32_t hi_base, uint32_t lo_base) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = arow + (uint32_t)(((ac + 2 * i) ^ (ar & 7)) * 16);
        float4 hi, lo;
        hi.x = rn_tf32(v[i].x); hi.y = rn_tf32(v[i].y); hi.z = rn_tf32(v[i].z); hi.w = rn_tf32(v[i].w);
        lo.x = rn_tf32(v[i].x - hi.x); lo.y = rn_tf32(v[i].y - hi.y);
        lo.z = rn_tf32(v[i].z - hi.z); lo.w = rn_tf32(v[i].w - hi.w);
        st_shared_v4(hi_base + off, hi);
        st_shared_v4(lo_base + off, lo);
      }
    };
    auto drain = [&](uint32_t c) {
      mbar_wait(rfull_bar0 + 8 * (c & 1), (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int q = 0; q < kAcc / 32; ++q) {
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
      mbar_arrive(rempty_bar0 + 8 * (c & 1));
    };

    float4 va[4], vb[4];
    load_t(va, a.A, a.lda, m0, a.M, kbeg);
    load_t(vb, a.B, a.ldb, n0, a.N, kbeg);
    for (int kt = 0; kt < n_slabs; ++kt) {
      const uint32_t st = (uint32_t)kt % kStages;
      const uint32_t stage_base = smem + st * kStage;
      if (kt >= kStages) mbar_wait(free_bar0 + 8 * st, (((uint32_t)kt / kStages) - 1) & 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) csum += (va[i].x + va[i].y) + (va[i].z + va[i].w);
      store_t(va, stage_base, stage_base + kATile);
      store_t(vb, stage_base + 2 * kATile, stage_base + 2 * kATile + kBTile);
      if (kt + 1 < n_slabs) {
        load_t(va, a.A, a.lda, m0, a.M, kbeg + (kt + 1) * BK);
        load_t(vb, a.B, a.ldb, n0, a.N, kbeg + (kt + 1) * BK);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(afull_bar0 + 8 * st);
      if (kt >= 1) drain((uint32_t)kt - 1);
    }
    drain((uint32_t)n_slabs - 1);

    if (a.colsum && blockIdx.x == 0 && m0 + ar < a.M && csum != 0.f) atomicAdd(a.colsum + m0 + ar, csum);

    // ---- epilogue: registers -> 32x33 smem transpose -> 128-byte red.global.add rows ----
    float* stg = staging + warp * (32 * 33);
#pragma unroll
    for (int q = 0; q < kAcc / 32; ++q) {
      const int c0 = col_half + 32 * q;
      if (c0 < BN) {
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = acc[32 * q + j];
        __syncwarp();
        const int col = n0 + c0 + lane;
        if (col < a.N) {
          const int mr0 = m0 + (warp & 3) * 32;
          float* __restrict__ cp = a.C + (size_t)mr0 * a.ldc + col;
#pragma unroll 8
          for (int r = 0; r < 32; ++r)
            if (mr0 + r < a.M) atomicAdd(cp + (size_t)r * a.ldc, stg[r * 33 + lane]);
        }
        __syncwarp();
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
  }
}

long long* g_dbg = nullptr;
int g_tile_n = 128;   // 128 or 256 (fgs_debug_gemm_tile_n)

template <int TN>
int configure_once() {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_w_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg<TN>::kSmemBytes);
    if (e != cudaSuccess) return fgs::check_cuda(e, "gemm_w smem attribute");
    configured = true;
  }
  return 0;
}

// 2-D tensor map of a K-major fp32 plane [rows_alloc][ld]; box = (32 floats, 256 rows) = one k-slab of
// 256 rows, written to shared memory as 128-byte rows with the 128B swizzle -- the UMMA K-major
// SWIZZLE_128B layout.  (A first version used a 4-D map producing the no-swizzle core-matrix layout;
// its 16-byte inner box made TMA the bottleneck: ~1.4 us per 32 KB slab.)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::map<std::tuple<const void*, int, int, int>, CUtensorMap> g_maps;

int get_map(const float* plane, int rows_alloc, int ld, int tn, CUtensorMap* out) {
  auto key = std::make_tuple((const void*)plane, rows_alloc, ld, tn);
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
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows_alloc};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};                           // bytes, dim 1
  const cuuint32_t box[2] = {BK, (cuuint32_t)tn};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)plane, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(fgs::g_err, sizeof(fgs::g_err), "fgs_b200: cuTensorMapEncodeTiled failed (%d) rows %d ld %d", (int)r,
             rows_alloc, ld);
    return FGS_E_BADARG;
  }
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
  // tile width: 128-wide tiles double the CTA count; pick them when that fills the last wave better
  const int n_tiles = fgs::cdiv(M, BM);
  const int tn = (N > 128 && g_tile_n != 256) ? 128 : 256;
  int rc = (tn == 128) ? configure_once<128>() : configure_once<256>();
  if (rc) return rc;
  CUtensorMap tm_hi, tm_lo;
  const int rows_alloc = (N + 7) & ~7;
  if ((rc = get_map(Bh, rows_alloc, ldb, tn, &tm_hi))) return rc;
  if ((rc = get_map(Bl, rows_alloc, ldb, tn, &tm_lo))) return rc;
  WArgs a{A, C, bias, bits_in, bits_out, rows_dev, M, N, K, lda, ldc, relu, g_dbg};
  const int nx = fgs::cdiv(N, tn);
  int ny = FGS_NUM_SMS / nx;
  if (ny > n_tiles) ny = n_tiles;
  if (ny < 1) ny = 1;
  if (tn == 128)
    gemm_w_kernel<128><<<dim3(nx, ny, 1), kThreadsWS, Cfg<128>::kSmemBytes, st>>>(a, tm_hi, tm_lo);
  else
    gemm_w_kernel<256><<<dim3(nx, ny, 1), kThreadsWS, Cfg<256>::kSmemBytes, st>>>(a, tm_hi, tm_lo);
  return fgs::check_launch("gemm_w", st);
}

// C[M,N] += A[K,M]^T B[K,N] over the first min(K, *rows_dev) rows; colsum[m] += sum_k A[k][m] (may be NULL)
int gemm_tn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            float* colsum, const int32_t* rows_dev, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg<128>::kSmemBytes);
    if (e != cudaSuccess) return fgs::check_cuda(e, "gemm_tn smem attribute");
    configured = true;
  }
  TnArgs a{A, B, C, colsum, rows_dev, M, N, K, lda, ldb, ldc};
  dim3 grid(fgs::cdiv(N, 128), fgs::cdiv(M, BM), 1);
  const int tiles = grid.x * grid.y;
  int splits = (FGS_NUM_SMS + tiles - 1) / tiles;
  const int max_splits = fgs::cdiv(K, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  grid.z = splits;
  gemm_tn_kernel<<<grid, kThreadsWS, Cfg<128>::kSmemBytes, st>>>(a);
  return fgs::check_launch("gemm_tn", st);
}

// ---- weight splitting ------------------------------------------------------------------------
int weight_split_jobs(const SplitJobHost* jobs, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n > 16) return fgs::fail_arg("at most 16 weight split jobs per launch");
  SplitJobs J;
  J.n = n;
  for (int i = 0; i < n; ++i) {
    J.j[i].W = jobs[i].W; J.j[i].hi = jobs[i].hi; J.j[i].lo = jobs[i].lo;
    J.j[i].rows = jobs[i].rows; J.j[i].cols = jobs[i].cols; J.j[i].ld = jobs[i].ld;
    J.j[i].ld_out = jobs[i].ld_out; J.j[i].transpose = jobs[i].transpose;
  }
  weight_split_kernel<<<dim3(32, n), 256, 0, st>>>(J);
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
  job.ld = ldb; job.ld_out = kp;
  if (mode == 0) { job.rows = N; job.cols = K; job.transpose = 0; }
  else { job.rows = K; job.cols = N; job.transpose = 1; }
  if ((rc = fine::weight_split_jobs(&job, 1, st))) return rc;
  return fine::gemm_w(M, N, K, A, lda, job.hi, job.lo, kp, C, ldc, bias, relu, bits_in, bits_out, rows_dev, st);
}

// Debug: device buffer of 2 int64 that the next gemm_w launches fill with the clock cycles worker 0 of
// CTA (0,0) spent in main loops and in epilogues (NULL switches it off).
extern "C" int fgs_debug_gemm_phases(void* buf) {
  g_dbg = (long long*)buf;
  return 0;
}

// Debug / tuning: output-tile width of the pre-split GEMM (128 or 256 columns per CTA).
extern "C" int fgs_debug_gemm_tile_n(int tn) {
  FGS_REQUIRE(tn == 128 || tn == 256, "tile width must be 128 or 256");
  g_tile_n = tn;
  return 0;
}

// Stand-alone entry for tests: C[M,N] += A[K,M]^T B[K,N] (warp-specialised split-K kernel).
extern "C" int fgs_gemm_tn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                           int ldc, float* colsum, const int32_t* rows_dev, fgs_stream_t stream) {
  FGS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
  return fine::gemm_tn(M, N, K, A, lda, B, ldb, C, ldc, colsum, rows_dev, (cudaStream_t)stream);
}
