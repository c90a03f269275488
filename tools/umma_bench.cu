// tcgen05.mma kind::tf32 issue/throughput microbenchmark for the no-swizzle K-major layout.
// One CTA, one thread issues `iters` x 6 MMAs (M128 x N x K8) on fixed smem operands, commits, waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__global__ void bench(int N, uint32_t lbo, uint32_t sbo, uint32_t layout, int iters, int kind_f16, int a_tmem, int mn, long long* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 200000);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 200016);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 50000; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid == 0) {
    const uint32_t fmt = kind_f16 ? 1u : 2u;   // bf16 : tf32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) |
                           ((mn & 1) ? (1u << 15) : 0u) | ((mn & 2) ? (1u << 16) : 0u);
    const uint32_t base = smem_u32(sm);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint64_t da = make_desc(base + (j & 1) * 320 + (j >> 1) * 16384, lbo, sbo, layout);
        const uint64_t db = make_desc(base + 65536 + (j & 1) * 320 + (j >> 1) * 32768, lbo, sbo, layout);
        if (a_tmem) {   // A operand from TMEM (lane = row, 8 columns = one K8 step), D in columns [0, N)
          const uint32_t ta = tmem + 256 + (uint32_t)j * 8;
          if (kind_f16)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(1u) : "memory");
        } else if (kind_f16)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (it & 1) * 256), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (it & 1) * 256), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
      }
    }
    t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 28) && !done; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0; out[2] = done;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 64); long long h[3];
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 201000);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  struct V { int N; uint32_t lbo, sbo, layout; int f16; const char* name; int a_tmem; int mn; } vs[] = {
    {256, 128, 512, 0, 0, "tf32 N=256 no-swizzle lbo128 sbo512"},
    {256, 160, 640, 0, 0, "tf32 N=256 no-swizzle lbo160 sbo640"},
    {128, 160, 640, 0, 0, "tf32 N=128 no-swizzle lbo160 sbo640"},
    {256, 16, 1024, 2, 0, "tf32 N=256 swizzle128B (sbo 1024)"},
    {256, 128, 512, 0, 1, "bf16 N=256 K16 no-swizzle lbo128 sbo512"},
    {256, 16, 1024, 2, 1, "bf16 N=256 K16 swizzle128B"},
    {128, 16, 1024, 2, 0, "tf32 N=128 swizzle128B"},
    {64, 16, 1024, 2, 0, "tf32 N=64 swizzle128B"},
    {256, 16, 1024, 2, 0, "tf32 N=256 swizzle128B, A from TMEM", 1},
    {128, 16, 1024, 2, 0, "tf32 N=128 swizzle128B, A from TMEM", 1},
    {256, 16, 1024, 2, 1, "bf16 N=256 K16 swizzle128B, A from TMEM", 1},
    {128, 16, 1024, 2, 1, "bf16 N=128 K16 swizzle128B, A from TMEM", 1},
    {128, 16, 1024, 2, 1, "bf16 N=128 K16 swizzle128B"},
    {256, 16, 512, 4, 0, "tf32 N=256 swizzle64B (sbo 512) A and B", 0, 0},
    {128, 4096, 512, 1, 0, "tf32 N=128 A+B MN-major SW128_BASE32B", 0, 3},
    {256, 4096, 512, 1, 0, "tf32 N=256 A+B MN-major SW128_BASE32B", 0, 3},
    {128, 4096, 512, 1, 0, "tf32 N=128 A MN-major only", 0, 1},
    {128, 4096, 512, 1, 0, "tf32 N=128 B MN-major only", 0, 2},
  };
  for (auto& v : vs) {
    const int iters = 2000;
    bench<<<1, 128, 201000>>>(v.N, v.lbo, v.sbo, v.layout, iters, v.f16, v.a_tmem, v.mn, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    const double n = 6.0 * iters;
    const double flop = 2.0 * 128 * v.N * (v.f16 ? 16 : 8);
    printf("%-44s issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.2f TFLOP/s/SM at %d MHz (x148 = %.0f TF) done=%lld\n", v.name,
           h[0] / n, h[1] / n, flop / (h[1] / n) * clk * 1e3 / 1e12, clk / 1000, flop / (h[1] / n) * clk * 1e3 / 1e12 * 148, h[2]);
  }
  return 0;
}
