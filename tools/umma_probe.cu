// Layout discovery for tcgen05.mma kind::tf32 shared-memory descriptors (no swizzle).
// A (K-major, known-good) is the 8x8 identity in rows 0..7, so D[m][n] = B(n, k=m): with the B
// tile filled with its own float index, D reveals which smem element the hardware reads for
// logical (n, k) under a given (major, LBO, SBO).  Build: nvcc -arch=sm_100a -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// which: 0 = probe B, 1 = probe A (then B is the identity: D[m][n] = A(m, k=n) for n < 8)
__global__ void probe(int which, int mn_major, uint32_t lbo, uint32_t sbo, int div, float* out, uint32_t layout) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* At = reinterpret_cast<float*>(sm);            // 16 KB
  float* Bt = reinterpret_cast<float*>(sm + 16384);    // 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 32768 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  float* probed = which ? At : Bt;
  float* ident = which ? Bt : At;
  for (int i = tid; i < 4096; i += 128) {
    probed[i] = (float)((i / div) % 2048);
    ident[i] = 0.f;
  }
  __syncthreads();
  // identity, K-major canonical: elem(r,k) at (r/8)*512B + (k/4)*128B + (r%8)*16B + (k%4)*4B  (LBO 128, SBO 512)
  if (tid < 8) ident[(tid / 4) * 32 + tid * 4 + (tid % 4)] = 1.f;   // r = k = tid
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(slot)) : "memory");
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
  const int N = 16;
  if (tid == 0) {
    const uint32_t a_major = which ? mn_major : 0, b_major = which ? 0 : mn_major;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a_major << 15) | (b_major << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t d_id = make_desc(smem_u32(ident), 128, 512);
    const uint64_t d_pr = make_desc(smem_u32(probed), lbo, sbo, layout);
    const uint64_t da = which ? d_pr : d_id, db = which ? d_id : d_pr;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db),
                 "r"(idesc), "r"(0u)
                 : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 24) && !done; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) out[tid * 16 + j] = done ? __uint_as_float(r[j]) : -1.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 16 * 4);
  float h[128 * 16], h2[128 * 16];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  struct V { int which, mn; uint32_t lbo, sbo; const char* name; uint32_t layout; } vs[] = {
      {0, 0, 128, 512, "B K-major  lbo=128 sbo=512 (known good)"},
      {0, 1, 2048, 128, "B MN-major lbo=2048 sbo=128"},
      {0, 1, 128, 2048, "B MN-major lbo=128 sbo=2048"},
      {0, 1, 1024, 256, "B MN-major lbo=1024 sbo=256"},
      {0, 1, 256, 1024, "B MN-major lbo=256 sbo=1024"},
      {1, 1, 4096, 128, "A MN-major lbo=4096 sbo=128"},
      {1, 1, 128, 4096, "A MN-major lbo=128 sbo=4096"},
      {0, 1, 4096, 1024, "B MN-major SW128 (layout 2) lbo=4096 sbo=1024", 2},
      {0, 1, 1024, 4096, "B MN-major SW128 (layout 2) lbo=1024 sbo=4096", 2},
      {0, 1, 4096, 1024, "B MN-major SW128_BASE32B (layout 1) lbo=4096 sbo=1024", 1},
      {1, 1, 4096, 1024, "A MN-major SW128 (layout 2) lbo=4096 sbo=1024", 2},
      {1, 1, 1024, 4096, "A MN-major SW128 (layout 2) lbo=1024 sbo=4096", 2},
      {1, 1, 4096, 1024, "A MN-major SW128_BASE32B (layout 1) lbo=4096 sbo=1024", 1},
      {0, 0, 16, 1024, "B K-major SW128 (layout 2) sbo=1024 (known good)", 2},
  };
  for (auto& v : vs) {
    for (int pass = 0; pass < 2; ++pass) {
      probe<<<1, 128, 40000>>>(v.which, v.mn, v.lbo, v.sbo, pass ? 2048 : 1, d, v.layout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(pass ? h2 : h, d, sizeof(h), cudaMemcpyDeviceToHost);
    }
    printf("== %s : float index in the tile read for logical (mn, k)\n", v.name);
    if (v.which == 0) {   // D[m=k][n]
      for (int n = 0; n < 16; ++n) {
        printf("  n=%2d:", n);
        for (int k = 0; k < 8; ++k) printf(" %5d", (int)h[k * 16 + n] + 2048 * (int)h2[k * 16 + n]);
        printf("\n");
      }
    } else {              // D[m][n=k], show m = 0..15 and 32..35
      for (int m = 0; m < 36; ++m) {
        if (m >= 16 && m < 32) continue;
        printf("  m=%2d:", m);
        for (int k = 0; k < 8; ++k) printf(" %5d", (int)h[m * 16 + k] + 2048 * (int)h2[m * 16 + k]);
        printf("\n");
      }
    }
  }
  return 0;
}
