// Shared helpers for the fgs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fgs_b200.h"

#define FGS_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids below are sized in multiples of this

namespace fgs {

extern thread_local char g_err[256];
extern int64_t g_launches;
extern bool g_prof_on;

inline int fail_arg(const char* msg) {
  snprintf(g_err, sizeof(g_err), "fgs_b200: bad argument: %s", msg);
  return FGS_E_BADARG;
}

inline int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "fgs_b200: %s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}

// Optional per-launch timing (fgs_prof_begin/end): one event per launch on the launching stream.
void prof_mark(const char* what, cudaStream_t st);

inline int check_launch(const char* what, cudaStream_t st) {
  const int rc = check_launch(what);
  prof_mark(what, st);
  return rc;
}

// Kernels that may run on a side stream NEXT TO the persistent tensor-core CTAs (which carve ~214 KB of the SM's
// 256 KB out as shared memory) ask for the same carve-out: an SM cannot change its L1 / shared split while CTAs are
// resident, so a kernel with the default preference waited until the tensor-core kernel had drained (a 12 us
// brick-list scan took 137 us).  Streaming kernels lose nothing with a small L1.
template <typename Kern>
inline void prefer_max_shared(Kern kern) {
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
#define FGS_PREFER_MAX_SHARED(kern)                \
  do {                                             \
    static bool done_ = false;                     \
    if (!done_) { fgs::prefer_max_shared(kern); done_ = true; } \
  } while (0)

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "fgs_b200: %s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Grid for a grid-stride streaming kernel: enough CTAs to fill the chip a few times over,
// rounded to a multiple of the SM count, never more than the work needs.
inline int stream_grid(int64_t work_items, int per_block, int ctas_per_sm = 8) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)FGS_NUM_SMS * ctas_per_sm;
  if (need >= cap) return (int)cap;
  return (int)(need < 1 ? 1 : need);
}

}  // namespace fgs

#define FGS_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return fgs::fail_arg(msg); \
  } while (0)

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= o) v += n;
  }
  return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 128-bit streaming load / store (read-once data: keep it out of L1)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
