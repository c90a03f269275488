// Masked / dense / per-voxel-lr Adam as bandwidth-bound streaming kernels (sm_100a).
//
// Replaces model/cuda/adam_upd_kernel.cu:8-58 (three scalar one-thread-per-element kernels)
// and the host step_size of :72,:94,:117.  Here: 128-bit loads/stores, grid sized to the 148
// SMs with a grid-stride loop, and a multi-tensor variant that updates all MLP tensors in one
// launch (the reference's MaskedAdam.step issues one launch per tensor, model/adam.py:190-221).
// Algorithmic traffic: dense 28 B/param (p,g,m,v read; p,m,v written); masked 4 B/param where
// grad==0 (only g is read) -- the masked kernel loads p/m/v only inside the taken branch.
//
// Arithmetic is the reference's SASS sequence (so p,m,v are bit-identical):
//   m = fma(beta1, m, (1-beta1)*g);  v = fma(beta2, v, ((1-beta2)*g)*g);
//   p = p - (step_size*m) / (sqrt(v) + eps)        (IEEE sqrt and divide)
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kBlock = 256;

struct AdamConst {
  float step_size, beta1, beta2, omb1, omb2, eps;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamConst& c,
                                          float lr_scale) {
  m = __fmaf_rn(m, c.beta1, __fmul_rn(g, c.omb1));
  v = __fmaf_rn(v, c.beta2, __fmul_rn(g, __fmul_rn(g, c.omb2)));
  const float num = __fmul_rn(lr_scale, m);
  p = __fsub_rn(p, __fdiv_rn(num, __fadd_rn(__fsqrt_rn(v), c.eps)));
}

// MODE 0 dense, 1 masked (skip g==0), 2 per-voxel lr (dense)
template <int MODE>
__global__ void __launch_bounds__(kBlock)
adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
            float* __restrict__ exp_avg_sq, const float* __restrict__ perlr, long long n,
            AdamConst c) {
  const long long n4 = n >> 2;
  const long long tid = blockIdx.x * (long long)kBlock + threadIdx.x;
  const long long nth = (long long)gridDim.x * kBlock;
  for (long long i = tid; i < n4; i += nth) {
    const float4 g = ld_stream4(grad + 4 * i);
    if (MODE == 1 && g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    float4 p = ld4(param + 4 * i), m = ld4(exp_avg + 4 * i), v = ld4(exp_avg_sq + 4 * i);
    float4 l = make_float4(c.step_size, c.step_size, c.step_size, c.step_size);
    if (MODE == 2) {
      // kernel.cu:56: step_size * perlr[index] * exp_avg[index] -> (step_size*perlr)*m
      const float4 q = ld_stream4(perlr + 4 * i);
      l.x = __fmul_rn(c.step_size, q.x); l.y = __fmul_rn(c.step_size, q.y);
      l.z = __fmul_rn(c.step_size, q.z); l.w = __fmul_rn(c.step_size, q.w);
    }
    if (MODE != 1 || g.x != 0.f) adam_elem(p.x, g.x, m.x, v.x, c, l.x);
    if (MODE != 1 || g.y != 0.f) adam_elem(p.y, g.y, m.y, v.y, c, l.y);
    if (MODE != 1 || g.z != 0.f) adam_elem(p.z, g.z, m.z, v.z, c, l.z);
    if (MODE != 1 || g.w != 0.f) adam_elem(p.w, g.w, m.w, v.w, c, l.w);
    st4(param + 4 * i, p);
    st4(exp_avg + 4 * i, m);
    st4(exp_avg_sq + 4 * i, v);
  }
  // tail (< 4 elements)
  for (long long i = (n4 << 2) + tid; i < n; i += nth) {
    const float g = grad[i];
    if (MODE == 1 && g == 0.f) continue;
    float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    const float l = (MODE == 2) ? __fmul_rn(c.step_size, perlr[i]) : c.step_size;
    adam_elem(p, g, m, v, c, l);
    param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
  }
}

// Scalar fallback for pointers that are not 16-byte aligned (tensor views).
template <int MODE>
__global__ void __launch_bounds__(kBlock)
adam_kernel_scalar(float* __restrict__ param, const float* __restrict__ grad,
                   float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                   const float* __restrict__ perlr, long long n, AdamConst c) {
  for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < n;
       i += (long long)gridDim.x * kBlock) {
    const float g = grad[i];
    if (MODE == 1 && g == 0.f) continue;
    float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    const float l = (MODE == 2) ? __fmul_rn(c.step_size, perlr[i]) : c.step_size;
    adam_elem(p, g, m, v, c, l);
    param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
  }
}

struct MultiArgs {
  float* p[FGS_MAX_MT];
  const float* g[FGS_MAX_MT];
  float* m[FGS_MAX_MT];
  float* v[FGS_MAX_MT];
  long long n[FGS_MAX_MT];
  float step_size[FGS_MAX_MT];
  int blk_start[FGS_MAX_MT + 1];  // first CTA of each tensor
  int count;
  int zero_grad;                  // clear every gradient element once it has been consumed
  float beta1, beta2, omb1, omb2, eps;
};

constexpr int kMtChunk = kBlock * 4;  // elements per CTA in the multi-tensor kernel

__global__ void __launch_bounds__(kBlock) adam_multi_kernel(const __grid_constant__ MultiArgs a) {
  int t = 0;
  while (t + 1 < a.count && (int)blockIdx.x >= a.blk_start[t + 1]) ++t;
  const long long base = (long long)(blockIdx.x - a.blk_start[t]) * kMtChunk;
  AdamConst c{a.step_size[t], a.beta1, a.beta2, a.omb1, a.omb2, a.eps};
  float* __restrict__ P = a.p[t];
  float* __restrict__ G = const_cast<float*>(a.g[t]);
  float* __restrict__ M = a.m[t];
  float* __restrict__ V = a.v[t];
  const long long n = a.n[t];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = base + k * kBlock + threadIdx.x;
    if (i < n) {
      float p = P[i], m = M[i], v = V[i];
      adam_elem(p, G[i], m, v, c, c.step_size);
      P[i] = p; M[i] = m; V[i] = v;
      if (a.zero_grad) G[i] = 0.f;
    }
  }
}

// adam_upd_kernel.cu:72 -- all in float, as the reference's host code evaluates it
inline float host_step_size(float lr, float beta1, float beta2, int step) {
  return lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));
}

inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

template <int MODE>
int launch_adam(float* param, const float* grad, float* m, float* v, const float* perlr,
                int64_t n, const AdamConst& c, cudaStream_t st) {
  const bool vec = aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(v) &&
                   (MODE != 2 || aligned16(perlr));
  if (vec) {
    const int grid = fgs::stream_grid((n + 3) / 4, kBlock, 8);
    adam_kernel<MODE><<<grid, kBlock, 0, st>>>(param, grad, m, v, perlr, n, c);
  } else {
    const int grid = fgs::stream_grid(n, kBlock, 8);
    adam_kernel_scalar<MODE><<<grid, kBlock, 0, st>>>(param, grad, m, v, perlr, n, c);
  }
  return fgs::check_launch("adam_upd", st);
}

}  // namespace

extern "C" {

int fgs_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                 const float* perlr, int64_t n, int step, float beta1, float beta2, float lr,
                 float eps, int mode, fgs_stream_t stream) {
  FGS_REQUIRE(n >= 0, "n < 0");
  FGS_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0,1,2");
  FGS_REQUIRE(mode != 2 || perlr != nullptr, "perlr is NULL");
  if (n == 0) return 0;
  AdamConst c{host_step_size(lr, beta1, beta2, step), beta1, beta2, 1 - beta1, 1 - beta2, eps};
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case 0: return launch_adam<0>(param, grad, exp_avg, exp_avg_sq, perlr, n, c, st);
    case 1: return launch_adam<1>(param, grad, exp_avg, exp_avg_sq, perlr, n, c, st);
    default: return launch_adam<2>(param, grad, exp_avg, exp_avg_sq, perlr, n, c, st);
  }
}

static int adam_upd_multi_impl(int n_tensors, float* const* params, const float* const* grads,
                               float* const* exp_avgs, float* const* exp_avg_sqs, const int64_t* numels,
                               const int* steps, const float* lrs, float beta1, float beta2, float eps,
                               int zero_grad, fgs_stream_t stream);

int fgs_adam_upd_multi(int n_tensors, float* const* params, const float* const* grads,
                       float* const* exp_avgs, float* const* exp_avg_sqs, const int64_t* numels,
                       const int* steps, const float* lrs, float beta1, float beta2, float eps,
                       fgs_stream_t stream) {
  return adam_upd_multi_impl(n_tensors, params, grads, exp_avgs, exp_avg_sqs, numels, steps, lrs, beta1, beta2, eps, 0, stream);
}

int fgs_adam_upd_multi_zero_grad(int n_tensors, float* const* params, float* const* grads,
                                 float* const* exp_avgs, float* const* exp_avg_sqs, const int64_t* numels,
                                 const int* steps, const float* lrs, float beta1, float beta2, float eps,
                                 fgs_stream_t stream) {
  return adam_upd_multi_impl(n_tensors, params, grads, exp_avgs, exp_avg_sqs, numels, steps, lrs, beta1, beta2, eps, 1, stream);
}

static int adam_upd_multi_impl(int n_tensors, float* const* params, const float* const* grads,
                               float* const* exp_avgs, float* const* exp_avg_sqs, const int64_t* numels,
                               const int* steps, const float* lrs, float beta1, float beta2, float eps,
                               int zero_grad, fgs_stream_t stream) {
  FGS_REQUIRE(n_tensors >= 0 && n_tensors <= FGS_MAX_MT, "n_tensors out of range");
  if (n_tensors == 0) return 0;
  MultiArgs a;
  a.count = n_tensors;
  a.zero_grad = zero_grad;
  a.beta1 = beta1; a.beta2 = beta2; a.omb1 = 1 - beta1; a.omb2 = 1 - beta2; a.eps = eps;
  int blk = 0;
  for (int t = 0; t < n_tensors; ++t) {
    a.p[t] = params[t]; a.g[t] = grads[t]; a.m[t] = exp_avgs[t]; a.v[t] = exp_avg_sqs[t];
    a.n[t] = numels[t];
    a.step_size[t] = host_step_size(lrs[t], beta1, beta2, steps[t]);
    a.blk_start[t] = blk;
    blk += fgs::cdiv(numels[t] > 0 ? numels[t] : 0, kMtChunk);
  }
  a.blk_start[n_tensors] = blk;
  if (blk == 0) return 0;
  FGS_PREFER_MAX_SHARED(adam_multi_kernel);
  adam_multi_kernel<<<blk, kBlock, 0, (cudaStream_t)stream>>>(a);
  return fgs::check_launch("adam_upd_multi", (cudaStream_t)stream);
}

}  // extern "C"

// ---- fused-step optimiser kernels ---------------------------------------------------------------
// Same arithmetic as above (adam_elem), different traffic:
//  * adam_dense_zero_kernel: dense Adam that also clears the gradient it has just consumed, so the
//    next step needs no separate memset pass over the grid (32 B/param instead of 28 + 4 + launch).
//  * adam_bricks_kernel: MASKED Adam driven by the touched-brick flags that the backward scatter
//    leaves behind (1 byte per 4x4x4-voxel brick).  The reference's masked_adam_upd reads the whole
//    gradient grid to find the few percent of voxels with grad != 0 (model/adam.py:211-214,
//    adam_upd_kernel.cu:25-40) after a full-size zero fill; here only flagged bricks are read, the
//    per-element `grad != 0` test of the reference is kept inside them (bit-identical p, m, v), and
//    the consumed gradients and flags are cleared on the way out.
namespace {

__global__ void __launch_bounds__(kBlock)
adam_dense_zero_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ exp_avg,
                       float* __restrict__ exp_avg_sq, long long n, AdamConst c) {
  const long long n4 = n >> 2;
  const long long tid = blockIdx.x * (long long)kBlock + threadIdx.x;
  const long long nth = (long long)gridDim.x * kBlock;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = tid; i < n4; i += nth) {
    const float4 g = ld4(grad + 4 * i);       // plain load: this kernel also WRITES grad (ld.global.nc would be UB)
    float4 p = ld4(param + 4 * i), m = ld4(exp_avg + 4 * i), v = ld4(exp_avg_sq + 4 * i);
    adam_elem(p.x, g.x, m.x, v.x, c, c.step_size);
    adam_elem(p.y, g.y, m.y, v.y, c, c.step_size);
    adam_elem(p.z, g.z, m.z, v.z, c, c.step_size);
    adam_elem(p.w, g.w, m.w, v.w, c, c.step_size);
    st4(param + 4 * i, p);
    st4(exp_avg + 4 * i, m);
    st4(exp_avg_sq + 4 * i, v);
    st4(grad + 4 * i, z4);
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) {
    float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    adam_elem(p, grad[i], m, v, c, c.step_size);
    param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v; grad[i] = 0.f;
  }
}

// touched-brick list in ASCENDING brick order (deterministic, identical on every rank that holds the same flags); the
// flags are cleared on the way.  list[0] = count, list[1..] = brick ids, list[1 + n_bricks ..] = per-CTA counts (scratch).
// Two launches over chunks of 4096 flags (a thread owns 16 consecutive flags = one 16-byte load): brick_count_kernel
// counts per chunk, brick_fill_kernel sums the counts of the chunks in front of its own (<= a few hundred values) and
// writes its ids in order.  (The first version was ONE CTA walking all flags in passes of 16 K: 50 us at 160^3 and
// 110 us at 256^3 -- twice per data-parallel step since the SDF gradient also travels as bricks.)
constexpr int kListThreads = 256;
constexpr int kListRun = 16;
constexpr int kListChunk = kListThreads * kListRun;

// the thread's 16 flags as 4 words with bit 8 j + 7 of word i set <=> flag 4 i + j nonzero; returns the count
__device__ __forceinline__ int list_load_run(const unsigned char* __restrict__ flags, int beg, int n_bricks, bool vec,
                                             uint32_t (&m)[4]) {
  if (vec && beg + kListRun <= n_bricks) {
    const uint4 v = *reinterpret_cast<const uint4*>(flags + beg);
    m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = beg + 4 * i + j;
        if (b < n_bricks) x |= (uint32_t)flags[b] << (8 * j);
      }
      m[i] = x;
    }
  }
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = ((((m[i] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m[i]) & 0x80808080u);
    cnt += __popc(m[i]);
  }
  return cnt;
}

__global__ void __launch_bounds__(kListThreads)
brick_count_kernel(const unsigned char* __restrict__ flags, int n_bricks, int* __restrict__ counts) {
  __shared__ int s_warp[kListThreads / 32];
  const bool vec = (((uintptr_t)flags) & 15) == 0;
  uint32_t m[4];
  int cnt = list_load_run(flags, blockIdx.x * kListChunk + threadIdx.x * kListRun, n_bricks, vec, m);
  cnt = warp_sum(cnt);
  if (lane_id() == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < kListThreads / 32; ++i) t += s_warp[i];
    counts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kListThreads)
brick_fill_kernel(unsigned char* __restrict__ flags, int n_bricks, const int* __restrict__ counts, int* __restrict__ list) {
  __shared__ int s_warp[kListThreads / 32];
  __shared__ int s_carry;
  const int w = threadIdx.x >> 5, lane = lane_id();
  const bool vec = (((uintptr_t)flags) & 15) == 0;
  const int beg = blockIdx.x * kListChunk + threadIdx.x * kListRun;
  uint32_t m[4];
  const int cnt = list_load_run(flags, beg, n_bricks, vec, m);
  // ids in front of this chunk
  int part = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += kListThreads) part += counts[i];
  part = warp_sum(part);
  const int incl = warp_incl_scan(cnt);
  __syncthreads();
  if (lane == 0) s_warp[w] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < kListThreads / 32; ++i) t += s_warp[i];
    s_carry = t;
  }
  __syncthreads();
  const int carry = s_carry;
  __syncthreads();
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int v = lane < kListThreads / 32 ? s_warp[lane] : 0;
    const int sc = warp_incl_scan(v);
    if (lane < kListThreads / 32) s_warp[lane] = sc - v;
    if (lane == 31 && blockIdx.x == gridDim.x - 1) list[0] = carry + sc;      // last chunk: the total
  }
  __syncthreads();
  if (cnt) {
    int o = 1 + carry + incl - cnt + s_warp[w];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t x = m[i];
      while (x) {
        const int bit = __ffs(x) - 1;               // 7, 15, 23 or 31
        x &= x - 1;
        list[o++] = beg + 4 * i + (bit >> 3);
      }
    }
    if (vec && beg + kListRun <= n_bricks) {
      *reinterpret_cast<uint4*>(flags + beg) = make_uint4(0u, 0u, 0u, 0u);
    } else {
      for (int b = beg; b < min(beg + kListRun, n_bricks); ++b) flags[b] = 0;
    }
  }
}

// gather / scatter of listed k0 gradient bricks (768 floats each, channels-last [X][Y][Z][12]) to and
// from a compact [count][768] buffer: the payload of the sparse multi-GPU gradient exchange.
template <bool PACK>
__global__ void __launch_bounds__(kBlock)
brick_copy_kernel(float* __restrict__ grid, float* __restrict__ compact, const int* __restrict__ list, int max_bricks,
                  int X, int Y, int Z) {
  const int nbY = (Y + 3) >> 2, nbZ = (Z + 3) >> 2;
  const int lane = lane_id();
  const int wpb = kBlock / 32;
  const int count = min(list[0], max_bricks);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = blockIdx.x * wpb + (threadIdx.x >> 5); e < count; e += gridDim.x * wpb) {
    const int br = list[1 + e];
    const int bz = br % nbZ, by = (br / nbZ) % nbY, bx = br / (nbZ * nbY);
    float* __restrict__ c = compact + (size_t)e * 768;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int q = lane + 32 * i;
      const int row = q / 12, f4 = q - row * 12;
      const int x = 4 * bx + (row >> 2), y = 4 * by + (row & 3), z = 4 * bz + f4 / 3;
      const bool ok = x < X && y < Y && z < Z;
      const size_t off = (((size_t)x * Y + y) * Z + 4 * bz) * 12 + 4 * f4;
      if (PACK) st4(c + 4 * q, ok ? ld4(grid + off) : z4);
      else if (ok) st4(grid + off, ld4(c + 4 * q));
    }
  }
}

// the same for a single-channel grid [X][Y][Z] (the SDF gradient): 64 floats per brick, 16 rows of one float4; two bricks
// per warp
template <bool PACK>
__global__ void __launch_bounds__(kBlock)
brick_copy1_kernel(float* __restrict__ grid, float* __restrict__ compact, const int* __restrict__ list, int max_bricks,
                   int X, int Y, int Z) {
  const int nbY = (Y + 3) >> 2, nbZ = (Z + 3) >> 2;
  const int lane = lane_id(), half = lane >> 4, row = lane & 15;
  const int wpb = kBlock / 32;
  const int count = min(list[0], max_bricks);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e0 = 2 * (blockIdx.x * wpb + (threadIdx.x >> 5)); e0 < count; e0 += 2 * gridDim.x * wpb) {
    const int e = e0 + half;
    if (e >= count) continue;
    const int br = list[1 + e];
    const int bz = br % nbZ, by = (br / nbZ) % nbY, bx = br / (nbZ * nbY);
    const int x = 4 * bx + (row >> 2), y = 4 * by + (row & 3);
    const bool ok = x < X && y < Y;
    const size_t off = ((size_t)x * Y + y) * Z + 4 * bz;
    float* __restrict__ c = compact + (size_t)e * 64 + 4 * row;
    if (PACK) st4(c, ok ? ld4(grid + off) : z4);
    else if (ok) st4(grid + off, ld4(c));
  }
}

// grid layout [X][Y][Z][12] (channels-last); brick b = ((x>>2)*nbY + (y>>2))*nbZ + (z>>2).
// One warp per listed brick: 16 (x,y) rows of 4 z-voxels x 12 channels = 192 float4 = 6 per lane,
// all gradient loads issued before any dependent work.
constexpr int kBrickBlock = 128;
// (128 threads x ~98 registers: small enough to be resident next to a tensor-core CTA on a side stream)
__device__ __forceinline__ void adam_one_brick(float* __restrict__ param, float* __restrict__ grad,
                                               float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, int br,
                                               int X, int Y, int Z, int nbY, int nbZ, int lane, const AdamConst& c) {
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int bz = br % nbZ, by = (br / nbZ) % nbY, bx = br / (nbZ * nbY);
  size_t off[6];
  float4 g[6];
  bool ok[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int q = lane + 32 * i;
    const int row = q / 12, f4 = q - row * 12;
    const int x = 4 * bx + (row >> 2), y = 4 * by + (row & 3), z = 4 * bz + f4 / 3;
    ok[i] = x < X && y < Y && z < Z;
    off[i] = (((size_t)x * Y + y) * Z + 4 * bz) * 12 + 4 * f4;
    g[i] = ok[i] ? ld4(grad + off[i]) : z4;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    if (ok[i] && (g[i].x != 0.f || g[i].y != 0.f || g[i].z != 0.f || g[i].w != 0.f)) {
      float4 p = ld4(param + off[i]), m = ld4(exp_avg + off[i]), v = ld4(exp_avg_sq + off[i]);
      if (g[i].x != 0.f) adam_elem(p.x, g[i].x, m.x, v.x, c, c.step_size);
      if (g[i].y != 0.f) adam_elem(p.y, g[i].y, m.y, v.y, c, c.step_size);
      if (g[i].z != 0.f) adam_elem(p.z, g[i].z, m.z, v.z, c, c.step_size);
      if (g[i].w != 0.f) adam_elem(p.w, g[i].w, m.w, v.w, c, c.step_size);
      st4(param + off[i], p);
      st4(exp_avg + off[i], m);
      st4(exp_avg_sq + off[i], v);
      st4(grad + off[i], z4);
    }
  }
}

__global__ void __launch_bounds__(kBrickBlock)
adam_bricks_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ exp_avg,
                   float* __restrict__ exp_avg_sq, const int* __restrict__ count_ptr, const int* __restrict__ ids,
                   int X, int Y, int Z, AdamConst c) {
  const int nbY = (Y + 3) >> 2, nbZ = (Z + 3) >> 2;
  const int lane = lane_id();
  const int wpb = kBrickBlock / 32;
  const int count = *count_ptr;
  for (int e = blockIdx.x * wpb + (threadIdx.x >> 5); e < count; e += gridDim.x * wpb)
    adam_one_brick(param, grad, exp_avg, exp_avg_sq, ids[e], X, Y, Z, nbY, nbZ, lane, c);
}

// Single GPU: nobody needs the ORDERED list (and its single-CTA scan).  brick_ulist_kernel compacts the flags over all
// SMs in whatever order the warps arrive -- a warp reads 32 consecutive flags, one atomicAdd per warp reserves its
// slots -- clears them, and zeroes the OTHER parity's counter for the next step; adam_bricks_kernel then runs one warp
// per listed brick as before.  ulist = [count(parity 0), count(parity 1), ids...].  (Driving the Adam kernel from the
// flags directly left 2 K warps with 2.7 serial bricks each: 88 us against 25 us with one brick per warp.)
__global__ void __launch_bounds__(256)
brick_ulist_kernel(unsigned char* __restrict__ flags, int n_bricks, int* __restrict__ ulist, int parity) {
  const int lane = lane_id();
  const int groups = (n_bricks + 31) >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0) ulist[1 - parity] = 0;
  for (int gidx = blockIdx.x * 8 + (threadIdx.x >> 5); gidx < groups; gidx += gridDim.x * 8) {
    const int b = gidx * 32 + lane;
    const bool set = b < n_bricks && flags[b] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, set);
    if (m == 0) continue;
    int base = 0;
    if (lane == 0) base = atomicAdd(ulist + parity, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (set) {
      ulist[2 + base + __popc(m & ((1u << lane) - 1u))] = b;
      flags[b] = 0;
    }
  }
}

// ---- voxel-granular touched list (data-parallel exchange of g_k0) -------------------------------------------------
// Only ~23 % of the voxels of a touched 4^3 brick carry a gradient, so the collective ships voxels (48 bytes each),
// not bricks.  One byte per voxel, set by the backward scatter; the list is ASCENDING (identical on every rank that
// holds the same flags) and built by two launches over all SMs: per-CTA counts, then every CTA sums the counts in
// front of it, re-scans its own flags and writes its ids (flags cleared on the way).  list[0] = count.
constexpr int kVoxPerThread = 64, kVoxThreads = 256, kVoxPerCta = kVoxPerThread * kVoxThreads;

__device__ __forceinline__ int vox_load_masks(const unsigned char* __restrict__ flags, long long beg, long long n,
                                              bool vec, uint32_t (&m)[16]) {
  if (vec && beg + kVoxPerThread <= n) {
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(flags + beg);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = src[q];
      m[4 * q] = v.x; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long b = beg + 4 * i + j;
        if (b < n) x |= (uint32_t)flags[b] << (8 * j);
      }
      m[i] = x;
    }
  }
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    m[i] = ((((m[i] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m[i]) & 0x80808080u);   // bit 8 j + 7 <=> flag 4 i + j nonzero
    cnt += __popc(m[i]);
  }
  return cnt;
}

__global__ void __launch_bounds__(kVoxThreads)
voxel_count_kernel(const unsigned char* __restrict__ flags, long long n, int* __restrict__ cta_counts) {
  __shared__ int s_warp[kVoxThreads / 32];
  uint32_t m[16];
  const long long beg = (long long)blockIdx.x * kVoxPerCta + (long long)threadIdx.x * kVoxPerThread;
  int cnt = beg < n ? vox_load_masks(flags, beg, n, (((uintptr_t)flags) & 15) == 0, m) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane_id() == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kVoxThreads / 32; ++w) t += s_warp[w];
    cta_counts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kVoxThreads)
voxel_write_kernel(unsigned char* __restrict__ flags, long long n, const int* __restrict__ cta_counts,
                   int* __restrict__ list) {
  __shared__ int s_warp[kVoxThreads / 32];
  __shared__ int s_base;
  // ids in front of this CTA (a few hundred counts: one warp sums them)
  if (threadIdx.x < 32) {
    int t = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += 32) t += cta_counts[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) s_base = t;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) list[0] = t + cta_counts[blockIdx.x];
  }
  uint32_t m[16];
  const bool vec = (((uintptr_t)flags) & 15) == 0;
  const long long beg = (long long)blockIdx.x * kVoxPerCta + (long long)threadIdx.x * kVoxPerThread;
  const int cnt = beg < n ? vox_load_masks(flags, beg, n, vec, m) : 0;
  const int incl = warp_incl_scan(cnt);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) s_warp[w] = incl;
  __syncthreads();
  int before = 0;
  for (int i = 0; i < w; ++i) before += s_warp[i];
  if (cnt) {
    int o = 1 + s_base + before + incl - cnt;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t x = m[i];
      while (x) {
        const int bit = __ffs(x) - 1;
        x &= x - 1;
        list[o++] = (int)(beg + 4 * i + (bit >> 3));
      }
    }
    if (vec && beg + kVoxPerThread <= n) {
      uint4* __restrict__ dst = reinterpret_cast<uint4*>(flags + beg);
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = z;
    } else {
      for (long long b = beg; b < min(beg + (long long)kVoxPerThread, n); ++b) flags[b] = 0;
    }
  }
}

// ---- bit-packed variant: the union of the ranks' touched voxels travels as an all-gather of bit masks --------------
// (a MAX all-reduce of the byte flags moves 8x the bytes through a reduction kernel).  Per rank: byte flags -> one
// word per 32 voxels (flags cleared); after the all-gather every rank ORs the `world` masks, derives the touched-brick
// byte flags from the union (they drive the brick Adam) and builds the ascending voxel list from the union words.
__global__ void __launch_bounds__(kBlock)
voxel_bits_pack_kernel(unsigned char* __restrict__ flags, uint32_t* __restrict__ bits, long long n_words) {
  for (long long w = blockIdx.x * (long long)kBlock + threadIdx.x; w < n_words; w += (long long)gridDim.x * kBlock) {
    uint4* __restrict__ src = reinterpret_cast<uint4*>(flags + 32 * w);
    const uint4 a = src[0], b = src[1];
    const uint32_t x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t m = (((x[i] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x[i]) & 0x80808080u;   // bit 8 j + 7 <=> byte j != 0
      word |= (((m >> 7) & 1u) | ((m >> 14) & 2u) | ((m >> 21) & 4u) | ((m >> 28) & 8u)) << (4 * i);
    }
    bits[w] = word;
    if (word) {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      src[0] = z;
      src[1] = z;
    }
  }
}

constexpr int kBitsPerCtaWords = 2048;     // words (of 32 voxels) per CTA of the union / list kernels: 8 per thread

// pass 1: union word = OR over the ranks (stored into slice 0), brick flags, per-CTA voxel counts
__global__ void __launch_bounds__(kBlock)
voxel_bits_union_kernel(uint32_t* __restrict__ bits_all, int world, long long n_words, int Y, int Z,
                        unsigned char* __restrict__ brick_flags, int* __restrict__ cta_counts) {
  __shared__ int s_warp[kBlock / 32];
  const int nbY = (Y + 3) >> 2, nbZ = (Z + 3) >> 2;
  int cnt = 0;
  for (int i = 0; i < kBitsPerCtaWords / kBlock; ++i) {
    const long long w = (long long)blockIdx.x * kBitsPerCtaWords + i * kBlock + threadIdx.x;
    if (w >= n_words) break;
    uint32_t u = 0;
    for (int r = 0; r < world; ++r) u |= bits_all[(long long)r * n_words + w];
    bits_all[w] = u;
    cnt += __popc(u);
    if (u && brick_flags) {
      // a nibble = 4 consecutive z voxels starting at a multiple of 4 (Z % 4 == 0): one brick
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if ((u >> (4 * q)) & 0xfu) {
          const long long v = 32 * w + 4 * q;
          const int z = (int)(v % Z);
          const long long xy = v / Z;
          const int y = (int)(xy % Y), x = (int)(xy / Y);
          brick_flags[((x >> 2) * nbY + (y >> 2)) * nbZ + (z >> 2)] = 1;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane_id() == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kBlock / 32; ++w) t += s_warp[w];
    cta_counts[blockIdx.x] = t;
  }
}

// pass 2: ascending voxel ids from the union words (slice 0)
__global__ void __launch_bounds__(kBlock)
voxel_bits_list_kernel(const uint32_t* __restrict__ bits, long long n_words, const int* __restrict__ cta_counts,
                       int* __restrict__ list) {
  __shared__ int s_warp[kBlock / 32];
  __shared__ int s_base;
  if (threadIdx.x < 32) {
    int t = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += 32) t += cta_counts[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) s_base = t;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) list[0] = t + cta_counts[blockIdx.x];
  }
  // thread t owns the CONSECUTIVE words [8 t, 8 t + 8) of the CTA's range (ids stay ascending)
  constexpr int kPer = kBitsPerCtaWords / kBlock;
  uint32_t u[kPer];
  int cnt = 0;
  const long long w0 = (long long)blockIdx.x * kBitsPerCtaWords + (long long)threadIdx.x * kPer;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    u[i] = (w0 + i < n_words) ? bits[w0 + i] : 0u;
    cnt += __popc(u[i]);
  }
  const int incl = warp_incl_scan(cnt);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) s_warp[w] = incl;
  __syncthreads();
  int before = 0;
  for (int i = 0; i < w; ++i) before += s_warp[i];
  if (cnt) {
    int o = 1 + s_base + before + incl - cnt;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      uint32_t x = u[i];
      while (x) {
        const int bit = __ffs(x) - 1;
        x &= x - 1;
        list[o++] = (int)(32 * (w0 + i) + bit);
      }
    }
  }
}

// gather / scatter of the listed voxels' 12 gradient channels (3 float4 per voxel) to / from a compact buffer
template <bool PACK>
__global__ void __launch_bounds__(kBlock)
voxel_copy_kernel(float* __restrict__ grid, float* __restrict__ compact, const int* __restrict__ list, int max_voxels) {
  const int count = min(list[0], max_voxels);
  const long long total = 3LL * count;
  for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < total; q += (long long)gridDim.x * kBlock) {
    const int e = (int)(q / 3), j = (int)(q - 3LL * e);
    const size_t off = (size_t)list[1 + e] * 12 + 4 * j;
    if (PACK) st4(compact + 4 * q, ld4(grid + off));
    else st4(grid + off, ld4(compact + 4 * q));
  }
}

}  // namespace

extern "C" {

int fgs_adam_upd_zero_grad(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int step,
                           float beta1, float beta2, float lr, float eps, fgs_stream_t stream) {
  FGS_REQUIRE(n >= 0, "n < 0");
  FGS_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq),
              "fgs_adam_upd_zero_grad needs 16-byte aligned tensors");
  if (n == 0) return 0;
  AdamConst c{host_step_size(lr, beta1, beta2, step), beta1, beta2, 1 - beta1, 1 - beta2, eps};
  FGS_PREFER_MAX_SHARED(adam_dense_zero_kernel);
  adam_dense_zero_kernel<<<fgs::stream_grid((n + 3) / 4, kBlock, 8), kBlock, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, c);
  return fgs::check_launch("adam_dense_zero", (cudaStream_t)stream);
}

int64_t fgs_brick_list_ints(int X, int Y, int Z) {
  const int64_t nb = (int64_t)((X + 3) >> 2) * ((Y + 3) >> 2) * ((Z + 3) >> 2);
  return 1 + nb + (nb + kListChunk - 1) / kListChunk;       // count, ids, per-chunk counts (scratch)
}

// brick_flags -> brick_list (count + ids), flags cleared
int fgs_brick_list(uint8_t* brick_flags, int X, int Y, int Z, int32_t* brick_list, fgs_stream_t stream) {
  FGS_REQUIRE(X > 0 && Y > 0 && Z > 0, "bad grid size");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_bricks = ((X + 3) >> 2) * ((Y + 3) >> 2) * ((Z + 3) >> 2);
  FGS_PREFER_MAX_SHARED(brick_count_kernel);
  FGS_PREFER_MAX_SHARED(brick_fill_kernel);
  const int chunks = fgs::cdiv(n_bricks, kListChunk);
  int* counts = brick_list + 1 + n_bricks;
  brick_count_kernel<<<chunks, kListThreads, 0, st>>>(brick_flags, n_bricks, counts);
  int rc = fgs::check_launch("brick_list", st);
  if (rc) return rc;
  brick_fill_kernel<<<chunks, kListThreads, 0, st>>>(brick_flags, n_bricks, counts, brick_list);
  return fgs::check_launch("brick_list", st);
}

// compact[e][768] <- grad bricks list[1+e] (pack != 0), or grad bricks <- compact (pack == 0); e < max_bricks
int64_t fgs_voxel_list_ints(int64_t n_voxels) {
  return 1 + n_voxels + (n_voxels + kVoxPerCta - 1) / kVoxPerCta;      // count, ids, per-CTA counts (scratch at the end)
}

int fgs_voxel_list(uint8_t* voxel_flags, int64_t n_voxels, int32_t* voxel_list, fgs_stream_t stream) {
  FGS_REQUIRE(n_voxels > 0 && n_voxels < (1LL << 31), "bad voxel count");
  cudaStream_t st = (cudaStream_t)stream;
  const int ctas = (int)((n_voxels + kVoxPerCta - 1) / kVoxPerCta);
  int* counts = voxel_list + 1 + n_voxels;
  FGS_PREFER_MAX_SHARED(voxel_count_kernel);
  FGS_PREFER_MAX_SHARED(voxel_write_kernel);
  voxel_count_kernel<<<ctas, kVoxThreads, 0, st>>>(voxel_flags, n_voxels, counts);
  int rc = fgs::check_launch("voxel_count", st);
  if (rc) return rc;
  voxel_write_kernel<<<ctas, kVoxThreads, 0, st>>>(voxel_flags, n_voxels, counts, voxel_list);
  return fgs::check_launch("voxel_list", st);
}

int fgs_voxel_bits_pack(uint8_t* voxel_flags, uint32_t* bits, int64_t n_voxels, fgs_stream_t stream) {
  FGS_REQUIRE(n_voxels > 0 && (n_voxels % 32) == 0, "voxel bit masks need a voxel count that is a multiple of 32");
  FGS_REQUIRE((((uintptr_t)voxel_flags) & 15) == 0, "voxel_flags must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  FGS_PREFER_MAX_SHARED(voxel_bits_pack_kernel);
  voxel_bits_pack_kernel<<<fgs::stream_grid(n_voxels / 32, kBlock, 8), kBlock, 0, st>>>(voxel_flags, bits, n_voxels / 32);
  return fgs::check_launch("voxel_bits_pack", st);
}

int fgs_voxel_bits_union_list(uint32_t* bits_all, int world, int X, int Y, int Z, uint8_t* brick_flags,
                              int32_t* voxel_list, fgs_stream_t stream) {
  const int64_t n_voxels = (int64_t)X * Y * Z;
  FGS_REQUIRE(world >= 1 && X > 0 && Y > 0 && Z > 0 && (Z % 4) == 0 && (n_voxels % 32) == 0,
              "voxel bit masks need Z % 4 == 0 and a voxel count that is a multiple of 32");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_words = n_voxels / 32;
  const int ctas = (int)((n_words + kBitsPerCtaWords - 1) / kBitsPerCtaWords);
  int* counts = voxel_list + 1 + n_voxels;                     // scratch tail of fgs_voxel_list_ints()
  FGS_PREFER_MAX_SHARED(voxel_bits_union_kernel);
  FGS_PREFER_MAX_SHARED(voxel_bits_list_kernel);
  voxel_bits_union_kernel<<<ctas, kBlock, 0, st>>>(bits_all, world, n_words, Y, Z, brick_flags, counts);
  int rc = fgs::check_launch("voxel_bits_union", st);
  if (rc) return rc;
  voxel_bits_list_kernel<<<ctas, kBlock, 0, st>>>(bits_all, n_words, counts, voxel_list);
  return fgs::check_launch("voxel_list", st);
}

int fgs_voxel_copy(float* grid, float* compact, const int32_t* voxel_list, int max_voxels, int channels, int pack,
                   fgs_stream_t stream) {
  FGS_REQUIRE(channels == 12, "voxel copy is specialised for the 12-channel channels-last k0 grid");
  if (max_voxels <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid_dim = fgs::stream_grid(3LL * max_voxels, kBlock, 8);
  FGS_PREFER_MAX_SHARED(voxel_copy_kernel<true>);
  FGS_PREFER_MAX_SHARED(voxel_copy_kernel<false>);
  if (pack) voxel_copy_kernel<true><<<grid_dim, kBlock, 0, st>>>(grid, compact, voxel_list, max_voxels);
  else voxel_copy_kernel<false><<<grid_dim, kBlock, 0, st>>>(grid, compact, voxel_list, max_voxels);
  return fgs::check_launch(pack ? "voxel_pack" : "voxel_unpack", st);
}

int fgs_brick_copy(float* grid, float* compact, const int32_t* brick_list, int max_bricks, int X, int Y, int Z,
                   int channels, int pack, fgs_stream_t stream) {
  FGS_REQUIRE((channels == 12 || channels == 1) && X > 0 && Y > 0 && Z > 0 && (Z % 4) == 0,
              "brick copy: 12-channel (channels-last) or single-channel grid, Z % 4 == 0");
  if (max_bricks <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (channels == 1) {
    const int grid1 = fgs::stream_grid((max_bricks + 1) / 2, kBlock / 32, 8);
    FGS_PREFER_MAX_SHARED(brick_copy1_kernel<true>);
    FGS_PREFER_MAX_SHARED(brick_copy1_kernel<false>);
    if (pack) brick_copy1_kernel<true><<<grid1, kBlock, 0, st>>>(grid, compact, brick_list, max_bricks, X, Y, Z);
    else brick_copy1_kernel<false><<<grid1, kBlock, 0, st>>>(grid, compact, brick_list, max_bricks, X, Y, Z);
    return fgs::check_launch(pack ? "brick_pack" : "brick_unpack", st);
  }
  const int grid_dim = fgs::stream_grid(max_bricks, kBlock / 32, 8);
  FGS_PREFER_MAX_SHARED(brick_copy_kernel<true>);
  FGS_PREFER_MAX_SHARED(brick_copy_kernel<false>);
  if (pack) brick_copy_kernel<true><<<grid_dim, kBlock, 0, st>>>(grid, compact, brick_list, max_bricks, X, Y, Z);
  else brick_copy_kernel<false><<<grid_dim, kBlock, 0, st>>>(grid, compact, brick_list, max_bricks, X, Y, Z);
  return fgs::check_launch(pack ? "brick_pack" : "brick_unpack", st);
}

int fgs_adam_upd_bricks(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const int32_t* brick_list,
                        int X, int Y, int Z, int channels, int step, float beta1, float beta2, float lr,
                        float eps, fgs_stream_t stream) {
  FGS_REQUIRE(channels == 12, "brick Adam is specialised for the 12-channel channels-last k0 grid");
  FGS_REQUIRE(X > 0 && Y > 0 && Z > 0 && (Z % 4) == 0, "brick Adam needs Z % 4 == 0");
  AdamConst c{host_step_size(lr, beta1, beta2, step), beta1, beta2, 1 - beta1, 1 - beta2, eps};
  FGS_PREFER_MAX_SHARED(adam_bricks_kernel);
  adam_bricks_kernel<<<FGS_NUM_SMS * 16, kBrickBlock, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq,
                                                                          brick_list, brick_list + 1, X, Y, Z, c);
  return fgs::check_launch("adam_bricks", (cudaStream_t)stream);
}

int fgs_adam_upd_brick_flags(float* param, float* grad, float* exp_avg, float* exp_avg_sq, uint8_t* brick_flags,
                             int32_t* brick_ulist, int X, int Y, int Z, int channels, int step, float beta1,
                             float beta2, float lr, float eps, fgs_stream_t stream) {
  FGS_REQUIRE(channels == 12, "brick Adam is specialised for the 12-channel channels-last k0 grid");
  FGS_REQUIRE(X > 0 && Y > 0 && Z > 0 && (Z % 4) == 0, "brick Adam needs Z % 4 == 0");
  FGS_REQUIRE(brick_flags && brick_ulist, "brick flags / list missing");
  AdamConst c{host_step_size(lr, beta1, beta2, step), beta1, beta2, 1 - beta1, 1 - beta2, eps};
  const int n_bricks = ((X + 3) / 4) * ((Y + 3) / 4) * ((Z + 3) / 4);
  const int parity = step & 1;
  cudaStream_t st = (cudaStream_t)stream;
  brick_ulist_kernel<<<fgs::stream_grid((n_bricks + 31) / 32, 8, 4), 256, 0, st>>>(brick_flags, n_bricks, brick_ulist, parity);
  int rc = fgs::check_launch("brick_ulist", st);
  if (rc) return rc;
  adam_bricks_kernel<<<FGS_NUM_SMS * 16, kBrickBlock, 0, st>>>(param, grad, exp_avg, exp_avg_sq, brick_ulist + parity,
                                                              brick_ulist + 2, X, Y, Z, c);
  return fgs::check_launch("adam_bricks", st);
}

}  // extern "C"
