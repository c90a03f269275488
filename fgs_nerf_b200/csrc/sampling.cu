// Ray sampling, occupancy lookup and density->alpha kernels (sm_100a).
//
// Replaces the reference's render_utils sampling family
//   model/cuda/render_utils_kernel.cu:12-79   (t_min/t_max, n_samples, start/dir)
//   model/cuda/render_utils_kernel.cu:144-242 (seg-start marks, 2x cumsum, step_id, points)
//   model/cuda/render_utils_kernel.cu:245-360 (ndc / background samplers)
//   model/cuda/render_utils_kernel.cu:374-424 (maskcache_lookup)
//   model/cuda/render_utils_kernel.cu:431-574 (raw2alpha family)
// The reference issues 6 tiny kernels + 2 CUB scans + a D2H sync per sample_pts_on_rays call;
// here it is one count kernel (per-ray setup fused with a block scan and a last-block-done
// scan of the block totals) and one warp-per-ray fill kernel with coalesced stores.
//
// Float semantics: ids and masks must be bit-exact, so every expression keeps the exact
// operation/rounding sequence of the reference's SASS (FMA placement included) through
// explicit __f*_rn intrinsics -- the compiler may not re-associate or re-contract them.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

struct RayIsect {
  float t_min, t_max;
};

// kernel.cu:23-33
__device__ __forceinline__ RayIsect ray_aabb(float ox, float oy, float oz, float dx, float dy,
                                             float dz, const float* __restrict__ mn,
                                             const float* __restrict__ mx, float near_,
                                             float far_) {
  const float vx = (dx == 0.f) ? (float)1e-6 : dx;
  const float vy = (dy == 0.f) ? (float)1e-6 : dy;
  const float vz = (dz == 0.f) ? (float)1e-6 : dz;
  const float ax = __fdiv_rn(__fsub_rn(mx[0], ox), vx);
  const float ay = __fdiv_rn(__fsub_rn(mx[1], oy), vy);
  const float az = __fdiv_rn(__fsub_rn(mx[2], oz), vz);
  const float bx = __fdiv_rn(__fsub_rn(mn[0], ox), vx);
  const float by = __fdiv_rn(__fsub_rn(mn[1], oy), vy);
  const float bz = __fdiv_rn(__fsub_rn(mn[2], oz), vz);
  RayIsect r;
  r.t_min = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far_), near_);
  r.t_max = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far_), near_);
  return r;
}

// kernel.cu:48-51 / 68-71: the reference's SASS evaluates dx*dx+dy*dy+dz*dz as
// fma(dz,dz, fma(dx,dx, dy*dy)), then an IEEE sqrt.
__device__ __forceinline__ float ray_norm(float dx, float dy, float dz) {
  return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
}

// kernel.cu:53
__device__ __forceinline__ long long ray_nsamples(float t_min, float t_max, float rnorm,
                                                  float stepdist) {
  const float c = ceilf(__fdiv_rn(__fmul_rn(__fsub_rn(t_max, t_min), rnorm), stepdist));
  return (long long)fmax((double)c, 1.);
}

__global__ void infer_t_minmax_kernel(const float* __restrict__ rays_o,
                                      const float* __restrict__ rays_d,
                                      const float* __restrict__ mn, const float* __restrict__ mx,
                                      float near_, float far_, int n_rays,
                                      float* __restrict__ t_min, float* __restrict__ t_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  const float* o = rays_o + 3 * (size_t)i;
  const float* d = rays_d + 3 * (size_t)i;
  RayIsect r = ray_aabb(o[0], o[1], o[2], d[0], d[1], d[2], mn, mx, near_, far_);
  t_min[i] = r.t_min;
  t_max[i] = r.t_max;
}

__global__ void infer_n_samples_kernel(const float* __restrict__ rays_d,
                                       const float* __restrict__ t_min,
                                       const float* __restrict__ t_max, float stepdist,
                                       int n_rays, long long* __restrict__ n_samples) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  const float* d = rays_d + 3 * (size_t)i;
  n_samples[i] = ray_nsamples(t_min[i], t_max[i], ray_norm(d[0], d[1], d[2]), stepdist);
}

__global__ void infer_ray_start_dir_kernel(const float* __restrict__ rays_o,
                                           const float* __restrict__ rays_d,
                                           const float* __restrict__ t_min, int n_rays,
                                           float* __restrict__ start, float* __restrict__ dir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  const float* o = rays_o + 3 * (size_t)i;
  const float* d = rays_d + 3 * (size_t)i;
  const float rn = ray_norm(d[0], d[1], d[2]);
  const float t = t_min[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    start[3 * (size_t)i + c] = __fmaf_rn(d[c], t, o[c]);
    dir[3 * (size_t)i + c] = __fdiv_rn(d[c], rn);
  }
}

// ---- fused count: per-ray setup + block scan + last-block-done scan of block totals ------
// scratch layout: [0] ticket (u64, must be 0 on entry; left 0), [1..nblk] block totals ->
// exclusive block offsets.
__global__ void __launch_bounds__(kBlock)
sample_count_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                    const float* __restrict__ mn, const float* __restrict__ mx, float near_,
                    float far_, float stepdist, int n_rays, float* __restrict__ t_min,
                    float* __restrict__ t_max, long long* __restrict__ n_steps,
                    long long* __restrict__ ray_offset, unsigned long long* __restrict__ scratch) {
  __shared__ long long s_warp[kBlock / 32];
  __shared__ bool s_last;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  long long n = 0;
  if (i < n_rays) {
    const float* o = rays_o + 3 * (size_t)i;
    const float* d = rays_d + 3 * (size_t)i;
    const float dx = d[0], dy = d[1], dz = d[2];
    RayIsect r = ray_aabb(o[0], o[1], o[2], dx, dy, dz, mn, mx, near_, far_);
    n = ray_nsamples(r.t_min, r.t_max, ray_norm(dx, dy, dz), stepdist);
    t_min[i] = r.t_min;
    t_max[i] = r.t_max;
    n_steps[i] = n;
  }
  const long long incl = warp_incl_scan(n);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    long long v = (lane_id() < kBlock / 32) ? s_warp[lane_id()] : 0;
    long long sc = warp_incl_scan(v);
    if (lane_id() < kBlock / 32) s_warp[lane_id()] = sc - v;  // exclusive warp offsets
    if (lane_id() == kBlock / 32 - 1) {
      scratch[1 + blockIdx.x] = (unsigned long long)sc;  // block total
    }
  }
  __syncthreads();
  if (i < n_rays) ray_offset[i] = incl - n + s_warp[w];  // block-local exclusive offset
  // last block done -> scan the block totals
  __threadfence();
  if (threadIdx.x == 0) {
    unsigned long long t = atomicAdd(scratch, 1ull);
    s_last = (t == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int nblk = gridDim.x;
  long long carry = 0;
  for (int base = 0; base < nblk; base += kBlock) {
    const int b = base + threadIdx.x;
    long long v = (b < nblk) ? (long long)((volatile unsigned long long*)scratch)[1 + b] : 0;
    const long long in = warp_incl_scan(v);
    __syncthreads();
    if (lane_id() == 31) s_warp[w] = in;
    __syncthreads();
    long long woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kBlock / 32; ++k) {
      const long long sv = s_warp[k];
      if (k < w) woff += sv;
      tot += sv;
    }
    if (b < nblk) scratch[1 + b] = (unsigned long long)(carry + woff + in - v);
    carry += tot;
  }
  if (threadIdx.x == 0) {
    ray_offset[n_rays] = carry;
    scratch[0] = 0ull;
  }
}

__global__ void sample_offset_finalize_kernel(long long* __restrict__ ray_offset,
                                              const unsigned long long* __restrict__ scratch,
                                              int n_rays) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rays) ray_offset[i] += (long long)scratch[1 + i / kBlock];
}

// ---- fill: one warp per ray, lanes stride over the ray's steps ----------------------------
__global__ void __launch_bounds__(kBlock)
sample_fill_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                   const float* __restrict__ mn, const float* __restrict__ mx,
                   const float* __restrict__ t_min, const long long* __restrict__ ray_offset,
                   float stepdist, int n_rays, long long total_len, float* __restrict__ ray_pts,
                   unsigned char* __restrict__ mask_outbbox, long long* __restrict__ ray_id,
                   long long* __restrict__ step_id) {
  const int warps_per_block = kBlock / 32;
  const float mn0 = mn[0], mn1 = mn[1], mn2 = mn[2];
  const float mx0 = mx[0], mx1 = mx[1], mx2 = mx[2];
  for (int ray = blockIdx.x * warps_per_block + (threadIdx.x >> 5); ray < n_rays;
       ray += gridDim.x * warps_per_block) {
    const long long off = ray_offset[ray];
    long long end = ray_offset[ray + 1];
    if (end > total_len) end = total_len;
    const int n = (int)(end - off);
    const float* o = rays_o + 3 * (size_t)ray;
    const float* d = rays_d + 3 * (size_t)ray;
    const float dx = d[0], dy = d[1], dz = d[2];
    const float t = t_min[ray];
    const float rn = ray_norm(dx, dy, dz);
    // kernel.cu:72-77
    const float sx = __fmaf_rn(dx, t, o[0]), sy = __fmaf_rn(dy, t, o[1]), sz = __fmaf_rn(dz, t, o[2]);
    const float ux = __fdiv_rn(dx, rn), uy = __fdiv_rn(dy, rn), uz = __fdiv_rn(dz, rn);
    for (int s = lane_id(); s < n; s += 32) {
      // kernel.cu:184-192
      const float dist = __fmul_rn(stepdist, (float)s);
      const float px = __fmaf_rn(dist, ux, sx);
      const float py = __fmaf_rn(dist, uy, sy);
      const float pz = __fmaf_rn(dist, uz, sz);
      const long long idx = off + s;
      ray_pts[3 * idx + 0] = px;
      ray_pts[3 * idx + 1] = py;
      ray_pts[3 * idx + 2] = pz;
      mask_outbbox[idx] = (mn0 > px) | (mn1 > py) | (mn2 > pz) | (mx0 < px) | (mx1 < py) | (mx2 < pz);
      ray_id[idx] = ray;
      step_id[idx] = s;
    }
  }
}

// kernel.cu:245-270
__global__ void sample_ndc_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                  const float* __restrict__ mn, const float* __restrict__ mx,
                                  int n_samples, int n_rays, float* __restrict__ ray_pts,
                                  unsigned char* __restrict__ mask_outbbox) {
  const long long total = (long long)n_samples * n_rays;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i_ray = (int)(idx / n_samples);
    const int i_step = (int)(idx % n_samples);
    const float dist = __fdiv_rn((float)i_step, (float)(n_samples - 1));
    const float px = __fmaf_rn(rays_d[3 * (size_t)i_ray + 0], dist, rays_o[3 * (size_t)i_ray + 0]);
    const float py = __fmaf_rn(rays_d[3 * (size_t)i_ray + 1], dist, rays_o[3 * (size_t)i_ray + 1]);
    const float pz = __fmaf_rn(rays_d[3 * (size_t)i_ray + 2], dist, rays_o[3 * (size_t)i_ray + 2]);
    ray_pts[3 * idx + 0] = px;
    ray_pts[3 * idx + 1] = py;
    ray_pts[3 * idx + 2] = pz;
    mask_outbbox[idx] = (mn[0] > px) | (mn[1] > py) | (mn[2] > pz) | (mx[0] < px) | (mx[1] < py) | (mx[2] < pz);
  }
}

// kernel.cu:301-340 (inverse-sphere background warp); mixed float/double exactly as the
// reference's promotion rules and SASS give it.
__global__ void sample_bg_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                 const float* __restrict__ t_max, float bg_preserve, int n_samples,
                                 int n_rays, float* __restrict__ ray_pts) {
  const long long total = (long long)n_samples * n_rays;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i_ray = (int)(idx / n_samples);
    const int i_step = (int)(idx % n_samples);
    const float t_inner = t_max[i_ray];
    const float frac = __fdiv_rn((float)i_step, (float)n_samples);
    const float t_outer0 =
        (float)__dadd_rn(__dadd_rn((double)t_inner, -1.), __ddiv_rn(1., __dadd_rn(1., -(double)frac)));
    const float x = __fmaf_rn(rays_d[3 * (size_t)i_ray + 0], t_outer0, rays_o[3 * (size_t)i_ray + 0]);
    const float y = __fmaf_rn(rays_d[3 * (size_t)i_ray + 1], t_outer0, rays_o[3 * (size_t)i_ray + 1]);
    const float z = __fmaf_rn(rays_d[3 * (size_t)i_ray + 2], t_outer0, rays_o[3 * (size_t)i_ray + 2]);
    const float t_outer = ray_norm(x, y, z);
    const float m = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    const float R = __fdiv_rn(t_outer, m);
    const float q = __fdiv_rn(__fmul_rn(R, R), __fmul_rn(t_outer, t_outer));
    const float r = __fmul_rn(__fdiv_rn(R, t_outer), bg_preserve);
    const float o2i = (float)__fma_rn((double)q, __dadd_rn(1., -(double)bg_preserve), (double)r);
    ray_pts[3 * idx + 0] = __fmul_rn(x, o2i);
    ray_pts[3 * idx + 1] = __fmul_rn(y, o2i);
    ray_pts[3 * idx + 2] = __fmul_rn(z, o2i);
  }
}

// kernel.cu:374-392
__global__ void maskcache_lookup_kernel(const unsigned char* __restrict__ world,
                                        const float* __restrict__ xyz,
                                        const float* __restrict__ scale,
                                        const float* __restrict__ shift, int sz_i, int sz_j,
                                        int sz_k, long long n_pts, unsigned char* __restrict__ out) {
  const float s0 = scale[0], s1 = scale[1], s2 = scale[2];
  const float h0 = shift[0], h1 = shift[1], h2 = shift[2];
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pts;
       p += (long long)gridDim.x * blockDim.x) {
    const int i = (int)roundf(__fmaf_rn(xyz[3 * p + 0], s0, h0));
    const int j = (int)roundf(__fmaf_rn(xyz[3 * p + 1], s1, h1));
    const int k = (int)roundf(__fmaf_rn(xyz[3 * p + 2], s2, h2));
    unsigned char v = 0;
    if (0 <= i && i < sz_i && 0 <= j && j < sz_j && 0 <= k && k < sz_k)
      v = world[(size_t)i * sz_j * sz_k + (size_t)j * sz_k + k];
    out[p] = v;
  }
}

// kernel.cu:431-458
__global__ void raw2alpha_kernel(const float* __restrict__ density, float shift, float interval,
                                 const float* __restrict__ interval_pts, long long n_pts,
                                 float* __restrict__ exp_d, float* __restrict__ alpha) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pts;
       p += (long long)gridDim.x * blockDim.x) {
    const float e = expf(__fadd_rn(density[p], shift));  // can be inf
    const float iv = interval_pts ? interval_pts[p] : interval;
    exp_d[p] = e;
    alpha[p] = __fsub_rn(1.f, powf(__fadd_rn(1.f, e), -iv));
  }
}

// kernel.cu:507-530: min(exp,1e10) * pow(1+exp, -interval-1) * interval * grad_back; the
// `min(float, 1e10)` is evaluated in double and narrowed back by the float multiply chain.
__global__ void raw2alpha_backward_kernel(const float* __restrict__ exp_d,
                                          const float* __restrict__ grad_back, float interval,
                                          const float* __restrict__ interval_pts, long long n_pts,
                                          float* __restrict__ grad) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pts;
       p += (long long)gridDim.x * blockDim.x) {
    const float e = exp_d[p];
    const float iv = interval_pts ? interval_pts[p] : interval;
    const double a = fmin((double)e, 1e10);
    const double b = (double)powf(__fadd_rn(1.f, e), __fsub_rn(-iv, 1.f));
    grad[p] = (float)__dmul_rn(__dmul_rn(__dmul_rn(a, b), (double)iv), (double)grad_back[p]);
  }
}

}  // namespace

// ---- C-ABI --------------------------------------------------------------------------------
extern "C" {

int fgs_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min,
                       const float* xyz_max, float near_, float far_, int n_rays, float* t_min,
                       float* t_max, fgs_stream_t stream) {
  FGS_REQUIRE(n_rays >= 0, "n_rays < 0");
  if (n_rays == 0) return 0;
  infer_t_minmax_kernel<<<fgs::cdiv(n_rays, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, xyz_min, xyz_max, near_, far_, n_rays, t_min, t_max);
  return fgs::check_launch("infer_t_minmax");
}

int fgs_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max,
                        float stepdist, int n_rays, int64_t* n_samples, fgs_stream_t stream) {
  FGS_REQUIRE(n_rays >= 0, "n_rays < 0");
  if (n_rays == 0) return 0;
  infer_n_samples_kernel<<<fgs::cdiv(n_rays, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      rays_d, t_min, t_max, stepdist, n_rays, (long long*)n_samples);
  return fgs::check_launch("infer_n_samples");
}

int fgs_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min,
                            int n_rays, float* rays_start, float* rays_dir, fgs_stream_t stream) {
  FGS_REQUIRE(n_rays >= 0, "n_rays < 0");
  if (n_rays == 0) return 0;
  infer_ray_start_dir_kernel<<<fgs::cdiv(n_rays, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, t_min, n_rays, rays_start, rays_dir);
  return fgs::check_launch("infer_ray_start_dir");
}

int64_t fgs_sample_scratch_bytes(int n_rays) {
  return 8 * (int64_t)(1 + fgs::cdiv(n_rays > 0 ? n_rays : 1, kBlock));
}

int fgs_sample_pts_count(const float* rays_o, const float* rays_d, const float* xyz_min,
                         const float* xyz_max, float near_, float far_, float stepdist,
                         int n_rays, float* t_min, float* t_max, int64_t* n_steps,
                         int64_t* ray_offset, void* scratch, fgs_stream_t stream) {
  FGS_REQUIRE(n_rays >= 0, "n_rays < 0");
  FGS_REQUIRE(stepdist > 0.f, "stepdist must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rays == 0) {
    return fgs::check_cuda(cudaMemsetAsync(ray_offset, 0, 8, st), "sample_pts_count memset");
  }
  const int nblk = fgs::cdiv(n_rays, kBlock);
  sample_count_kernel<<<nblk, kBlock, 0, st>>>(rays_o, rays_d, xyz_min, xyz_max, near_, far_,
                                               stepdist, n_rays, t_min, t_max,
                                               (long long*)n_steps, (long long*)ray_offset,
                                               (unsigned long long*)scratch);
  int rc = fgs::check_launch("sample_count");
  if (rc) return rc;
  sample_offset_finalize_kernel<<<nblk, kBlock, 0, st>>>(
      (long long*)ray_offset, (const unsigned long long*)scratch, n_rays);
  return fgs::check_launch("sample_offset_finalize");
}

int fgs_sample_pts_fill(const float* rays_o, const float* rays_d, const float* xyz_min,
                        const float* xyz_max, const float* t_min, const int64_t* ray_offset,
                        float stepdist, int n_rays, int64_t total_len, float* ray_pts,
                        uint8_t* mask_outbbox, int64_t* ray_id, int64_t* step_id,
                        fgs_stream_t stream) {
  FGS_REQUIRE(n_rays >= 0 && total_len >= 0, "negative size");
  if (n_rays == 0 || total_len == 0) return 0;
  const int grid = fgs::stream_grid(n_rays, kBlock / 32, 16);
  sample_fill_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, xyz_min, xyz_max, t_min, (const long long*)ray_offset, stepdist, n_rays,
      total_len, ray_pts, mask_outbbox, (long long*)ray_id, (long long*)step_id);
  return fgs::check_launch("sample_fill");
}

int fgs_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d, const float* xyz_min,
                               const float* xyz_max, int n_samples, int n_rays, float* ray_pts,
                               uint8_t* mask_outbbox, fgs_stream_t stream) {
  FGS_REQUIRE(n_rays >= 0 && n_samples >= 0, "negative size");
  const int64_t total = (int64_t)n_rays * n_samples;
  if (total == 0) return 0;
  sample_ndc_kernel<<<fgs::stream_grid(total, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, xyz_min, xyz_max, n_samples, n_rays, ray_pts, mask_outbbox);
  return fgs::check_launch("sample_ndc");
}

int fgs_sample_bg_pts_on_rays(const float* rays_o, const float* rays_d, const float* t_max,
                              float bg_preserve, int n_samples, int n_rays, float* ray_pts,
                              fgs_stream_t stream) {
  FGS_REQUIRE(n_rays >= 0 && n_samples >= 0, "negative size");
  const int64_t total = (int64_t)n_rays * n_samples;
  if (total == 0) return 0;
  sample_bg_kernel<<<fgs::stream_grid(total, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, t_max, bg_preserve, n_samples, n_rays, ray_pts);
  return fgs::check_launch("sample_bg");
}

int fgs_maskcache_lookup(const uint8_t* world, const float* xyz, const float* xyz2ijk_scale,
                         const float* xyz2ijk_shift, int sz_i, int sz_j, int sz_k,
                         int64_t n_pts, uint8_t* out, fgs_stream_t stream) {
  FGS_REQUIRE(n_pts >= 0, "n_pts < 0");
  if (n_pts == 0) return 0;
  maskcache_lookup_kernel<<<fgs::stream_grid(n_pts, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      world, xyz, xyz2ijk_scale, xyz2ijk_shift, sz_i, sz_j, sz_k, n_pts, out);
  return fgs::check_launch("maskcache_lookup");
}

int fgs_raw2alpha(const float* density, float shift, float interval, const float* interval_pts,
                  int64_t n_pts, float* exp_d, float* alpha, fgs_stream_t stream) {
  FGS_REQUIRE(n_pts >= 0, "n_pts < 0");
  if (n_pts == 0) return 0;
  raw2alpha_kernel<<<fgs::stream_grid(n_pts, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      density, shift, interval, interval_pts, n_pts, exp_d, alpha);
  return fgs::check_launch("raw2alpha");
}

int fgs_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval,
                           const float* interval_pts, int64_t n_pts, float* grad,
                           fgs_stream_t stream) {
  FGS_REQUIRE(n_pts >= 0, "n_pts < 0");
  if (n_pts == 0) return 0;
  raw2alpha_backward_kernel<<<fgs::stream_grid(n_pts, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      exp_d, grad_back, interval, interval_pts, n_pts, grad);
  return fgs::check_launch("raw2alpha_backward");
}

}  // extern "C"
