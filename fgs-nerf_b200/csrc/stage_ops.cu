// Between-step operators of the training loop: progressive grid scaling and the batch gather.
// Both are single-pass streaming kernels: HBM-bound, 128-bit stores, grids sized in multiples of the SM count.
#include "common.cuh"
#include "fine_common.cuh"
#include "../../include/fgs_b200_train.h"

namespace {

constexpr int kBlock = 256;

// ATen area_pixel_compute_scale<float>(in, out, align_corners=true): (float)(in - 1) / (out - 1), 0 if out <= 1
inline float resize_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

struct Axis {
  int i0, ip;        // lower source index, step to the upper one (0 on the last plane)
  float l0, l1;      // weights of the lower / upper sample
};

// upsample_trilinear3d (align_corners): src = scale * dst; i0 = (int)src; l1 = src - i0; l0 = 1 - l1
__device__ __forceinline__ Axis axis_of(int dst, float scale, int in) {
  Axis a;
  const float r = __fmul_rn(scale, (float)dst);
  a.i0 = (int)r;
  a.ip = (a.i0 < in - 1) ? 1 : 0;
  a.l1 = __fsub_rn(r, (float)a.i0);
  a.l0 = __fsub_rn(1.f, a.l1);
  return a;
}

// val = d0 * (h0 * (w0 * v000 + w1 * v001) + h1 * (w0 * v010 + w1 * v011)) + d1 * (same on the upper plane)
__device__ __forceinline__ float blend(const Axis& d, const Axis& h, const Axis& w, float v000, float v001, float v010,
                                       float v011, float v100, float v101, float v110, float v111) {
  const float lo = h.l0 * (w.l0 * v000 + w.l1 * v001) + h.l1 * (w.l0 * v010 + w.l1 * v011);
  const float hi = h.l0 * (w.l0 * v100 + w.l1 * v101) + h.l1 * (w.l0 * v110 + w.l1 * v111);
  return d.l0 * lo + d.l1 * hi;
}

// NCDHW: one thread = VEC consecutive outputs along Z2 of one (c, x2, y2) row.
template <int VEC>
__global__ void __launch_bounds__(kBlock) resize_ncdhw_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int C, int X, int Y, int Z, int X2, int Y2, int Z2,
                                                              float sx, float sy, float sz, long long n_items) {
  const int Zq = Z2 / VEC;
  for (long long it = blockIdx.x * (long long)kBlock + threadIdx.x; it < n_items;
       it += (long long)gridDim.x * kBlock) {
    const int zq = (int)(it % Zq);
    long long r = it / Zq;
    const int y2 = (int)(r % Y2);
    r /= Y2;
    const int x2 = (int)(r % X2);
    const int c = (int)(r / X2);
    const Axis ax = axis_of(x2, sx, X), ay = axis_of(y2, sy, Y);
    const float* p00 = src + (((long long)c * X + ax.i0) * Y + ay.i0) * Z;
    const float* p01 = p00 + (long long)ay.ip * Z;
    const float* p10 = p00 + (long long)ax.ip * Y * Z;
    const float* p11 = p10 + (long long)ay.ip * Z;
    float out[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const Axis az = axis_of(zq * VEC + e, sz, Z);
      const int z0 = az.i0, z1 = az.i0 + az.ip;
      out[e] = blend(ax, ay, az, __ldg(p00 + z0), __ldg(p00 + z1), __ldg(p01 + z0), __ldg(p01 + z1), __ldg(p10 + z0),
                     __ldg(p10 + z1), __ldg(p11 + z0), __ldg(p11 + z1));
    }
    float* o = dst + (((long long)c * X2 + x2) * Y2 + y2) * Z2 + (long long)zq * VEC;
    if (VEC == 4) {
      st4(o, make_float4(out[0], out[1], out[2], out[3]));
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = out[e];
    }
  }
}

// channels-last ([X][Y][Z][C]): one thread = 4 channels of one output voxel; the 8 corners are 128-bit loads.
__global__ void __launch_bounds__(kBlock) resize_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                                                           int X, int Y, int Z, int X2, int Y2, int Z2, float sx,
                                                           float sy, float sz, long long n_items) {
  const int Cq = C >> 2;
  for (long long it = blockIdx.x * (long long)kBlock + threadIdx.x; it < n_items;
       it += (long long)gridDim.x * kBlock) {
    const int cq = (int)(it % Cq);
    long long r = it / Cq;
    const int z2 = (int)(r % Z2);
    r /= Z2;
    const int y2 = (int)(r % Y2);
    const int x2 = (int)(r / Y2);
    const Axis ax = axis_of(x2, sx, X), ay = axis_of(y2, sy, Y), az = axis_of(z2, sz, Z);
    const long long sZ = C, sY = (long long)Z * C, sX = (long long)Y * Z * C;
    const float* p = src + ax.i0 * sX + ay.i0 * sY + az.i0 * sZ + cq * 4;
    const long long dx = ax.ip * sX, dy = ay.ip * sY, dz = az.ip * sZ;
    const float4 v000 = ld4(p), v001 = ld4(p + dz), v010 = ld4(p + dy), v011 = ld4(p + dy + dz);
    const float4 v100 = ld4(p + dx), v101 = ld4(p + dx + dz), v110 = ld4(p + dx + dy), v111 = ld4(p + dx + dy + dz);
    float4 o;
    o.x = blend(ax, ay, az, v000.x, v001.x, v010.x, v011.x, v100.x, v101.x, v110.x, v111.x);
    o.y = blend(ax, ay, az, v000.y, v001.y, v010.y, v011.y, v100.y, v101.y, v110.y, v111.y);
    o.z = blend(ax, ay, az, v000.z, v001.z, v010.z, v011.z, v100.z, v101.z, v110.z, v111.z);
    o.w = blend(ax, ay, az, v000.w, v001.w, v010.w, v011.w, v100.w, v101.w, v110.w, v111.w);
    st4(dst + (((long long)x2 * Y2 + y2) * Z2 + z2) * C + cq * 4, o);
  }
}

// one thread = one (ray, array) pair: 3 floats in, 3 floats out
__global__ void __launch_bounds__(kBlock) gather_rays_kernel(const float* __restrict__ o_tr, const float* __restrict__ d_tr,
                                                             const float* __restrict__ v_tr, const float* __restrict__ c_tr,
                                                             long long n_total, const int64_t* __restrict__ sel, int n,
                                                             float* __restrict__ o, float* __restrict__ d,
                                                             float* __restrict__ v, float* __restrict__ c,
                                                             int32_t* __restrict__ bad) {
  const int t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= 4 * n) return;
  const int which = t / n, i = t - which * n;      // array-major: a warp reads neighbouring sel entries
  const float* in = which == 0 ? o_tr : which == 1 ? d_tr : which == 2 ? v_tr : c_tr;
  float* out = which == 0 ? o : which == 1 ? d : which == 2 ? v : c;
  const long long s = sel[i];
  float a = 0.f, b = 0.f, e = 0.f;
  if (s >= 0 && s < n_total) {
    a = __ldg(in + 3 * s);
    b = __ldg(in + 3 * s + 1);
    e = __ldg(in + 3 * s + 2);
  } else if (bad) {
    *bad = 1;
  }
  out[3 * (long long)i] = a;
  out[3 * (long long)i + 1] = b;
  out[3 * (long long)i + 2] = e;
}

// SDF field on a regular lattice (extract_fields): one thread per lattice point, z fastest (coalesced stores);
// the coordinates come from the three 1-D arrays the caller built with torch.linspace, the sample is ATen's
// grid_sampler_3d arithmetic (ray_math.cuh), so the values equal the reference's query_func bit for bit.
__global__ void __launch_bounds__(kBlock) lattice_sample_kernel(const float* __restrict__ grid, fine::GridGeo g,
                                                                const float* __restrict__ xs, int nx,
                                                                const float* __restrict__ ys, int ny,
                                                                const float* __restrict__ zs, int nz, float sign,
                                                                float* __restrict__ out, long long n) {
  for (long long p = blockIdx.x * (long long)kBlock + threadIdx.x; p < n; p += (long long)gridDim.x * kBlock) {
    const int k = (int)(p % nz);
    const long long r = p / nz;
    const int j = (int)(r % ny), i = (int)(r / ny);
    float iw, ih, id;
    fine::cont_index(g, __ldg(xs + i), __ldg(ys + j), __ldg(zs + k), iw, ih, id);
    const Tri t = tri_setup(iw, ih, id);
    out[p] = __fmul_rn(sign, tri_sample1(grid, t, g.X, g.Y, g.Z));
  }
}

// Dense k^3 convolution of a single-channel volume with replicate padding (nn.Conv3d(1, 1, k, padding=k//2,
// padding_mode='replicate'), the reference's Gaussian smoother).  A CTA computes a 4 x 8 x 32 block of outputs from a
// shared-memory tile with its halo (each input voxel is read from HBM ~1.5-2x instead of k^3 times); the k^3 weights
// travel as a kernel parameter (constant bank, uniform index).  pre_div / post_div fold the reference's
// `m(sdf0 / reduce) / reduce` into the same pass the way torch's CUDA kernel divides a tensor by a python scalar:
// one fp32 reciprocal on the host, a multiplication per element (skipped when the divisor is 1).
constexpr int kConvMaxK = 7;
constexpr int CTX = 4, CTY = 8, CTZ = 32;
struct ConvWeights { float w[kConvMaxK * kConvMaxK * kConvMaxK]; };

template <int K>
__global__ void __launch_bounds__(kBlock) conv3d_replicate_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                 int X, int Y, int Z, ConvWeights cw, float pre_inv,
                                                                 float post_inv, int tiles_y, int tiles_z) {
  constexpr int H = K / 2;
  constexpr int SX = CTX + 2 * H, SY = CTY + 2 * H, SZ = CTZ + 2 * H;
  __shared__ float tile[SX * SY * SZ];
  const int tz = blockIdx.x % tiles_z, ty = (blockIdx.x / tiles_z) % tiles_y, tx = blockIdx.x / (tiles_z * tiles_y);
  const int x0 = tx * CTX, y0 = ty * CTY, z0 = tz * CTZ;
  for (int i = threadIdx.x; i < SX * SY * SZ; i += kBlock) {
    const int lz = i % SZ, ly = (i / SZ) % SY, lx = i / (SZ * SY);
    const int gx = min(max(x0 + lx - H, 0), X - 1), gy = min(max(y0 + ly - H, 0), Y - 1), gz = min(max(z0 + lz - H, 0), Z - 1);
    float v = __ldg(src + ((long long)gx * Y + gy) * Z + gz);
    if (pre_inv != 1.f) v = __fmul_rn(v, pre_inv);
    tile[i] = v;
  }
  __syncthreads();
  const int lz = threadIdx.x % CTZ, ly = threadIdx.x / CTZ;       // 32 x 8 threads, each walks the CTX planes
#pragma unroll 1
  for (int lx = 0; lx < CTX; ++lx) {
    float acc = 0.f;
#pragma unroll 1
    for (int a = 0; a < K; ++a)
#pragma unroll
      for (int b = 0; b < K; ++b)
#pragma unroll
        for (int c = 0; c < K; ++c)
          acc = __fmaf_rn(cw.w[(a * K + b) * K + c], tile[((lx + a) * SY + (ly + b)) * SZ + lz + c], acc);
    const int gx = x0 + lx, gy = y0 + ly, gz = z0 + lz;
    if (gx < X && gy < Y && gz < Z) {
      if (post_inv != 1.f) acc = __fmul_rn(acc, post_inv);
      dst[((long long)gx * Y + gy) * Z + gz] = acc;
    }
  }
}

template <int K>
int launch_conv(const float* src, float* dst, int X, int Y, int Z, const float* w, float pre_div, float post_div,
                cudaStream_t st) {
  ConvWeights cw;
  for (int i = 0; i < K * K * K; ++i) cw.w[i] = w[i];
  const int tx = fgs::cdiv(X, CTX), ty = fgs::cdiv(Y, CTY), tz = fgs::cdiv(Z, CTZ);
  const long long blocks = (long long)tx * ty * tz;
  if (blocks > 0x7fffffffLL) return fgs::fail_arg("conv3d_replicate: volume too large");
  // ATen's div_true_kernel_cuda with a CPU-scalar divisor: inv_b = 1.0f / b, then a * inv_b
  const float pre_inv = pre_div == 1.f ? 1.f : 1.0f / pre_div, post_inv = post_div == 1.f ? 1.f : 1.0f / post_div;
  conv3d_replicate_kernel<K><<<(unsigned)blocks, kBlock, 0, st>>>(src, dst, X, Y, Z, cw, pre_inv, post_inv, ty, tz);
  return fgs::check_launch("conv3d_replicate", st);
}

}  // namespace

extern "C" int fgs_conv3d_replicate(const float* src, float* dst, int X, int Y, int Z, const float* weights, int ksize,
                                    float pre_div, float post_div, fgs_stream_t stream) {
  FGS_REQUIRE(X > 0 && Y > 0 && Z > 0, "conv3d_replicate: sizes must be positive");
  FGS_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5 || ksize == 7, "conv3d_replicate: ksize must be 1, 3, 5 or 7");
  FGS_REQUIRE(src && dst && src != dst && weights, "conv3d_replicate: src / dst must be distinct non-null buffers");
  FGS_REQUIRE(pre_div != 0.f && post_div != 0.f, "conv3d_replicate: zero divisor");
  cudaStream_t st = (cudaStream_t)stream;
  switch (ksize) {
    case 1: return launch_conv<1>(src, dst, X, Y, Z, weights, pre_div, post_div, st);
    case 3: return launch_conv<3>(src, dst, X, Y, Z, weights, pre_div, post_div, st);
    case 5: return launch_conv<5>(src, dst, X, Y, Z, weights, pre_div, post_div, st);
    default: return launch_conv<7>(src, dst, X, Y, Z, weights, pre_div, post_div, st);
  }
}

extern "C" int fgs_grid_sample_lattice(const float* grid, int X, int Y, int Z, const float* xyz_min,
                                       const float* xyz_max, const float* xs, int nx, const float* ys, int ny,
                                       const float* zs, int nz, float sign, float* out, fgs_stream_t stream) {
  FGS_REQUIRE(X > 0 && Y > 0 && Z > 0 && nx >= 0 && ny >= 0 && nz >= 0, "grid_sample_lattice: bad size");
  const long long n = (long long)nx * ny * nz;
  if (n == 0) return 0;
  FGS_REQUIRE(grid && xs && ys && zs && out && xyz_min && xyz_max, "grid_sample_lattice: null pointer");
  fine::GridGeo g;
  for (int c = 0; c < 3; ++c) {
    g.mn[c] = xyz_min[c];
    g.ext[c] = xyz_max[c] - xyz_min[c];  // fp32 subtraction, as torch evaluates (xyz_max - xyz_min)
  }
  g.X = X; g.Y = Y; g.Z = Z;
  cudaStream_t st = (cudaStream_t)stream;
  lattice_sample_kernel<<<fgs::stream_grid(n, kBlock, 8), kBlock, 0, st>>>(grid, g, xs, nx, ys, ny, zs, nz, sign, out, n);
  return fgs::check_launch("grid_sample_lattice", st);
}

extern "C" int fgs_grid_resize_trilinear(const float* src, int C, int X, int Y, int Z, float* dst, int X2, int Y2,
                                         int Z2, int channels_last, fgs_stream_t stream) {
  FGS_REQUIRE(C >= 0 && X > 0 && Y > 0 && Z > 0 && X2 > 0 && Y2 > 0 && Z2 > 0, "grid_resize: sizes must be positive");
  if (C == 0) return 0;
  FGS_REQUIRE(src && dst && src != dst, "grid_resize: src / dst must be distinct non-null buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const float sx = resize_scale(X, X2), sy = resize_scale(Y, Y2), sz = resize_scale(Z, Z2);
  if (channels_last) {
    FGS_REQUIRE(C % 4 == 0, "grid_resize: channels_last needs C % 4 == 0");
    FGS_REQUIRE(((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0, "grid_resize: buffers must be 16-byte aligned");
    const long long n = (long long)X2 * Y2 * Z2 * (C >> 2);
    resize_cl_kernel<<<fgs::stream_grid(n, kBlock, 8), kBlock, 0, st>>>(src, dst, C, X, Y, Z, X2, Y2, Z2, sx, sy, sz, n);
    return fgs::check_launch("grid_resize_cl", st);
  }
  if (Z2 % 4 == 0 && (((uintptr_t)dst) & 15) == 0) {
    const long long n = (long long)C * X2 * Y2 * (Z2 / 4);
    resize_ncdhw_kernel<4><<<fgs::stream_grid(n, kBlock, 8), kBlock, 0, st>>>(src, dst, C, X, Y, Z, X2, Y2, Z2, sx, sy,
                                                                              sz, n);
  } else {
    const long long n = (long long)C * X2 * Y2 * Z2;
    resize_ncdhw_kernel<1><<<fgs::stream_grid(n, kBlock, 8), kBlock, 0, st>>>(src, dst, C, X, Y, Z, X2, Y2, Z2, sx, sy,
                                                                              sz, n);
  }
  return fgs::check_launch("grid_resize", st);
}

extern "C" int fgs_gather_rays(const float* rays_o_tr, const float* rays_d_tr, const float* viewdirs_tr,
                               const float* rgb_tr, int64_t n_total, const int64_t* sel, int n, float* rays_o,
                               float* rays_d, float* viewdirs, float* target, int32_t* bad, fgs_stream_t stream) {
  FGS_REQUIRE(n >= 0 && n_total >= 0, "gather_rays: negative size");
  if (n == 0) return 0;
  FGS_REQUIRE(rays_o_tr && rays_d_tr && viewdirs_tr && rgb_tr && sel && rays_o && rays_d && viewdirs && target,
              "gather_rays: null pointer");
  FGS_REQUIRE(n <= (1 << 28), "gather_rays: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  gather_rays_kernel<<<fgs::cdiv(4ll * n, kBlock), kBlock, 0, st>>>(rays_o_tr, rays_d_tr, viewdirs_tr, rgb_tr, n_total,
                                                                    sel, n, rays_o, rays_d, viewdirs, target, bad);
  return fgs::check_launch("gather_rays", st);
}
